/* dsrt.h -- C ABI of the B200-native path-tracing core (libdsrt.so).
 *
 * This is the drop-in boundary for the reference's GPU render path: it replaces the C++ class
 * CUDAPathTracer of libGPUAccel.so (reference cuda_src/setup.h:90-148), which is driven from
 * Application::startGPURayTracing (reference src/application.cpp:766-786).  Plain pointers and sizes
 * only; every call returns a status code (DSRT_OK == 0) and never terminates the process (the reference
 * prints and exit()s, cuda_src/setup.cu:139-143).  The message for the last failure of a context is
 * available from dsrt_last_error().
 *
 * Call order for one render (mirrors CUDAPathTracer::init, cuda_src/setup.cu:181-201):
 *   dsrt_create -> dsrt_set_scene -> dsrt_set_bvh (or dsrt_build_bvh2 + dsrt_set_bvh) -> dsrt_set_camera
 *   -> dsrt_set_params -> dsrt_build_accel -> dsrt_render / dsrt_render_device -> dsrt_destroy
 *
 * Ownership: the caller owns every host array for the duration of the call only (the library copies);
 * the library owns all device memory; output buffers are caller-allocated.  No cudaDeviceReset, no
 * process-global state: contexts are independent (the reference keeps scene state in __constant__
 * globals, cuda_src/kernel.cu:16-20, and is limited to 20 lights / 20 BSDFs, kernel.cu:3-4).
 * One context is used by one host thread at a time.
 */
#ifndef DSRT_H
#define DSRT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DSRT_OK 0
#define DSRT_ERR_INVALID 1   /* bad argument / call order */
#define DSRT_ERR_CUDA 2      /* CUDA runtime failure (no device, out of memory, launch failure, ...) */
#define DSRT_ERR_LIMIT 3     /* scene exceeds an internal limit */
#define DSRT_CANCELLED 4     /* dsrt_render stopped early by dsrt_cancel: the frame holds the samples rendered so far */

typedef struct dsrt_ctx dsrt_ctx;

/* Flattened static scene = what CUDAPathTracer::loadPrimitives / loadLights read out of
 * PathTracer::primitives, Primitive::get_bsdf() and scene->lights (cuda_src/setup.cu:249-402, 689-774).
 * Primitive i is PathTracer::primitives[i] (object order, src/pathtracer.cpp:230-234) -- that index is
 * the primitive id reported by dsrt_primary_hits. */
typedef struct {
  int32_t n_prims;
  const int32_t* prim_type;   /* 1 triangle, 0 sphere (triangle.h:74, sphere.h:85) */
  const int32_t* prim_bsdf;   /* index into the BSDF table */
  const double* tri_pos;      /* 9 per prim: p1,p2,p3 world space (Triangle v1,v2,v3; object.cpp:36-41) */
  const double* tri_nrm;      /* 9 per prim: vertex normals n1,n2,n3 (halfEdgeMesh.h:492-515) */
  const double* sphere;       /* 4 per prim: centre xyz, radius (sphere.h:98-99) */
  int32_t n_bsdf;
  const int32_t* bsdf_type;   /* 0 diffuse 1 mirror 2 refraction 3 glass 4 emission (bsdf.h:123-236) */
  const float* bsdf_param;    /* 8 per BSDF: a[3] albedo|reflectance|radiance, b[3] transmittance, ior, 0 */
  int32_t n_lights;
  const int32_t* light_type;  /* 0 directional 1 infinite hemisphere 2 point 3 area (light.h:24-99) */
  const double* light_param;  /* 28 per light: radiance[3], dirToLight|position[3], direction[3], dim_x[3],
                                 dim_y[3], area, sampleToWorld[9] column-major at offset 16 */
} dsrt_scene;

/* Host-built binary SAH BVH (BVHAccel, src/bvh.cpp:21-202), flattened; node 0 is the root. */
typedef struct {
  int32_t n_nodes;
  const double* node_bbox;    /* 6 per node: min xyz, max xyz */
  const int32_t* node_start;  /* first slot in prim_order */
  const int32_t* node_range;  /* number of slots */
  const int32_t* node_left;   /* -1 when absent */
  const int32_t* node_right;
  const int32_t* prim_order;  /* n_prims: BVH slot -> primitive id (BVHAccel::primitives after the build) */
} dsrt_bvh2;

/* Counters of one render call (device-side atomics; the metric unit is the path SEGMENT = one BVH query). */
typedef struct {
  uint64_t camera_samples;    /* camera rays generated */
  uint64_t extend_rays;       /* closest-hit queries (BVHAccel::intersect(ray, isect) calls) */
  uint64_t shadow_rays;       /* any-hit queries (BVHAccel::intersect(ray) calls) */
  /* fetch counters, only filled when dsrt_set_option("count_traversal", 1): wide-BVH nodes (80 B of information each) and primitive
   * records (48 B) fetched by the extend (closest-hit) and connect (any-hit) kernels */
  uint64_t extend_nodes, extend_prims;
  uint64_t connect_nodes, connect_prims;
  double gpu_seconds;         /* CUDA-event time of the whole call on the render stream */
  double extend_seconds;      /* summed CUDA-event time of the extend (closest-hit) launches */
  double connect_seconds;     /* summed CUDA-event time of the connect (any-hit) launches */
  double shade_seconds;       /* summed CUDA-event time of generate + shade launches */
  uint32_t kernel_launches;   /* kernels launched by this call */
  uint32_t batches;
  uint64_t null_shadow_rays;  /* option "skip_null_shadow": shadow rays (counted in shadow_rays) that were NOT traced because
                               * their contribution is exactly zero; 0 when the option is off (the reference traces them) */
} dsrt_stats;

/* ---- lifetime ------------------------------------------------------------------------------------ */
int dsrt_create(int device, dsrt_ctx** out);          /* replaces new CUDAPathTracer + init(), setup.cu:77-95,181-201 */
/* One context driving several GPUs of one box from one process (the `pathtracer -g N` path): the scene is
 * replicated, dsrt_render gives GPU r the samples k = r (mod N), and device 0 combines the partial framebuffers
 * by reading its peers' memory over NVLink inside the resolve kernel.  (One process per GPU + one NCCL reduce,
 * as bench.py does under torchrun, uses dsrt_create + dsrt_render_device instead.) */
int dsrt_create_multi(int n_devices, const int* devices, dsrt_ctx** out);
int dsrt_device_count(const dsrt_ctx* ctx);
int dsrt_destroy(dsrt_ctx* ctx);                      /* replaces ~CUDAPathTracer, setup.cu:97-115 */
const char* dsrt_last_error(const dsrt_ctx* ctx);     /* never NULL */
const char* dsrt_version(void);

/* ---- scene upload -------------------------------------------------------------------------------- */
int dsrt_set_scene(dsrt_ctx* ctx, const dsrt_scene* scene);   /* loadPrimitives + loadLights, setup.cu:249-402,689-774 */
int dsrt_set_bvh(dsrt_ctx* ctx, const dsrt_bvh2* bvh);        /* loadBVH, setup.cu:415-476 */
/* pos[3], c2w[9] column-major, frame size, screenDist -- Camera::generate_ray inputs (camera.cpp:113-129);
 * replaces loadCamera, setup.cu:221-247 */
int dsrt_set_camera(dsrt_ctx* ctx, const double* pos, const double* c2w, int32_t width, int32_t height,
                    double screen_dist);
/* Tile partitioning: restricts the render calls that follow to the pixels [x0, x0+width) x [y0, y0+height) of the frame
 * (buffers keep the frame's size; pixels outside stay untouched / zero).  Disjoint windows rendered on different GPUs add up
 * to the frame exactly like disjoint sample sets do.  width == 0 clears it; dsrt_set_camera clears it too.  The reference's
 * own tiling is its WorkQueue of image tiles, src/pathtracer.cpp:585-647. */
int dsrt_set_window(dsrt_ctx* ctx, int32_t x0, int32_t y0, int32_t width, int32_t height);
/* Lat-long environment map (row 0 = +y pole, width*height*3 floats): PathTracer's `envmap` constructor argument
 * (src/pathtracer.cpp:41-45) -> EnvironmentLight (src/static_scene/environment_light.cpp).  The light is appended after
 * the scene's lights; rays that leave the scene with includeLe see the map.  width == height == 0 removes it.
 * Call before dsrt_build_accel. */
int dsrt_set_envmap(dsrt_ctx* ctx, int32_t width, int32_t height, const float* rgb);
/* ns_aa (-s), ns_area_light (-l), max_ray_depth (-m); replaces loadParameters, setup.cu:777-811 */
int dsrt_set_params(dsrt_ctx* ctx, int32_t ns_aa, int32_t ns_area_light, int32_t max_ray_depth, uint32_t seed);
/* named knobs: "count_traversal" (0/1), "batch_spp" (camera samples per pixel per wavefront batch),
 * "stage_timing" (0/1: per-stage CUDA events), "postpone_min_lanes" (primitive tests wait until this many lanes
 * of a warp have some pending; 0 = test at once; default 8), "pool_batches" (how many consecutive batches share one
 * deep-path pool: paths that survive depth 0 are gathered and advanced together; default 8, 1 = per batch),
 * "coop_min_pairs" (any-hit kernel: when a warp has at least this many pending (ray, primitive) pairs they are
 * dealt out one per lane; default 6, a huge value disables the cooperative test), "postpone_wait_mode" (0: primitives are
 * tested as soon as one lane has nothing else to do; 1: only when no lane opened a node; K >= 2: when K lanes wait),
 * "refill_busy_lanes" (a warp fetches new rays for its idle lanes when at most this many are busy; default 18),
 * "refill_hi_lanes" / "refill_patience" (... or already when at most refill_hi_lanes (26) are busy, once the warp has run
 * refill_patience (6) traversal iterations since its last refill: long rays -- tens of node visits in the triangle soups --
 * make an early refill worth its fixed cost, short rays never get there),
 * "max_ctas_per_sm" (caps the persistent grid; 0 = what fits), "smem_carveout_pct" (shared-memory carve-out of the
 * traversal kernels, -1 = driver default, which measured best), "collapse_prim_cost_pct" (SAH cost of a primitive test
 * relative to a wide-node visit in the collapse, percent; default 100; applies at the next dsrt_build_accel),
 * "drop_coplanar_mates" (0/1, default 1; applies at the next dsrt_build_accel: leaf slots whose two or three triangles lie
 * in one plane -- the halves of a wall quad -- are marked in the node, and a ray that starts on one of them skips the others
 * like it skips its source: it meets their plane at t = 0 only),
 * "regroup_top" (0/1, default 0; host builder only; applies at the next dsrt_build_accel: where an internal child of one of the top wide nodes
 * covers most of its parent -- the reference's binned SAH leaves the scene-sized wall triangles of a Cornell box in a subtree
 * whose box is the whole scene, which every ray then has to open -- its children and its siblings are regrouped so that the
 * summed area of the internal nodes drops: walls become direct children of the root, the mesh gets a node of its own; hits are
 * unchanged, node visits per segment fall by 9-26 % on the Cornell scenes; measured on B200 the closest-hit stage gains 1-8 %
 * and the any-hit stage loses 2 %, -1.2 % / -0.2 % per frame on the two 1080p stand-ins: the rays that cross the mesh get one
 * level more and a warp walks the union of its lanes' nodes, so the visits saved on the short rays do not shorten it),
 * "light_aligned_grid" (0/1, default 1; applies at the next dsrt_build_accel: a wide node that holds a flat child in the
 * plane of an axis-aligned area light shifts its quantisation grid by a fraction of a quantum so that the plane facing the
 * arriving shadow rays is tight -- they stop 0.1 % short of the light, src/pathtracer.cpp:486-504, and then miss the box of
 * the emissive quad under the light instead of testing its triangles; results are unchanged, boxes stay conservative),
 * "skip_null_shadow" (0/1, default 0: the reference traces every shadow ray before it evaluates the BSDF and the cosine,
 * src/pathtracer.cpp:497-519; 1 = shadow rays whose contribution is exactly zero -- light behind the surface, non-diffuse
 * BSDF, emitter facing away -- keep their queue slot but are not traced; the image is identical, dsrt_stats.null_shadow_rays
 * says how many), "wavefront_budget_mb" (cap on the wavefront + pool memory, 0 = 80 % of the free device memory: the batch
 * and the pool group shrink to fit, whatever -l asks for), "device_build" (0/1, default 0; applies at the next
 * dsrt_build_accel, which then needs no dsrt_set_bvh: Morton codes, radix sort, Karras' binary radix tree, bottom-up boxes and
 * the collapse to the same 8-wide layout all run on the first GPU -- what the reference's disabled PARALLEL_BUILD_BVH path set
 * out to do, cuda_src/setup.cu:478-686 -- for scenes of tens of millions of primitives; hit results do not depend on the builder) */
int dsrt_set_option(dsrt_ctx* ctx, const char* name, int64_t value);

/* Host SAH builder = BVHAccel::BVHAccel + buildBVH (src/bvh.cpp:21-202: 32 buckets, max leaf 4, with the
 * bucket-index clamp).  Arrays must hold 2*n_prims nodes / n_prims slots; returns the node count in *n_nodes. */
int dsrt_build_bvh2(const dsrt_scene* scene, double* node_bbox, int32_t* node_start, int32_t* node_range,
                    int32_t* node_left, int32_t* node_right, int32_t* prim_order, int32_t* n_nodes);

/* Collapse the binary BVH into the compressed 8-wide SoA BVH, reorder primitives leaf-contiguously,
 * upload.  Outputs sizes for the roofline accounting (may be NULL). */
int dsrt_build_accel(dsrt_ctx* ctx);
/* Re-sends the flattened scene built by dsrt_build_accel to the GPU(s): a pure host->device copy (what loadPrimitives /
 * loadBVH / loadLights do with cudaMemcpy in the reference, setup.cu:375-402, 415-476, 745-774). */
int dsrt_upload_accel(dsrt_ctx* ctx);
int dsrt_accel_bytes(const dsrt_ctx* ctx, int64_t* h2d_bytes);
int dsrt_accel_info(const dsrt_ctx* ctx, int64_t* n_wide_nodes, int64_t* node_bytes, int64_t* prim_bytes,
                    int32_t* max_depth);

/* ---- rendering ----------------------------------------------------------------------------------- */
/* Renders camera samples spp_begin, spp_begin+spp_stride, ... (spp_count of them) of every pixel and writes
 * the radiance SUM scaled by 1/ns_aa (so a full render with spp_count == ns_aa equals the reference's
 * sampleBuffer, src/pathtracer.cpp:571-581).  rgb_out: host, width*height*3 floats, row 0 = bottom
 * (image.h:113-117).  Replaces startRayTracingPT + updateHostSampleBuffer, setup.cu:147-179, 813-827. */
int dsrt_render(dsrt_ctx* ctx, int32_t spp_begin, int32_t spp_count, int32_t spp_stride, float* rgb_out,
                dsrt_stats* stats);
/* Same frame, plus the tone-mapped RGBA8 image (HDRImageBuffer::toColor, image.h:174-189: 0xAABBGGRR, row 0 = bottom) from the
 * fused reduce + resolve kernel; rgb_out may be NULL.  Replaces updateHostSampleBuffer + PathTracer::updateBufferFromGPU,
 * setup.cu:813-843. */
int dsrt_render_tonemapped(dsrt_ctx* ctx, int32_t spp_begin, int32_t spp_count, int32_t spp_stride, float* rgb_out,
                           uint32_t* rgba8_out, dsrt_stats* stats);
/* PathTracer::stop (src/pathtracer.cpp:148-171): callable from another thread while dsrt_render runs.  The render stops after
 * the chunk of samples in flight (about 0.2 s of GPU work at 1080p), returns DSRT_CANCELLED, and its frame is the mean of the
 * samples rendered so far (dsrt_stats.camera_samples / pixels).  The flag is cleared at the start of every dsrt_render. */
int dsrt_cancel(dsrt_ctx* ctx);
/* Same, but ADDS un-normalised radiance sums into a DEVICE buffer (width*height*3 floats) on the given
 * cudaStream_t (NULL = the context's stream) without synchronising: the multi-GPU path reduces these
 * partial sums with one NCCL reduce and then calls dsrt_resolve_device. */
int dsrt_render_device(dsrt_ctx* ctx, int32_t spp_begin, int32_t spp_count, int32_t spp_stride,
                       float* d_accum, void* stream, dsrt_stats* stats);
/* d_rgb = d_accum / ns_aa; optional RGBA8 tone map (HDRImageBuffer::toColor, image.h:174-189) when d_rgba8 != NULL */
int dsrt_resolve_device(dsrt_ctx* ctx, const float* d_accum, float* d_rgb, uint32_t* d_rgba8, void* stream);
int dsrt_sync(dsrt_ctx* ctx);
/* Fills *stats from the counters/events of the last dsrt_render_device call (synchronises). */
int dsrt_collect_stats(dsrt_ctx* ctx, dsrt_stats* stats);

/* Primary closest hits at pixel centres (px=(x+.5)/w): primitive id (-1 miss) and t.  mode 0 = production
 * float kernel; mode 1 = parity kernel (same wide-BVH traversal, conservative box tests, fp64 leaf tests in the
 * reference's operation order: triangle.cpp:55-104, sphere.cpp:10-77) -- bit-exact vs BVHAccel::intersect. */
int dsrt_primary_hits(dsrt_ctx* ctx, int32_t mode, int32_t* prim_id, double* t);

/* Arbitrary ray batches through the production kernels (tests / roofline counting):
 * o,d: n*3 floats; tmax: n floats (NULL = inf).  closest: prim_id (-1 miss), t, u, v.  any: hit 0/1. */
int dsrt_trace_closest(dsrt_ctx* ctx, int64_t n, const float* o, const float* d, const float* tmax,
                       int32_t* prim_id, float* t);
int dsrt_trace_any(dsrt_ctx* ctx, int64_t n, const float* o, const float* d, const float* tmax, int32_t* hit);

/* HDRImageBuffer::toColor + ImageBuffer::update_pixel on the host result (image.h:49-58,174-189). */
int dsrt_tonemap(dsrt_ctx* ctx, const float* rgb, int64_t n_pixels, uint32_t* rgba8);

/* Roofline denominator for the L2-resident regime: every SM sweeps the same read-only buffer of `bytes`
 * (rounded down to 16 B) `repeats` times with 128-bit loads; *gb_per_s = bytes * repeats / kernel time (CUDA events,
 * first sweep excluded).  A 32 MiB buffer stays in the B200's L2, so the figure is the L2 -> SM read bandwidth; a buffer
 * several times the L2 size gives the HBM read bandwidth. */
int dsrt_measure_read_bandwidth(dsrt_ctx* ctx, int64_t bytes, int32_t repeats, double* gb_per_s);

#ifdef __cplusplus
}
#endif
#endif /* DSRT_H */
