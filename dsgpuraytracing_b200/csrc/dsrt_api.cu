// dsrt_api.cu -- the C ABI (include/dsrt.h): contexts, host copies of the inputs, flattening and upload of the scene, and the
// launch sequence of the wavefront (render_impl).  The kernels are in kernels.cuh, the traversal / shading device code in
// traverse.cuh / shade.cuh, the host builders in bvh2_sah.cpp / wide_bvh.cpp.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dsrt.h"
#include "kernels.cuh"
#include "wide_bvh.h"
#include "device_build.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <nvtx3/nvToolsExt.h>          // header-only: ranges show up in nsys / ncu timelines, no-ops otherwise

// ================================================================================================ host side
using namespace dsrt;

// everything that lives on ONE GPU
struct DevState {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int trace_blocks = 148 * 6;
  void* d_nodes = nullptr; void* d_prims = nullptr; void* d_shade = nullptr; void* d_prims64 = nullptr;
  void* d_bsdf = nullptr; void* d_lights = nullptr;
  float* d_env_rgb = nullptr; float* d_env_tp = nullptr; float* d_env_t = nullptr; float* d_env_pgt = nullptr;
  size_t cap_paths = 0, cap_shadow = 0;
  PathState ps{}; ShadowQueue sq{};
  uint32_t* queue[2] = {nullptr, nullptr};
  size_t cap_pool = 0, cap_pool_shadow = 0;
  PathState pool{}; ShadowQueue pool_sq{}; uint32_t* pool_queue[2] = {nullptr, nullptr};
  PoolCounters* d_pool_counters = nullptr; int n_pool_counter_blocks = 0;
  Counters* d_counters = nullptr; int n_counter_blocks = 0;
  Totals* d_totals = nullptr;
  float* d_accum_own = nullptr; size_t accum_pixels = 0;
  float* d_stage = nullptr; size_t stage_floats = 0;      // device 0 only: staging for peers without P2P access
  uint32_t* d_rgba8 = nullptr; size_t rgba8_pixels = 0;   // device 0 only: tone-mapped frame of dsrt_render_tonemapped
  bool peer_ok = true;                                    // device 0 can read this device's memory directly
  std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
  struct Span { size_t e0, e1; int kind; };
  std::vector<Span> spans;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  uint32_t launches = 0, batches = 0;
  int64_t carveout_set = -1;
};

struct dsrt_ctx {
  std::vector<DevState> devs;
  std::string err;
  std::mutex err_mutex;                   // dsrt_render drives its devices from one host thread each
  std::atomic<int> cancel{0};             // dsrt_cancel: checked by dsrt_render between chunks of samples
  // host copies of the inputs
  bool have_scene = false, have_bvh = false, have_cam = false, have_accel = false;
  int n_prims = 0;
  PodVec<int32_t> prim_type, prim_bsdf;          // filled by a parallel copy (64 Mi triangles = 11.8 GB of caller arrays)
  PodVec<double> tri_pos, tri_nrm, sphere;
  std::vector<Bsdf> bsdfs;
  std::vector<int32_t> light_type;
  std::vector<double> light_param;
  std::vector<double> node_bbox;
  std::vector<int32_t> node_start, node_range, node_left, node_right, prim_order;
  Camera cam{};
  int ns_aa = 1, ns_area_light = 4, max_depth = 1;
  uint32_t seed = 0;
  int64_t opt_count = 0, opt_batch_spp = 0, opt_stage_timing = 0, opt_skip_null = 0, opt_tri_min = 8, opt_refill = 18, opt_refill_hi = 26, opt_refill_patience = 6, opt_wait_mode = 0, opt_pool_batches = 8, opt_coop_min = 6, opt_max_ctas = 0, opt_carveout = -1, opt_prim_cost = 100, opt_mem_budget_mb = 0, opt_device_build = 0, opt_light_phase = 1, opt_regroup = 0, opt_flat_slots = 1;
  WideBVH wide;
  std::vector<PrimRecord> recs; std::vector<ShadeRecord> shd; std::vector<PrimRecord64> r64; std::vector<Light> lights;
  int env_w = 0, env_h = 0;
  std::vector<float> env_rgb, env_tp, env_t, env_pgt;
  double scene_diag = 1.0;
  bool recs_on_device = false;             // device_build: the primitive / shading records live on the first GPU only (no host copy)
  int win[4] = {0, 0, 0, 0};               // dsrt_set_window: x0, y0, width, height (width 0 = the whole frame)
  float bsphere[4] = {0.f, 0.f, 0.f, 0.f};   // centre + radius of a sphere around all primitives (Accel::bcx..brad)
  int n_lights = 0, n_light_samples = 0;
};

namespace {

// DSRT_HOST_TRACE=1: host-side wall time between named points of a dsrt_render call, printed to stderr at its end (where an
// end-to-end step spends time outside the GPU window that dsrt_stats.gpu_seconds reports)
struct HostTrace { bool on = false; std::vector<std::pair<const char*, double>> pts; };
thread_local HostTrace g_host_trace;
inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define TP(name) do { if (g_host_trace.on) g_host_trace.pts.emplace_back(name, now_s()); } while (0)

int fail(dsrt_ctx* c, int code, const std::string& msg) {
  if (c) { std::lock_guard<std::mutex> g(c->err_mutex); c->err = msg; }
  return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, DSRT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

template <typename T> int dev_alloc(dsrt_ctx* ctx, T** p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (n == 0) n = 1;
  CK(cudaMalloc((void**)p, n * sizeof(T)));
  return DSRT_OK;
}
void dev_free(void* p) { if (p) cudaFree(p); }

// scratch device allocations / events of one API call: released on every return path
struct Scratch {
  std::vector<void*> ptrs; std::vector<cudaEvent_t> events;
  ~Scratch() { for (void* p : ptrs) cudaFree(p); for (cudaEvent_t e : events) cudaEventDestroy(e); }
  template <typename T> cudaError_t alloc(T** p, size_t bytes) { cudaError_t e = cudaMalloc((void**)p, bytes); if (e == cudaSuccess) ptrs.push_back(*p); return e; }
  cudaError_t event(cudaEvent_t* ev) { cudaError_t e = cudaEventCreate(ev); if (e == cudaSuccess) events.push_back(*ev); return e; }
};

cudaEvent_t next_event(DevState& D) {
  if (D.ev_used == D.ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); D.ev_pool.push_back(e); }
  return D.ev_pool[D.ev_used++];
}

Accel make_accel(const dsrt_ctx* ctx, const DevState& D, bool parity) {
  Accel A;
  A.nodes = (const uint4*)D.d_nodes; A.prims = (const float4*)D.d_prims;
  A.prims64 = (const double*)D.d_prims64;
  A.pad = parity ? (float)(1e-5 * ctx->scene_diag) : 0.f;
  A.one_bits = 0x3f800000u;
  A.bcx = ctx->bsphere[0]; A.bcy = ctx->bsphere[1]; A.bcz = ctx->bsphere[2]; A.brad = ctx->bsphere[3];
  return A;
}

// shared-memory traversal stack: one node group per wide-BVH level per lane (whatever is not used stays L1 cache)
int stack_entries(const dsrt_ctx* ctx) { return std::max(ctx->wide.max_depth, 1) + DSRT_STACK_SLACK; }
// dynamic shared memory of k_trace: traversal stacks + the lanes' ray blocks; the any-hit kernel adds pair tables and hit flags
size_t stack_bytes(const dsrt_ctx* ctx, bool any = true) {
  const size_t stack = (size_t)stack_entries(ctx) * kTraceThreads * sizeof(uint2);
  if (!any) return stack + (size_t)kRayBlockClosest * kTraceThreads * sizeof(float);
  return stack + (size_t)kRayBlock * kTraceThreads * sizeof(float) +
         (size_t)(kTraceThreads / 32) * kPairCap * kPairBytes + kTraceThreads + (kTraceThreads / 32) * sizeof(uint32_t);
}

// persistent grid = resident CTAs per SM (registers / shared-memory stack) x SM count
int size_trace_grid(dsrt_ctx* ctx, DevState& D) {
  if (ctx->opt_carveout != D.carveout_set) {      // experiment knob: shared-memory carve-out (percent of the 228 KB maximum)
    const int pct = (int)ctx->opt_carveout;
    CK(cudaFuncSetAttribute(k_trace<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    CK(cudaFuncSetAttribute(k_trace<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    D.carveout_set = ctx->opt_carveout;
  }
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace<true, false>, kTraceThreads, stack_bytes(ctx)));
  if (ctx->opt_max_ctas > 0) per_sm = std::min(per_sm, (int)ctx->opt_max_ctas);
  D.trace_blocks = D.sm_count * std::max(per_sm, 1);
  return DSRT_OK;
}

int init_device(dsrt_ctx* ctx, DevState& D, int device) {
  D.device = device;
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  D.sm_count = prop.multiProcessorCount;
  CK(cudaMalloc((void**)&D.d_totals, sizeof(Totals)));
  CK(cudaEventCreate(&D.ev_begin)); CK(cudaEventCreate(&D.ev_end));
  return DSRT_OK;
}

void free_device(DevState& D) {
  cudaSetDevice(D.device);
  if (D.stream) cudaStreamSynchronize(D.stream);
  dev_free(D.d_nodes); dev_free(D.d_prims); dev_free(D.d_shade); dev_free(D.d_prims64); dev_free(D.d_bsdf); dev_free(D.d_lights);
  dev_free(D.d_env_rgb); dev_free(D.d_env_tp); dev_free(D.d_env_t); dev_free(D.d_env_pgt);
  dev_free(D.ps.ray_o); dev_free(D.ps.ray_d); dev_free(D.ps.hit); dev_free(D.ps.thr); dev_free(D.ps.pixel); dev_free(D.ps.sample);
  dev_free(D.queue[0]); dev_free(D.queue[1]); dev_free(D.sq.a); dev_free(D.sq.b); dev_free(D.sq.c);
  dev_free(D.pool.ray_o); dev_free(D.pool.ray_d); dev_free(D.pool.hit); dev_free(D.pool.thr); dev_free(D.pool.pixel); dev_free(D.pool.sample);
  dev_free(D.pool_queue[0]); dev_free(D.pool_queue[1]); dev_free(D.pool_sq.a); dev_free(D.pool_sq.b); dev_free(D.pool_sq.c); dev_free(D.d_pool_counters);
  dev_free(D.d_counters); dev_free(D.d_totals); dev_free(D.d_accum_own); dev_free(D.d_stage); dev_free(D.d_rgba8);
  for (cudaEvent_t e : D.ev_pool) cudaEventDestroy(e);
  if (D.ev_begin) cudaEventDestroy(D.ev_begin);
  if (D.ev_end) cudaEventDestroy(D.ev_end);
  if (D.stream) cudaStreamDestroy(D.stream);
}

int ensure_wavefront(dsrt_ctx* ctx, DevState& D, size_t paths, size_t shadow) {
  if (paths > D.cap_paths) {
    int rc;
    if ((rc = dev_alloc(ctx, &D.ps.ray_o, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.ps.ray_d, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.ps.hit, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.ps.thr, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.ps.pixel, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.ps.sample, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.queue[0], paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.queue[1], paths))) return rc;
    D.cap_paths = paths;
  }
  if (shadow > D.cap_shadow) {
    int rc;
    if ((rc = dev_alloc(ctx, &D.sq.a, shadow))) return rc;
    if ((rc = dev_alloc(ctx, &D.sq.b, shadow))) return rc;
    if ((rc = dev_alloc(ctx, &D.sq.c, shadow))) return rc;
    D.cap_shadow = shadow;
  }
  return DSRT_OK;
}

int ensure_pool(dsrt_ctx* ctx, DevState& D, size_t paths, size_t shadow) {
  if (paths > D.cap_pool) {
    int rc;
    if ((rc = dev_alloc(ctx, &D.pool.ray_o, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.pool.ray_d, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.pool.hit, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.pool.thr, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.pool.pixel, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.pool.sample, paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.pool_queue[0], paths))) return rc;
    if ((rc = dev_alloc(ctx, &D.pool_queue[1], paths))) return rc;
    D.cap_pool = paths;
  }
  if (shadow > D.cap_pool_shadow) {
    int rc;
    if ((rc = dev_alloc(ctx, &D.pool_sq.a, shadow))) return rc;
    if ((rc = dev_alloc(ctx, &D.pool_sq.b, shadow))) return rc;
    if ((rc = dev_alloc(ctx, &D.pool_sq.c, shadow))) return rc;
    D.cap_pool_shadow = shadow;
  }
  return DSRT_OK;
}

int create_impl(int n, const int* devices, dsrt_ctx** out) {
  if (!out || n < 1 || n > 16 || !devices) return DSRT_ERR_INVALID;
  *out = nullptr;
  dsrt_ctx* ctx = new dsrt_ctx();
  ctx->devs.resize(n);
  *out = ctx;     // returned even on failure so the caller can read the message; every later call fails loudly
  for (int i = 0; i < n; i++) {
    int rc = init_device(ctx, ctx->devs[i], devices[i]);
    if (rc) { ctx->err = "dsrt_create: " + ctx->err; return rc; }
  }
  // let device 0 read its peers' partial framebuffers directly (fused reduce + resolve)
  for (int i = 1; i < n; i++) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, ctx->devs[0].device, ctx->devs[i].device);
    if (can) {
      cudaSetDevice(ctx->devs[0].device);
      cudaError_t e = cudaDeviceEnablePeerAccess(ctx->devs[i].device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
      cudaGetLastError();
    }
    ctx->devs[i].peer_ok = can != 0;
  }
  return DSRT_OK;
}

}  // namespace

extern "C" {

int dsrt_upload_accel(dsrt_ctx* ctx);

const char* dsrt_version(void) { return "dsrt 0.2 (sm_100a wavefront path tracer)"; }

int dsrt_create(int device, dsrt_ctx** out) { return create_impl(1, &device, out); }
int dsrt_create_multi(int n_devices, const int* devices, dsrt_ctx** out) { return create_impl(n_devices, devices, out); }
int dsrt_device_count(const dsrt_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int dsrt_destroy(dsrt_ctx* ctx) {
  if (!ctx) return DSRT_ERR_INVALID;
  for (DevState& D : ctx->devs) free_device(D);
  delete ctx;
  return DSRT_OK;
}

const char* dsrt_last_error(const dsrt_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int dsrt_set_scene(dsrt_ctx* ctx, const dsrt_scene* s) {
  if (!ctx || !s) return DSRT_ERR_INVALID;
  // a failed call must not leave a half-updated context usable: nothing is valid until this call succeeds, and every
  // input is checked before the first byte of ctx changes
  ctx->have_scene = false; ctx->have_bvh = false; ctx->have_accel = false;
  if (s->n_prims < 0 || s->n_bsdf < 0 || s->n_lights < 0) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_scene: negative count");
  if (s->n_prims > 0 && (!s->prim_type || !s->prim_bsdf || !s->tri_pos || !s->tri_nrm || !s->sphere))
    return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_scene: null primitive array");
  if (s->n_bsdf > 0 && (!s->bsdf_type || !s->bsdf_param)) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_scene: null BSDF array");
  if (s->n_lights > 0 && (!s->light_type || !s->light_param)) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_scene: null light array");
  const size_t n = (size_t)s->n_prims;
  for (size_t i = 0; i < n; i++) {
    if (s->prim_bsdf[i] < 0 || s->prim_bsdf[i] >= s->n_bsdf) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_scene: prim_bsdf out of range");
    if (s->prim_type[i] != 0 && s->prim_type[i] != 1) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_scene: prim_type must be 0 or 1");
  }
  for (int i = 0; i < s->n_bsdf; i++)
    if (s->bsdf_type[i] < 0 || s->bsdf_type[i] > 4) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_scene: unknown BSDF type");
  for (int i = 0; i < s->n_lights; i++)
    if (s->light_type[i] < 0 || s->light_type[i] > 3) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_scene: unsupported light type (spot/sphere/mesh lights are empty stubs in the reference, light.cpp:61-115)");
  ctx->n_prims = s->n_prims;
  assign_parallel(ctx->prim_type, s->prim_type, n);
  assign_parallel(ctx->prim_bsdf, s->prim_bsdf, n);
  assign_parallel(ctx->tri_pos, s->tri_pos, 9 * n);
  assign_parallel(ctx->tri_nrm, s->tri_nrm, 9 * n);
  assign_parallel(ctx->sphere, s->sphere, 4 * n);
  ctx->bsdfs.resize(s->n_bsdf);
  for (int i = 0; i < s->n_bsdf; i++) {
    Bsdf& b = ctx->bsdfs[i]; const float* q = s->bsdf_param + 8 * i;
    for (int k = 0; k < 3; k++) { b.a[k] = q[k]; b.b[k] = q[3 + k]; }
    b.ior = q[6]; b.type = s->bsdf_type[i];
  }
  ctx->light_type.assign(s->light_type, s->light_type + s->n_lights);
  ctx->light_param.assign(s->light_param, s->light_param + 28 * (size_t)s->n_lights);
  ctx->have_scene = true;
  return DSRT_OK;
}

int dsrt_set_bvh(dsrt_ctx* ctx, const dsrt_bvh2* b) {
  if (!ctx || !b) return DSRT_ERR_INVALID;
  if (!ctx->have_scene) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_bvh: call dsrt_set_scene first");
  ctx->have_bvh = false; ctx->have_accel = false;
  if (b->n_nodes < 0) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_bvh: negative node count");
  if (b->n_nodes > 0 && (!b->node_bbox || !b->node_start || !b->node_range || !b->node_left || !b->node_right))
    return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_bvh: null node array");
  if (ctx->n_prims > 0 && !b->prim_order) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_bvh: null prim_order");
  std::vector<char> seen(ctx->n_prims, 0);
  for (int i = 0; i < ctx->n_prims; i++) {
    const int p = b->prim_order[i];
    if (p < 0 || p >= ctx->n_prims || seen[p]) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_bvh: prim_order is not a permutation");
    seen[p] = 1;
  }
  const size_t m = (size_t)b->n_nodes;
  ctx->node_bbox.assign(b->node_bbox, b->node_bbox + 6 * m);
  ctx->node_start.assign(b->node_start, b->node_start + m);
  ctx->node_range.assign(b->node_range, b->node_range + m);
  ctx->node_left.assign(b->node_left, b->node_left + m);
  ctx->node_right.assign(b->node_right, b->node_right + m);
  ctx->prim_order.assign(b->prim_order, b->prim_order + ctx->n_prims);
  ctx->have_bvh = true;
  return DSRT_OK;
}

int dsrt_set_camera(dsrt_ctx* ctx, const double* pos, const double* c2w, int32_t width, int32_t height, double screen_dist) {
  if (!ctx || !pos || !c2w) return DSRT_ERR_INVALID;
  if (width <= 0 || height <= 0 || !(screen_dist > 0)) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_camera: bad frame size / screenDist");
  Camera& c = ctx->cam;
  for (int k = 0; k < 3; k++) { c.pos[k] = (float)pos[k]; c.pos64[k] = pos[k]; }
  for (int k = 0; k < 9; k++) { c.c2w[k] = (float)c2w[k]; c.c2w64[k] = c2w[k]; }
  c.width = width; c.height = height;
  c.W64 = width; c.H64 = height; c.dist64 = screen_dist;
  c.w_over_dist = (float)(width / screen_dist); c.h_over_dist = (float)(height / screen_dist);
  ctx->have_cam = true;
  ctx->win[0] = ctx->win[1] = ctx->win[2] = ctx->win[3] = 0;       // a new frame size: back to the whole frame
  return DSRT_OK;
}

// Tile partitioning (north_star: "disjoint slice of samples per pixel (or image tiles)"): the render calls that follow only
// generate camera samples for the pixels of [x0, x0+width) x [y0, y0+height); buffers keep the frame's size and indexing, so
// partial frames of disjoint windows add up exactly like partial frames of disjoint sample sets.  width == 0 clears the window.
int dsrt_set_window(dsrt_ctx* ctx, int32_t x0, int32_t y0, int32_t width, int32_t height) {
  if (!ctx) return DSRT_ERR_INVALID;
  if (!ctx->have_cam) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_window: call dsrt_set_camera first");
  if (width == 0) { ctx->win[0] = ctx->win[1] = ctx->win[2] = ctx->win[3] = 0; return DSRT_OK; }
  if (x0 < 0 || y0 < 0 || width < 0 || height <= 0 || x0 + width > ctx->cam.width || y0 + height > ctx->cam.height)
    return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_window: window outside the frame");
  ctx->win[0] = x0; ctx->win[1] = y0; ctx->win[2] = width; ctx->win[3] = height;
  return DSRT_OK;
}

int dsrt_set_params(dsrt_ctx* ctx, int32_t ns_aa, int32_t ns_area_light, int32_t max_ray_depth, uint32_t seed) {
  if (!ctx) return DSRT_ERR_INVALID;
  if (ns_aa < 1 || ns_area_light < 1 || max_ray_depth < 0) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_params: ns_aa/ns_area_light must be >= 1, max_ray_depth >= 0");
  if (max_ray_depth + 1 >= kMaxDepthSlots) return fail(ctx, DSRT_ERR_LIMIT, "dsrt_set_params: max_ray_depth too large");
  if (ns_area_light != ctx->ns_area_light) ctx->have_accel = false;   // light sample layout is baked at build_accel
  ctx->ns_aa = ns_aa; ctx->ns_area_light = ns_area_light; ctx->max_depth = max_ray_depth; ctx->seed = seed;
  return DSRT_OK;
}

// PathTracer's envmap constructor argument (pathtracer.cpp:41-45): the EnvironmentLight is appended after the scene's
// lights and also lights camera / specular rays that miss the scene.  width == 0 removes it.
int dsrt_set_envmap(dsrt_ctx* ctx, int32_t width, int32_t height, const float* rgb) {
  if (!ctx) return DSRT_ERR_INVALID;
  if (width < 0 || height < 0 || ((width > 0) != (height > 0)) || (width > 0 && !rgb)) return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_envmap: bad size / null data");
  ctx->env_w = width; ctx->env_h = height;
  if (width > 0) {
    ctx->env_rgb.assign(rgb, rgb + (size_t)width * height * 3);
    build_env_tables(width, height, ctx->env_rgb.data(), ctx->env_tp, ctx->env_t, ctx->env_pgt);
  } else { ctx->env_rgb.clear(); ctx->env_tp.clear(); ctx->env_t.clear(); ctx->env_pgt.clear(); }
  ctx->have_accel = false;
  return DSRT_OK;
}

int dsrt_set_option(dsrt_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return DSRT_ERR_INVALID;
  std::string n(name);
  if (n == "count_traversal") ctx->opt_count = value;
  else if (n == "batch_spp") ctx->opt_batch_spp = value;
  else if (n == "stage_timing") ctx->opt_stage_timing = value;
  else if (n == "skip_null_shadow") ctx->opt_skip_null = value;
  else if (n == "postpone_min_lanes") ctx->opt_tri_min = value;
  else if (n == "refill_busy_lanes") ctx->opt_refill = value;
  else if (n == "refill_hi_lanes") ctx->opt_refill_hi = value;
  else if (n == "refill_patience") ctx->opt_refill_patience = std::max<int64_t>(1, value);
  else if (n == "postpone_wait_mode") ctx->opt_wait_mode = value;
  else if (n == "pool_batches") ctx->opt_pool_batches = std::max<int64_t>(1, value);
  else if (n == "coop_min_pairs") ctx->opt_coop_min = value;
  else if (n == "max_ctas_per_sm") ctx->opt_max_ctas = value;
  // the three tree options take effect at the next dsrt_build_accel (a changed value invalidates the current structure)
  else if (n == "drop_coplanar_mates") { if (ctx->opt_flat_slots != (value != 0)) ctx->have_accel = false; ctx->opt_flat_slots = value != 0; }
  else if (n == "regroup_top") { if (ctx->opt_regroup != (value != 0)) ctx->have_accel = false; ctx->opt_regroup = value != 0; }
  else if (n == "light_aligned_grid") { if (ctx->opt_light_phase != (value != 0)) ctx->have_accel = false; ctx->opt_light_phase = value != 0; }
  else if (n == "collapse_prim_cost_pct") { ctx->opt_prim_cost = std::max<int64_t>(1, value); ctx->have_accel = false; }   // takes effect at the next dsrt_build_accel
  else if (n == "device_build") { ctx->opt_device_build = value; ctx->have_accel = false; }   // takes effect at the next dsrt_build_accel
  else if (n == "wavefront_budget_mb") ctx->opt_mem_budget_mb = value;   // 0: 80 % of the free device memory; > 0: additional cap (tests)
  else if (n == "smem_carveout_pct") ctx->opt_carveout = value;        // -1: driver default          // 0: whatever fits
  else return fail(ctx, DSRT_ERR_INVALID, "dsrt_set_option: unknown option " + n);
  return DSRT_OK;
}


// Option "device_build": the wide BVH, the leaf-contiguous primitive order and the primitive / shading records are built on the
// first device (device_build.cuh) from the caller's scene arrays and read back into the same host containers the host SAH path
// fills, so everything downstream (uploads to every GPU, parity records, statistics) is shared.
static int build_wide_bvh_device(dsrt_ctx* ctx, DevState& D, float scene_box[6]) {
  const int n = ctx->n_prims;
  const bool timing = std::getenv("DSRT_BUILD_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  CK(cudaSetDevice(D.device));
  cudaStream_t st = D.stream;
  Scratch tmp;
  int32_t *d_type = nullptr, *d_bsdf = nullptr; double *d_pos = nullptr, *d_nrm = nullptr, *d_sph = nullptr;
  CK(tmp.alloc(&d_type, (size_t)n * 4)); CK(tmp.alloc(&d_bsdf, (size_t)n * 4));
  CK(tmp.alloc(&d_pos, (size_t)n * 72)); CK(tmp.alloc(&d_nrm, (size_t)n * 72)); CK(tmp.alloc(&d_sph, (size_t)n * 32));
  CK(cudaMemcpyAsync(d_type, ctx->prim_type.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_bsdf, ctx->prim_bsdf.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_pos, ctx->tri_pos.data(), (size_t)n * 72, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_nrm, ctx->tri_nrm.data(), (size_t)n * 72, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_sph, ctx->sphere.data(), (size_t)n * 32, cudaMemcpyHostToDevice, st));
  DbScene sc; sc.prim_type = d_type; sc.prim_bsdf = d_bsdf; sc.tri_pos = d_pos; sc.tri_nrm = d_nrm; sc.sphere = d_sph; sc.n = n;
  float *d_pbox = nullptr, *d_ibox = nullptr; uint32_t* d_scene = nullptr;
  uint64_t *d_keys = nullptr, *d_keys2 = nullptr; uint32_t *d_vals = nullptr, *d_sorted = nullptr;
  int *d_left = nullptr, *d_right = nullptr, *d_first = nullptr, *d_last = nullptr, *d_pint = nullptr, *d_pleaf = nullptr, *d_visit = nullptr;
  const size_t ni = (size_t)std::max(n - 1, 1);
  CK(tmp.alloc(&d_pbox, (size_t)n * 24)); CK(tmp.alloc(&d_ibox, ni * 24)); CK(tmp.alloc(&d_scene, 32));
  CK(tmp.alloc(&d_keys, (size_t)n * 8)); CK(tmp.alloc(&d_keys2, (size_t)n * 8)); CK(tmp.alloc(&d_vals, (size_t)n * 4)); CK(tmp.alloc(&d_sorted, (size_t)n * 4));
  CK(tmp.alloc(&d_left, ni * 4)); CK(tmp.alloc(&d_right, ni * 4)); CK(tmp.alloc(&d_first, ni * 4)); CK(tmp.alloc(&d_last, ni * 4));
  CK(tmp.alloc(&d_pint, ni * 4)); CK(tmp.alloc(&d_pleaf, (size_t)n * 4)); CK(tmp.alloc(&d_visit, ni * 4));
  const uint32_t scene_init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  CK(cudaMemcpyAsync(d_scene, scene_init, sizeof(scene_init), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d_visit, 0, ni * 4, st));
  const int B = 256, grid_n = (n + B - 1) / B;
  k_db_boxes<<<grid_n, B, 0, st>>>(sc, d_pbox, d_scene);
  k_db_morton<<<grid_n, B, 0, st>>>(n, d_pbox, d_scene, d_keys, d_vals);
  size_t cub_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, d_keys, d_keys2, d_vals, d_sorted, n, 0, 63, st));
  void* d_cub = nullptr; CK(tmp.alloc(&d_cub, cub_bytes + 16));
  CK(cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, d_keys, d_keys2, d_vals, d_sorted, n, 0, 63, st));
  if (n > 1) {
    k_db_tree<<<(n - 1 + B - 1) / B, B, 0, st>>>(n, d_keys2, d_left, d_right, d_first, d_last, d_pint, d_pleaf);
    k_db_refit<<<grid_n, B, 0, st>>>(n, d_left, d_right, d_pint, d_pleaf, d_pbox, d_sorted, d_ibox, d_visit);
  }
  CK(cudaGetLastError());
  // collapse, one level of wide nodes per launch
  const unsigned node_cap = (unsigned)n + 8u;      // every wide node below the root holds >= 2 primitives and has >= 2 children
  WideNode* d_nodes = nullptr; int32_t* d_slot = nullptr; int2 *d_items[2] = {nullptr, nullptr}; unsigned int* d_cnt = nullptr;
  CK(tmp.alloc(&d_nodes, (size_t)node_cap * sizeof(WideNode))); CK(tmp.alloc(&d_slot, (size_t)n * 4));
  CK(tmp.alloc(&d_items[0], (size_t)node_cap * 8)); CK(tmp.alloc(&d_items[1], (size_t)node_cap * 8)); CK(tmp.alloc(&d_cnt, 16));
  const unsigned int cnt_init[4] = {1u, 0u, 0u, 0u};               // the root is wide node 0
  CK(cudaMemcpyAsync(d_cnt, cnt_init, sizeof(cnt_init), cudaMemcpyHostToDevice, st));
  const int2 root_item = make_int2(n > 1 ? 0 : -1, 0);              // a single primitive: sorted leaf 0
  CK(cudaMemcpyAsync(d_items[0], &root_item, sizeof(root_item), cudaMemcpyHostToDevice, st));
  DbTree T; T.left = d_left; T.right = d_right; T.first = d_first; T.last = d_last; T.ibox = d_ibox; T.pbox = d_pbox; T.sorted = d_sorted;
  CK(cudaStreamSynchronize(st));
  const double t1 = now();
  DbEndPlanes ends; std::memset(&ends, 0, sizeof(ends));
  if (ctx->opt_light_phase) {
    for (const EndPlane& e : light_end_planes((int)ctx->light_type.size(), ctx->light_type.data(), ctx->light_param.data())) {
      if (ends.n == kDbMaxEndPlanes) break;
      const int i = ends.n++;
      ends.axis[i] = e.axis; ends.from_low[i] = e.from_low ? 1 : 0; ends.coord[i] = e.coord;
      for (int k = 0; k < 3; k++) { ends.lo[i][k] = e.lo[k]; ends.hi[i][k] = e.hi[k]; }
    }
  }
  unsigned n_items = 1; int levels = 0, cur = 0;
  while (n_items > 0) {
    levels++;
    if (levels > kStackEntries) return fail(ctx, DSRT_ERR_LIMIT, "dsrt_build_accel(device_build): wide BVH deeper than the traversal stack");
    k_db_collapse<<<(n_items + 127) / 128, 128, 0, st>>>(T, d_items[cur], (int)n_items, d_items[cur ^ 1], d_cnt, d_nodes, d_slot, node_cap, ends);
    unsigned int h[4];
    CK(cudaMemcpyAsync(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h[3]) return fail(ctx, DSRT_ERR_LIMIT, "dsrt_build_accel(device_build): wide node pool exhausted");
    n_items = h[2];
    const unsigned int zero = 0;
    CK(cudaMemcpyAsync(d_cnt + 2, &zero, 4, cudaMemcpyHostToDevice, st));
    cur ^= 1;
  }
  unsigned int h[4];
  CK(cudaMemcpy(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost));
  { uint32_t so[6]; CK(cudaMemcpy(so, d_scene, sizeof(so), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 6; k++) { const uint32_t o = so[k]; const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o; std::memcpy(&scene_box[k], &u, 4); } }
  const size_t n_wide = h[0], n_slots = h[1];
  if (n_slots != (size_t)n) return fail(ctx, DSRT_ERR_LIMIT, "dsrt_build_accel(device_build): primitive slots do not add up");
  const double t2 = now();
  // The 48-byte primitive / shading records (the bulk of the structure) are written straight into this GPU's final arrays and
  // never visit the host: the other GPUs get them GPU -> GPU (dsrt_upload_accel).  The nodes and the slot -> primitive map are
  // read back (statistics, id mapping, the parity kernel's fp64 records).
  dev_free(D.d_prims); dev_free(D.d_shade); D.d_prims = D.d_shade = nullptr;
  CK(cudaMalloc(&D.d_prims, n_slots * sizeof(PrimRecord))); CK(cudaMalloc(&D.d_shade, n_slots * sizeof(ShadeRecord)));
  k_db_flatten<<<(unsigned)((n_slots + B - 1) / B), B, 0, st>>>(sc, (int)n_slots, d_slot, (PrimRecord*)D.d_prims, (ShadeRecord*)D.d_shade);
  CK(cudaGetLastError());
  if (ctx->opt_flat_slots) { k_db_mark_flat<<<(unsigned)((n_wide + B - 1) / B), B, 0, st>>>(sc, d_nodes, (int)n_wide, d_slot); CK(cudaGetLastError()); }
  ctx->wide.nodes.resize(n_wide); ctx->wide.slot_prim.resize(n_slots); ctx->wide.max_depth = levels;
  ctx->recs.clear(); ctx->recs.shrink_to_fit(); ctx->shd.clear(); ctx->shd.shrink_to_fit(); ctx->recs_on_device = true;
  CK(cudaMemcpyAsync(ctx->wide.nodes.data(), d_nodes, n_wide * sizeof(WideNode), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ctx->wide.slot_prim.data(), d_slot, n_slots * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (timing) fprintf(stderr, "device_build: %d prims, upload + boxes + morton + sort + tree + refit %.3f s, collapse (%d levels, %zu wide nodes) %.3f s, records + read-back %.3f s\n",
                      n, t1 - t0, levels, n_wide, t2 - t1, now() - t2);
  return DSRT_OK;
}

int dsrt_build_accel(dsrt_ctx* ctx) {
  if (!ctx) return DSRT_ERR_INVALID;
  struct Range { Range(const char* n) { nvtxRangePushA(n); } ~Range() { nvtxRangePop(); } } nvtx_range("dsrt_build_accel");
  const bool on_device = ctx->opt_device_build != 0 && ctx->n_prims > 0;
  if (!ctx->have_scene || (!ctx->have_bvh && !on_device)) return fail(ctx, DSRT_ERR_INVALID, "dsrt_build_accel: scene and BVH must be set first");
  dsrt_scene s{}; s.n_prims = ctx->n_prims; s.prim_type = ctx->prim_type.data(); s.prim_bsdf = ctx->prim_bsdf.data();
  s.tri_pos = ctx->tri_pos.data(); s.tri_nrm = ctx->tri_nrm.data(); s.sphere = ctx->sphere.data();
  const bool timing = std::getenv("DSRT_BUILD_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tb0 = now();
  std::vector<Box3> pbox;
  Box3 all; all.reset();
  if (on_device) {          // the device builder makes its own primitive boxes and hands back the scene box
    float sb[6];
    int rcd = build_wide_bvh_device(ctx, ctx->devs[0], sb);
    if (rcd) return rcd;
    for (int k = 0; k < 3; k++) { all.lo[k] = sb[k]; all.hi[k] = sb[3 + k]; }
  } else {
  primitive_boxes(&s, pbox);
  for (const Box3& p : pbox) all.grow(p);
  dsrt_bvh2 b{}; b.n_nodes = (int)ctx->node_start.size(); b.node_bbox = ctx->node_bbox.data(); b.node_start = ctx->node_start.data();
  b.node_range = ctx->node_range.data(); b.node_left = ctx->node_left.data(); b.node_right = ctx->node_right.data();
  b.prim_order = ctx->prim_order.data();
  std::string err;
  const std::vector<EndPlane> ends = ctx->opt_light_phase ? light_end_planes((int)ctx->light_type.size(), ctx->light_type.data(), ctx->light_param.data()) : std::vector<EndPlane>();
  int rc = build_wide_bvh(b, pbox, ctx->n_prims, ctx->wide, err, (double)ctx->opt_prim_cost / 100.0, &ends, ctx->opt_regroup != 0);
  if (!rc && ctx->opt_regroup && ctx->wide.max_depth > kStackEntries)      // regrouping may add levels: never at the price of a scene that no longer fits the stack
    rc = build_wide_bvh(b, pbox, ctx->n_prims, ctx->wide, err, (double)ctx->opt_prim_cost / 100.0, &ends, false);
  if (rc) return fail(ctx, rc, err);
  if (ctx->opt_flat_slots) mark_flat_slots(s, ctx->wide);
  }
  if (ctx->wide.max_depth > kStackEntries) return fail(ctx, DSRT_ERR_LIMIT, "dsrt_build_accel: wide BVH deeper than the traversal stack");
  if (ctx->wide.slot_prim.size() >= ((size_t)1 << kOwnerShift)) return fail(ctx, DSRT_ERR_LIMIT, "dsrt_build_accel: more than 2^27 primitives");
  const double tb1 = now();
  double dg = 0; for (int k = 0; k < 3; k++) { double e = ctx->n_prims ? all.hi[k] - all.lo[k] : 0; double m = ctx->n_prims ? std::fmax(std::fabs(all.lo[k]), std::fabs(all.hi[k])) : 0; dg += (e + m) * (e + m); }
  ctx->scene_diag = std::sqrt(dg) + 1.0;
  bounding_sphere(all, ctx->n_prims, ctx->bsphere);

  if (!on_device) { flatten_records(s, ctx->wide, ctx->recs, ctx->shd); ctx->recs_on_device = false; }
  ctx->r64.clear(); ctx->r64.shrink_to_fit();     // fp64 records: built and uploaded by the first dsrt_primary_hits(mode 1)
  ctx->n_light_samples = flatten_lights((int)ctx->light_type.size(), ctx->light_type.data(), ctx->light_param.data(), ctx->ns_area_light, ctx->env_w > 0, ctx->lights);
  ctx->n_lights = (int)ctx->lights.size();
  for (DevState& D : ctx->devs) {      // new sizes: (re)allocate the device copies
    CK(cudaSetDevice(D.device));
    const bool keep_records = ctx->recs_on_device && &D == &ctx->devs[0];      // written in place by the device builder
    dev_free(D.d_nodes); dev_free(D.d_prims64); dev_free(D.d_bsdf); dev_free(D.d_lights);
    if (!keep_records) { dev_free(D.d_prims); dev_free(D.d_shade); D.d_prims = D.d_shade = nullptr; }
    D.d_nodes = D.d_prims64 = D.d_bsdf = D.d_lights = nullptr;
    const size_t n = ctx->wide.slot_prim.size();
    CK(cudaMalloc(&D.d_nodes, std::max<size_t>(ctx->wide.nodes.size() * sizeof(WideNode), 16)));
    if (!keep_records) {
      CK(cudaMalloc(&D.d_prims, std::max<size_t>(n * sizeof(PrimRecord), 16)));
      CK(cudaMalloc(&D.d_shade, std::max<size_t>(n * sizeof(ShadeRecord), 16)));
    }
    CK(cudaMalloc(&D.d_bsdf, std::max<size_t>(ctx->bsdfs.size() * sizeof(Bsdf), 16)));
    CK(cudaMalloc(&D.d_lights, std::max<size_t>(ctx->lights.size() * sizeof(Light), 16)));
    dev_free(D.d_env_rgb); dev_free(D.d_env_tp); dev_free(D.d_env_t); dev_free(D.d_env_pgt);
    D.d_env_rgb = D.d_env_tp = D.d_env_t = D.d_env_pgt = nullptr;
    if (ctx->env_w > 0) {
      const size_t np = (size_t)ctx->env_w * ctx->env_h;
      CK(cudaMalloc((void**)&D.d_env_rgb, np * 3 * sizeof(float))); CK(cudaMalloc((void**)&D.d_env_tp, np * sizeof(float)));
      CK(cudaMalloc((void**)&D.d_env_t, (size_t)ctx->env_h * sizeof(float))); CK(cudaMalloc((void**)&D.d_env_pgt, np * sizeof(float)));
    }
  }
  ctx->have_accel = true;
  const double tb2 = now();
  const int rcu = dsrt_upload_accel(ctx);
  if (timing) fprintf(stderr, "dsrt_build_accel: wide BVH %.3f s (%s), records + light tables + device allocation %.3f s, upload to %zu GPU(s) %.3f s\n",
                      tb1 - tb0, on_device ? "device_build" : "host collapse", tb2 - tb1, ctx->devs.size(), now() - tb2);
  return rcu;
}

// Copy of the flattened scene (wide nodes, primitive / shading records, BSDF and light tables, environment tables) to every
// GPU: ONE host-to-device copy, to the first device, then device-to-device copies from there to the others over NVLink
// (the scene is replicated, fed from GPU 0).  dsrt_build_accel calls it; callers may call it again to re-send
// the same scene (bench.py's e2e step).
int dsrt_upload_accel(dsrt_ctx* ctx) {
  if (!ctx) return DSRT_ERR_INVALID;
  if (!ctx->have_accel) return fail(ctx, DSRT_ERR_INVALID, "dsrt_upload_accel: call dsrt_build_accel first");
  const size_t n = ctx->wide.slot_prim.size();
  const size_t np = (size_t)ctx->env_w * ctx->env_h;
  auto parts_of = [&](DevState& D, std::vector<std::pair<void*, std::pair<const void*, size_t>>>& out) {
    out.clear();
    auto add = [&](void* d, const void* s, size_t b) { if (b) out.push_back({d, {s, b}}); };
    add(D.d_nodes, ctx->wide.nodes.data(), ctx->wide.nodes.size() * sizeof(WideNode));
    add(D.d_prims, ctx->recs_on_device ? nullptr : ctx->recs.data(), n * sizeof(PrimRecord));      // no host source: resident on device 0
    add(D.d_shade, ctx->recs_on_device ? nullptr : ctx->shd.data(), n * sizeof(ShadeRecord));
    add(D.d_bsdf, ctx->bsdfs.data(), ctx->bsdfs.size() * sizeof(Bsdf));
    add(D.d_lights, ctx->lights.data(), ctx->lights.size() * sizeof(Light));
    if (ctx->env_w > 0) {
      add(D.d_env_rgb, ctx->env_rgb.data(), np * 3 * sizeof(float));
      add(D.d_env_tp, ctx->env_tp.data(), np * sizeof(float));
      add(D.d_env_t, ctx->env_t.data(), (size_t)ctx->env_h * sizeof(float));
      add(D.d_env_pgt, ctx->env_pgt.data(), np * sizeof(float));
    }
  };
  std::vector<std::pair<void*, std::pair<const void*, size_t>>> p0, pr;
  DevState& D0 = ctx->devs[0];
  CK(cudaSetDevice(D0.device));
  parts_of(D0, p0);
  for (auto& p : p0) if (p.second.first) CK(cudaMemcpyAsync(p.first, p.second.first, p.second.second, cudaMemcpyHostToDevice, D0.stream));
  CK(cudaStreamSynchronize(D0.stream));
  for (size_t r = 1; r < ctx->devs.size(); r++) {       // peers: device 0's copy is the source, all peers in flight together
    DevState& D = ctx->devs[r];
    CK(cudaSetDevice(D.device));
    parts_of(D, pr);
    for (size_t k = 0; k < pr.size(); k++)
      CK(cudaMemcpyPeerAsync(pr[k].first, D.device, p0[k].first, D0.device, pr[k].second.second, D.stream));
  }
  for (size_t r = 1; r < ctx->devs.size(); r++) { CK(cudaSetDevice(ctx->devs[r].device)); CK(cudaStreamSynchronize(ctx->devs[r].stream)); }
  return DSRT_OK;
}

int dsrt_accel_bytes(const dsrt_ctx* ctx, int64_t* h2d_bytes) {
  if (!ctx || !ctx->have_accel || !h2d_bytes) return DSRT_ERR_INVALID;
  const size_t n = ctx->wide.slot_prim.size();
  *h2d_bytes = (int64_t)(ctx->wide.nodes.size() * sizeof(WideNode) + (ctx->recs_on_device ? 0 : n * (sizeof(PrimRecord) + sizeof(ShadeRecord))) +
                         ctx->bsdfs.size() * sizeof(Bsdf) + ctx->lights.size() * sizeof(Light) +
                         (size_t)ctx->env_w * ctx->env_h * 5 * sizeof(float) + (size_t)ctx->env_h * sizeof(float));
  return DSRT_OK;
}

int dsrt_accel_info(const dsrt_ctx* ctx, int64_t* n_wide_nodes, int64_t* node_bytes, int64_t* prim_bytes, int32_t* max_depth) {
  if (!ctx || !ctx->have_accel) return DSRT_ERR_INVALID;
  if (n_wide_nodes) *n_wide_nodes = (int64_t)ctx->wide.nodes.size();
  if (node_bytes) *node_bytes = (int64_t)(ctx->wide.nodes.size() * sizeof(WideNode));
  if (prim_bytes) *prim_bytes = (int64_t)(ctx->wide.slot_prim.size() * sizeof(PrimRecord));
  if (max_depth) *max_depth = ctx->wide.max_depth;
  return DSRT_OK;
}

// ------------------------------------------------------------------------------------------------ rendering
// Samples per wavefront batch and batches per deep-path pool for a frame of npp (padded) pixels on device D.
// Wavefront + pool memory = paths x (80 B state and queues + 48 B per light sample) x (1 + pooled batches).  The reference
// accepts any -l / frame size, so the batch and the pool group shrink until the working set fits in what the device has
// free (plus what this context already holds); only a single-sample batch that still does not fit is an error.
static int plan_batches(dsrt_ctx* ctx, DevState& D, int npp, int spp_count, int* batch_spp_out, int* pool_group_out) {
  const int nls = ctx->n_light_samples;
  int batch_spp = (int)ctx->opt_batch_spp;
  if (batch_spp <= 0) batch_spp = std::max(1, (int)((16u << 20) / (unsigned)npp));   // ~16M paths per batch (measured: 4 / 8 / 16 M paths -> 6.38 / 6.59 / 6.71 Grays/s)
  batch_spp = std::min(batch_spp, std::max(1, spp_count));
  int pool_group = (int)ctx->opt_pool_batches;
  // A frame like the previous one finds its buffers in place: nothing to size, and no cudaMemGetInfo (the query takes
  // 10-70 ms every few calls once tens of GB are allocated -- measured as the only host-side stall of an end-to-end step)
  if (ctx->opt_mem_budget_mb == 0) {
    const size_t P = (size_t)npp * (size_t)batch_spp, ls = (size_t)std::max(nls, 1);
    const size_t grp = (size_t)std::max(1, std::min((std::max(spp_count, 1) + batch_spp - 1) / batch_spp, pool_group));
    if (P <= D.cap_paths && P * ls <= D.cap_shadow && (ctx->max_depth <= 0 || (P * grp <= D.cap_pool && P * grp * ls <= D.cap_pool_shadow))) {
      *batch_spp_out = batch_spp; *pool_group_out = pool_group;
      return DSRT_OK;
    }
  }
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  TP("cudaMemGetInfo");
  const double held = (double)D.cap_paths * 80.0 + (double)D.cap_shadow * 48.0 + (double)D.cap_pool * 80.0 + (double)D.cap_pool_shadow * 48.0;
  double budget = 0.8 * ((double)free_b + held);
  if (ctx->opt_mem_budget_mb > 0) budget = std::min(budget, (double)ctx->opt_mem_budget_mb * 1048576.0);
  const double per_path = 80.0 + 48.0 * (double)std::max(nls, 1);
  auto need = [&](int bspp, int grp) { return (double)npp * bspp * per_path * (1.0 + (ctx->max_depth > 0 ? grp : 0)); };
  while (need(batch_spp, pool_group) > budget) {
    if (pool_group > 2) pool_group = (pool_group + 1) / 2;
    else if (batch_spp > 1) batch_spp = (batch_spp + 1) / 2;
    else if (pool_group > 1) pool_group = 1;
    else return fail(ctx, DSRT_ERR_LIMIT, "dsrt_render: one sample per pixel of this frame with " + std::to_string(nls) +
                     " light samples needs " + std::to_string((long long)(need(1, 1) / 1048576.0)) + " MiB of wavefront state; the device has " +
                     std::to_string((long long)(budget / 1048576.0)) + " MiB available");
  }
  *batch_spp_out = batch_spp; *pool_group_out = pool_group;
  return DSRT_OK;
}

// Enqueues the whole wavefront for samples spp_begin + k*spp_stride (k < spp_count) on one device; no host sync.
// first == false: a later chunk of the same frame (dsrt_render splits a frame so that dsrt_cancel can take effect): counters,
// launch totals and the begin event carry over.
static int render_impl(dsrt_ctx* ctx, DevState& D, int spp_begin, int spp_count, int spp_stride, float* d_accum, cudaStream_t st, bool first = true) {
  if (!ctx->have_accel || !ctx->have_cam) return fail(ctx, DSRT_ERR_INVALID, "dsrt_render: call dsrt_build_accel and dsrt_set_camera first");
  if (spp_count < 0 || spp_stride < 1 || spp_begin < 0) return fail(ctx, DSRT_ERR_INVALID, "dsrt_render: bad sample range");
  CK(cudaSetDevice(D.device));
  struct Range { Range(const char* n) { nvtxRangePushA(n); } ~Range() { nvtxRangePop(); } } nvtx_range("dsrt render_impl (enqueue wavefront)");
  { int rc0 = size_trace_grid(ctx, D); if (rc0) return rc0; }
  TP("trace grid");
  const bool windowed = ctx->win[2] > 0;
  const int wx0 = windowed ? ctx->win[0] : 0, wy0 = windowed ? ctx->win[1] : 0;
  const int W = windowed ? ctx->win[2] : ctx->cam.width, H = windowed ? ctx->win[3] : ctx->cam.height;     // extent rendered by this call
  if (wx0 < 0 || wy0 < 0 || wx0 + W > ctx->cam.width || wy0 + H > ctx->cam.height) return fail(ctx, DSRT_ERR_INVALID, "dsrt_render: window outside the frame");
  const int blocks_x = (W + 7) / 8, blocks_y = (H + 3) / 4;
  const int npp = blocks_x * blocks_y * 32;
  const int aligned = (W % 8 == 0 && H % 4 == 0) ? 1 : 0;
  const int nls = ctx->n_light_samples;
  int batch_spp = 1, pool_group = 1;
  { int rcp = plan_batches(ctx, D, npp, spp_count, &batch_spp, &pool_group); if (rcp) return rcp; }
  const size_t P = (size_t)npp * batch_spp;
  int rc = ensure_wavefront(ctx, D, P, P * (size_t)std::max(nls, 1));
  if (rc) return rc;
  TP("plan + wavefront buffers");
  const int n_batches = (spp_count + batch_spp - 1) / batch_spp;
  if (n_batches > D.n_counter_blocks) {
    if ((rc = dev_alloc(ctx, &D.d_counters, (size_t)std::max(n_batches, 1)))) return rc;
    D.n_counter_blocks = std::max(n_batches, 1);
  }
  CK(cudaMemsetAsync(D.d_counters, 0, sizeof(Counters) * (size_t)std::max(n_batches, 1), st));
  if (first) {
    CK(cudaMemsetAsync(D.d_totals, 0, sizeof(Totals), st));
    D.ev_used = 0; D.spans.clear(); D.launches = 0; D.batches = 0;
  }
  D.batches += (uint32_t)n_batches;

  SceneDev sc; sc.bsdf = (const Bsdf*)D.d_bsdf; sc.lights = (const Light*)D.d_lights; sc.shade = (const float4*)D.d_shade;
  sc.n_lights = ctx->n_lights; sc.n_light_samples = nls;
  sc.env.rgb = D.d_env_rgb; sc.env.pThetaPhi = D.d_env_tp; sc.env.pTheta = D.d_env_t; sc.env.pPhiGivenTheta = D.d_env_pgt;
  sc.env.w = ctx->env_w; sc.env.h = ctx->env_h;
  RenderParams rp; rp.cam = ctx->cam; rp.seed = ctx->seed; rp.max_depth = ctx->max_depth; rp.spp_begin = spp_begin; rp.spp_stride = spp_stride;
  rp.n_pix_padded = npp; rp.blocks_x = blocks_x; rp.skip_null_shadow = (int)ctx->opt_skip_null; rp.batch_first_sample = 0;
  rp.win_x0 = wx0; rp.win_y0 = wy0; rp.win_x1 = wx0 + W; rp.win_y1 = wy0 + H;
  const Accel A = make_accel(ctx, D, false);
  const bool count = ctx->opt_count != 0, timing = ctx->opt_stage_timing != 0;
  const int tgrid = D.trace_blocks;
  const int tri_min = (int)ctx->opt_tri_min, refill_busy = (int)ctx->opt_refill, wait_mode = (int)ctx->opt_wait_mode, coop_min = (int)ctx->opt_coop_min, refill_hi = (int)ctx->opt_refill_hi, refill_patience = (int)ctx->opt_refill_patience;
  const size_t sbytes = stack_bytes(ctx), sbytes_closest = stack_bytes(ctx, false);

  auto span_begin = [&](int kind) { if (timing) { DevState::Span s; s.kind = kind; s.e0 = D.ev_used; cudaEventRecord(next_event(D), st); s.e1 = 0; D.spans.push_back(s); } };
  auto span_end = [&]() { if (timing) { D.spans.back().e1 = D.ev_used; cudaEventRecord(next_event(D), st); } };

  // pool: survivors of up to `group` consecutive batches (worst case: every path survives, e.g. a mirror box)
  const int group = std::max(1, std::min(n_batches, pool_group));
  const int n_groups = (n_batches + group - 1) / group;
  const size_t pool_cap = P * (size_t)group;
  if (ctx->max_depth > 0) {
    if ((rc = ensure_pool(ctx, D, pool_cap, pool_cap * (size_t)std::max(nls, 1)))) return rc;
    if (n_groups > D.n_pool_counter_blocks) {
      if ((rc = dev_alloc(ctx, &D.d_pool_counters, (size_t)n_groups))) return rc;
      D.n_pool_counter_blocks = n_groups;
    }
    CK(cudaMemsetAsync(D.d_pool_counters, 0, sizeof(PoolCounters) * (size_t)n_groups, st));
  }
  TP("pool buffers + memsets");
  if (first) CK(cudaEventRecord(D.ev_begin, st));     // after every (re)allocation: gpu_seconds covers the kernels of the frame only
  auto trace = [&](bool any, const float4* ro, const float4* rd, const uint32_t* q, const uint32_t* n_ptr, uint32_t* work, float4* hits, const float4* contrib) {
    span_begin(any ? 1 : 0);
    if (any) {
      if (count) k_trace<true, true><<<tgrid, kTraceThreads, sbytes, st>>>(A, ro, rd, q, n_ptr, work, nullptr, contrib, d_accum, D.d_totals, tri_min, refill_busy, wait_mode, stack_entries(ctx), coop_min, refill_hi, refill_patience);
      else k_trace<true, false><<<tgrid, kTraceThreads, sbytes, st>>>(A, ro, rd, q, n_ptr, work, nullptr, contrib, d_accum, D.d_totals, tri_min, refill_busy, wait_mode, stack_entries(ctx), coop_min, refill_hi, refill_patience);
    } else {
      if (count) k_trace<false, true><<<tgrid, kTraceThreads, sbytes_closest, st>>>(A, ro, rd, q, n_ptr, work, hits, nullptr, nullptr, D.d_totals, tri_min, refill_busy, wait_mode, stack_entries(ctx), coop_min, refill_hi, refill_patience);
      else k_trace<false, false><<<tgrid, kTraceThreads, sbytes_closest, st>>>(A, ro, rd, q, n_ptr, work, hits, nullptr, nullptr, D.d_totals, tri_min, refill_busy, wait_mode, stack_entries(ctx), coop_min, refill_hi, refill_patience);
    }
    span_end();
    D.launches++;
  };
  for (int bi = 0; bi < n_batches; bi++) {
    const int s0 = bi * batch_spp, ns = std::min(batch_spp, spp_count - s0);
    const int n_paths = npp * ns;
    Counters* C = D.d_counters + bi;
    PoolCounters* PC = ctx->max_depth > 0 ? D.d_pool_counters + bi / group : nullptr;
    rp.batch_first_sample = s0;
    // ---- depth 0 of this batch: generate, extend, shade (survivors -> pool), connect
    span_begin(2);
    k_generate<<<(n_paths + 255) / 256, 256, 0, st>>>(D.ps, rp, n_paths, D.queue[0], &C->q_count[0], aligned);
    span_end();
    D.launches++;
    const uint32_t* q0 = aligned ? nullptr : D.queue[0];
    trace(false, D.ps.ray_o, D.ps.ray_d, q0, &C->q_count[0], &C->work_extend[0], D.ps.hit, nullptr);
    span_begin(2);
    k_shade<<<(n_paths + 127) / 128, 128, 0, st>>>(D.ps, (const float4*)D.d_prims, sc, rp, q0, &C->q_count[0], D.pool, nullptr,
                                                   PC ? &PC->q_count[0] : &C->q_count[1], (uint32_t)pool_cap, D.sq, &C->s_count[0], d_accum, D.d_totals);
    span_end();
    D.launches++;
    if (nls > 0) trace(true, D.sq.a, D.sq.b, nullptr, &C->s_count[0], &C->work_connect[0], nullptr, D.sq.c);
    k_tally<<<1, 32, 0, st>>>(C, D.d_totals, (uint32_t)((size_t)W * H * ns));
    D.launches++;
    // ---- after the last batch of a group: advance the pooled paths together, one bounce per iteration
    if (PC && ((bi + 1) % group == 0 || bi + 1 == n_batches)) {
      const int pgrid = std::min((int)((pool_cap + 127) / 128), D.sm_count * 16);
      int cur = 0;
      for (int it = 0; it < ctx->max_depth; it++) {
        const uint32_t* q = it == 0 ? nullptr : D.pool_queue[cur];
        trace(false, D.pool.ray_o, D.pool.ray_d, q, &PC->q_count[it], &PC->work_extend[it], D.pool.hit, nullptr);
        span_begin(2);
        k_shade<<<pgrid, 128, 0, st>>>(D.pool, (const float4*)D.d_prims, sc, rp, q, &PC->q_count[it], D.pool, D.pool_queue[cur ^ 1],
                                       &PC->q_count[it + 1], 0u, D.pool_sq, &PC->s_count[it], d_accum, D.d_totals);
        span_end();
        D.launches++;
        if (nls > 0) trace(true, D.pool_sq.a, D.pool_sq.b, nullptr, &PC->s_count[it], &PC->work_connect[it], nullptr, D.pool_sq.c);
        cur ^= 1;
      }
      k_tally_pool<<<1, 32, 0, st>>>(PC, D.d_totals, ctx->max_depth);
      D.launches++;
    }
  }
  CK(cudaEventRecord(D.ev_end, st));
  CK(cudaGetLastError());
  TP("launches enqueued");
  return DSRT_OK;
}

// sums the counters of all devices; times are the maximum over devices
int dsrt_collect_stats(dsrt_ctx* ctx, dsrt_stats* stats) {
  if (!ctx) return DSRT_ERR_INVALID;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  for (DevState& D : ctx->devs) {
    CK(cudaSetDevice(D.device));
    CK(cudaEventSynchronize(D.ev_end));
    if (!stats) continue;
    Totals t;
    CK(cudaMemcpy(&t, D.d_totals, sizeof(t), cudaMemcpyDeviceToHost));
    stats->camera_samples += t.camera; stats->extend_rays += t.extend; stats->shadow_rays += t.shadow;
    stats->null_shadow_rays += t.null_shadow;
    stats->extend_nodes += t.nodes[0]; stats->extend_prims += t.prims[0]; stats->connect_nodes += t.nodes[1]; stats->connect_prims += t.prims[1];
    float ms = 0; CK(cudaEventElapsedTime(&ms, D.ev_begin, D.ev_end));
    stats->gpu_seconds = std::max(stats->gpu_seconds, (double)ms * 1e-3);
    double e = 0, c = 0, sh = 0;
    for (const auto& s : D.spans) {
      float m = 0; cudaEventElapsedTime(&m, D.ev_pool[s.e0], D.ev_pool[s.e1]);
      if (s.kind == 0) e += m * 1e-3; else if (s.kind == 1) c += m * 1e-3; else sh += m * 1e-3;
    }
    stats->extend_seconds = std::max(stats->extend_seconds, e); stats->connect_seconds = std::max(stats->connect_seconds, c);
    stats->shade_seconds = std::max(stats->shade_seconds, sh);
    stats->kernel_launches += D.launches; stats->batches += D.batches;
  }
  return DSRT_OK;
}

int dsrt_render_device(dsrt_ctx* ctx, int32_t spp_begin, int32_t spp_count, int32_t spp_stride, float* d_accum, void* stream, dsrt_stats* stats) {
  if (!ctx || !d_accum) return DSRT_ERR_INVALID;
  if (ctx->devs.size() != 1) return fail(ctx, DSRT_ERR_INVALID, "dsrt_render_device: single-device contexts only (one process per GPU)");
  DevState& D = ctx->devs[0];
  cudaStream_t st = stream ? (cudaStream_t)stream : D.stream;
  int rc = render_impl(ctx, D, spp_begin, spp_count, spp_stride, d_accum, st);
  if (rc) return rc;
  if (stats) return dsrt_collect_stats(ctx, stats);
  return DSRT_OK;
}

int dsrt_resolve_device(dsrt_ctx* ctx, const float* d_accum, float* d_rgb, uint32_t* d_rgba8, void* stream) {
  if (!ctx || !d_accum || !ctx->have_cam) return DSRT_ERR_INVALID;
  DevState& D = ctx->devs[0];
  CK(cudaSetDevice(D.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : D.stream;
  const int n = ctx->cam.width * ctx->cam.height;
  k_resolve<<<(n + 255) / 256, 256, 0, st>>>(d_accum, d_rgb, d_rgba8, n, 1.0f / (float)ctx->ns_aa);
  CK(cudaGetLastError());
  return DSRT_OK;
}

int dsrt_sync(dsrt_ctx* ctx) {
  if (!ctx) return DSRT_ERR_INVALID;
  for (DevState& D : ctx->devs) { CK(cudaSetDevice(D.device)); CK(cudaStreamSynchronize(D.stream)); }
  return DSRT_OK;
}

// dsrt_render / dsrt_render_tonemapped.  The frame is enqueued in chunks of a few pool groups (about 0.2 s of GPU work at
// 1080p) with at most two chunks in flight, so that dsrt_cancel -- PathTracer::stop (src/pathtracer.cpp:148-171) -- takes
// effect within a chunk; a frame that fits in one chunk (every test-sized frame) is enqueued exactly as before.
static int render_host(dsrt_ctx* ctx, int32_t spp_begin, int32_t spp_count, int32_t spp_stride, float* rgb_out, uint32_t* rgba8_out, dsrt_stats* stats) {
  if (!ctx || (!rgb_out && !rgba8_out)) return DSRT_ERR_INVALID;
  if (!ctx->have_cam) return fail(ctx, DSRT_ERR_INVALID, "dsrt_render: call dsrt_set_camera first");
  if (!ctx->have_accel) return fail(ctx, DSRT_ERR_INVALID, "dsrt_render: call dsrt_build_accel first");
  if (spp_count < 0 || spp_stride < 1 || spp_begin < 0) return fail(ctx, DSRT_ERR_INVALID, "dsrt_render: bad sample range");
  ctx->cancel.store(0);
  static const bool host_trace = std::getenv("DSRT_HOST_TRACE") != nullptr;
  g_host_trace.on = host_trace; g_host_trace.pts.clear();
  TP("enter");
  const size_t npix = (size_t)ctx->cam.width * ctx->cam.height;
  const int G = (int)ctx->devs.size();
  const int npp = (((ctx->win[2] > 0 ? ctx->win[2] : ctx->cam.width) + 7) / 8) * (((ctx->win[2] > 0 ? ctx->win[3] : ctx->cam.height) + 3) / 4) * 32;
  std::vector<int> done((size_t)G, 0);
  // GPU r renders samples k with k mod G == r of the requested list (load is balanced whatever the image content)
  // one host thread per device: a frame is a few thousand launches, and enqueueing them device after device from one
  // thread would start the last GPU tens of milliseconds late
  auto enqueue = [&](int r) -> int {
    DevState& D = ctx->devs[r];
    CK(cudaSetDevice(D.device));
    if (npix > D.accum_pixels) { int rc = dev_alloc(ctx, &D.d_accum_own, npix * 3); if (rc) return rc; D.accum_pixels = npix; }
    CK(cudaMemsetAsync(D.d_accum_own, 0, npix * 3 * sizeof(float), D.stream));
    const int cnt = spp_count > r ? (spp_count - r + G - 1) / G : 0;
    int bspp = 1, grp = 1;
    { int rc = plan_batches(ctx, D, npp, cnt, &bspp, &grp); if (rc) return rc; }
    TP("accum memset + plan");
    const int per_chunk = std::max(1, bspp * grp * 2);
    Scratch fences; cudaEvent_t fence[2];
    CK(fences.event(&fence[0])); CK(fences.event(&fence[1]));
    int i = 0;
    // experiment (DSRT_STAGGER=1, several streams on one GPU): odd streams start with half a batch, so that their
    // bandwidth-bound stages (generate / shade) fall into the other stream's issue-bound traversal stages
    static const bool stagger = std::getenv("DSRT_STAGGER") != nullptr;
    do {                                   // at least one call, so that an empty sample list still resets the counters
      int c = std::min(per_chunk, cnt - done[r]);
      if (stagger && i == 0 && (r & 1) && bspp >= 2) c = std::min(c, bspp / 2);
      int rc = render_impl(ctx, D, spp_begin + (r + done[r] * G) * spp_stride, c, spp_stride * G, D.d_accum_own, D.stream, i == 0);
      if (rc) return rc;
      done[r] += c;
      if (done[r] >= cnt) break;
      CK(cudaEventRecord(fence[i & 1], D.stream));
      if (i >= 1) CK(cudaEventSynchronize(fence[(i - 1) & 1]));
      i++;
    } while (!ctx->cancel.load());
    return DSRT_OK;
  };
  if (G == 1) { int rc = enqueue(0); if (rc) return rc; }
  else {
    std::vector<int> rcs((size_t)G, DSRT_OK);
    std::vector<std::thread> workers;
    for (int r = 0; r < G; r++) workers.emplace_back([&, r] { rcs[r] = enqueue(r); });
    for (std::thread& t : workers) t.join();
    for (int r = 0; r < G; r++) if (rcs[r]) return rcs[r];
  }
  int total_done = 0; for (int r = 0; r < G; r++) total_done += done[r];
  const bool cancelled = total_done < spp_count;
  // a cancelled frame is normalised by the samples that were rendered (an unbiased, noisier image), a complete one by ns_aa
  const float inv_spp = cancelled ? 1.0f / (float)std::max(total_done, 1) : 1.0f / (float)ctx->ns_aa;
  DevState& D0 = ctx->devs[0];
  CK(cudaSetDevice(D0.device));
  PeerPtrs pp; pp.n = G; pp.p[0] = D0.d_accum_own;
  for (int r = 1; r < G; r++) {
    DevState& D = ctx->devs[r];
    CK(cudaStreamWaitEvent(D0.stream, D.ev_end, 0));
    if (D.peer_ok) pp.p[r] = D.d_accum_own;
    else {
      const size_t need = npix * 3 * (size_t)(G - 1);
      if (need > D0.stage_floats) { int rc = dev_alloc(ctx, &D0.d_stage, need); if (rc) return rc; D0.stage_floats = need; }
      float* dst = D0.d_stage + npix * 3 * (size_t)(r - 1);
      CK(cudaMemcpyPeerAsync(dst, D0.device, D.d_accum_own, D.device, npix * 3 * sizeof(float), D0.stream));
      pp.p[r] = dst;
    }
  }
  if (rgba8_out && npix > D0.rgba8_pixels) { int rc = dev_alloc(ctx, &D0.d_rgba8, npix); if (rc) return rc; D0.rgba8_pixels = npix; }
  k_resolve_peers<<<(unsigned)((npix + 255) / 256), 256, 0, D0.stream>>>(pp, D0.d_accum_own, rgba8_out ? D0.d_rgba8 : nullptr, (int)npix, inv_spp);
  CK(cudaGetLastError());
  TP("resolve enqueued");
  if (host_trace) { CK(cudaEventSynchronize(D0.ev_end)); TP("kernels finished"); }
  if (rgb_out) CK(cudaMemcpyAsync(rgb_out, D0.d_accum_own, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, D0.stream));
  if (rgba8_out) CK(cudaMemcpyAsync(rgba8_out, D0.d_rgba8, npix * sizeof(uint32_t), cudaMemcpyDeviceToHost, D0.stream));
  CK(cudaStreamSynchronize(D0.stream));
  TP("resolve + D2H + sync");
  if (stats) { int rc = dsrt_collect_stats(ctx, stats); if (rc) return rc; }
  if (host_trace) {
    TP("stats");
    std::string line = "[dsrt host trace]";
    char buf[96];
    for (size_t i = 1; i < g_host_trace.pts.size(); i++) {
      std::snprintf(buf, sizeof(buf), "  %s %.4f", g_host_trace.pts[i].first, g_host_trace.pts[i].second - g_host_trace.pts[i - 1].second);
      line += buf;
    }
    std::snprintf(buf, sizeof(buf), "  | total %.4f s", g_host_trace.pts.back().second - g_host_trace.pts.front().second);
    std::fprintf(stderr, "%s%s\n", line.c_str(), buf);
    g_host_trace.on = false;
  }
  if (cancelled) { fail(ctx, DSRT_CANCELLED, "dsrt_render: cancelled after " + std::to_string(total_done) + " of " + std::to_string(spp_count) + " samples per pixel"); return DSRT_CANCELLED; }
  return DSRT_OK;
}

int dsrt_render(dsrt_ctx* ctx, int32_t spp_begin, int32_t spp_count, int32_t spp_stride, float* rgb_out, dsrt_stats* stats) {
  if (!rgb_out) return DSRT_ERR_INVALID;
  return render_host(ctx, spp_begin, spp_count, spp_stride, rgb_out, nullptr, stats);
}

int dsrt_render_tonemapped(dsrt_ctx* ctx, int32_t spp_begin, int32_t spp_count, int32_t spp_stride, float* rgb_out, uint32_t* rgba8_out, dsrt_stats* stats) {
  if (!rgba8_out) return DSRT_ERR_INVALID;
  return render_host(ctx, spp_begin, spp_count, spp_stride, rgb_out, rgba8_out, stats);
}

int dsrt_cancel(dsrt_ctx* ctx) {
  if (!ctx) return DSRT_ERR_INVALID;
  ctx->cancel.store(1);
  return DSRT_OK;
}

// Camera::generate_ray in double, exactly the reference's operation order (camera.cpp:113-129 with
// Matrix3x3::operator*, matrix3x3.cpp:138-142, and Vector3D::normalize, vector3D.h:95-131).  Host code is
// compiled without FMA contraction (see Makefile), so these rays are bit-identical to the reference's.
static void host_generate_ray64(const Camera& c, double x, double y, double* out6) {
  const double sp[3] = {-(x - 0.5) * c.W64 / c.dist64, -(y - 0.5) * c.H64 / c.dist64, 1.0};
  double w[3], dir[3];
  for (int k = 0; k < 3; k++) {
    w[k] = (sp[0] * c.c2w64[k] + sp[1] * c.c2w64[3 + k]) + sp[2] * c.c2w64[6 + k];
    dir[k] = ((-sp[0]) * c.c2w64[k] + (-sp[1]) * c.c2w64[3 + k]) + (-sp[2]) * c.c2w64[6 + k];
  }
  const double inv = 1. / std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  for (int k = 0; k < 3; k++) { out6[k] = w[k] + c.pos64[k]; out6[3 + k] = dir[k] * inv; }
}

int dsrt_primary_hits(dsrt_ctx* ctx, int32_t mode, int32_t* prim_id, double* t) {
  if (!ctx || !prim_id) return DSRT_ERR_INVALID;
  if (!ctx->have_accel || !ctx->have_cam) return fail(ctx, DSRT_ERR_INVALID, "dsrt_primary_hits: call dsrt_build_accel and dsrt_set_camera first");
  DevState& D = ctx->devs[0];
  CK(cudaSetDevice(D.device));
  { int rc0 = size_trace_grid(ctx, D); if (rc0) return rc0; }
  const int W = ctx->cam.width, H = ctx->cam.height, n = W * H;
  std::vector<int32_t> slots(n); std::vector<double> ts(n);
  cudaStream_t st = D.stream;
  if (mode == 1) {
    if (!D.d_prims64) {     // the parity kernel's fp64 primitive records are not part of the render path: built here, once
      dsrt_scene s{}; s.n_prims = ctx->n_prims; s.prim_type = ctx->prim_type.data(); s.prim_bsdf = ctx->prim_bsdf.data();
      s.tri_pos = ctx->tri_pos.data(); s.tri_nrm = ctx->tri_nrm.data(); s.sphere = ctx->sphere.data();
      flatten_records64(s, ctx->wide, ctx->r64);
      CK(cudaMalloc(&D.d_prims64, ctx->r64.size() * sizeof(PrimRecord64)));
      CK(cudaMemcpy(D.d_prims64, ctx->r64.data(), ctx->r64.size() * sizeof(PrimRecord64), cudaMemcpyHostToDevice));
      ctx->r64.clear(); ctx->r64.shrink_to_fit();
    }
    std::vector<double> rays((size_t)n * 6);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) host_generate_ray64(ctx->cam, (x + 0.5) / W, (y + 0.5) / H, &rays[6 * ((size_t)y * W + x)]);
    Scratch tmp;
    double* d_rays = nullptr; int32_t* d_slot = nullptr; double* d_t = nullptr;
    CK(tmp.alloc(&d_rays, rays.size() * sizeof(double)));
    CK(tmp.alloc(&d_slot, n * sizeof(int32_t)));
    CK(tmp.alloc(&d_t, n * sizeof(double)));
    CK(cudaMemcpyAsync(d_rays, rays.data(), rays.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    k_primary_parity<<<(n + kTraceThreads - 1) / kTraceThreads, kTraceThreads, stack_bytes(ctx), st>>>(make_accel(ctx, D, true), d_rays, n, d_slot, d_t);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(slots.data(), d_slot, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ts.data(), d_t, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  } else {
    int rc = ensure_wavefront(ctx, D, (size_t)n, 1);
    if (rc) return rc;
    if (D.n_counter_blocks < 1) { if ((rc = dev_alloc(ctx, &D.d_counters, (size_t)1))) return rc; D.n_counter_blocks = 1; }
    CK(cudaMemsetAsync(D.d_counters, 0, sizeof(Counters), st));
    RenderParams rp; std::memset(&rp, 0, sizeof(rp)); rp.cam = ctx->cam; rp.win_x1 = ctx->cam.width; rp.win_y1 = ctx->cam.height;
    const int tri_min = (int)ctx->opt_tri_min, refill_busy = (int)ctx->opt_refill, wait_mode = (int)ctx->opt_wait_mode, coop_min = (int)ctx->opt_coop_min, refill_hi = (int)ctx->opt_refill_hi, refill_patience = (int)ctx->opt_refill_patience;
    k_generate_centres<<<(n + 255) / 256, 256, 0, st>>>(D.ps, rp, n);
    k_set_u32<<<1, 1, 0, st>>>(&D.d_counters->q_count[0], (uint32_t)n);
    k_trace<false, false><<<D.trace_blocks, kTraceThreads, stack_bytes(ctx, false), st>>>(make_accel(ctx, D, false), D.ps.ray_o, D.ps.ray_d, nullptr, &D.d_counters->q_count[0],
                                                                            &D.d_counters->work_extend[0], D.ps.hit, nullptr, nullptr, D.d_totals, tri_min, refill_busy, wait_mode, stack_entries(ctx), coop_min, refill_hi, refill_patience);
    CK(cudaGetLastError());
    std::vector<float4> hits(n);
    CK(cudaMemcpyAsync(hits.data(), D.ps.hit, n * sizeof(float4), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++) { int sl; std::memcpy(&sl, &hits[i].w, 4); slots[i] = sl; ts[i] = sl >= 0 ? (double)hits[i].x : (double)INFINITY; }
  }
  for (int i = 0; i < n; i++) { prim_id[i] = slots[i] >= 0 ? ctx->wide.slot_prim[slots[i]] : -1; if (t) t[i] = ts[i]; }
  return DSRT_OK;
}

static int trace_batch(dsrt_ctx* ctx, bool any, int64_t n, const float* o, const float* d, const float* tmax, int32_t* out_id, float* out_t) {
  if (!ctx->have_accel) return fail(ctx, DSRT_ERR_INVALID, "dsrt_trace: call dsrt_build_accel first");
  if (n < 0 || (n > 0 && (!o || !d))) return DSRT_ERR_INVALID;
  if (n == 0) return DSRT_OK;
  DevState& D = ctx->devs[0];
  CK(cudaSetDevice(D.device));
  { int rc0 = size_trace_grid(ctx, D); if (rc0) return rc0; }
  int rc = ensure_wavefront(ctx, D, (size_t)n, 1);
  if (rc) return rc;
  if (D.n_counter_blocks < 1) { if ((rc = dev_alloc(ctx, &D.d_counters, (size_t)1))) return rc; D.n_counter_blocks = 1; }
  cudaStream_t st = D.stream;
  std::vector<float4> ho(n), hd(n);
  for (int64_t i = 0; i < n; i++) {
    ho[i] = make_float4(o[3 * i], o[3 * i + 1], o[3 * i + 2], tmax ? tmax[i] : INFINITY);
    int m1 = -1; float f; std::memcpy(&f, &m1, 4);
    hd[i] = make_float4(d[3 * i], d[3 * i + 1], d[3 * i + 2], f);
  }
  CK(cudaMemsetAsync(D.d_counters, 0, sizeof(Counters), st));
  CK(cudaMemsetAsync(D.d_totals, 0, sizeof(Totals), st));
  CK(cudaMemcpyAsync(D.ps.ray_o, ho.data(), n * sizeof(float4), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(D.ps.ray_d, hd.data(), n * sizeof(float4), cudaMemcpyHostToDevice, st));
  k_set_u32<<<1, 1, 0, st>>>(&D.d_counters->q_count[0], (uint32_t)n);
  const Accel A = make_accel(ctx, D, false);
  const int tri_min = (int)ctx->opt_tri_min, refill_busy = (int)ctx->opt_refill, wait_mode = (int)ctx->opt_wait_mode, coop_min = (int)ctx->opt_coop_min, refill_hi = (int)ctx->opt_refill_hi, refill_patience = (int)ctx->opt_refill_patience;
  if (any) k_trace<true, true><<<D.trace_blocks, kTraceThreads, stack_bytes(ctx), st>>>(A, D.ps.ray_o, D.ps.ray_d, nullptr, &D.d_counters->q_count[0], &D.d_counters->work_extend[0], D.ps.hit, nullptr, nullptr, D.d_totals, tri_min, refill_busy, wait_mode, stack_entries(ctx), coop_min, refill_hi, refill_patience);
  else k_trace<false, true><<<D.trace_blocks, kTraceThreads, stack_bytes(ctx, false), st>>>(A, D.ps.ray_o, D.ps.ray_d, nullptr, &D.d_counters->q_count[0], &D.d_counters->work_extend[0], D.ps.hit, nullptr, nullptr, D.d_totals, tri_min, refill_busy, wait_mode, stack_entries(ctx), coop_min, refill_hi, refill_patience);
  CK(cudaGetLastError());
  std::vector<float4> hits(n);
  CK(cudaMemcpyAsync(hits.data(), D.ps.hit, n * sizeof(float4), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int64_t i = 0; i < n; i++) {
    int sl; std::memcpy(&sl, &hits[i].w, 4);
    if (any) out_id[i] = sl >= 0 ? 1 : 0;
    else { out_id[i] = sl >= 0 ? ctx->wide.slot_prim[sl] : -1; if (out_t) out_t[i] = sl >= 0 ? hits[i].x : INFINITY; }
  }
  return DSRT_OK;
}

int dsrt_trace_closest(dsrt_ctx* ctx, int64_t n, const float* o, const float* d, const float* tmax, int32_t* prim_id, float* t) {
  if (!ctx || !prim_id) return DSRT_ERR_INVALID;
  return trace_batch(ctx, false, n, o, d, tmax, prim_id, t);
}
int dsrt_trace_any(dsrt_ctx* ctx, int64_t n, const float* o, const float* d, const float* tmax, int32_t* hit) {
  if (!ctx || !hit) return DSRT_ERR_INVALID;
  return trace_batch(ctx, true, n, o, d, tmax, hit, nullptr);
}

int dsrt_measure_read_bandwidth(dsrt_ctx* ctx, int64_t bytes, int32_t repeats, double* gb_per_s) {
  if (!ctx || !gb_per_s || bytes < 16 || repeats < 1) return DSRT_ERR_INVALID;
  DevState& D = ctx->devs[0];
  CK(cudaSetDevice(D.device));
  const size_t n_vec = (size_t)bytes / 16;
  Scratch tmp;
  uint4* d_buf = nullptr; uint32_t* d_sink = nullptr;
  CK(tmp.alloc(&d_buf, n_vec * 16));
  CK(tmp.alloc(&d_sink, 16));
  CK(cudaMemsetAsync(d_buf, 1, n_vec * 16, D.stream));
  const int grid = D.sm_count * 8;                 // 8 CTAs of 256 threads per SM: the full 2048 resident threads
  k_read_sweep<<<grid, 256, 0, D.stream>>>(d_buf, n_vec, 1, d_sink);          // warm: pulls the buffer into L2
  cudaEvent_t e0, e1; CK(tmp.event(&e0)); CK(tmp.event(&e1));
  CK(cudaEventRecord(e0, D.stream));
  k_read_sweep<<<grid, 256, 0, D.stream>>>(d_buf, n_vec, repeats, d_sink);
  CK(cudaEventRecord(e1, D.stream));
  CK(cudaGetLastError());
  CK(cudaEventSynchronize(e1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  *gb_per_s = ms > 0 ? (double)n_vec * 16.0 * repeats / ((double)ms * 1e-3) / 1e9 : 0.0;
  return DSRT_OK;
}

int dsrt_tonemap(dsrt_ctx* ctx, const float* rgb, int64_t n_pixels, uint32_t* rgba8) {
  if (!ctx || !rgb || !rgba8 || n_pixels < 0) return DSRT_ERR_INVALID;
  DevState& D = ctx->devs[0];
  CK(cudaSetDevice(D.device));
  Scratch tmp;
  float* d_in = nullptr; uint32_t* d_out = nullptr;
  CK(tmp.alloc(&d_in, (size_t)n_pixels * 3 * sizeof(float) + 16));
  CK(tmp.alloc(&d_out, (size_t)n_pixels * sizeof(uint32_t) + 16));
  CK(cudaMemcpyAsync(d_in, rgb, (size_t)n_pixels * 3 * sizeof(float), cudaMemcpyHostToDevice, D.stream));
  k_resolve<<<(unsigned)((n_pixels + 255) / 256), 256, 0, D.stream>>>(d_in, nullptr, d_out, (int)n_pixels, 1.0f);
  CK(cudaMemcpyAsync(rgba8, d_out, (size_t)n_pixels * sizeof(uint32_t), cudaMemcpyDeviceToHost, D.stream));
  CK(cudaStreamSynchronize(D.stream));
  return DSRT_OK;
}

}  // extern "C"
