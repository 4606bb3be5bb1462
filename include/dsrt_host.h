/* dsrt_host.h -- C entry points of the host side (libdsrt_host.so): COLLADA import and the PathTracer mirror.
 * The host side is C++ (dsgpuraytracing_b200/csrc/host) like the reference's; these wrappers exist so that tests and
 * foreign callers can reach it without C++ types.
 *   dsrth_load_dae      = ColladaParser::load + Application::load + loadCamera
 *                         (reference src/collada/collada.cpp:131-225, src/application.cpp:223-299, 823-853)
 *   dsrth_render_file   = main.cpp:71-188 headless GPU path: load -> PathTracer::set_camera/set_scene/set_frame_size
 *                         -> start_raytracing -> (save_image), through the class in csrc/host/pathtracer.h
 */
#ifndef DSRT_HOST_H
#define DSRT_HOST_H
#include "dsrt.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsrth_scene dsrth_scene;

/* cam_info may be NULL (default orbit camera of Application::load).  On failure returns non-zero and writes a
 * message to err (the reference exit()s). */
int dsrth_load_dae(const char* path, int32_t width, int32_t height, const char* cam_info, dsrth_scene** out,
                   char* err, int32_t err_len);
void dsrth_free(dsrth_scene* s);
/* fills *out with pointers INTO the handle (valid until dsrth_free) */
int dsrth_get_scene(const dsrth_scene* s, dsrt_scene* out);
/* pos[3], c2w[9] column-major, screenW, screenH, screenDist, hFov, vFov */
int dsrth_get_camera(const dsrth_scene* s, double* cam17);

/* Full headless render through the PathTracer class.  rgb_out (w*h*3 floats, row 0 = bottom) and png_path may be
 * NULL.  bvh_seconds / render_seconds may be NULL. */
int dsrth_render_file(const char* dae_path, const char* cam_info, int32_t width, int32_t height, int32_t ns_aa,
                      int32_t ns_area_light, int32_t max_ray_depth, int32_t n_gpus, uint32_t seed, float* rgb_out,
                      const char* png_path, dsrt_stats* stats, double* bvh_seconds, double* render_seconds,
                      char* err, int32_t err_len);

/* Loader options: "direct_triangles" (0/1, default 0): meshes the half-edge builder rejects (non-manifold, inconsistently
 * oriented; the reference exit(1)s, src/halfEdgeMesh.cpp:165-175) are imported as plain indexed triangles, polygons
 * fan-triangulated.  Returns non-zero for an unknown option. */
int dsrth_set_loader_option(const char* name, int32_t value);

/* Environment map of the -e option (reference src/main.cpp:30-67, 99-101: load_exr through tinyexr): scan-line OpenEXR
 * (NONE / RLE / ZIPS / ZIP / PIZ; R, G, B channels of HALF / FLOAT / UINT) or .pfm.  Call with rgb == NULL to get the size, then
 * with a buffer of width*height*3 floats (top row first, the layout dsrt_set_envmap takes). */
int dsrth_load_envmap(const char* path, int32_t* width, int32_t* height, float* rgb, int64_t rgb_capacity_floats,
                      char* err, int32_t err_len);

#ifdef __cplusplus
}
#endif
#endif /* DSRT_HOST_H */
