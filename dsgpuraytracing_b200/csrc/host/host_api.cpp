// host_api.cpp -- C wrappers declared in include/dsrt_host.h
#include <cstring>
#include <exception>
#include <memory>
#include <string>

#include "../../../include/dsrt_host.h"
#include "image_io.h"
#include "pathtracer.h"
#include "scene_loader.h"

using namespace dsrt_host;

struct dsrth_scene { FlatScene scene; HostCamera camera; };

static void set_err(char* err, int32_t n, const std::string& m) {
  if (err && n > 0) { std::strncpy(err, m.c_str(), (size_t)n - 1); err[n - 1] = 0; }
}

// No C++ exception may cross the C boundary (a ctypes / cgo caller would be terminated): every entry point that parses
// files or allocates from file-supplied sizes runs its body through guard().
template <class F> static int guard(char* err, int32_t err_len, F&& body) {
  try { return body(); }
  catch (const std::exception& e) { set_err(err, err_len, std::string("exception: ") + e.what()); return DSRT_ERR_INVALID; }
  catch (...) { set_err(err, err_len, "unknown exception"); return DSRT_ERR_INVALID; }
}

extern "C" {

int dsrth_load_dae(const char* path, int32_t width, int32_t height, const char* cam_info, dsrth_scene** out, char* err, int32_t err_len) {
  if (!path || !out || width <= 0 || height <= 0) { set_err(err, err_len, "bad arguments"); return DSRT_ERR_INVALID; }
  *out = nullptr;
  return guard(err, err_len, [&]() -> int {
    std::unique_ptr<dsrth_scene> s(new dsrth_scene());
    std::string e;
    if (!load_collada(path, (size_t)width, (size_t)height, s->scene, s->camera, e)) { set_err(err, err_len, e); return DSRT_ERR_INVALID; }
    if (cam_info && *cam_info && !s->camera.load_info(cam_info, e)) { set_err(err, err_len, e); return DSRT_ERR_INVALID; }
    *out = s.release();
    return DSRT_OK;
  });
}

void dsrth_free(dsrth_scene* s) { delete s; }

int dsrth_get_scene(const dsrth_scene* s, dsrt_scene* o) {
  if (!s || !o) return DSRT_ERR_INVALID;
  const FlatScene& f = s->scene;
  o->n_prims = f.n_prims(); o->prim_type = f.prim_type.data(); o->prim_bsdf = f.prim_bsdf.data();
  o->tri_pos = f.tri_pos.data(); o->tri_nrm = f.tri_nrm.data(); o->sphere = f.sphere.data();
  o->n_bsdf = (int32_t)f.bsdf_type.size(); o->bsdf_type = f.bsdf_type.data(); o->bsdf_param = f.bsdf_param.data();
  o->n_lights = (int32_t)f.light_type.size(); o->light_type = f.light_type.data(); o->light_param = f.light_param.data();
  return DSRT_OK;
}

int dsrth_get_camera(const dsrth_scene* s, double* c) {
  if (!s || !c) return DSRT_ERR_INVALID;
  const HostCamera& k = s->camera;
  for (int i = 0; i < 3; i++) c[i] = k.pos[i];
  for (int i = 0; i < 9; i++) c[3 + i] = k.c2w[i];
  c[12] = (double)k.screenW; c[13] = (double)k.screenH; c[14] = k.screenDist; c[15] = k.hFov; c[16] = k.vFov;
  return DSRT_OK;
}

int dsrth_render_file(const char* dae_path, const char* cam_info, int32_t width, int32_t height, int32_t ns_aa, int32_t ns_area_light,
                      int32_t max_ray_depth, int32_t n_gpus, uint32_t seed, float* rgb_out, const char* png_path, dsrt_stats* stats,
                      double* bvh_seconds, double* render_seconds, char* err, int32_t err_len) {
  if (!dae_path || width <= 0 || height <= 0) { set_err(err, err_len, "bad arguments"); return DSRT_ERR_INVALID; }
  return guard(err, err_len, [&]() -> int {
  FlatScene scene; HostCamera camera; std::string e;
  if (!load_collada(dae_path, (size_t)width, (size_t)height, scene, camera, e)) { set_err(err, err_len, e); return DSRT_ERR_INVALID; }
  PathTracer pt((size_t)ns_aa, (size_t)max_ray_depth, (size_t)ns_area_light, 1, 1, 1, 1, nullptr);
  pt.set_gpus(n_gpus); pt.set_seed(seed);
  pt.set_camera(&camera);                       // Application::set_up_pathtracer, application.cpp:624-633
  pt.set_scene(&scene);
  pt.set_frame_size((size_t)width, (size_t)height);
  if (cam_info && *cam_info && !camera.load_info(cam_info, e)) { set_err(err, err_len, e); return DSRT_ERR_INVALID; }   // main.cpp:163-165
  pt.start_raytracing();
  if (!pt.is_done()) { set_err(err, err_len, pt.last_error()); return DSRT_ERR_CUDA; }
  if (rgb_out) std::memcpy(rgb_out, pt.sampleBuffer.data.data(), pt.sampleBuffer.data.size() * sizeof(float));
  if (stats) *stats = pt.stats();
  if (bvh_seconds) *bvh_seconds = pt.bvh_build_seconds;
  if (render_seconds) *render_seconds = pt.render_seconds;
  if (png_path && *png_path && !pt.save_image(png_path)) { set_err(err, err_len, pt.last_error()); return DSRT_ERR_INVALID; }
  return DSRT_OK;
  });
}

int dsrth_set_loader_option(const char* name, int32_t value) {
  if (!name) return DSRT_ERR_INVALID;
  if (std::string(name) == "direct_triangles") { set_direct_triangle_fallback(value != 0); return DSRT_OK; }
  return DSRT_ERR_INVALID;
}

int dsrth_load_envmap(const char* path, int32_t* width, int32_t* height, float* rgb, int64_t cap, char* err, int32_t err_len) {
  if (!path || !width || !height) { set_err(err, err_len, "bad arguments"); return DSRT_ERR_INVALID; }
  return guard(err, err_len, [&]() -> int {
  HDRImageBuffer img; std::string e;
  if (!load_envmap(path, img, e)) { set_err(err, err_len, e); return DSRT_ERR_INVALID; }
  *width = (int32_t)img.w; *height = (int32_t)img.h;
  if (rgb) {
    if (cap < (int64_t)img.data.size()) { set_err(err, err_len, "buffer too small"); return DSRT_ERR_INVALID; }
    std::memcpy(rgb, img.data.data(), img.data.size() * sizeof(float));
  }
  return DSRT_OK;
  });
}

}  // extern "C"
