// pathtracer.h -- host-side mirror of the reference's renderer state machine (reference src/pathtracer.h:63-268)
// for the GPU render path.  Same surface -- constructor arguments, set_scene / set_camera / set_frame_size /
// build_accel / start_raytracing / stop / clear / save_image, the INIT..DONE states, sampleBuffer / frameBuffer --
// but start_raytracing() drives the B200 core through the C ABI of include/dsrt.h instead of spawning CPU worker
// threads (pathtracer.cpp:192-221) or the CUDAPathTracer class (application.cpp:766-786).  There is no CPU fallback.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/dsrt.h"
#include "scene_loader.h"

namespace dsrt_host {

// HDRImageBuffer (reference src/image.h:85-199): linear float RGB, row 0 = bottom
struct HDRImageBuffer {
  size_t w = 0, h = 0;
  std::vector<float> data;   // w*h*3
  void resize(size_t w_, size_t h_) { w = w_; h = h_; data.assign(w * h * 3, 0.f); }
  void clear() { std::fill(data.begin(), data.end(), 0.f); }
  bool is_empty() const { return w == 0 && h == 0; }
};
// ImageBuffer (image.h:16-77): RGBA8 packed 0xAABBGGRR
struct ImageBuffer {
  size_t w = 0, h = 0;
  std::vector<uint32_t> data;
  void resize(size_t w_, size_t h_) { w = w_; h = h_; data.assign(w * h, 0u); }
  void clear() { std::fill(data.begin(), data.end(), 0u); }
};

using Scene = FlatScene;      // StaticScene::Scene, flattened
using Camera = HostCamera;

class PathTracer {
 public:
  PathTracer(size_t ns_aa = 1, size_t max_ray_depth = 4, size_t ns_area_light = 1, size_t ns_diff = 1, size_t ns_glsy = 1,
             size_t ns_refr = 1, size_t num_threads = 1, HDRImageBuffer* envmap = nullptr);
  ~PathTracer();

  void set_scene(Scene* scene);                      // pathtracer.cpp:76-98 (takes the scene, runs build_accel)
  void set_camera(Camera* camera);                   // :100-111
  void set_frame_size(size_t width, size_t height);  // :113-122
  void build_accel();                                // :224-248: collect primitives, build the SAH BVH (+ upload)
  void start_raytracing();                           // :192-221: here synchronous, like Application::startGPURayTracing
  void stop();                                       // :148-171
  void clear();                                      // :173-183
  bool is_done() const { return state == DONE; }
  void save_image();                                 // :649-674 ("Screen Shot GPU <ctime>.png")
  bool save_image(const std::string& filename);
  void updateBufferFromGPU(const float* gpuBuffer);  // cuda_src/setup.cu:829-843

  enum State { INIT, READY, VISUALIZE, RENDERING, DONE };

  // extensions of this port
  void set_gpus(int n) { n_gpus = n < 1 ? 1 : n; }
  void set_seed(uint32_t s) { seed = s; }
  // dsrt_set_option passthrough (include/dsrt.h: "skip_null_shadow", "regroup_top", "wavefront_budget_mb", ...); applied when the
  // GPU context is created, before the scene is handed over
  void set_option(const std::string& name, int64_t value) { options.emplace_back(name, value); }
  const std::string& last_error() const { return error; }
  const dsrt_stats& stats() const { return last_stats; }

  HDRImageBuffer* envMap = nullptr;   // lat-long environment map (row 0 = +y pole), optional
  bool useCPU = false;
  State state = INIT;
  Scene* scene = nullptr;
  Camera* camera = nullptr;
  size_t max_ray_depth, ns_aa, ns_area_light, ns_diff, ns_glsy, ns_refr;
  HDRImageBuffer sampleBuffer;
  ImageBuffer frameBuffer;
  size_t numWorkerThreads;
  size_t n_primitives = 0;          // PathTracer::primitives.size()
  // the reference-identical binary SAH BVH (BVHAccel), flattened
  std::vector<double> node_bbox; std::vector<int32_t> node_start, node_range, node_left, node_right, prim_order;
  double bvh_build_seconds = 0, render_seconds = 0;

 private:
  bool has_valid_configuration() const;
  bool fail(const std::string& what);
  dsrt_ctx* ctx = nullptr;
  bool accel_uploaded = false;
  int n_gpus = 1;
  uint32_t seed = 0;
  std::vector<std::pair<std::string, int64_t>> options;
  std::string error;
  dsrt_stats last_stats{};
};

// 8-bit RGBA PNG writer (stands in for the vendored lodepng::encode the reference calls, pathtracer.cpp:671)
bool write_png_rgba8(const std::string& path, const uint8_t* rgba, size_t w, size_t h, std::string& err);

}  // namespace dsrt_host
