// integration/cuda_path_tracer_shim.h -- the reference-side binding of libdsrt.so.
//
// Drop-in for the reference's cuda_src/setup.h:90-148: the SAME class name and the SAME methods that
// Application::transferToGPU / startGPURayTracing call (src/application.cpp:219-221, 766-786), implemented on the C ABI of
// include/dsrt.h instead of the reference's own kernels.  A maintainer replaces cuda_src/setup.{h,cu} by this pair and links
// `pathtracer` against libdsrt.so instead of libGPUAccel.so (src/CMakeLists.txt:78); nothing else in the reference changes.
//
// This file is compiled and tested: oracle/build_ref.sh builds it against the reference's real PathTracer (headers and
// translation units where they lie) into oracle/_ref/ref_gpu_driver (integration/ref_gpu_driver.cpp restates
// Application::startGPURayTracing), and tests/test_gpu_workloads.py::test_reference_pathtracer_drives_libdsrt_through_the_shim
// compares its frame with the product's own host path on a B200.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "pathtracer.h"                 // the reference's src/pathtracer.h (cuda_src/setup.h includes "../src/pathtracer.h")
#include "static_scene/sphere.h"
#include "static_scene/triangle.h"
#include "static_scene/light.h"

#include "dsrt.h"                       // include/dsrt.h of this repository

class CUDAPathTracer {
 public:
  explicit CUDAPathTracer(CMU462::PathTracer* _pathTracer);      // setup.cu:77-95
  ~CUDAPathTracer();                                             // setup.cu:97-115

  void loadCamera();          // setup.cu:221-247  -> dsrt_set_camera
  void loadPrimitives();      // setup.cu:249-402  -> first half of dsrt_set_scene (primitives + BSDF table)
  void loadLights();          // setup.cu:689-774  -> second half of dsrt_set_scene
  void loadBVH();             // setup.cu:415-476  -> dsrt_set_bvh with the reference's own BVHAccel, flattened pre-order
  void loadParameters();      // setup.cu:777-811  -> dsrt_set_params
  void createFrameBuffer();   // setup.cu:203-219  -> host-side frame (the device frame lives inside the dsrt context)
  void init();                // setup.cu:181-201: the same call order, then dsrt_build_accel
  void startRayTracing() { startRayTracingPT(); }
  void startRayTracingPT();   // setup.cu:147-179  -> dsrt_render
  void updateHostSampleBuffer();   // setup.cu:813-827 -> PathTracer::updateBufferFromGPU

  // additions (not in the reference's class): status instead of exit(EXIT_FAILURE), counters, seed
  bool ok() const { return status == 0; }
  const std::string& error() const { return message; }
  const dsrt_stats& stats() const { return last_stats; }
  void setSeed(uint32_t s) { seed = s; }

 private:
  void check(int rc, const char* what);
  CMU462::PathTracer* pathTracer;
  dsrt_ctx* ctx = nullptr;
  int status = 0; std::string message;
  int screenW = 0, screenH = 0;
  uint32_t seed = 0;
  std::vector<float> frame;                       // what the reference's cudaMemcpy D->H fills (setup.cu:814-817)
  dsrt_stats last_stats{};
  std::map<CMU462::StaticScene::Primitive*, int> primMap;        // setup.h:109
  // staging for dsrt_set_scene (the C ABI takes primitives and lights in one call)
  std::vector<int32_t> prim_type, prim_bsdf, bsdf_type, light_type;
  std::vector<double> tri_pos, tri_nrm, sphere, light_param;
  std::vector<float> bsdf_param;
  bool have_prims = false, have_lights = false;
  void commitScene();
};
