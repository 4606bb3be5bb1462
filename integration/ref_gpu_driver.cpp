// integration/ref_gpu_driver.cpp -- TEST INFRASTRUCTURE for the reference-side binding.
//
// The reference's own ColladaParser / DynamicScene / PathTracer (compiled where they lie, oracle/build_ref.sh) set up a scene
// exactly as its `pathtracer` binary does (oracle/ref_app.h), and then this file restates Application::startGPURayTracing
// (src/application.cpp:766-786) with the CUDAPathTracer of integration/cuda_path_tracer_shim.h -- i.e. the reference
// application rendering through libdsrt.so.  Writes the linear float frame (PathTracer::sampleBuffer) and the counters.
//
//   ref_gpu_driver [-s -l -m -w -h -f cam.info] [--seed N] --raw frame.f32 scene.dae
#include "../oracle/ref_app.h"
#include "cuda_path_tracer_shim.h"

int main(int argc, char** argv) {
  Args a; std::string raw;
  for (int i = 1; i < argc; i++) {
    std::string s = argv[i];
    auto next = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", s.c_str()); exit(2); } return argv[++i]; };
    if (s == "-s") a.spp = atoi(next());
    else if (s == "-l") a.nl = atoi(next());
    else if (s == "-m") a.depth = atoi(next());
    else if (s == "-w") a.w = atoi(next());
    else if (s == "-h") a.h = atoi(next());
    else if (s == "-f") a.cam = next();
    else if (s == "--seed") a.seed = (unsigned)strtoul(next(), 0, 10);
    else if (s == "--raw") raw = next();
    else if (s == "--envmap") { a.env_w = atoi(next()); a.env_h = atoi(next()); }
    else a.scene = s;
  }
  if (a.scene.empty()) { fprintf(stderr, "usage: ref_gpu_driver [-s -l -m -w -h -f] [--seed N] [--envmap W H] --raw out.f32 scene.dae\n"); return 2; }
  RefApp app;
  if (int rc = ref_app_setup(a, app)) return rc;
  PathTracer* pathtracer = app.pt;
  pathtracer->useCPU = false;

  // ---- Application::startGPURayTracing, application.cpp:766-786
  CUDAPathTracer* cuPathTracer = new CUDAPathTracer(pathtracer);
  cuPathTracer->setSeed(a.seed);
  cuPathTracer->init();                                            // transferToGPU(), application.cpp:219-221
  if (!cuPathTracer->ok()) return 4;
  pathtracer->state = PathTracer::RENDERING;
  pathtracer->continueRaytracing = true;
  pathtracer->sampleBuffer.clear();
  pathtracer->frameBuffer.clear();
  pathtracer->timer.start();
  cuPathTracer->startRayTracingPT();
  pathtracer->timer.stop();
  if (!cuPathTracer->ok()) return 4;
  fprintf(stdout, "GPU ray tracing done! (%.4f sec)\n", pathtracer->timer.duration());
  cuPathTracer->updateHostSampleBuffer();
  const dsrt_stats st = cuPathTracer->stats();
  fprintf(stdout, "segments %llu + %llu\n", (unsigned long long)st.extend_rays, (unsigned long long)st.shadow_rays);
  delete cuPathTracer;

  if (!raw.empty()) {
    FILE* f = fopen(raw.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", raw.c_str()); return 3; }
    fwrite(pathtracer->sampleBuffer.data.data(), sizeof(Spectrum), pathtracer->sampleBuffer.data.size(), f);
    fclose(f);
  }
  fflush(stdout); fflush(stderr);
  _exit(0);
}
