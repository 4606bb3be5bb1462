// main.cpp -- `pathtracer`: command-line front end with the reference's flags (reference src/main.cpp:71-188).
//   -s <spp>  -l <area-light samples>  -m <max ray depth>  -t <threads, accepted and ignored>  -w <width>  -h <height>
//   -f <cam_*.info>  -c (CPU render: refused, there is no CPU fallback)  -v (viewer: not part of this port)
//   -e <environment map>: the reference declares -e but leaves it out of its getopt string (main.cpp:85, 99-101) and
//      loads .exr through the vendored tinyexr; here -e works and reads scan-line OpenEXR (NONE / RLE / ZIPS / ZIP / PIZ, host/image_io.cpp)
//      or a binary .pfm (PF, little endian) lat-long map
// additions: -g <number of GPUs>  -o <output.png>  -S <seed>  -r <raw float dump of the linear buffer>
//            -T tolerate non-manifold meshes (direct indexed-triangle import; the reference exit(1)s on them)
//            -j print one JSON line per run (counters, seconds per stage, Mrays/s) on stdout after the human-readable lines
// As in the reference, -h is the frame HEIGHT and the default frame is 1000x1000 (main.cpp:79-80).
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "image_io.h"
#include "pathtracer.h"
#include "scene_loader.h"

using namespace dsrt_host;

static void usage(const char* bin) {
  printf("Usage: %s [options] <scenefile.dae>\n", bin);
  printf("  -s <INT>  camera rays per pixel (default 1)\n  -l <INT>  samples per area light (default 4)\n");
  printf("  -t <INT>  render threads (ignored: the render runs on the GPU)\n  -m <INT>  maximum ray depth (default 1)\n");
  printf("  -w <INT>  frame width (default 1000)\n  -h <INT>  frame height (default 1000)\n  -f <FILE> camera .info file\n");
  printf("  -g <INT>  number of GPUs (default 1)\n  -o <FILE> output PNG (default \"Screen Shot GPU <time>.png\")\n");
  printf("  -e <FILE> lat-long environment map (.exr or .pfm)\n");
  printf("  -T        import meshes the half-edge builder rejects (non-manifold ...) as plain indexed triangles\n");
  printf("  -O <NAME=INT> option of the GPU core (include/dsrt.h dsrt_set_option: skip_null_shadow, regroup_top, wavefront_budget_mb, ...); repeatable\n");
  printf("  -S <INT>  Philox seed (default 0)\n  -r <FILE> also dump the linear float RGB buffer\n  -j        one JSON line of run statistics\n");
}

int main(int argc, char** argv) {
  size_t ns_aa = 1, ns_area_light = 4, max_ray_depth = 1, num_threads = 1;   // application.h:45-58
  int screenW = 1000, screenH = 1000, n_gpus = 1;
  unsigned seed = 0;
  bool useCPU = false, json = false;
  std::string camFileName, outName, rawName, envName;
  std::vector<std::pair<std::string, long long>> coreOptions;
  int opt;
  while ((opt = getopt(argc, argv, "s:l:t:m:f:w:h:g:o:O:S:r:e:vcTj")) != -1) {
    switch (opt) {
      case 's': ns_aa = (size_t)atoi(optarg); break;
      case 'l': ns_area_light = (size_t)atoi(optarg); break;
      case 't': num_threads = (size_t)atoi(optarg); break;
      case 'm': max_ray_depth = (size_t)atoi(optarg); break;
      case 'w': screenW = atoi(optarg); break;
      case 'h': screenH = atoi(optarg); break;
      case 'f': camFileName = optarg; break;
      case 'g': n_gpus = atoi(optarg); break;
      case 'o': outName = optarg; break;
      case 'O': {
        const char* eq = strchr(optarg, '=');
        if (!eq || eq == optarg) { fprintf(stderr, "-O expects NAME=INT\n"); return 1; }
        coreOptions.emplace_back(std::string(optarg, (size_t)(eq - optarg)), atoll(eq + 1));
        break;
      }
      case 'S': seed = (unsigned)strtoul(optarg, nullptr, 10); break;
      case 'r': rawName = optarg; break;
      case 'e': envName = optarg; break;
      case 'c': useCPU = true; break;
      case 'T': set_direct_triangle_fallback(true); break;
      case 'j': json = true; break;
      case 'v': fprintf(stderr, "the interactive viewer is not part of this port\n"); return 1;
      default: usage(argv[0]); return 1;
    }
  }
  if (optind >= argc) { usage(argv[0]); return 1; }
  if (useCPU) { fprintf(stderr, "-c: this port has no CPU fallback; run the reference for a CPU render\n"); return 1; }
  const std::string sceneFilePath = argv[optind];
  printf("Input scene file: %s\n", sceneFilePath.c_str());

  FlatScene scene; HostCamera camera; std::string err;
  if (!load_collada(sceneFilePath, (size_t)screenW, (size_t)screenH, scene, camera, err)) { fprintf(stderr, "error: %s\n", err.c_str()); return 1; }
  HDRImageBuffer envmap;
  if (!envName.empty() && !load_envmap(envName, envmap, err)) { fprintf(stderr, "error: %s\n", err.c_str()); return 1; }
  PathTracer pathtracer(ns_aa, max_ray_depth, ns_area_light, 1, 1, 1, num_threads, envName.empty() ? nullptr : &envmap);
  pathtracer.set_gpus(n_gpus); pathtracer.set_seed(seed);
  for (const auto& o : coreOptions) pathtracer.set_option(o.first, o.second);
  pathtracer.set_camera(&camera);
  pathtracer.set_scene(&scene);
  pathtracer.set_frame_size((size_t)screenW, (size_t)screenH);
  if (!camFileName.empty() && !camera.load_info(camFileName, err)) { fprintf(stderr, "error: %s\n", err.c_str()); return 1; }
  pathtracer.start_raytracing();
  if (!pathtracer.is_done()) { fprintf(stderr, "render failed: %s\n", pathtracer.last_error().c_str()); return 2; }
  const dsrt_stats& st = pathtracer.stats();
  printf("[PathTracer] %llu camera samples, %llu extend + %llu shadow segments, %.4f s on the GPU, %.1f Mrays/s\n",
         (unsigned long long)st.camera_samples, (unsigned long long)st.extend_rays, (unsigned long long)st.shadow_rays, st.gpu_seconds,
         (double)(st.extend_rays + st.shadow_rays) / st.gpu_seconds / 1e6);
  if (json)
    printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %zu, \"light_samples\": %zu, \"max_depth\": %zu, \"gpus\": %d, \"seed\": %u, "
           "\"camera_samples\": %llu, \"extend_rays\": %llu, \"shadow_rays\": %llu, \"gpu_seconds\": %.6f, \"render_call_seconds\": %.6f, "
           "\"bvh_build_seconds\": %.6f, \"kernel_launches\": %u, \"batches\": %u, \"mrays_per_s\": %.2f}\n",
           sceneFilePath.c_str(), screenW, screenH, ns_aa, ns_area_light, max_ray_depth, n_gpus, seed, (unsigned long long)st.camera_samples,
           (unsigned long long)st.extend_rays, (unsigned long long)st.shadow_rays, st.gpu_seconds, pathtracer.render_seconds, pathtracer.bvh_build_seconds,
           st.kernel_launches, st.batches, (double)(st.extend_rays + st.shadow_rays) / st.gpu_seconds / 1e6);
  if (!rawName.empty()) {
    FILE* f = fopen(rawName.c_str(), "wb");
    if (f) { fwrite(pathtracer.sampleBuffer.data.data(), sizeof(float), pathtracer.sampleBuffer.data.size(), f); fclose(f); }
  }
  if (outName.empty()) pathtracer.save_image(); else if (!pathtracer.save_image(outName)) return 3;
  return 0;
}
