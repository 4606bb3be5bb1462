// pathtracer.cpp -- see pathtracer.h.
#include "pathtracer.h"

#include <zlib.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>

namespace dsrt_host {

PathTracer::PathTracer(size_t ns_aa_, size_t max_ray_depth_, size_t ns_area_light_, size_t ns_diff_, size_t ns_glsy_,
                       size_t ns_refr_, size_t num_threads, HDRImageBuffer* envmap) {
  state = INIT;
  ns_aa = ns_aa_; max_ray_depth = max_ray_depth_; ns_area_light = ns_area_light_;
  ns_diff = ns_diff_; ns_glsy = ns_diff_;          // sic: pathtracer.cpp:39 assigns ns_diff
  ns_refr = ns_refr_;
  (void)ns_glsy_;
  numWorkerThreads = num_threads;
  envMap = envmap;                                 // pathtracer.cpp:41-45: becomes an EnvironmentLight appended to the scene
}

PathTracer::~PathTracer() { if (ctx) dsrt_destroy(ctx); }

bool PathTracer::fail(const std::string& what) {
  error = what + (ctx ? std::string(": ") + dsrt_last_error(ctx) : std::string());
  fprintf(stderr, "[PathTracer] error: %s\n", error.c_str());
  return false;
}

bool PathTracer::has_valid_configuration() const { return scene && camera && !sampleBuffer.is_empty(); }

void PathTracer::set_scene(Scene* s) {
  if (state != INIT) return;
  scene = s;
  build_accel();
  if (has_valid_configuration()) state = READY;
}

void PathTracer::set_camera(Camera* c) {
  camera = c;
  if (has_valid_configuration()) state = READY;
}

void PathTracer::set_frame_size(size_t width, size_t height) {
  if (state != INIT && state != READY) stop();
  sampleBuffer.resize(width, height);
  frameBuffer.resize(width, height);
  if (has_valid_configuration()) state = READY;
}

void PathTracer::build_accel() {
  fprintf(stdout, "[PathTracer] Collecting primitives... "); fflush(stdout);
  n_primitives = (size_t)scene->n_prims();
  fprintf(stdout, "Done! (%zu primitives)\n", n_primitives);
  fprintf(stdout, "[PathTracer] Building BVH... "); fflush(stdout);
  auto t0 = std::chrono::steady_clock::now();
  dsrt_scene sc{};
  sc.n_prims = scene->n_prims(); sc.prim_type = scene->prim_type.data(); sc.prim_bsdf = scene->prim_bsdf.data();
  sc.tri_pos = scene->tri_pos.data(); sc.tri_nrm = scene->tri_nrm.data(); sc.sphere = scene->sphere.data();
  sc.n_bsdf = (int)scene->bsdf_type.size(); sc.bsdf_type = scene->bsdf_type.data(); sc.bsdf_param = scene->bsdf_param.data();
  sc.n_lights = (int)scene->light_type.size(); sc.light_type = scene->light_type.data(); sc.light_param = scene->light_param.data();
  const size_t cap = 2 * std::max<size_t>(n_primitives, 1);
  node_bbox.assign(cap * 6, 0); node_start.assign(cap, 0); node_range.assign(cap, 0); node_left.assign(cap, 0); node_right.assign(cap, 0);
  prim_order.assign(std::max<size_t>(n_primitives, 1), 0);
  int32_t n_nodes = 0;
  if (dsrt_build_bvh2(&sc, node_bbox.data(), node_start.data(), node_range.data(), node_left.data(), node_right.data(), prim_order.data(), &n_nodes)) {
    fail("dsrt_build_bvh2"); return;
  }
  node_bbox.resize((size_t)n_nodes * 6); node_start.resize(n_nodes); node_range.resize(n_nodes); node_left.resize(n_nodes); node_right.resize(n_nodes);
  prim_order.resize(n_primitives);
  bvh_build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  fprintf(stdout, "Done! (%.4f sec, %d nodes)\n", bvh_build_seconds, n_nodes);
  accel_uploaded = false;
}

// pathtracer.cpp:148-171.  Callable from another thread while start_raytracing() is blocked in the render: the core stops
// after the chunk of samples in flight (dsrt_cancel) and start_raytracing() returns with the frame rendered so far.
void PathTracer::stop() {
  if (state == RENDERING && ctx) { dsrt_cancel(ctx); return; }      // start_raytracing() moves the state to READY when it returns
  if (state == DONE || state == VISUALIZE) state = READY;
}

void PathTracer::clear() {
  if (state != READY) return;
  scene = nullptr; camera = nullptr;
  sampleBuffer.resize(0, 0); frameBuffer.resize(0, 0);
  node_bbox.clear(); node_start.clear(); node_range.clear(); node_left.clear(); node_right.clear(); prim_order.clear();
  accel_uploaded = false;
  state = INIT;
}

void PathTracer::start_raytracing() {
  if (state != READY) return;
  if (useCPU) { fail("useCPU requested: this port has no CPU fallback (run the reference for -c)"); return; }
  state = RENDERING;
  sampleBuffer.clear(); frameBuffer.clear();
  auto tp = std::chrono::steady_clock::now();
  auto lap = [&tp]() { auto n = std::chrono::steady_clock::now(); double s = std::chrono::duration<double>(n - tp).count(); tp = n; return s; };
  const bool first_use = !ctx || !accel_uploaded;
  if (!ctx) {
    std::vector<int> devs(n_gpus); for (int i = 0; i < n_gpus; i++) devs[i] = i;
    int rc = n_gpus > 1 ? dsrt_create_multi(n_gpus, devs.data(), &ctx) : dsrt_create(0, &ctx);
    if (rc) { fail("dsrt_create"); if (ctx) { dsrt_destroy(ctx); ctx = nullptr; } state = READY; return; }
    for (const auto& o : options)
      if (dsrt_set_option(ctx, o.first.c_str(), o.second)) { fail("dsrt_set_option(" + o.first + ")"); dsrt_destroy(ctx); ctx = nullptr; state = READY; return; }
  }
  if (!accel_uploaded) {
    dsrt_scene sc{};
    sc.n_prims = scene->n_prims(); sc.prim_type = scene->prim_type.data(); sc.prim_bsdf = scene->prim_bsdf.data();
    sc.tri_pos = scene->tri_pos.data(); sc.tri_nrm = scene->tri_nrm.data(); sc.sphere = scene->sphere.data();
    sc.n_bsdf = (int)scene->bsdf_type.size(); sc.bsdf_type = scene->bsdf_type.data(); sc.bsdf_param = scene->bsdf_param.data();
    sc.n_lights = (int)scene->light_type.size(); sc.light_type = scene->light_type.data(); sc.light_param = scene->light_param.data();
    dsrt_bvh2 b{}; b.n_nodes = (int)node_start.size(); b.node_bbox = node_bbox.data(); b.node_start = node_start.data();
    b.node_range = node_range.data(); b.node_left = node_left.data(); b.node_right = node_right.data(); b.prim_order = prim_order.data();
    if (dsrt_set_scene(ctx, &sc)) { fail("dsrt_set_scene"); state = READY; return; }
    if (envMap && dsrt_set_envmap(ctx, (int)envMap->w, (int)envMap->h, envMap->data.data())) { fail("dsrt_set_envmap"); state = READY; return; }
    if (dsrt_set_bvh(ctx, &b)) { fail("dsrt_set_bvh"); state = READY; return; }
  }
  const double t_ctx = lap();
  if (dsrt_set_params(ctx, (int)ns_aa, (int)ns_area_light, (int)max_ray_depth, seed)) { fail("dsrt_set_params"); state = READY; return; }
  if (!accel_uploaded) { if (dsrt_build_accel(ctx)) { fail("dsrt_build_accel"); state = READY; return; } accel_uploaded = true; }
  if (first_use) fprintf(stdout, "[PathTracer] GPU context + scene copy %.4f sec, wide-BVH collapse + upload %.4f sec\n", t_ctx, lap());
  // generate_ray reads the camera's own screenW/H/screenDist (camera.cpp:113-129); the buffers have the frame size
  if (camera->screenW != sampleBuffer.w || camera->screenH != sampleBuffer.h) {
    fprintf(stderr, "[PathTracer] warning: camera is configured for %zux%zu, frame is %zux%zu (scene without a camera node?)\n",
            camera->screenW, camera->screenH, sampleBuffer.w, sampleBuffer.h);
  }
  // the core renders a (camera.screenW x camera.screenH)-parameterised ray grid over the frame's pixels
  const double sx = (double)camera->screenW / (double)sampleBuffer.w, sy = (double)camera->screenH / (double)sampleBuffer.h;
  if (sx != 1.0 || sy != 1.0) { fail("camera/frame size mismatch is not supported"); state = READY; return; }
  if (dsrt_set_camera(ctx, camera->pos, camera->c2w, (int)sampleBuffer.w, (int)sampleBuffer.h, camera->screenDist)) { fail("dsrt_set_camera"); state = READY; return; }
  fprintf(stdout, "[PathTracer] Rendering... "); fflush(stdout);
  auto t0 = std::chrono::steady_clock::now();
  // sampleBuffer and the tone-mapped frameBuffer both come straight from the fused reduce + resolve kernel
  const int rc = dsrt_render_tonemapped(ctx, 0, (int)ns_aa, 1, sampleBuffer.data.data(), frameBuffer.data.data(), &last_stats);
  render_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc == DSRT_CANCELLED) {            // stop(): keep what was rendered, go back to READY like the reference
    fprintf(stdout, "stopped (%.4f sec, %llu camera samples)\n", render_seconds, (unsigned long long)last_stats.camera_samples);
    state = READY; return;
  }
  if (rc) { fail("dsrt_render"); state = READY; return; }
  fprintf(stdout, "GPU ray tracing done! (%.4f sec)\n", render_seconds);
  state = DONE;
}

// HDRImageBuffer::update_pixel + toColor + ImageBuffer::update_pixel (image.h:113-117, 174-189, 49-58) on the host: kept for
// callers that hand in their own linear frame (cuda_src/setup.cu:829-843); start_raytracing() itself takes the tone-mapped
// frame from the device (dsrt_render_tonemapped).
void PathTracer::updateBufferFromGPU(const float* gpuBuffer) {
  const size_t n = sampleBuffer.w * sampleBuffer.h;
  std::memcpy(sampleBuffer.data.data(), gpuBuffer, n * 3 * sizeof(float));
  const float gamma = 2.2f, level = 1.0f;
  const float one_over_gamma = 1.0f / gamma;
  const float exposure = std::sqrt(std::pow(2.0f, level));
  for (size_t i = 0; i < n; i++) {
    uint32_t p = 255u << 24;
    for (int k = 0; k < 3; k++) {
      float c = std::pow(gpuBuffer[3 * i + k] * exposure, one_over_gamma);
      c = (c < 1.f) ? c : 1.f;              // clamp(0.f, 1.f, c) with the reference's argument order == min(1, c)
      p += ((uint32_t)(c * 255)) << (8 * k);
    }
    frameBuffer.data[i] = p;
  }
}

bool PathTracer::save_image(const std::string& filename) {
  const size_t w = frameBuffer.w, h = frameBuffer.h;
  std::vector<uint32_t> out(w * h);
  for (size_t i = 0; i < h; ++i) std::memcpy(&out[i * w], &frameBuffer.data[(h - i - 1) * w], 4 * w);   // pathtracer.cpp:666-668
  fprintf(stderr, "[PathTracer] Saving to file: %s... ", filename.c_str());
  std::string err;
  if (!write_png_rgba8(filename, (const uint8_t*)out.data(), w, h, err)) { fprintf(stderr, "failed: %s\n", err.c_str()); error = err; return false; }
  fprintf(stderr, "Done!\n");
  return true;
}

void PathTracer::save_image() {
  time_t rawtime; time(&rawtime);
  std::string filename = "Screen Shot ";
  filename += useCPU ? "CPU " : "GPU ";
  filename += std::string(ctime(&rawtime));
  filename.erase(filename.end() - 1);
  filename += ".png";
  save_image(filename);
}

// ---- PNG -----------------------------------------------------------------------------------------------------------
static void put_u32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
static void chunk(std::vector<uint8_t>& png, const char* type, const std::vector<uint8_t>& data) {
  put_u32(png, (uint32_t)data.size());
  std::vector<uint8_t> body(type, type + 4);
  body.insert(body.end(), data.begin(), data.end());
  png.insert(png.end(), body.begin(), body.end());
  put_u32(png, (uint32_t)crc32(0L, body.data(), (uInt)body.size()));
}
bool write_png_rgba8(const std::string& path, const uint8_t* rgba, size_t w, size_t h, std::string& err) {
  std::vector<uint8_t> raw; raw.reserve(h * (w * 4 + 1));
  for (size_t y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), rgba + y * w * 4, rgba + (y + 1) * w * 4); }
  uLongf clen = compressBound((uLong)raw.size());
  std::vector<uint8_t> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) { err = "zlib compress2 failed"; return false; }
  comp.resize(clen);
  std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  std::vector<uint8_t> ihdr; put_u32(ihdr, (uint32_t)w); put_u32(ihdr, (uint32_t)h);
  ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  chunk(png, "IHDR", ihdr); chunk(png, "IDAT", comp); chunk(png, "IEND", {});
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { err = "cannot write " + path; return false; }
  fwrite(png.data(), 1, png.size(), f); fclose(f);
  return true;
}

}  // namespace dsrt_host
