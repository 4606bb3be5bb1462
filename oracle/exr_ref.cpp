// oracle/exr_ref.cpp -- TEST INFRASTRUCTURE (never linked into the product).
// Drives the reference's vendored tinyexr (CMU462/include/CMU462/tinyexr.h, compiled where it lies) so that the
// product's own OpenEXR reader (csrc/host/image_io.cpp) can be pinned against what the reference's load_exr
// (src/main.cpp:30-67) would hand to EnvironmentLight:
//   exr_ref load  in.exr out.bin     -> int32 w, int32 h, then w*h*3 floats (R,G,B per pixel, top row first), restating
//                                       main.cpp:37-64: half channels requested as float, R = images[2], G = [1], B = [0]
//   exr_ref save  out.exr w h half|float seed  -> a procedural RGB image written by tinyexr's own SaveMultiChannelEXRToFile
//                                       (ZIP, 16-line blocks): fixtures that were NOT produced by the product's writer
#define TINYEXR_IMPLEMENTATION
#include "tinyexr.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static unsigned short float_to_half(float f) {   // round to nearest even, no denormal / inf care needed for the fixtures
  uint32_t x; std::memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u; int32_t e = (int32_t)((x >> 23) & 0xff) - 127 + 15; uint32_t m = x & 0x7fffffu;
  if (e <= 0) return (unsigned short)sign;
  if (e >= 31) return (unsigned short)(sign | 0x7c00u);
  uint32_t h = sign | ((uint32_t)e << 10) | (m >> 13);
  if ((m & 0x1fffu) > 0x1000u || ((m & 0x1fffu) == 0x1000u && (h & 1u))) h++;
  return (unsigned short)h;
}

int main(int argc, char** argv) {
  if (argc >= 4 && !std::strcmp(argv[1], "load")) {
    const char* err = nullptr;
    EXRImage exr; InitEXRImage(&exr);
    if (ParseMultiChannelEXRHeaderFromFile(&exr, argv[2], &err) != 0) { std::fprintf(stderr, "parse: %s\n", err ? err : "?"); return 2; }
    for (int i = 0; i < exr.num_channels; i++)
      if (exr.pixel_types[i] == TINYEXR_PIXELTYPE_HALF) exr.requested_pixel_types[i] = TINYEXR_PIXELTYPE_FLOAT;
    if (LoadMultiChannelEXRFromFile(&exr, argv[2], &err) != 0) { std::fprintf(stderr, "load: %s\n", err ? err : "?"); return 3; }
    if (exr.num_channels < 3) { std::fprintf(stderr, "fewer than 3 channels\n"); return 4; }
    const float* r = (const float*)exr.images[2]; const float* g = (const float*)exr.images[1]; const float* b = (const float*)exr.images[0];
    FILE* f = std::fopen(argv[3], "wb"); if (!f) return 5;
    int32_t wh[2] = {exr.width, exr.height}; std::fwrite(wh, 4, 2, f);
    for (size_t i = 0; i < (size_t)exr.width * exr.height; i++) { float px[3] = {r[i], g[i], b[i]}; std::fwrite(px, 4, 3, f); }
    std::fclose(f);
    return 0;
  }
  if (argc >= 7 && !std::strcmp(argv[1], "save")) {
    const int w = std::atoi(argv[3]), h = std::atoi(argv[4]); const bool half = !std::strcmp(argv[5], "half");
    uint32_t s = (uint32_t)std::strtoul(argv[6], nullptr, 10) * 2654435761u + 12345u;
    std::vector<float> ch[3];
    for (int c = 0; c < 3; c++) ch[c].resize((size_t)w * h);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) for (int c = 0; c < 3; c++) {
      s = s * 1664525u + 1013904223u;
      const float noise = (float)(s >> 8) / 16777216.0f;
      float v = 0.25f + 0.5f * std::sin(0.37f * x + 0.11f * y + c) * std::cos(0.05f * y - 0.21f * c) + 0.1f * noise;
      if (x == w / 3 && y == h / 4) v = 900.0f + 10.0f * c;    // a "sun"
      ch[c][(size_t)y * w + x] = v < 0 ? 0.f : v;
    }
    std::vector<unsigned short> hc[3];
    EXRImage img; InitEXRImage(&img);
    img.num_channels = 3;
    const char* names[3] = {"B", "G", "R"};          // alphabetical, as OpenEXR stores them
    img.channel_names = names;
    unsigned char* ptrs[3]; int ptype[3], rtype[3];
    for (int c = 0; c < 3; c++) {
      const std::vector<float>& src = ch[2 - c];     // names[c]: B <- ch[2], G <- ch[1], R <- ch[0]
      if (half) { hc[c].resize(src.size()); for (size_t i = 0; i < src.size(); i++) hc[c][i] = float_to_half(src[i]); ptrs[c] = (unsigned char*)hc[c].data(); }
      else ptrs[c] = (unsigned char*)src.data();
      ptype[c] = half ? TINYEXR_PIXELTYPE_HALF : TINYEXR_PIXELTYPE_FLOAT; rtype[c] = ptype[c];
    }
    img.images = ptrs; img.pixel_types = ptype; img.requested_pixel_types = rtype;
    img.width = w; img.height = h;
    const char* err = nullptr;
    if (SaveMultiChannelEXRToFile(&img, argv[2], &err) != 0) { std::fprintf(stderr, "save: %s\n", err ? err : "?"); return 6; }
    return 0;
  }
  std::fprintf(stderr, "usage: exr_ref load in.exr out.bin | exr_ref save out.exr w h half|float seed\n");
  return 1;
}
