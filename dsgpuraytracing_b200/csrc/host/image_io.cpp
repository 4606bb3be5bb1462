// image_io.cpp -- OpenEXR / PFM readers (see image_io.h).  Written against the OpenEXR file layout document: magic,
// version word, attribute list, line-offset table, chunks of 1 (NONE, RLE, ZIPS) or 16 (ZIP) scan lines; a compressed
// chunk is zlib / run-length data of the byte-planar, delta-predicted pixel bytes.
#include "image_io.h"

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace dsrt_host {
namespace {

struct Reader {
  const std::vector<uint8_t>& b; size_t p = 0; bool ok = true;
  explicit Reader(const std::vector<uint8_t>& bytes) : b(bytes) {}
  bool need(size_t n) { if (p + n > b.size()) { ok = false; return false; } return true; }
  uint8_t u8() { if (!need(1)) return 0; return b[p++]; }
  uint32_t u32() { if (!need(4)) return 0; uint32_t v; std::memcpy(&v, &b[p], 4); p += 4; return v; }
  int32_t i32() { return (int32_t)u32(); }
  uint64_t u64() { if (!need(8)) return 0; uint64_t v; std::memcpy(&v, &b[p], 8); p += 8; return v; }
  std::string str() { std::string s; while (p < b.size() && b[p]) s.push_back((char)b[p++]); if (p >= b.size()) ok = false; else p++; return s; }
};

float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16; uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu, x;
  if (e == 0) {
    if (m == 0) x = sign;
    else { e = 113; while (!(m & 0x400u)) { m <<= 1; e--; } x = sign | (e << 23) | ((m & 0x3ffu) << 13); }   // subnormal half
  } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
  else x = sign | ((e + 112u) << 23) | (m << 13);
  float f; std::memcpy(&f, &x, 4); return f;
}

struct Channel { std::string name; int type = 0, xs = 1, ys = 1; };

// undo the predictor and the even/odd byte split applied before ZIP / RLE compression
void unpredict(std::vector<uint8_t>& t, std::vector<uint8_t>& out) {
  for (size_t i = 1; i < t.size(); i++) t[i] = (uint8_t)(t[i - 1] + t[i] - 128);
  out.resize(t.size());
  const size_t half = (t.size() + 1) / 2;
  for (size_t i = 0; i < t.size(); i++) out[i] = (i & 1) ? t[half + i / 2] : t[i / 2];
}

bool rle_decode(const uint8_t* src, size_t n, std::vector<uint8_t>& out, size_t expect) {
  out.clear(); out.reserve(expect);
  size_t i = 0;
  while (i < n) {
    const int c = (int8_t)src[i++];
    if (c < 0) { const size_t k = (size_t)(-c); if (i + k > n) return false; out.insert(out.end(), src + i, src + i + k); i += k; }
    else { if (i >= n) return false; out.insert(out.end(), (size_t)c + 1, src[i]); i++; }
    if (out.size() > expect) return false;
  }
  return out.size() == expect;
}

}  // namespace

bool load_exr(const std::string& path, HDRImageBuffer& img, std::string& err) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open " + path; return false; }
  std::vector<uint8_t> bytes;
  { uint8_t buf[1 << 16]; size_t n; while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) bytes.insert(bytes.end(), buf, buf + n); }
  std::fclose(f);
  Reader r(bytes);
  if (r.u32() != 20000630u) { err = "not an OpenEXR file: " + path; return false; }
  const uint32_t version = r.u32();
  if ((version & 0xffu) != 2u) { err = "unsupported OpenEXR version"; return false; }
  if (version & 0x1a00u) { err = "tiled, deep and multi-part OpenEXR files are not supported (scan-line images only)"; return false; }
  std::vector<Channel> chans; int compression = -1, line_order = 0; int32_t dw[4] = {0, 0, -1, -1}; bool have_dw = false;
  while (r.ok) {
    const std::string name = r.str();
    if (name.empty()) break;
    const std::string type = r.str();
    const int32_t size = r.i32();
    if (size < 0 || !r.need((size_t)size)) { err = "truncated OpenEXR header"; return false; }
    const size_t end = r.p + (size_t)size;
    if (name == "channels") {
      while (r.p < end) {
        Channel c; c.name = r.str();
        if (c.name.empty()) break;
        c.type = r.i32(); r.p += 4; c.xs = r.i32(); c.ys = r.i32();
        chans.push_back(c);
      }
    } else if (name == "compression") compression = r.u8();
    else if (name == "dataWindow") { for (int k = 0; k < 4; k++) dw[k] = r.i32(); have_dw = true; }
    else if (name == "lineOrder") line_order = r.u8();
    r.p = end;
  }
  if (!r.ok || chans.empty() || !have_dw || compression < 0) { err = "malformed OpenEXR header"; return false; }
  // Chunks carry their own y coordinate, so per the file-format document lineOrder only describes the order of the chunks
  // in the file.  The reference's tinyexr, however, mirrors the image vertically when lineOrder is DECREASING_Y
  // (tinyexr.h:9320-9334, 9434-9447: row = height - 1 - line); kept, so that the same file lights the scene the same way.
  const bool flip_rows = line_order == 1;
  const int64_t W = (int64_t)dw[2] - dw[0] + 1, H = (int64_t)dw[3] - dw[1] + 1;
  if (W <= 0 || H <= 0 || W > (1 << 20) || H > (1 << 20)) { err = "bad OpenEXR data window"; return false; }
  int lines_per_chunk;
  switch (compression) {
    case 0: case 1: case 2: lines_per_chunk = 1; break;
    case 3: lines_per_chunk = 16; break;
    case 4: err = "PIZ-compressed OpenEXR is not supported: re-save the map with ZIP, ZIPS, RLE or no compression"; return false;
    default: err = "unsupported OpenEXR compression " + std::to_string(compression); return false;
  }
  int ci[3] = {-1, -1, -1};
  size_t row_bytes = 0; std::vector<size_t> ch_off(chans.size());
  for (size_t c = 0; c < chans.size(); c++) {
    if (chans[c].xs != 1 || chans[c].ys != 1) { err = "sub-sampled OpenEXR channels are not supported"; return false; }
    if (chans[c].type < 0 || chans[c].type > 2) { err = "bad OpenEXR pixel type"; return false; }
    ch_off[c] = row_bytes; row_bytes += (size_t)W * (chans[c].type == 1 ? 2u : 4u);
    if (chans[c].name == "R") ci[0] = (int)c; else if (chans[c].name == "G") ci[1] = (int)c; else if (chans[c].name == "B") ci[2] = (int)c;
  }
  if (ci[0] < 0 || ci[1] < 0 || ci[2] < 0) { err = "OpenEXR file has no R, G, B channels"; return false; }
  const int64_t n_chunks = (H + lines_per_chunk - 1) / lines_per_chunk;
  std::vector<uint64_t> offsets((size_t)n_chunks);
  for (int64_t i = 0; i < n_chunks; i++) offsets[(size_t)i] = r.u64();
  if (!r.ok) { err = "truncated OpenEXR offset table"; return false; }
  img.resize((size_t)W, (size_t)H);
  std::vector<uint8_t> tmp, raw; std::vector<char> seen((size_t)H, 0);
  for (int64_t i = 0; i < n_chunks; i++) {
    Reader c(bytes); c.p = (size_t)offsets[(size_t)i];
    const int32_t y0 = c.i32(); const int32_t sz = c.i32();
    if (!c.ok || sz < 0 || !c.need((size_t)sz) || y0 < dw[1] || y0 > dw[3]) { err = "corrupt OpenEXR chunk"; return false; }
    const int64_t lines = std::min<int64_t>(lines_per_chunk, (int64_t)dw[3] - y0 + 1);
    const size_t expect = row_bytes * (size_t)lines;
    const uint8_t* src = &bytes[c.p]; const uint8_t* px = nullptr;
    if ((size_t)sz == expect || compression == 0) {           // stored raw (also when compression did not shrink the chunk)
      if ((size_t)sz != expect) { err = "corrupt OpenEXR chunk size"; return false; }
      px = src;
    } else if (compression == 1) {
      if (!rle_decode(src, (size_t)sz, tmp, expect)) { err = "corrupt RLE data in OpenEXR chunk"; return false; }
      unpredict(tmp, raw); px = raw.data();
    } else {
      tmp.resize(expect); uLongf dst_len = (uLongf)expect;
      if (uncompress(tmp.data(), &dst_len, src, (uLong)sz) != Z_OK || dst_len != expect) { err = "corrupt zlib data in OpenEXR chunk"; return false; }
      unpredict(tmp, raw); px = raw.data();
    }
    for (int64_t l = 0; l < lines; l++) {
      size_t row = (size_t)(y0 - dw[1] + l);
      if (flip_rows) row = (size_t)H - 1 - row;
      seen[row] = 1;
      for (int k = 0; k < 3; k++) {
        const Channel& ch = chans[(size_t)ci[k]];
        const uint8_t* q = px + (size_t)l * row_bytes + ch_off[(size_t)ci[k]];
        float* dst = &img.data[row * (size_t)W * 3 + (size_t)k];
        for (int64_t x = 0; x < W; x++) {
          float v;
          if (ch.type == 1) { uint16_t hv; std::memcpy(&hv, q + 2 * x, 2); v = half_to_float(hv); }
          else if (ch.type == 2) std::memcpy(&v, q + 4 * x, 4);
          else { uint32_t u; std::memcpy(&u, q + 4 * x, 4); v = (float)u; }
          dst[3 * x] = v;
        }
      }
    }
  }
  for (char s : seen) if (!s) { err = "OpenEXR file is missing scan lines"; return false; }
  return true;
}

bool load_pfm(const std::string& path, HDRImageBuffer& img, std::string& err) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open " + path; return false; }
  char magic[3] = {0, 0, 0}; int w = 0, h = 0; float scale = 0;
  if (std::fscanf(f, "%2s %d %d %f", magic, &w, &h, &scale) != 4 || std::string(magic) != "PF" || w <= 0 || h <= 0) { std::fclose(f); err = "not a colour .pfm: " + path; return false; }
  std::fgetc(f);
  if (scale > 0) { std::fclose(f); err = "big-endian .pfm is not supported"; return false; }
  img.resize((size_t)w, (size_t)h);
  for (int y = h - 1; y >= 0; y--)      // file is bottom-up; the environment map wants row 0 = +y pole
    if (std::fread(&img.data[(size_t)y * w * 3], sizeof(float), (size_t)w * 3, f) != (size_t)w * 3) { std::fclose(f); err = "truncated .pfm"; return false; }
  std::fclose(f);
  return true;
}

bool load_envmap(const std::string& path, HDRImageBuffer& img, std::string& err) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open " + path; return false; }
  uint8_t m[4] = {0, 0, 0, 0}; const size_t n = std::fread(m, 1, 4, f); std::fclose(f);
  if (n == 4 && m[0] == 0x76 && m[1] == 0x2f && m[2] == 0x31 && m[3] == 0x01) return load_exr(path, img, err);
  if (n >= 2 && m[0] == 'P' && m[1] == 'F') return load_pfm(path, img, err);
  err = "unknown environment map format (OpenEXR or .pfm expected): " + path;
  return false;
}

}  // namespace dsrt_host
