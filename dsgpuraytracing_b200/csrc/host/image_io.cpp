// image_io.cpp -- OpenEXR / PFM readers (see image_io.h).  Scope is frozen (scan-line NONE / RLE / ZIPS / ZIP / PIZ): the -e
// option is off the render hot path.  Written against the OpenEXR file layout document: magic,
// version word, attribute list, line-offset table, chunks of 1 (NONE, RLE, ZIPS), 16 (ZIP) or 32 (PIZ) scan lines; a ZIP / RLE
// chunk is zlib / run-length data of the byte-planar, delta-predicted pixel bytes, a PIZ chunk is Huffman-coded wavelet data.
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
  // p and n can both come from untrusted 64-bit file fields: compare without forming p + n (which could wrap)
  bool need(size_t n) { if (p > b.size() || n > b.size() - p) { ok = false; return false; } return true; }
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


// ---- PIZ: 16-bit range compaction (bitmap + lookup table), 2-D Haar-like wavelet per channel, canonical Huffman coding
// with a run-length symbol.  The algorithms are those of OpenEXR's ImfPizCompressor / ImfHuf / ImfWav (Industrial Light & Magic,
// BSD-3-Clause), which the file format fixes bit for bit; the reference reads such files through its vendored tinyexr
// (CMU462/include/CMU462/tinyexr.h, BSD-3-Clause, itself derived from the ILM sources).  This is a re-implementation of
// those published algorithms for this reader, checked against tinyexr's output (tests/golden/exr), not original design.
constexpr int kHufEncBits = 16, kHufDecBits = 14;
constexpr int kHufEncSize = (1 << kHufEncBits) + 1, kHufDecSize = 1 << kHufDecBits, kHufDecMask = kHufDecSize - 1;
constexpr int kShortZeroRun = 59, kLongZeroRun = 63, kShortestLongRun = 2 + kLongZeroRun - kShortZeroRun;

struct BitReader {
  const uint8_t* p; const uint8_t* end; uint64_t c = 0; int lc = 0; bool ok = true;
  uint32_t get(int n) {
    while (lc < n) { if (p >= end) { ok = false; return 0; } c = (c << 8) | *p++; lc += 8; }
    lc -= n;
    return (uint32_t)((c >> lc) & ((1ull << n) - 1));
  }
};

struct HufDec { uint32_t len = 0, lit = 0; std::vector<uint32_t> longs; };

bool huf_uncompress(const uint8_t* src, size_t n_src, std::vector<uint16_t>& out, size_t n_out) {
  out.assign(n_out, 0);
  if (n_src == 0) return n_out == 0;
  if (n_src < 20) return false;
  auto rd32 = [&](size_t o) { uint32_t v; std::memcpy(&v, src + o, 4); return v; };
  uint32_t im = rd32(0); const uint32_t iM = rd32(4); const uint32_t n_bits = rd32(12);
  if (im >= (uint32_t)kHufEncSize || iM >= (uint32_t)kHufEncSize) return false;
  // code lengths, 6 bits each, with short / long runs of zero lengths
  std::vector<uint64_t> code((size_t)kHufEncSize, 0);
  BitReader br{src + 20, src + n_src};
  for (uint32_t i = im; i <= iM; i++) {
    const uint32_t l = br.get(6);
    if (!br.ok) return false;
    code[i] = l;
    if (l == (uint32_t)kLongZeroRun || l >= (uint32_t)kShortZeroRun) {
      uint32_t run = l == (uint32_t)kLongZeroRun ? br.get(8) + (uint32_t)kShortestLongRun : l - (uint32_t)kShortZeroRun + 2;
      if (!br.ok || i + run > iM + 1) return false;
      while (run--) code[i++] = 0;
      i--;
    }
  }
  const uint8_t* data = br.p;                             // the table ends on a byte boundary
  if ((uint64_t)n_bits > 8ull * (uint64_t)(src + n_src - data)) return false;
  // canonical codes from the lengths: code = length | (value << 6)
  {
    uint64_t n[59] = {0};
    for (int i = 0; i < kHufEncSize; i++) { if (code[i] > 58) return false; n[code[i]]++; }
    uint64_t c = 0;
    for (int i = 58; i > 0; --i) { const uint64_t nc = (c + n[i]) >> 1; n[i] = c; c = nc; }
    for (int i = 0; i < kHufEncSize; i++) { const uint64_t l = code[i]; if (l > 0) code[i] = l | (n[l]++ << 6); }
  }
  // decoding table: 14-bit primary index; longer codes are searched in a per-entry list
  std::vector<HufDec> dec((size_t)kHufDecSize);
  for (uint32_t i = im; i <= iM; i++) {
    const uint64_t c = code[i] >> 6; const int l = (int)(code[i] & 63);
    if (c >> l) return false;
    if (l > kHufDecBits) {
      HufDec& pl = dec[(size_t)(c >> (l - kHufDecBits))];
      if (pl.len) return false;
      pl.longs.push_back(i);
    } else if (l) {
      HufDec* pl = &dec[(size_t)(c << (kHufDecBits - l))];
      for (uint64_t k = 1ull << (kHufDecBits - l); k > 0; k--, pl++) {
        if (pl->len || !pl->longs.empty()) return false;
        pl->len = (uint32_t)l; pl->lit = i;
      }
    }
  }
  // decode
  const uint8_t* in = data; const uint8_t* ie = data + (n_bits + 7) / 8;
  uint64_t c = 0; int lc = 0; size_t o = 0;
  auto emit = [&](uint32_t sym) -> bool {
    if (sym == iM) {                                      // run-length symbol: repeat the previous value `count` times
      if (lc < 8) { if (in >= ie) return false; c = (c << 8) | *in++; lc += 8; }
      lc -= 8;
      uint32_t cs = (uint32_t)((c >> lc) & 0xff);
      if (o + cs > n_out || o == 0) return false;
      const uint16_t s = out[o - 1];
      while (cs-- > 0) out[o++] = s;
      return true;
    }
    if (o >= n_out) return false;
    out[o++] = (uint16_t)sym;
    return true;
  };
  while (in < ie) {
    c = (c << 8) | *in++; lc += 8;
    while (lc >= kHufDecBits) {
      const HufDec& pl = dec[(size_t)((c >> (lc - kHufDecBits)) & (uint64_t)kHufDecMask)];
      if (pl.len) { lc -= (int)pl.len; if (!emit(pl.lit)) return false; }
      else {
        if (pl.longs.empty()) return false;
        size_t j = 0;
        for (; j < pl.longs.size(); j++) {
          const int l = (int)(code[pl.longs[j]] & 63);
          while (lc < l && in < ie) { c = (c << 8) | *in++; lc += 8; }
          if (lc >= l && (code[pl.longs[j]] >> 6) == ((c >> (lc - l)) & ((1ull << l) - 1))) {
            lc -= l; if (!emit(pl.longs[j])) return false;
            break;
          }
        }
        if (j == pl.longs.size()) return false;
      }
    }
  }
  const int pad = (8 - (int)n_bits) & 7;                  // the last byte is padded with zero bits
  c >>= pad; lc -= pad;
  while (lc > 0) {
    const HufDec& pl = dec[(size_t)((c << (kHufDecBits - lc)) & (uint64_t)kHufDecMask)];
    if (!pl.len || (int)pl.len > lc) return false;
    lc -= (int)pl.len; if (!emit(pl.lit)) return false;
  }
  return o == n_out;
}

// The PIZ wavelet is fixed by the OpenEXR file format (ILM's ImfWav.cpp, BSD-3-Clause; the reference vendors a copy inside
// CMU462/include/CMU462/tinyexr.h:7213-7470).  Each step undoes one "average / difference" pair (lo, hi) -> (a, b): the
// 14-bit form when the range-compacted data fits in 14 bits, otherwise the modulo-2^16 form.
inline void unpair14(uint16_t lo, uint16_t hi, uint16_t& a, uint16_t& b) {
  const int d = (int16_t)hi, s = (int16_t)lo + (d & 1) + (d >> 1);
  a = (uint16_t)(int16_t)s; b = (uint16_t)(int16_t)(s - d);
}
inline void unpair16(uint16_t lo, uint16_t hi, uint16_t& a, uint16_t& b) {
  const int y = ((int)lo - ((int)hi >> 1)) & 0xffff;
  b = (uint16_t)y; a = (uint16_t)(((int)hi + y - 0x8000) & 0xffff);
}
// inverse 2-D wavelet, in place, on an nx x ny grid of 16-bit values (element (x, y) at v[x * sx + y * sy]); max_value =
// largest value after range compaction.  Levels run from the coarsest cell size (largest power of two <= min(nx, ny))
// down to 2; a level first undoes the 2x2 cells, then the left-over column / row when the size has that bit set.
void wav2_decode(uint16_t* v, int nx, int sx, int ny, int sy, uint16_t max_value) {
  const bool small = max_value < (1 << 14);
  auto at = [&](int x, int y) -> uint16_t& { return v[(ptrdiff_t)x * sx + (ptrdiff_t)y * sy]; };
  auto unpair = [&](uint16_t lo, uint16_t hi, uint16_t& a, uint16_t& b) { if (small) unpair14(lo, hi, a, b); else unpair16(lo, hi, a, b); };
  int cell = 1;
  while (cell * 2 <= (nx < ny ? nx : ny)) cell *= 2;
  for (int half = cell / 2; half >= 1; cell = half, half /= 2) {
    const int x_end = ((nx - cell) / cell + 1) * cell, y_end = ((ny - cell) / cell + 1) * cell;   // extent covered by whole cells
    for (int y = 0; y < y_end; y += cell) {
      for (int x = 0; x < x_end; x += cell) {
        uint16_t t00, t01, t10, t11;
        unpair(at(x, y), at(x, y + half), t00, t10); unpair(at(x + half, y), at(x + half, y + half), t01, t11);   // columns
        unpair(t00, t01, at(x, y), at(x + half, y)); unpair(t10, t11, at(x, y + half), at(x + half, y + half));     // rows
      }
      if (nx & half) { uint16_t t; unpair(at(x_end, y), at(x_end, y + half), t, at(x_end, y + half)); at(x_end, y) = t; }
    }
    if (ny & half)
      for (int x = 0; x < x_end; x += cell) { uint16_t t; unpair(at(x, y_end), at(x + half, y_end), t, at(x + half, y_end)); at(x, y_end) = t; }
  }
}

// one PIZ chunk -> the same byte layout an uncompressed chunk has (per scan line, per channel, a row of samples)
bool piz_decode(const uint8_t* src, size_t n_src, const std::vector<Channel>& chans, int64_t W, int64_t lines, std::vector<uint8_t>& raw) {
  if (n_src < 4) return false;
  uint16_t min_nz, max_nz; std::memcpy(&min_nz, src, 2); std::memcpy(&max_nz, src + 2, 2);
  size_t p = 4;
  std::vector<uint8_t> bitmap(8192, 0);
  if (max_nz >= 8192) return false;
  if (min_nz <= max_nz) {
    const size_t nb = (size_t)max_nz - min_nz + 1;
    if (p + nb > n_src) return false;
    std::memcpy(&bitmap[min_nz], src + p, nb); p += nb;
  }
  std::vector<uint16_t> lut(65536, 0);
  int k = 0;
  for (int i = 0; i < 65536; i++) if (i == 0 || (bitmap[(size_t)i >> 3] & (1 << (i & 7)))) lut[(size_t)k++] = (uint16_t)i;
  const uint16_t max_value = (uint16_t)(k - 1);
  if (p + 4 > n_src) return false;
  int32_t length; std::memcpy(&length, src + p, 4); p += 4;
  if (length < 0 || p + (size_t)length > n_src) return false;
  size_t total = 0; std::vector<size_t> start(chans.size()), words(chans.size());
  for (size_t c = 0; c < chans.size(); c++) { words[c] = chans[c].type == 1 ? 1 : 2; start[c] = total; total += (size_t)W * (size_t)lines * words[c]; }
  std::vector<uint16_t> tmp;
  if (!huf_uncompress(src + p, (size_t)length, tmp, total)) return false;
  for (size_t c = 0; c < chans.size(); c++)
    for (size_t j = 0; j < words[c]; j++)
      wav2_decode(&tmp[start[c] + j], (int)W, (int)words[c], (int)lines, (int)((size_t)W * words[c]), max_value);
  for (uint16_t& v : tmp) v = lut[v];
  raw.resize(total * 2);
  uint8_t* o = raw.data(); std::vector<size_t> cur = start;
  for (int64_t y = 0; y < lines; y++)
    for (size_t c = 0; c < chans.size(); c++) {
      const size_t n = (size_t)W * words[c];
      std::memcpy(o, &tmp[cur[c]], n * 2); o += n * 2; cur[c] += n;
    }
  return true;
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
    case 4: lines_per_chunk = 32; break;
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
    if (offsets[(size_t)i] >= (uint64_t)bytes.size()) { err = "corrupt OpenEXR chunk offset"; return false; }
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
    } else if (compression == 4) {
      if (!piz_decode(src, (size_t)sz, chans, W, lines, raw) || raw.size() != expect) { err = "corrupt PIZ data in OpenEXR chunk"; return false; }
      px = raw.data();
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
