// tests/cpu_walk/cpu_walk.cu -- TEST INFRASTRUCTURE (a checker; the product never links or loads it).
//
// Compiles the product's own __host__ __device__ traversal / shading functions (traverse.cuh, shade.cuh,
// rng.cuh) and its host flattening code (wide_bvh.cpp, bvh2_sah.cpp) for the HOST, and walks the same compressed
// wide BVH sequentially on the CPU.  Purpose: cross-check the device node/primitive fetch
// counters with a CPU walk of the same flattened BVH over the same rays, and validate the collapse, the
// quantisation and the estimator logic on a box without a GPU.  Results are float like the GPU's, and differ
// from it only by libm-vs-intrinsic rounding (sincospi, rsqrt) and FMA contraction.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../dsgpuraytracing_b200/csrc/layout.h"
#include "../../dsgpuraytracing_b200/csrc/rng.cuh"
#include "../../dsgpuraytracing_b200/csrc/shade.cuh"
#include "../../dsgpuraytracing_b200/csrc/traverse.cuh"
#include "../../dsgpuraytracing_b200/csrc/wide_bvh.h"

using namespace dsrt;

struct Walk {
  WideBVH wide;
  std::vector<PrimRecord> recs; std::vector<ShadeRecord> shd; std::vector<PrimRecord64> r64; std::vector<Light> lights;
  std::vector<Bsdf> bsdfs;
  int n_light_samples = 0;
  int env_w = 0, env_h = 0; std::vector<float> env_rgb, env_tp, env_t, env_pgt;
  double scene_diag = 1;
  float bsphere[4] = {0, 0, 0, 0};
  Camera cam;
  std::string err;
};

static Accel accel_of(const Walk* w, bool parity) {
  Accel A; A.nodes = (const uint4*)w->wide.nodes.data(); A.prims = (const float4*)w->recs.data();
  A.prims64 = (const double*)w->r64.data(); A.pad = parity ? (float)(1e-5 * w->scene_diag) : 0.f; A.one_bits = 0x3f800000u;
  A.bcx = w->bsphere[0]; A.bcy = w->bsphere[1]; A.bcz = w->bsphere[2]; A.brad = w->bsphere[3];
  return A;
}

extern "C" {

Walk* cw_create(const dsrt_scene* s, const dsrt_bvh2* b, int ns_area_light, int env_w, int env_h, const float* env_rgb) {
  Walk* w = new Walk();
  if (env_w > 0) { w->env_w = env_w; w->env_h = env_h; w->env_rgb.assign(env_rgb, env_rgb + (size_t)env_w * env_h * 3); build_env_tables(env_w, env_h, w->env_rgb.data(), w->env_tp, w->env_t, w->env_pgt); }
  std::vector<Box3> pbox; primitive_boxes(s, pbox);
  const std::vector<EndPlane> ends = light_end_planes(s->n_lights, s->light_type, s->light_param);
  if (build_wide_bvh(*b, pbox, s->n_prims, w->wide, w->err, 1.0, std::getenv("CW_NO_LIGHT_GRID") ? nullptr : &ends, std::getenv("CW_REGROUP") != nullptr)) { fprintf(stderr, "cw_create: %s\n", w->err.c_str()); delete w; return nullptr; }
  if (!std::getenv("CW_NO_FLAT_SLOTS")) mark_flat_slots(*s, w->wide);
  flatten_records(*s, w->wide, w->recs, w->shd); flatten_records64(*s, w->wide, w->r64);
  w->n_light_samples = flatten_lights(s->n_lights, s->light_type, s->light_param, ns_area_light, env_w > 0, w->lights);
  w->bsdfs.resize(s->n_bsdf);
  for (int i = 0; i < s->n_bsdf; i++) { Bsdf& q = w->bsdfs[i]; const float* p = s->bsdf_param + 8 * i; for (int k = 0; k < 3; k++) { q.a[k] = p[k]; q.b[k] = p[3 + k]; } q.ior = p[6]; q.type = s->bsdf_type[i]; }
  Box3 all; all.reset(); for (auto& p : pbox) all.grow(p);
  double dg = 0; for (int k = 0; k < 3; k++) { double e = s->n_prims ? all.hi[k] - all.lo[k] : 0, m = s->n_prims ? fmax(fabs(all.lo[k]), fabs(all.hi[k])) : 0; dg += (e + m) * (e + m); }
  w->scene_diag = sqrt(dg) + 1.0;
  bounding_sphere(all, s->n_prims, w->bsphere);
  return w;
}
void cw_destroy(Walk* w) { delete w; }
int cw_node_bytes() { return (int)sizeof(WideNode); }
void cw_nodes(Walk* w, void* out) { std::memcpy(out, w->wide.nodes.data(), w->wide.nodes.size() * sizeof(WideNode)); }   // raw nodes (sizeof(WideNode) bytes each, cw_node_bytes)
void cw_info(Walk* w, int64_t* n_nodes, int32_t* depth) { *n_nodes = (int64_t)w->wide.nodes.size(); *depth = w->wide.max_depth; }
void cw_slot_prim(Walk* w, int32_t* out) { memcpy(out, w->wide.slot_prim.data(), 4 * w->wide.slot_prim.size()); }

void cw_set_camera(Walk* w, const double* pos, const double* c2w, int W, int H, double dist) {
  Camera& c = w->cam;
  for (int k = 0; k < 3; k++) { c.pos[k] = (float)pos[k]; c.pos64[k] = pos[k]; }
  for (int k = 0; k < 9; k++) { c.c2w[k] = (float)c2w[k]; c.c2w64[k] = c2w[k]; }
  c.width = W; c.height = H; c.W64 = W; c.H64 = H; c.dist64 = dist;
  c.w_over_dist = (float)(W / dist); c.h_over_dist = (float)(H / dist);
}

// closest / any hit for float rays through the production code path; counters = nodes, prims fetched
void cw_trace(Walk* w, int any, int64_t n, const float* o, const float* d, const float* tmax, int32_t* prim_id, float* t, uint64_t* counters) {
  const Accel A = accel_of(w, false);
  std::vector<uint2> stack(kStackEntries);
  uint64_t cn = 0, cp = 0;
  for (int64_t i = 0; i < n; i++) {
    TraceRay r; r.ox = o[3 * i]; r.oy = o[3 * i + 1]; r.oz = o[3 * i + 2]; r.dx = d[3 * i]; r.dy = d[3 * i + 1]; r.dz = d[3 * i + 2];
    r.tmax = tmax ? tmax[i] : kInfF; r.src_slot = -1;
    TraceHit h; TraceCounters c; c.nodes = c.prims = 0;
    if (any) trace_ray<true, false, true>(A, r, nullptr, stack.data(), 1, h, nullptr, &c);
    else trace_ray<false, false, true>(A, r, nullptr, stack.data(), 1, h, nullptr, &c);
    cn += c.nodes; cp += c.prims;
    if (any) prim_id[i] = h.slot >= 0 ? 1 : 0;
    else { prim_id[i] = h.slot >= 0 ? w->wide.slot_prim[h.slot] : -1; if (t) t[i] = h.slot >= 0 ? h.t : kInfF; }
  }
  if (counters) { counters[0] = cn; counters[1] = cp; }
}

// brute force over all slots with the production primitive tests (no BVH): separates culling from intersection bugs
// the PRODUCT's Philox4x32-10 (rng.cuh, host build) on a raw counter / key: Random123 known-answer vectors
void cw_philox_raw(const uint32_t* ctr4, uint32_t k0, uint32_t k1, uint32_t* out4) {
  const uint4 r = philox4x32_10(make_uint4(ctr4[0], ctr4[1], ctr4[2], ctr4[3]), k0, k1);
  out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}

// the node-word helpers of traverse.cuh on raw inputs (tests/test_host.py checks them against a plain restatement)
// dir: any direction with the wanted signs; out_nodes: the node indices next_child hands out for the group (child_base, hits & inner | inner >> 3), in order
int cw_open_order(const float* dir, uint32_t child_base, uint32_t hits, uint32_t inner, int ordered, uint32_t* out_nodes) {
  TraceRay r; r.ox = r.oy = r.oz = 0.f; r.dx = dir[0]; r.dy = dir[1]; r.dz = dir[2]; r.tmax = kInfF; r.src_slot = -1;
  const NodeFrame fr = make_frame(r);
  const uint32_t open = hits & inner;
  uint2 g = make_uint2(child_base, (ordered ? order_children(open, fr) : open) | (inner >> 3));
  int n = 0;
  while (g.y & kHitBits) { bool more; out_nodes[n++] = ordered ? next_child<true>(g, fr, more) : next_child<false>(g, fr, more); }
  return n;
}
uint32_t cw_drop_source(uint32_t prim_mask, uint32_t prim_base, uint32_t valid, int src) { return drop_source(prim_mask, prim_base, valid, src); }
int cw_prim_slot(uint32_t prim_base, uint32_t valid, uint32_t bit) { return prim_slot(prim_base, valid, bit); }

void cw_trace_brute(Walk* w, int64_t n, const float* o, const float* d, int32_t* prim_id, float* t) {
  const Accel A = accel_of(w, false);
  for (int64_t i = 0; i < n; i++) {
    TraceRay r; r.ox = o[3 * i]; r.oy = o[3 * i + 1]; r.oz = o[3 * i + 2]; r.dx = d[3 * i]; r.dy = d[3 * i + 1]; r.dz = d[3 * i + 2];
    r.tmax = kInfF; r.src_slot = -1;
    const WatertightRay wr = make_watertight(r);
    float best = kInfF; int bs = -1;
    for (size_t sl = 0; sl < w->wide.slot_prim.size(); sl++) {
      const float4* pp = A.prims + sl * 3;
      float tt, u, v; bool h;
      if (pp[1].w != 0.0f) h = hit_triangle(r, wr, pp[0], pp[1], pp[2], best, tt, u, v);
      else h = hit_sphere(r, pp[0], pp[1], false, false, best, tt);
      if (h) { best = tt; bs = (int)sl; }
    }
    prim_id[i] = bs >= 0 ? w->wide.slot_prim[bs] : -1; t[i] = best;
  }
}

static void gen_ray64(const Camera& c, double x, double y, Ray64& r) {
  const double sp[3] = {-(x - 0.5) * c.W64 / c.dist64, -(y - 0.5) * c.H64 / c.dist64, 1.0};
  double wv[3], dir[3];
  for (int k = 0; k < 3; k++) {
    wv[k] = (sp[0] * c.c2w64[k] + sp[1] * c.c2w64[3 + k]) + sp[2] * c.c2w64[6 + k];
    dir[k] = ((-sp[0]) * c.c2w64[k] + (-sp[1]) * c.c2w64[3 + k]) + (-sp[2]) * c.c2w64[6 + k];
  }
  const double inv = 1. / sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  r.ox = wv[0] + c.pos64[0]; r.oy = wv[1] + c.pos64[1]; r.oz = wv[2] + c.pos64[2];
  r.dx = dir[0] * inv; r.dy = dir[1] * inv; r.dz = dir[2] * inv;
}

// pixel-centre primary hits: mode 1 = parity path (fp64 leaves), mode 0 = production float path
void cw_primary_hits(Walk* w, int mode, int32_t* prim_id, double* t) {
  const int W = w->cam.width, H = w->cam.height;
  const Accel A = accel_of(w, mode == 1);
  std::vector<uint2> stack(kStackEntries);
  for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
    const size_t i = (size_t)y * W + x;
    TraceHit h; double t64 = 0;
    if (mode == 1) {
      Ray64 r64; gen_ray64(w->cam, (x + 0.5) / W, (y + 0.5) / H, r64);
      TraceRay r; r.ox = (float)r64.ox; r.oy = (float)r64.oy; r.oz = (float)r64.oz; r.dx = (float)r64.dx; r.dy = (float)r64.dy; r.dz = (float)r64.dz;
      r.tmax = kInfF; r.src_slot = -1;
      trace_ray<false, true, false>(A, r, &r64, stack.data(), 1, h, &t64, nullptr);
    } else {
      V3 o, d; generate_ray(w->cam, ((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H, &o, &d);
      TraceRay r; r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z; r.tmax = kInfF; r.src_slot = -1;
      trace_ray<false, false, false>(A, r, nullptr, stack.data(), 1, h, nullptr, nullptr);
      t64 = h.t;
    }
    prim_id[i] = h.slot >= 0 ? w->wide.slot_prim[h.slot] : -1;
    if (t) t[i] = h.slot >= 0 ? t64 : (double)kInfF;
  }
}

struct ImmediateSink {
  const Walk* w; Accel A; uint2* stack; float* px; uint64_t* n_shadow; uint64_t* cn; uint64_t* cp;
  void shadow_ray(float4 a, float4 b, float4 c) {
    TraceRay r; r.ox = a.x; r.oy = a.y; r.oz = a.z; r.tmax = a.w; r.dx = b.x; r.dy = b.y; r.dz = b.z; r.src_slot = hd_f2i(b.w);
    TraceHit h; TraceCounters k; k.nodes = k.prims = 0;
    trace_ray<true, false, true>(A, r, nullptr, stack, 1, h, nullptr, &k);
    (*n_shadow)++; *cn += k.nodes; *cp += k.prims;
    if (h.slot < 0) { px[0] += c.x; px[1] += c.y; px[2] += c.z; }
  }
  void shadow(int, float4 a, float4 b, float4 c) { shadow_ray(a, b, c); }
};

// sequential emulation of the wavefront (same per-vertex code as k_shade); counters = camera, extend, shadow, nodes, prims
void cw_render(Walk* w, int spp_begin, int spp_count, int spp_stride, int spp_total, int max_depth, uint32_t seed, float* rgb, uint64_t* counters) {
  const int W = w->cam.width, H = w->cam.height;
  const Accel A = accel_of(w, false);
  std::vector<uint2> stack(kStackEntries);
  SceneDev sc; sc.bsdf = w->bsdfs.data(); sc.lights = w->lights.data(); sc.shade = (const float4*)w->shd.data();
  sc.n_lights = (int)w->lights.size(); sc.n_light_samples = w->n_light_samples;
  sc.env.rgb = w->env_rgb.data(); sc.env.pThetaPhi = w->env_tp.data(); sc.env.pTheta = w->env_t.data(); sc.env.pPhiGivenTheta = w->env_pgt.data();
  sc.env.w = w->env_w; sc.env.h = w->env_h;
  uint64_t camera = 0, extend = 0, shadow = 0, cn = 0, cp = 0;
  for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
    float px[3] = {0, 0, 0};
    const uint32_t pix = (uint32_t)(y * W + x);
    for (int si = 0; si < spp_count; si++) {
      const uint32_t smp = (uint32_t)(spp_begin + si * spp_stride);
      const float4 u = rng_block(seed, pix, smp, 0u, kBlockCamera);
      V3 o, d; generate_ray(w->cam, ((float)x + u.x) / (float)W, ((float)y + u.y) / (float)H, &o, &d);
      PathIn in; in.ray_o = make_float4(o.x, o.y, o.z, kInfF); in.ray_d = make_float4(d.x, d.y, d.z, hd_i2f(-1));
      in.thr = make_float4(1.f, 1.f, 1.f, hd_i2f(0 | (1 << 8))); in.pix = pix; in.smp = smp;
      camera++;
      for (int depth = 0; depth <= max_depth; depth++) {
        TraceRay r; r.ox = in.ray_o.x; r.oy = in.ray_o.y; r.oz = in.ray_o.z; r.tmax = in.ray_o.w;
        r.dx = in.ray_d.x; r.dy = in.ray_d.y; r.dz = in.ray_d.z; r.src_slot = hd_f2i(in.ray_d.w);
        TraceHit h; TraceCounters k; k.nodes = k.prims = 0;
        trace_ray<false, false, true>(A, r, nullptr, stack.data(), 1, h, nullptr, &k);
        extend++; cn += k.nodes; cp += k.prims;
        if (h.slot < 0) {
          if (sc.env.w > 0 && ((hd_f2i(in.thr.w) >> 8) & 1)) {
            const V3 e = v3(in.thr.x, in.thr.y, in.thr.z) * env_sample_dir(sc.env, v3(in.ray_d.x, in.ray_d.y, in.ray_d.z));
            px[0] += e.x; px[1] += e.y; px[2] += e.z;
          }
          break;
        }
        in.hit = make_float4(h.t, h.u, h.v, hd_i2f(h.slot));
        PathOut out; ImmediateSink sink{w, A, stack.data(), px, &shadow, &cn, &cp};
        shade_path(in, A.prims, sc, seed, max_depth, depth, out, sink);
        if (out.has_emission) { px[0] += out.emission.x; px[1] += out.emission.y; px[2] += out.emission.z; }
        if (!out.cont) break;
        in.ray_o = out.new_o; in.ray_d = out.new_d; in.thr = out.new_thr;
      }
    }
    const float inv = 1.0f / (float)spp_total;
    rgb[3 * (size_t)pix] = px[0] * inv; rgb[3 * (size_t)pix + 1] = px[1] * inv; rgb[3 * (size_t)pix + 2] = px[2] * inv;
  }
  if (counters) { counters[0] = camera; counters[1] = extend; counters[2] = shadow; counters[3] = cn; counters[4] = cp; }
}

}  // extern "C"
