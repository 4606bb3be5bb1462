// traverse.cuh -- per-lane traversal of the compressed 8-wide BVH (layout.h) for sm_100a.
//
// Replaces BVHAccel::intersect closest / any (reference src/bvh.cpp:227-363) and Triangle/Sphere::intersect
// (src/static_scene/triangle.cpp:25-104, sphere.cpp:10-77); the reference GPU fork's counterpart is
// cuda_src/traversal.cu:3-222 + intersect.cu (pointer-linked binary tree, 64-entry local-memory stack).
//
// Node fetch = five 128-bit loads (80 B), primitive fetch = three 128-bit loads (48 B), both through the
// read-only path.  The traversal stack holds node GROUPS (child_base, pending-children mask): one entry per
// tree level, kept in shared memory (entry-major, so a warp's accesses are conflict free).
//
// Production mode: float slabs + watertight ray/triangle test (Woop, Benthin, Wald 2013: shear to ray space,
// antisymmetric edge functions, double-precision fallback on exact zeros).
// PARITY mode (template flag): identical traversal, but boxes are padded conservatively and the leaf tests run
// in fp64 with the reference's exact operation order and no FMA contraction, so primary-hit ids are bit-exact
// against the reference's double-precision Moller-Trumbore (gate 1 of SURVEY.md 8d).
#pragma once
#include <stdint.h>

#include "hd.h"
#include "layout.h"

namespace dsrt {

constexpr int kStackEntries = 32;    // upper bound on the levels of wide nodes (checked at dsrt_build_accel); the
                                     // kernels get exactly wide.max_depth entries of shared memory per lane


struct TraceRay {
  float ox, oy, oz, dx, dy, dz;
  float tmax;
  int src_slot;   // where the ray starts (see hit-point note in shade.cuh): -1 nowhere (camera rays), slot >= 0 of the TRIANGLE it
                  // leaves (never tested: dropped from the leaf's hit mask), -(slot + 2) of the SPHERE it leaves (tested, near root resolved)
};
// source codes of rays that leave a primitive
DSRT_HD int source_code(int slot, bool is_triangle) { return is_triangle ? slot : -(slot + 2); }
DSRT_HD bool leaves_sphere(int src, int slot) { return src < -1 && slot == -(src + 2); }
// Removes the source triangle from a node's primitive mask: a ray always passes the box of the triangle it starts on, so
// without this every secondary ray would fetch and test its own source once.  prim_mask / valid are in the node's nibble
// format (layout.h): record prim_base + r is the r-th set bit of `valid`.  The source lies in this node for about one node
// visit per ray, so the bit is looked up in a (rare) branch instead of with straight-line code on every visit.
// The node's `flat` word (layout.h) extends this to the source's coplanar slot mates -- the other half of a wall quad, which
// every ray leaving a Cornell wall would otherwise fetch and test (0.85 of the 3.1 primitive tests per shadow ray on the bench
// scene): a ray that starts in a triangle's plane meets that plane at t = 0 only.  The word arrives with the node's first
// 256-bit load (NodeRegs::flat), so the branch needs no memory access.
DSRT_HD uint32_t drop_source(uint32_t prim_mask, uint32_t prim_base, uint32_t valid, int src, uint32_t flat = 0u) {
  const uint32_t rel = (uint32_t)src - prim_base;     // wraps to a huge value for src < prim_base (incl. -1 and sphere codes)
  if (rel < 24u && rel < (uint32_t)hd_popc(valid)) {      // (a node holds at most 24 primitives: the first test is the cheap one)
    uint32_t v = valid;
#pragma unroll 1
    for (uint32_t i = 0; i < rel; i++) v &= v - 1u;
    const uint32_t bit = v & (0u - v);
    prim_mask &= ~(bit | (flat & (0xfu << ((31u - (uint32_t)hd_clz(bit)) & 0x1cu))));
  }
  return prim_mask;
}
// record index of the primitive whose (one-hot) bit in a node's primitive mask is `bit` (layout.h)
DSRT_HD int prim_slot(uint32_t prim_base, uint32_t valid, uint32_t bit) { return (int)(prim_base + (uint32_t)hd_popc(valid & (bit - 1u))); }
constexpr uint32_t kHitBits = 0x88888888u;        // node-group word: bit 4p+3 = pending internal child of priority p
constexpr uint32_t kInnerBits = 0x11111111u;      //                  bit 4s   = slot s is an internal child
struct TraceHit {
  float t, u, v;  // u,v = barycentric weights of p2,p3 (the reference's u,v, triangle.cpp:69-70)
  int slot;       // -1 = miss
};
struct TraceCounters { uint32_t nodes, prims; };

// double-precision ray for the parity kernel
struct Ray64 { double ox, oy, oz, dx, dy, dz; };

// ---- production primitive tests -------------------------------------------------------------------------
// Watertight ray/triangle test (Woop, Benthin, Wald 2013).  The per-ray shear + axis permutation
//   Ax = A[kx] - Sx*A[kz],  Ay = A[ky] - Sy*A[kz],  Az = Sz*A[kz]
// is folded into three per-ray vectors bx, by, bz (entries 1, -Sx | -Sy, 0 / Sz in permuted positions), so a
// vertex is mapped to ray space with three dot products and the test is free of data-dependent selects
// (which the compiler turns into divergent branches).
struct WatertightRay { float bxx, bxy, bxz, byx, byy, byz, bzx, bzy, bzz; };

DSRT_HD WatertightRay make_watertight(const TraceRay& r) {
  const float ax = fabsf(r.dx), ay = fabsf(r.dy), az = fabsf(r.dz);
  int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  const float dz = kz == 0 ? r.dx : (kz == 1 ? r.dy : r.dz);
  if (dz < 0.0f) { int t = kx; kx = ky; ky = t; }   // preserve winding
  const float dx = kx == 0 ? r.dx : (kx == 1 ? r.dy : r.dz), dy = ky == 0 ? r.dx : (ky == 1 ? r.dy : r.dz);
  const float Sz = hd_rcp(dz), Sx = dx * Sz, Sy = dy * Sz;
  WatertightRay w;
  w.bxx = (kx == 0 ? 1.0f : 0.0f) - (kz == 0 ? Sx : 0.0f); w.bxy = (kx == 1 ? 1.0f : 0.0f) - (kz == 1 ? Sx : 0.0f); w.bxz = (kx == 2 ? 1.0f : 0.0f) - (kz == 2 ? Sx : 0.0f);
  w.byx = (ky == 0 ? 1.0f : 0.0f) - (kz == 0 ? Sy : 0.0f); w.byy = (ky == 1 ? 1.0f : 0.0f) - (kz == 1 ? Sy : 0.0f); w.byz = (ky == 2 ? 1.0f : 0.0f) - (kz == 2 ? Sy : 0.0f);
  w.bzx = kz == 0 ? Sz : 0.0f; w.bzy = kz == 1 ? Sz : 0.0f; w.bzz = kz == 2 ? Sz : 0.0f;
  return w;
}

// returns true and updates (t,u,v) when the triangle is hit in (0, tmax)
DSRT_HD bool hit_triangle(const TraceRay& r, const WatertightRay& w, const float4 a, const float4 b,
                                             const float4 c, float tmax, float& t_out, float& u_out, float& v_out) {
  const float A0 = a.x - r.ox, A1 = a.y - r.oy, A2 = a.z - r.oz;
  const float B0 = b.x - r.ox, B1 = b.y - r.oy, B2 = b.z - r.oz;
  const float C0 = c.x - r.ox, C1 = c.y - r.oy, C2 = c.z - r.oz;
  const float Ax = hd_fma(A2, w.bxz, hd_fma(A1, w.bxy, A0 * w.bxx)), Ay = hd_fma(A2, w.byz, hd_fma(A1, w.byy, A0 * w.byx));
  const float Bx = hd_fma(B2, w.bxz, hd_fma(B1, w.bxy, B0 * w.bxx)), By = hd_fma(B2, w.byz, hd_fma(B1, w.byy, B0 * w.byx));
  const float Cx = hd_fma(C2, w.bxz, hd_fma(C1, w.bxy, C0 * w.bxx)), Cy = hd_fma(C2, w.byz, hd_fma(C1, w.byy, C0 * w.byx));
  // edge functions: products rounded separately, so swapping the two vertices of a shared edge negates the
  // value exactly (no cracks between adjacent triangles)
  float U = hd_sub(hd_mul(Cx, By), hd_mul(Cy, Bx));
  float V = hd_sub(hd_mul(Ax, Cy), hd_mul(Ay, Cx));
  float W = hd_sub(hd_mul(Bx, Ay), hd_mul(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = (float)hd_dsub(hd_dmul((double)Cx, (double)By), hd_dmul((double)Cy, (double)Bx));
    V = (float)hd_dsub(hd_dmul((double)Ax, (double)Cy), hd_dmul((double)Ay, (double)Cx));
    W = (float)hd_dsub(hd_dmul((double)Bx, (double)Ay), hd_dmul((double)By, (double)Ax));
  }
  if (fminf(fminf(U, V), W) < 0.0f && fmaxf(fmaxf(U, V), W) > 0.0f) return false;
  const float det = U + V + W;
  if (det == 0.0f) return false;
  const float Az = hd_fma(A2, w.bzz, hd_fma(A1, w.bzy, A0 * w.bzx));
  const float Bz = hd_fma(B2, w.bzz, hd_fma(B1, w.bzy, B0 * w.bzx));
  const float Cz = hd_fma(C2, w.bzz, hd_fma(C1, w.bzy, C0 * w.bzx));
  const float T = U * Az + V * Bz + W * Cz;
  const float rdet = hd_rcp(det);
  const float t = T * rdet;
  if (!(t > 0.0f && t < tmax)) return false;
  t_out = t; u_out = V * rdet; v_out = W * rdet;
  return true;
}

// Any-hit form of the same test for the cooperative shadow-ray path (DSRT_TRI_FAST): the translation by the ray origin is
// folded into the start of each FMA chain (po = -(o . basis row), computed once per ray), which removes the nine subtractions
// per triangle, and the division is replaced by a sign-corrected comparison 0 < T sgn(det) < tmax |det|.  A vertex still maps
// to the same (x, y) whichever triangle it is reached through, and the edge products are still rounded separately, so
// adjacent triangles stay crack-free; only (t, u, v), which an any-hit query does not return, lose the benefit of the
// exact vertex - origin difference.
DSRT_HD bool hit_triangle_any(float pox, float poy, float poz, const WatertightRay& w, const float4 a, const float4 b,
                              const float4 c, float tmax) {
  const float Ax = hd_fma(a.z, w.bxz, hd_fma(a.y, w.bxy, hd_fma(a.x, w.bxx, pox))), Ay = hd_fma(a.z, w.byz, hd_fma(a.y, w.byy, hd_fma(a.x, w.byx, poy)));
  const float Bx = hd_fma(b.z, w.bxz, hd_fma(b.y, w.bxy, hd_fma(b.x, w.bxx, pox))), By = hd_fma(b.z, w.byz, hd_fma(b.y, w.byy, hd_fma(b.x, w.byx, poy)));
  const float Cx = hd_fma(c.z, w.bxz, hd_fma(c.y, w.bxy, hd_fma(c.x, w.bxx, pox))), Cy = hd_fma(c.z, w.byz, hd_fma(c.y, w.byy, hd_fma(c.x, w.byx, poy)));
  float U = hd_sub(hd_mul(Cx, By), hd_mul(Cy, Bx));
  float V = hd_sub(hd_mul(Ax, Cy), hd_mul(Ay, Cx));
  float W = hd_sub(hd_mul(Bx, Ay), hd_mul(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = (float)hd_dsub(hd_dmul((double)Cx, (double)By), hd_dmul((double)Cy, (double)Bx));
    V = (float)hd_dsub(hd_dmul((double)Ax, (double)Cy), hd_dmul((double)Ay, (double)Cx));
    W = (float)hd_dsub(hd_dmul((double)Bx, (double)Ay), hd_dmul((double)By, (double)Ax));
  }
  if (fminf(fminf(U, V), W) < 0.0f && fmaxf(fmaxf(U, V), W) > 0.0f) return false;
  const float det = U + V + W;
  if (det == 0.0f) return false;
  const float Az = hd_fma(a.z, w.bzz, hd_fma(a.y, w.bzy, hd_fma(a.x, w.bzx, poz)));
  const float Bz = hd_fma(b.z, w.bzz, hd_fma(b.y, w.bzy, hd_fma(b.x, w.bzx, poz)));
  const float Cz = hd_fma(c.z, w.bzz, hd_fma(c.y, w.bzy, hd_fma(c.x, w.bzx, poz)));
  const float T = U * Az + V * Bz + W * Cz;
  const float Ts = hd_u2f(hd_f2u(T) ^ (hd_f2u(det) & 0x80000000u));     // T * sgn(det)
  return Ts > 0.0f && Ts < tmax * fabsf(det);
}

// Sphere::test / intersect semantics (sphere.cpp:10-77) in float.  A ray that STARTS on this sphere (src)
// has one root at ~0: the reference rejects it with t > min_t thanks to its 1e-11 origin offset; in float
// that root is resolved analytically (inward -> far root 2b, outward -> miss).
DSRT_HD bool hit_sphere(const TraceRay& r, const float4 a, const float4 b, bool is_src, bool any_hit,
                                           float tmax, float& t_out) {
  const float mx = a.x - r.ox, my = a.y - r.oy, mz = a.z - r.oz;
  const float bq = mx * r.dx + my * r.dy + mz * r.dz;
  if (is_src) {
    const float t2 = 2.0f * bq;
    if (!(t2 > 0.0f && t2 < tmax)) return false;
    t_out = t2;
    return true;
  }
  const float cq = mx * mx + my * my + mz * mz - b.y;
  const float delta = bq * bq - cq;
  if (delta < 0.0f) return false;
  const float sq = sqrtf(delta);
  const float t1 = bq - sq, t2 = bq + sq;
  if (any_hit) {                       // sphere.cpp:42-44: both roots alias the far root
    return !(t2 >= tmax || t2 <= 0.0f);
  }
  if (t1 >= tmax || t2 <= 0.0f) return false;
  const float t = (t1 <= 0.0f) ? t2 : t1;
  if (!(t < tmax)) return false;       // closest-hit semantics (the reference omits this test, sphere.cpp:64-72)
  t_out = t;
  return true;
}

// ---- parity (fp64, reference operation order, no FMA) ------------------------------------------------------
DSRT_HD double dm(double a, double b) { return hd_dmul(a, b); }
DSRT_HD double da(double a, double b) { return hd_dadd(a, b); }
DSRT_HD double ds(double a, double b) { return hd_dsub(a, b); }
DSRT_HD double ddot(double ax, double ay, double az, double bx, double by, double bz) {
  return da(da(dm(ax, bx), dm(ay, by)), dm(az, bz));
}
// Triangle::intersect(r, i), triangle.cpp:62-84
DSRT_HD bool hit_triangle64(const Ray64& r, const double* __restrict__ p, double tmax, double& t_out,
                                               double& u_out, double& v_out) {
  const double e1x = ds(p[3], p[0]), e1y = ds(p[4], p[1]), e1z = ds(p[5], p[2]);
  const double e2x = ds(p[6], p[0]), e2y = ds(p[7], p[1]), e2z = ds(p[8], p[2]);
  const double sx = ds(r.ox, p[0]), sy = ds(r.oy, p[1]), sz = ds(r.oz, p[2]);
  // cross(e1, d)
  const double c1x = ds(dm(e1y, r.dz), dm(e1z, r.dy)), c1y = ds(dm(e1z, r.dx), dm(e1x, r.dz)), c1z = ds(dm(e1x, r.dy), dm(e1y, r.dx));
  const double f = ddot(c1x, c1y, c1z, e2x, e2y, e2z);
  if (f == 0) return false;
  // cross(s, d)
  const double c2x = ds(dm(sy, r.dz), dm(sz, r.dy)), c2y = ds(dm(sz, r.dx), dm(sx, r.dz)), c2z = ds(dm(sx, r.dy), dm(sy, r.dx));
  const double u = hd_ddiv(ddot(c2x, c2y, c2z, e2x, e2y, e2z), f);
  const double v = hd_ddiv(ddot(c1x, c1y, c1z, sx, sy, sz), f);
  // cross(e1, -s)
  const double nsx = -sx, nsy = -sy, nsz = -sz;
  const double c3x = ds(dm(e1y, nsz), dm(e1z, nsy)), c3y = ds(dm(e1z, nsx), dm(e1x, nsz)), c3z = ds(dm(e1x, nsy), dm(e1y, nsx));
  const double t = hd_ddiv(ddot(c3x, c3y, c3z, e2x, e2y, e2z), f);
  if (!(u >= 0 && v >= 0 && da(u, v) <= 1 && t > 0.0 && t < tmax)) return false;
  t_out = t; u_out = u; v_out = v;
  return true;
}
// Sphere::intersect(r, i), sphere.cpp:10-77 (no t < i->t test: the hit is overwritten)
DSRT_HD bool hit_sphere64(const Ray64& r, const double* __restrict__ p, double tmax, double& t_out) {
  const double mx = ds(p[0], r.ox), my = ds(p[1], r.oy), mz = ds(p[2], r.oz);
  const double b = ddot(mx, my, mz, r.dx, r.dy, r.dz);
  const double c = ds(ddot(mx, my, mz, mx, my, mz), p[4]);
  const double delta = ds(dm(b, b), c);
  if (delta < 0) return false;
  const double sq = hd_dsqrt(delta);
  const double t1 = ds(b, sq), t2 = da(b, sq);
  if (t1 >= tmax || t2 <= 0.0) return false;
  t_out = (t1 <= 0.0) ? t2 : t1;
  return true;
}

// ---- node fetch -----------------------------------------------------------------------------------------------
struct NodeRegs {
  uint4 n0;            // ox, oy, oz (float bits), exponent bytes | imask
  uint4 n1;            // prim_base, valid, child_base, inner
  uint4 n2, n3, n4;    // quantised planes: (qlox, qloy) (qloz, qhix) (qhiy, qhiz), 8 bytes each
  float sx, sy, sz;    // plane scales 2^15 * 2^(e-127)
  uint32_t flat;       // leaf slots of coplanar triangles (layout.h); arrives with the first 256-bit load
};
DSRT_HD NodeRegs load_node(const uint4* __restrict__ nodes, uint32_t node) {
  NodeRegs n;
  const uint4* np = nodes + (size_t)node * kNodeQuads;
#if DSRT_NODE96
#ifdef __CUDA_ARCH__
  uint32_t sx, sy, sz, pad;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(n.n0.x), "=r"(n.n0.y), "=r"(n.n0.z), "=r"(n.n0.w), "=r"(sx), "=r"(sy), "=r"(sz), "=r"(pad) : "l"(np));
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(n.n1.x), "=r"(n.n1.y), "=r"(n.n1.z), "=r"(n.n1.w), "=r"(n.n2.x), "=r"(n.n2.y), "=r"(n.n2.z), "=r"(n.n2.w) : "l"(np + 2));
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(n.n3.x), "=r"(n.n3.y), "=r"(n.n3.z), "=r"(n.n3.w), "=r"(n.n4.x), "=r"(n.n4.y), "=r"(n.n4.z), "=r"(n.n4.w) : "l"(np + 4));
  n.sx = __uint_as_float(sx); n.sy = __uint_as_float(sy); n.sz = __uint_as_float(sz); n.flat = pad;
#else
  const uint4 s4 = np[1];
  n.n0 = np[0]; n.n1 = np[2]; n.n2 = np[3]; n.n3 = np[4]; n.n4 = np[5];
  n.sx = hd_u2f(s4.x); n.sy = hd_u2f(s4.y); n.sz = hd_u2f(s4.z); n.flat = s4.w;
#endif
#else
  n.n0 = hd_ldg(np); n.n1 = hd_ldg(np + 1); n.n2 = hd_ldg(np + 2); n.n3 = hd_ldg(np + 3); n.n4 = hd_ldg(np + 4);
  // the node stores its exponent bytes biased up by 15 (layout.h): a byte moved into the exponent field IS 2^15 * 2^(e-127)
  n.sx = hd_u2f((n.n0.w << 23) & 0x7f800000u); n.sy = hd_u2f((n.n0.w << 15) & 0x7f800000u); n.sz = hd_u2f((n.n0.w << 7) & 0x7f800000u);
  n.flat = 0u;
#endif
  return n;
}
// the (prim_base, valid) word pair of a node (layout.h)
DSRT_HD uint2 load_node_prims(const uint4* __restrict__ nodes, uint32_t node) {
  const uint2* p = reinterpret_cast<const uint2*>(nodes + (size_t)node * kNodeQuads + (DSRT_NODE96 ? 2 : 1));
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

// ---- node test -----------------------------------------------------------------------------------------------
struct NodeFrame {   // per-ray constants
  float idx, idy, idz;   // reciprocal direction (clamped away from 0)
  uint32_t oct4;         // 4 * (7 - octant): the XOR that turns a slot's nibble position into its visiting priority's
  uint32_t bytesel;      // PRMT selector that moves byte b of a word to byte b ^ (octant code >> 1) (order_children)
  bool nx, ny, nz;       // direction component is negative: the near plane of a slab is its HIGH plane
};

// t_scale != 1: the node test then works in units of 1 / t_scale (the saturating any-hit test uses t_scale = 1 / tmax, so
// that the ray's valid range [0, tmax] becomes [0, 1])
DSRT_HD NodeFrame make_frame(const TraceRay& r, float t_scale = 1.0f) {
  NodeFrame f;
  const float eps = 1.0e-24f;
  const float dx = fabsf(r.dx) > eps ? r.dx : copysignf(eps, r.dx);
  const float dy = fabsf(r.dy) > eps ? r.dy : copysignf(eps, r.dy);
  const float dz = fabsf(r.dz) > eps ? r.dz : copysignf(eps, r.dz);
  f.idx = hd_rcp(dx) * t_scale; f.idy = hd_rcp(dy) * t_scale; f.idz = hd_rcp(dz) * t_scale;
  f.nx = dx < 0.0f; f.ny = dy < 0.0f; f.nz = dz < 0.0f;
  const uint32_t oct = (f.nx ? 1u : 0u) | (f.ny ? 2u : 0u) | (f.nz ? 4u : 0u);
  f.oct4 = (7u - oct) << 2;
  f.bytesel = 0x3210u ^ ((f.oct4 >> 3) * 0x1111u);
  return f;
}

// byte i of w -> the float 1 + q * 2^-15 (q placed in mantissa bits 8..15 by ONE byte-permute, no int->float
// conversion on the XU pipe).  The node test then evaluates plane = fma(1 + q 2^-15, 2^15 s, b - 2^15 s) = q s + b.
// `one` is 0x3f800000 handed in through the kernel parameters: kept out of the instruction's single immediate
// slot so that the byte selector can be the immediate (otherwise every PRMT costs an extra register move).
DSRT_HD float byte_unit(uint32_t w, int i, uint32_t one) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__byte_perm(w, one, 0x7604u | ((uint32_t)i << 4)));
#else
  return hd_u2f(one | (((w >> (8 * i)) & 0xffu) << 8));
#endif
}

// Tests the 8 quantised child boxes of one node; returns one nibble per SLOT (bits 4s..4s+3 all set when the ray meets the
// box of slot s; the caller ANDs with the node's `valid` / `inner` words, layout.h -- empty slots drop out there).  Near / far
// planes are picked per axis from the ray's sign (no per-child min/max).  Float rounding of the dequantised planes (<= 2^-9
// quantum) is covered by the 1/64-quantum margin the host puts on every quantised plane (wide_bvh.cpp), so no widening is
// needed here.  pad > 0 only in parity mode (extra conservative slabs).
//
// SAT (any-hit rays, DSRT_SAT_SLAB): the frame is pre-scaled by 1 / tmax (make_frame's t_scale), so the ray's valid range is
// [0, 1] and every plane distance is evaluated with ONE saturating FMA: the clamp to [0, 1] replaces max(near, 0) and
// min(far, tmax), i.e. two of the five min/max/compare instructions per child (the ALU pipe is the busiest pipe of k_trace).
// The comparison becomes strict (near < far): a box entirely behind the origin (far clamps to 0) or entirely beyond tmax
// (near clamps to 1) then fails as it must.  Strictness cannot lose a real hit: every quantised box contains its exact box
// with >= 1/64 quantum to spare on each side, so a ray that touches the contents has near < far by >= 1/32 quantum of t.
template <bool PARITY, bool SAT = false>
DSRT_HD uint32_t test_children(const TraceRay& r, const NodeFrame& fr, const NodeRegs& nd, float tmax, float pad, uint32_t one) {
  const uint4 n0 = nd.n0, n2 = nd.n2, n3 = nd.n3, n4 = nd.n4;
  const float ox = hd_u2f(n0.x), oy = hd_u2f(n0.y), oz = hd_u2f(n0.z);
  // 2^15 * 2^(e-127) * idir
  const float sx = nd.sx * fr.idx, sy = nd.sy * fr.idy, sz = nd.sz * fr.idz;
  float bnx, bny, bnz, bfx, bfy, bfz, sfx = sx, sfy = sy, sfz = sz;
  if (PARITY) {
    const float k = 1.0001f;
    bnx = ((fr.nx ? ox + pad : ox - pad) - r.ox) * fr.idx - sx; bfx = (((fr.nx ? ox - pad : ox + pad) - r.ox) * fr.idx - sx) * k + pad;
    bny = ((fr.ny ? oy + pad : oy - pad) - r.oy) * fr.idy - sy; bfy = (((fr.ny ? oy - pad : oy + pad) - r.oy) * fr.idy - sy) * k + pad;
    bnz = ((fr.nz ? oz + pad : oz - pad) - r.oz) * fr.idz - sz; bfz = (((fr.nz ? oz - pad : oz + pad) - r.oz) * fr.idz - sz) * k + pad;
    sfx = sx * k; sfy = sy * k; sfz = sz * k;
  } else {
    bnx = bfx = (ox - r.ox) * fr.idx - sx; bny = bfy = (oy - r.oy) * fr.idy - sy; bnz = bfz = (oz - r.oz) * fr.idz - sz;
  }
  uint32_t mask = 0;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const uint32_t qlx = half ? n2.y : n2.x, qly = half ? n2.w : n2.z;
    const uint32_t qlz = half ? n3.y : n3.x, qhx = half ? n3.w : n3.z;
    const uint32_t qhy = half ? n4.y : n4.x, qhz = half ? n4.w : n4.z;
    const uint32_t nearx = fr.nx ? qhx : qlx, farx = fr.nx ? qlx : qhx;
    const uint32_t neary = fr.ny ? qhy : qly, fary = fr.ny ? qly : qhy;
    const uint32_t nearz = fr.nz ? qhz : qlz, farz = fr.nz ? qlz : qhz;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      bool hit;
      if (SAT && !PARITY) {
        const float ax = hd_fma_sat(byte_unit(nearx, i, one), sx, bnx), bx = hd_fma_sat(byte_unit(farx, i, one), sfx, bfx);
        const float ay = hd_fma_sat(byte_unit(neary, i, one), sy, bny), by = hd_fma_sat(byte_unit(fary, i, one), sfy, bfy);
        const float az = hd_fma_sat(byte_unit(nearz, i, one), sz, bnz), bz = hd_fma_sat(byte_unit(farz, i, one), sfz, bfz);
        hit = fmaxf(fmaxf(ax, ay), az) < fminf(fminf(bx, by), bz);
      } else {
        const float ax = hd_fma(byte_unit(nearx, i, one), sx, bnx), bx = hd_fma(byte_unit(farx, i, one), sfx, bfx);
        const float ay = hd_fma(byte_unit(neary, i, one), sy, bny), by = hd_fma(byte_unit(fary, i, one), sfy, bfy);
        const float az = hd_fma(byte_unit(nearz, i, one), sz, bnz), bz = hd_fma(byte_unit(farz, i, one), sfz, bfz);
        const float tn = fmaxf(fmaxf(ax, ay), fmaxf(az, 0.0f));
        const float tf = fminf(fminf(bx, by), fminf(bz, tmax));
        hit = tn <= tf;
      }
      if (hit) mask |= 0xfu << (16 * half + 4 * i);
    }
  }
  return mask;
}

// Internal-child hits (bit 4s+3 of slot s) -> visiting priority: slot s gets priority p = s ^ octinv (the host places
// children in octant-ordered slots, wide_bvh.cpp), i.e. bit 4p+3.  XOR of the nibble index = swap of adjacent nibbles
// (octinv bit 0; a bit-select of the word shifted up and down) + a byte permutation (octinv bits 1, 2; one PRMT).
DSRT_HD uint32_t order_children(uint32_t hits, const NodeFrame& fr) {
  const uint32_t sh = fr.oct4 & 4u;
  const uint32_t up = hits << sh, down = hits >> sh;
  const uint32_t x = (up & 0x80808080u) | (down & ~0x80808080u);      // sh == 0: x == hits
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(r) : "r"(x), "r"(fr.bytesel));      // (__byte_perm would mask the selector first)
  return r;
#else
  uint32_t r = 0;
  for (int b = 0; b < 4; b++) r |= ((x >> (8 * ((fr.bytesel >> (4 * b)) & 3u))) & 0xffu) << (8 * b);
  return r;
#endif
}

// ---- the traversal loop --------------------------------------------------------------------------------------
// One lane = one ray.  `stack` points at this thread's column of the shared-memory stack; entry e lives at
// stack[e * stride].
struct Accel {
  const uint4* __restrict__ nodes;         // 5 uint4 per node
  const float4* __restrict__ prims;        // 3 float4 per slot
  const double* __restrict__ prims64;      // 12 doubles per slot (parity only)
  float pad;                               // parity slab padding
  uint32_t one_bits;                       // 0x3f800000, see byte_unit()
  float bcx, bcy, bcz, brad;               // sphere around the root box: bounds the reach of rays with tmax = inf
};

#ifndef DSRT_TRI_FAST
#define DSRT_TRI_FAST 1                    // any-hit triangle test with the origin folded into the FMA chains (hit_triangle_any)
#endif
#ifndef DSRT_SAT_SLAB
#define DSRT_SAT_SLAB 1                    // saturating node test for any-hit rays (see test_children)
#endif
#ifndef DSRT_SAT_CLOSEST
#define DSRT_SAT_CLOSEST 1                 // ... and for closest-hit rays: the frame is rescaled by t_old / t_new whenever a closer hit is found
#endif
// 1 / (effective tmax) of a ray: tmax itself when it is finite (area / point light shadow rays), else the far side of the
// scene's bounding sphere (camera and bounce rays, directional / hemisphere / environment lights).  A ray with tmax <= 0 gets a
// huge scale: every box then clamps to a miss.
DSRT_HD float any_hit_scale(const Accel& A, const TraceRay& r) {
  if (r.tmax < 1.0e30f) return hd_rcp(fmaxf(r.tmax, 1.0e-30f));
  const float cx = A.bcx - r.ox, cy = A.bcy - r.oy, cz = A.bcz - r.oz;
  // in units of |d| (directions handed to dsrt_trace_any need not be normalised)
  const float reach = (sqrtf(cx * cx + cy * cy + cz * cz) + A.brad) * hd_rsqrt(r.dx * r.dx + r.dy * r.dy + r.dz * r.dz) * 1.0001f;
  return hd_rcp(fmaxf(reach, 1.0e-30f));
}
// closest hit: the node test works in units of the current best distance; a closer hit at t rescales the frame
DSRT_HD void rescale_frame(NodeFrame& fr, float t_scaled_to, float t_new) {
  const float k = t_scaled_to * hd_rcp(t_new);
  fr.idx *= k; fr.idy *= k; fr.idz *= k;
}

// DSRT_ANY_ORDERED: shadow rays open internal children front to back like closest-hit rays (1) or in slot order (0, saves
// the reordering; an any-hit query is correct in any order).  Front to back was 0.9 % faster while every shadow ray fetched
// the light quad and its wall mate; on the refined tree slot order wins: +2.0 % bench scene, +4.5 % glass stand-in, +0.4 %
// soups (r2c37, r2c39)
#ifndef DSRT_ANY_ORDERED
#define DSRT_ANY_ORDERED 0
#endif

// One node step, shared by trace_ray and k_trace: pops the highest-priority pending child of `ngroup` (the caller has checked
// ngroup.y & kHitBits), returns its node index; `more` tells whether the group still has pending children (push it back).
template <bool ORDERED>
DSRT_HD uint32_t next_child(uint2& ngroup, const NodeFrame& fr, bool& more) {
  const uint32_t bit = 31u - (uint32_t)hd_clz(ngroup.y & kHitBits);
  ngroup.y &= ~(1u << bit);
  more = (ngroup.y & kHitBits) != 0u;
  const uint32_t slot4 = ORDERED ? ((bit & 0x1cu) ^ fr.oct4) : (bit & 0x1cu);       // 4 * slot
  return ngroup.x + (uint32_t)hd_popc(ngroup.y & kInnerBits & ((1u << slot4) - 1u));
}
// node test results -> (children to open, primitives to test).  n1 = (prim_base, valid, child_base, inner), layout.h
template <bool ORDERED>
DSRT_HD void split_hits(uint32_t m, const uint4 n1, uint32_t flat, const NodeFrame& fr, int src_slot, uint32_t node, uint2& ngroup, uint2& tgroup) {
  const uint32_t open = m & n1.w;
  ngroup = make_uint2(n1.z, (ORDERED ? order_children(open, fr) : open) | (n1.w >> 3));
  tgroup = make_uint2(node, drop_source(m & n1.y, n1.x, n1.y, src_slot, flat));
}

template <bool ANY, bool PARITY, bool COUNT>
DSRT_HD void trace_ray(const Accel& A, const TraceRay& ray, const Ray64* ray64, uint2* stack, int stride,
                                          TraceHit& hit, double* t64_out, TraceCounters* cnt) {
  constexpr bool SAT = !PARITY && (ANY ? DSRT_SAT_SLAB : DSRT_SAT_CLOSEST);
  constexpr bool ORDERED = !ANY || DSRT_ANY_ORDERED;
  const float scale0 = SAT ? any_hit_scale(A, ray) : 1.0f;
  NodeFrame fr = make_frame(ray, scale0);
  float t_unit = SAT ? hd_rcp(scale0) : 1.0f;          // the distance that maps to 1 in the saturating node test
  const WatertightRay wr = make_watertight(ray);
  constexpr bool FAST = ANY && !PARITY && DSRT_TRI_FAST;
  const float pox = FAST ? -(ray.ox * wr.bxx + ray.oy * wr.bxy + ray.oz * wr.bxz) : 0.f;
  const float poy = FAST ? -(ray.ox * wr.byx + ray.oy * wr.byy + ray.oz * wr.byz) : 0.f;
  const float poz = FAST ? -(ray.ox * wr.bzx + ray.oy * wr.bzy + ray.oz * wr.bzz) : 0.f;
  float tbest = ray.tmax;
  double tbest64 = PARITY ? (double)ray.tmax : 0.0;
  hit.slot = -1; hit.t = ray.tmax; hit.u = 0.f; hit.v = 0.f;
  int sp = 0;
  uint2 ngroup = make_uint2(0u, 0x80000000u);          // "child 0 of nothing": the root
  uint2 tgroup = make_uint2(0u, 0u);
  uint32_t prim_base = 0, valid = 0;
  while (true) {
    if (ngroup.y & kHitBits) {
      bool more;
      const uint32_t node = next_child<ORDERED>(ngroup, fr, more);
      if (more) { stack[sp * stride] = ngroup; sp++; }
      const NodeRegs nd = load_node(A.nodes, node);
      if (COUNT) cnt->nodes++;
      const uint32_t m = test_children<PARITY, SAT>(ray, fr, nd, tbest, A.pad, A.one_bits);
      split_hits<ORDERED>(m, nd.n1, nd.flat, fr, ray.src_slot, node, ngroup, tgroup);
      prim_base = nd.n1.x; valid = nd.n1.y;
    } else {
      tgroup = make_uint2(0u, 0u);
    }
    while (tgroup.y) {
      const uint32_t low = tgroup.y & (0u - tgroup.y);        // lowest pending primitive first
      tgroup.y ^= low;
      const int slot = prim_slot(prim_base, valid, low);
      if (COUNT) cnt->prims++;
      if (PARITY) {
        const double* p = A.prims64 + (size_t)slot * 12;
        double t, u = 0, v = 0; bool h;
        const bool tri = p[11] != 0.0;
        if (tri) h = hit_triangle64(*ray64, p, tbest64, t, u, v);
        else h = hit_sphere64(*ray64, p, tbest64, t);
        if (h) { tbest64 = t; tbest = hd_d2f_ru(t); hit.slot = slot; hit.t = (float)t; hit.u = (float)u; hit.v = (float)v; }
      } else {
        const float4* pp = A.prims + (size_t)slot * 3;
        const float4 a = hd_ldg(pp), b = hd_ldg(pp + 1);
        float t, u = 0.f, v = 0.f; bool h;
        if (b.w != 0.0f) {
          const float4 c = hd_ldg(pp + 2);
          // (the source triangle never gets here: drop_source took it out of the node's primitive mask)
          if (FAST) { h = hit_triangle_any(pox, poy, poz, wr, a, b, c, tbest); t = 0.f; }
          else h = hit_triangle(ray, wr, a, b, c, tbest, t, u, v);
        } else {
          h = hit_sphere(ray, a, b, leaves_sphere(ray.src_slot, slot), ANY, tbest, t);
        }
        if (h) {
          if (ANY) { hit.slot = slot; hit.t = t; return; }
          tbest = t; hit.slot = slot; hit.t = t; hit.u = u; hit.v = v;
          if (SAT) { rescale_frame(fr, t_unit, t); t_unit = t; }
        }
      }
    }
    if (!(ngroup.y & kHitBits)) {
      if (sp == 0) break;
      sp--;
      ngroup = stack[sp * stride];
    }
  }
  if (PARITY && t64_out) *t64_out = tbest64;
}

}  // namespace dsrt
