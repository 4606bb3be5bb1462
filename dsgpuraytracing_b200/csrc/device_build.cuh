// device_build.cuh -- device-side construction of the compressed 8-wide BVH (option "device_build").
//
// What the reference attempted with its (disabled) parallel builder, cuda_src/setup.cu:478-686 + cuda_src/kernel.cu:358-493
// (Morton codes, radix sort, Karras' binary radix tree, bottom-up boxes), taken to the end here: the binary radix tree is
// collapsed ON THE DEVICE into the same 80-byte wide nodes, leaf-contiguous primitive order and 48-byte records that the host
// SAH path (bvh2_sah.cpp + wide_bvh.cpp) produces, so the traversal kernels, the parity kernel and the multi-GPU replication
// are unchanged.  It exists for the regime of SURVEY.md 8f-2 (tens of millions of primitives: the host SAH build + collapse of
// the 64 Mi-triangle soup takes ~37 s); the default remains the host SAH builder, whose topology matches the reference's.
// Closest / any-hit RESULTS do not depend on the tree, so gate 1 (bit-exact primary ids) holds for both builders (tested).
//
//   k_db_boxes     primitive boxes (float, rounded outwards from the double inputs) + scene box
//   k_db_morton    63-bit Morton code of each box centre
//   (cub radix sort of (code, primitive))
//   k_db_tree      Karras 2012 binary radix tree over the sorted codes (ties broken by index)
//   k_db_refit     bottom-up boxes, one thread per leaf, second arrival at a node continues upwards
//   k_db_collapse  one tree level per launch: a wide node takes the two children of its binary root and keeps replacing the
//                  child of largest surface area by its two children until it has 8 (a subtree of <= 3 primitives whose box
//                  is tight stays whole as one leaf child); octant-ordered slots and outward quantisation exactly as wide_bvh.cpp
//   k_db_mark_flat leaf slots of coplanar triangles (WideNode::flat)
//   k_db_flatten   48-byte primitive / shading records in slot order
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.h"

namespace dsrt {

struct DbScene {            // the caller's double-precision arrays, copied to the device as they are
  const int32_t* prim_type; const int32_t* prim_bsdf;
  const double* tri_pos; const double* tri_nrm; const double* sphere;
  int n;
};

// order-preserving float <-> uint map, for atomicMin / atomicMax on floats
__device__ __forceinline__ uint32_t db_f2o(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float db_o2f(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

// pbox: 6 floats per primitive (lo xyz, hi xyz); scene: 6 ordered uints (min lo xyz, max hi xyz)
__global__ void k_db_boxes(DbScene sc, float* __restrict__ pbox, uint32_t* scene) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3], hi[3];
  const bool live = i < sc.n;
  if (live) {
    double dlo[3], dhi[3];
    if (sc.prim_type[i] == 1) {
      const double* q = sc.tri_pos + 9 * (size_t)i;
      for (int k = 0; k < 3; k++) { dlo[k] = fmin(fmin(q[k], q[3 + k]), q[6 + k]); dhi[k] = fmax(fmax(q[k], q[3 + k]), q[6 + k]); }
    } else {
      const double* q = sc.sphere + 4 * (size_t)i;
      for (int k = 0; k < 3; k++) { dlo[k] = q[k] - q[3]; dhi[k] = q[k] + q[3]; }
    }
    for (int k = 0; k < 3; k++) { lo[k] = __double2float_rd(dlo[k]); hi[k] = __double2float_ru(dhi[k]); pbox[6 * (size_t)i + k] = lo[k]; pbox[6 * (size_t)i + 3 + k] = hi[k]; }
  } else {
    for (int k = 0; k < 3; k++) { lo[k] = kInfF; hi[k] = -kInfF; }
  }
  for (int k = 0; k < 3; k++) {
    float a = lo[k], b = hi[k];
    for (int o = 16; o; o >>= 1) { a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(scene + k, db_f2o(a)); atomicMax(scene + 3 + k, db_f2o(b)); }
  }
}

__device__ __forceinline__ uint64_t db_spread21(uint32_t v) {   // 21 bits -> every third bit
  uint64_t x = v & 0x1fffffu;
  x = (x | (x << 32)) & 0x1f00000000ffffull;
  x = (x | (x << 16)) & 0x1f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

__global__ void k_db_morton(int n, const float* __restrict__ pbox, const uint32_t* __restrict__ scene, uint64_t* keys, uint32_t* vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t code = 0;
  for (int k = 0; k < 3; k++) {
    const double lo = (double)db_o2f(scene[k]), hi = (double)db_o2f(scene[3 + k]);
    const double c = 0.5 * ((double)pbox[6 * (size_t)i + k] + (double)pbox[6 * (size_t)i + 3 + k]);
    double u = hi > lo ? (c - lo) / (hi - lo) : 0.0;
    u = fmin(fmax(u, 0.0), 1.0);
    const uint32_t q = (uint32_t)fmin(u * 2097152.0, 2097151.0);
    code |= db_spread21(q) << k;
  }
  keys[i] = code; vals[i] = (uint32_t)i;
}

// ---- Karras 2012 ---------------------------------------------------------------------------------------------------------
// internal nodes 0 .. n-2 (0 = root); a child reference >= 0 is an internal node, c < 0 is sorted leaf -(c + 1)
__device__ __forceinline__ int db_delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const uint64_t a = keys[i], b = keys[j];
  return a == b ? 64 + __clz(i ^ j) : __clzll((long long)(a ^ b));
}

__global__ void k_db_tree(int n, const uint64_t* __restrict__ keys, int* left, int* right, int* first, int* last, int* parent_int, int* parent_leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = db_delta(keys, n, i, i + 1) - db_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
  const int dmin = db_delta(keys, n, i, i - d);
  int lmax = 2;
  while (db_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2) if (db_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = db_delta(keys, n, i, j);
  int s = 0, t = l;
  do { t = (t + 1) / 2; if (db_delta(keys, n, i, i + (s + t) * d) > dnode) s += t; } while (t > 1);
  const int gamma = i + s * d + (d < 0 ? -1 : 0);
  const int lo = min(i, j), hi = max(i, j);
  const int L = lo == gamma ? -(gamma + 1) : gamma, R = hi == gamma + 1 ? -(gamma + 2) : gamma + 1;
  left[i] = L; right[i] = R; first[i] = lo; last[i] = hi;
  if (L >= 0) parent_int[L] = i; else parent_leaf[-(L + 1)] = i;
  if (R >= 0) parent_int[R] = i; else parent_leaf[-(R + 1)] = i;
  if (i == 0) parent_int[0] = -1;
}

__device__ __forceinline__ void db_child_box(int c, const float* __restrict__ ibox, const float* __restrict__ pbox, const uint32_t* __restrict__ sorted, float* b) {
  const float* src = c >= 0 ? ibox + 6 * (size_t)c : pbox + 6 * (size_t)sorted[-(c + 1)];
  for (int k = 0; k < 6; k++) b[k] = src[k];
}

// one thread per leaf; the second thread to arrive at an internal node computes its box and goes on
__global__ void k_db_refit(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ parent_int,
                           const int* __restrict__ parent_leaf, const float* __restrict__ pbox, const uint32_t* __restrict__ sorted,
                           float* ibox, int* visit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = parent_leaf[i];
  while (p >= 0) {
    __threadfence();
    if (atomicAdd(visit + p, 1) == 0) return;
    __threadfence();
    // the sibling's box was written by another SM just before its atomic: read it through L2 (ld.global.cg), never through a
    // possibly stale L1 line
    float a[6], b[6];
    const int cl = left[p], cr = right[p];
    for (int k = 0; k < 6; k++) {
      a[k] = cl >= 0 ? __ldcg(ibox + 6 * (size_t)cl + k) : pbox[6 * (size_t)sorted[-(cl + 1)] + k];
      b[k] = cr >= 0 ? __ldcg(ibox + 6 * (size_t)cr + k) : pbox[6 * (size_t)sorted[-(cr + 1)] + k];
    }
    volatile float* out = ibox + 6 * (size_t)p;
    for (int k = 0; k < 3; k++) { out[k] = fminf(a[k], b[k]); out[3 + k] = fmaxf(a[3 + k], b[3 + k]); }
    p = parent_int[p];
  }
}

// ---- collapse ---------------------------------------------------------------------------------------------------------------
struct DbKid { int ref; int first, count; float b[6]; };

__device__ __forceinline__ void db_load_kid(int ref, const float* __restrict__ ibox, const float* __restrict__ pbox, const uint32_t* __restrict__ sorted,
                                            const int* __restrict__ first, const int* __restrict__ last, DbKid& k) {
  k.ref = ref;
  if (ref >= 0) { k.first = first[ref]; k.count = last[ref] - first[ref] + 1; }
  else { k.first = -(ref + 1); k.count = 1; }
  db_child_box(ref, ibox, pbox, sorted, k.b);
}
__device__ __forceinline__ float db_area(const DbKid& k) {
  const float x = k.b[3] - k.b[0], y = k.b[4] - k.b[1], z = k.b[5] - k.b[2];
  return x * y + y * z + z * x;
}

struct DbTree {
  const int* left; const int* right; const int* first; const int* last;
  const float* ibox; const float* pbox; const uint32_t* sorted;
};

// A subtree of 2-3 primitives stays whole as ONE leaf child only when its box is tight: a ray that enters the group's box pays
// for `count` primitive tests, against one test per primitive box entered when the primitives get a box each.  Mesh neighbours
// share most of their boxes (kept together); unrelated primitives that merely sort next to each other (triangle soups) do not.
#ifndef DSRT_DB_GROUP_ALPHA
#define DSRT_DB_GROUP_ALPHA 1.3f
#endif
__device__ __forceinline__ bool db_is_leaf_child(const DbKid& k, const DbTree& T) {
  if (k.ref < 0) return true;
  if (k.count > 3) return false;
  float sum = 0.f;
  for (int j = 0; j < k.count; j++) {
    const float* b = T.pbox + 6 * (size_t)T.sorted[k.first + j];
    const float x = b[3] - b[0], y = b[4] - b[1], z = b[5] - b[2];
    sum += x * y + y * z + z * x;
  }
  return db_area(k) * (float)k.count <= DSRT_DB_GROUP_ALPHA * sum;
}

// axis-aligned area lights (wide_bvh.h EndPlane), by value to the collapse kernel: a node holding a flat child in such a plane
// shifts its quantisation grid exactly like wide_bvh.cpp emit_node does
constexpr int kDbMaxEndPlanes = 4;
struct DbEndPlanes { int n; int axis[kDbMaxEndPlanes]; int from_low[kDbMaxEndPlanes]; double coord[kDbMaxEndPlanes]; double lo[kDbMaxEndPlanes][3], hi[kDbMaxEndPlanes][3]; };

// items: (binary reference, wide node index).  counters: [0] wide nodes allocated, [1] primitive slots allocated, [2] items of
// the next level, [3] error flag
__global__ void k_db_collapse(DbTree T, const int2* __restrict__ items, int n_items, int2* next_items, unsigned int* counters,
                              WideNode* nodes, int32_t* slot_prim, unsigned int node_cap, DbEndPlanes ends) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_items) return;
  const int2 it = items[t];
  DbKid kids[8]; int nk = 0; unsigned inner = 0;      // inner: bit k set = kids[k] becomes a wide node of its own
  {
    DbKid root; db_load_kid(it.x, T.ibox, T.pbox, T.sorted, T.first, T.last, root);
    if (db_is_leaf_child(root, T)) { kids[nk++] = root; }                             // the whole scene is one leaf child
    else {
      db_load_kid(T.left[root.ref], T.ibox, T.pbox, T.sorted, T.first, T.last, kids[0]);
      db_load_kid(T.right[root.ref], T.ibox, T.pbox, T.sorted, T.first, T.last, kids[1]);
      nk = 2;
      inner = (db_is_leaf_child(kids[0], T) ? 0u : 1u) | (db_is_leaf_child(kids[1], T) ? 0u : 2u);
      while (nk < 8) {
        int best = -1; float best_a = -1.f;
        for (int k = 0; k < nk; k++) if ((inner >> k) & 1) { const float a = db_area(kids[k]); if (a > best_a) { best_a = a; best = k; } }
        if (best < 0) break;
        const int r = kids[best].ref;
        db_load_kid(T.left[r], T.ibox, T.pbox, T.sorted, T.first, T.last, kids[best]);
        db_load_kid(T.right[r], T.ibox, T.pbox, T.sorted, T.first, T.last, kids[nk]);
        inner &= ~(1u << best);
        if (!db_is_leaf_child(kids[best], T)) inner |= 1u << best;
        if (!db_is_leaf_child(kids[nk], T)) inner |= 1u << nk;
        nk++;
      }
    }
  }
  // node box
  double nlo[3], nhi[3];
  for (int k = 0; k < 3; k++) { nlo[k] = kids[0].b[k]; nhi[k] = kids[0].b[3 + k]; }
  for (int i = 1; i < nk; i++) for (int k = 0; k < 3; k++) { nlo[k] = fmin(nlo[k], (double)kids[i].b[k]); nhi[k] = fmax(nhi[k], (double)kids[i].b[3 + k]); }
  // octant-ordered slots: greedy assignment maximising (child centre - node centre) . octant direction (wide_bvh.cpp emit_node)
  int slot_of[8]; unsigned used = 0, done = 0;
  for (int round = 0; round < nk; round++) {
    int bi = -1, bs = -1; double bc = -1e300;
    for (int i = 0; i < nk; i++) if (!((done >> i) & 1)) for (int s = 0; s < 8; s++) if (!((used >> s) & 1)) {
      double c = 0;
      for (int k = 0; k < 3; k++) c += (0.5 * ((double)kids[i].b[k] + (double)kids[i].b[3 + k]) - 0.5 * (nlo[k] + nhi[k])) * (((s >> k) & 1) ? 1.0 : -1.0);
      if (c > bc || bi < 0) { bc = c; bi = i; bs = s; }
    }
    slot_of[bi] = bs; done |= 1u << bi; used |= 1u << bs;
  }
  int kid_at[8]; for (int s = 0; s < 8; s++) kid_at[s] = -1;
  for (int i = 0; i < nk; i++) kid_at[slot_of[i]] = i;
  // quantisation frame (wide_bvh.cpp emit_node, step 4)
  WideNode w; memset(&w, 0, sizeof(w));
  float org[3]; double scale[3];
  for (int k = 0; k < 3; k++) {
    const double ext = nhi[k] - nlo[k];
    const double mag = fmax(fmax(fabs(nlo[k]), fabs(nhi[k])), 1e-30);
    int e = ext > 0 ? (int)ceil(log2(ext / 252.0)) : -126;
    e = max(e, ilogb(mag) - 18);
    e = min(110, max(-120, e));
    while (true) {
      scale[k] = ldexp(1.0, e);
      org[k] = __double2float_rd(nlo[k] - scale[k]);
      if (ceil((nhi[k] - (double)org[k]) / scale[k] + 1.0 / 64) <= 255.0 || e >= 110) break;
      e++;
    }
    // light-aligned grid (wide_bvh.cpp emit_node): a flat child in the plane of an area light gets its light-facing plane
    // 3/128 quantum past a grid line
    for (int p = 0; p < ends.n; p++) {
      if (ends.axis[p] != k) continue;
      const double tol = 1e-6 * fmax(1.0, fabs(ends.coord[p]));
      bool found = false; double at = 0;
      for (int i = 0; i < nk && !found; i++) {
        const float* b = kids[i].b;
        if (fabs((double)b[k] - ends.coord[p]) > tol || fabs((double)b[3 + k] - ends.coord[p]) > tol) continue;
        bool overlap = true;
        for (int a = 0; a < 3; a++) if (a != k && ((double)b[3 + a] < ends.lo[p][a] || (double)b[a] > ends.hi[p][a])) overlap = false;
        if (overlap) { found = true; at = ends.from_low[p] ? (double)b[k] : (double)b[3 + k]; }
      }
      if (!found) continue;
      const double want = ends.from_low[p] ? 3.0 / 128 : 1.0 - 3.0 / 128;
      const double x0 = (at - (nlo[k] - scale[k])) / scale[k];
      double d = want - (x0 - floor(x0));
      if (d < 0) d += 1.0;
      const float org_a = __double2float_rd(nlo[k] - scale[k] * (1.0 + d));
      const double xa = (at - (double)org_a) / scale[k], fa = xa - floor(xa);
      if (fabs(fa - want) <= 1.0 / 256 && (double)org_a <= nlo[k] - scale[k] && ceil((nhi[k] - (double)org_a) / scale[k] + 1.0 / 64) <= 255.0) org[k] = org_a;
      break;
    }
    (&w.ex)[k] = (uint8_t)(e + 127 + 15);
#if DSRT_NODE96
    (&w.sx)[k] = __uint_as_float((uint32_t)(e + 127 + 15) << 23);
#endif
  }
  w.ox = org[0]; w.oy = org[1]; w.oz = org[2];
  int n_internal = 0, n_leaf_prims = 0;
  for (int s = 0; s < 8; s++) if (kid_at[s] >= 0) { const DbKid& c = kids[kid_at[s]]; if ((inner >> kid_at[s]) & 1) n_internal++; else n_leaf_prims += c.count; }
  const unsigned child_base = n_internal ? atomicAdd(counters + 0, (unsigned)n_internal) : 0u;
  const unsigned prim_base = n_leaf_prims ? atomicAdd(counters + 1, (unsigned)n_leaf_prims) : 0u;
  const unsigned next_base = n_internal ? atomicAdd(counters + 2, (unsigned)n_internal) : 0u;
  if (child_base + (unsigned)n_internal > node_cap) { atomicExch(counters + 3, 1u); return; }
  w.child_base = child_base; w.prim_base = prim_base;
  int rank = 0, prim_off = 0;
  for (int s = 0; s < 8; s++) {
    if (kid_at[s] < 0) continue;
    const DbKid& c = kids[kid_at[s]];
    uint8_t q[6];
    for (int k = 0; k < 3; k++) {
      const double lo = floor(((double)c.b[k] - (double)org[k]) / scale[k] - 1.0 / 64);
      const double hi = ceil(((double)c.b[3 + k] - (double)org[k]) / scale[k] + 1.0 / 64);
      q[k] = (uint8_t)fmin(255.0, fmax(0.0, lo)); q[3 + k] = (uint8_t)fmin(255.0, fmax(0.0, hi));
    }
    w.qlox[s] = q[0]; w.qloy[s] = q[1]; w.qloz[s] = q[2]; w.qhix[s] = q[3]; w.qhiy[s] = q[4]; w.qhiz[s] = q[5];
    if ((inner >> kid_at[s]) & 1) {
      w.inner |= 8u << (4 * s);
      w.imask |= (uint8_t)(1 << s);
      next_items[next_base + rank] = make_int2(c.ref, (int)(child_base + rank));
      rank++;
    } else {
      w.valid |= ((1u << c.count) - 1u) << (4 * s);
      for (int j = 0; j < c.count; j++) slot_prim[prim_base + prim_off + j] = (int32_t)T.sorted[c.first + j];
      prim_off += c.count;
    }
  }
  nodes[it.y] = w;
}

// WideNode::flat (layout.h), as mark_flat_slots (wide_bvh.cpp) sets it: leaf slots whose 2-3 triangles lie in one plane
__global__ void k_db_mark_flat(DbScene sc, WideNode* nodes, int n_nodes, const int32_t* __restrict__ slot_prim) {
#if DSRT_NODE96
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  const uint32_t valid = nodes[n].valid, prim_base = nodes[n].prim_base;
  uint32_t flat = 0, rank = 0;
  for (int s = 0; s < 8; s++) {
    const int c = __popc((valid >> (4 * s)) & 0xfu);
    const uint32_t first = prim_base + rank; rank += (uint32_t)c;
    if (c < 2) continue;
    const double* best = nullptr; double best_n2 = 0, nx = 0, ny = 0, nz = 0, ext = 0; bool tris = true;
    for (int i = 0; i < c; i++) {
      const int p = slot_prim[first + i];
      if (sc.prim_type[p] != 1) { tris = false; break; }
      const double* v = sc.tri_pos + 9 * (size_t)p;
      const double ax = v[3] - v[0], ay = v[4] - v[1], az = v[5] - v[2], bx = v[6] - v[0], by = v[7] - v[1], bz = v[8] - v[2];
      const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx, n2 = cx * cx + cy * cy + cz * cz;
      ext = fmax(ext, sqrt(fmax(ax * ax + ay * ay + az * az, bx * bx + by * by + bz * bz)));
      if (n2 > best_n2) { best_n2 = n2; best = v; nx = cx; ny = cy; nz = cz; }
    }
    if (!tris || !best || !(best_n2 > 0) || !isfinite(best_n2)) continue;
    const double inv = 1.0 / sqrt(best_n2);
    double dev = 0;
    for (int i = 0; i < c; i++) {
      const double* v = sc.tri_pos + 9 * (size_t)slot_prim[first + i];
      for (int k = 0; k < 3; k++) {
        const double dx = v[3 * k] - best[0], dy = v[3 * k + 1] - best[1], dz = v[3 * k + 2] - best[2];
        dev = fmax(dev, fabs((dx * nx + dy * ny + dz * nz) * inv));
        ext = fmax(ext, sqrt(dx * dx + dy * dy + dz * dz));
      }
    }
    if (dev <= 1e-6 * ext) flat |= 0xfu << (4 * s);
  }
  nodes[n].flat = flat;
#endif
}

__global__ void k_db_flatten(DbScene sc, int n_slots, const int32_t* __restrict__ slot_prim, PrimRecord* recs, ShadeRecord* shd) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const int p = slot_prim[s];
  PrimRecord r; ShadeRecord h; memset(&r, 0, sizeof(r)); memset(&h, 0, sizeof(h));
  r.prim_id = p; r.bsdf = sc.prim_bsdf[p];
  if (sc.prim_type[p] == 1) {
    const double* q = sc.tri_pos + 9 * (size_t)p; const double* nn = sc.tri_nrm + 9 * (size_t)p;
    r.ax = (float)q[0]; r.ay = (float)q[1]; r.az = (float)q[2]; r.bx = (float)q[3]; r.by = (float)q[4]; r.bz = (float)q[5];
    r.cx = (float)q[6]; r.cy = (float)q[7]; r.cz = (float)q[8]; r.is_tri = 1.0f;
    h.n1x = (float)nn[0]; h.n1y = (float)nn[1]; h.n1z = (float)nn[2]; h.n2x = (float)nn[3]; h.n2y = (float)nn[4]; h.n2z = (float)nn[5];
    h.n3x = (float)nn[6]; h.n3y = (float)nn[7]; h.n3z = (float)nn[8];
  } else {
    const double* q = sc.sphere + 4 * (size_t)p;
    r.ax = (float)q[0]; r.ay = (float)q[1]; r.az = (float)q[2]; r.bx = (float)q[3]; r.by = (float)(q[3] * q[3]); r.is_tri = 0.0f;
  }
  recs[s] = r; shd[s] = h;
}

}  // namespace dsrt
