// wide_bvh.cpp -- collapse the reference-identical binary SAH tree into the compressed 8-wide BVH the
// sm_100a traversal kernels consume (layout.h), and emit primitives in leaf-contiguous order.
//
//  1. copy the binary tree; split any leaf with more than 3 primitives into halves (the reference allows
//     4 per leaf, and leaves of arbitrary size after a degenerate split, bvh.cpp:143-173), because the
//     24-bit primitive part of the hit mask gives each of the 8 children at most 3 primitives;
//  2. top-down collapse: start from a node's two children and repeatedly open the internal child with the
//     largest surface area until there are 8 children or only leaves left;
//  3. place children in octant-ordered slots (greedy auction on dot(child centre - node centre, octant
//     direction)) so that slot ^ (7 - ray octant) is a front-to-back visiting priority;
//  4. quantise child boxes to 8 bits per plane, outwards, relative to a float origin rounded down and
//     per-axis power-of-two scales.
// Wide nodes are laid out breadth-first, so the top of the tree is contiguous in memory and the internal
// children of a node are adjacent (one child_base per node).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <queue>

#include "wide_bvh.h"

namespace dsrt {
namespace {

struct BNode { Box3 box; int start, range, l, r; };

float round_down(double v) {
  float f = (float)v;
  if ((double)f > v) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
  return f;
}

}  // namespace

int build_wide_bvh(const dsrt_bvh2& b2, const std::vector<Box3>& pbox, int n_prims, WideBVH& out, std::string& err) {
  out.nodes.clear(); out.slot_prim.clear(); out.max_depth = 0;
  if (b2.n_nodes <= 0 || n_prims <= 0) {
    // empty scene: a single node with no children
    WideNode w; std::memset(&w, 0, sizeof(w)); w.ex = w.ey = w.ez = 1;
    out.nodes.push_back(w); out.max_depth = 1;
    return DSRT_OK;
  }
  // 1. working copy, unary nodes spliced out, big leaves split
  std::vector<BNode> T((size_t)b2.n_nodes);
  for (int i = 0; i < b2.n_nodes; i++) {
    BNode& n = T[i];
    for (int k = 0; k < 3; k++) { n.box.lo[k] = b2.node_bbox[6 * i + k]; n.box.hi[k] = b2.node_bbox[6 * i + 3 + k]; }
    n.start = b2.node_start[i]; n.range = b2.node_range[i]; n.l = b2.node_left[i]; n.r = b2.node_right[i];
    if (n.start < 0 || n.range < 0 || n.start + n.range > n_prims || n.l >= b2.n_nodes || n.r >= b2.n_nodes) {
      err = "dsrt_set_bvh: node " + std::to_string(i) + " is out of range"; return DSRT_ERR_INVALID;
    }
  }
  auto resolve = [&](int id) {            // follow single-child chains (bvh.cpp:238-243 does the same at run time)
    while (id >= 0) {
      const BNode& n = T[id];
      if (n.l >= 0 && n.r < 0) id = n.l; else if (n.r >= 0 && n.l < 0) id = n.r; else break;
    }
    return id;
  };
  {
    std::vector<int> work;
    for (int i = 0; i < b2.n_nodes; i++) if (T[i].l < 0 && T[i].r < 0 && T[i].range > 3) work.push_back(i);
    while (!work.empty()) {
      int id = work.back(); work.pop_back();
      int s = T[id].start, r = T[id].range, h = r / 2;
      BNode a, b; a.box.reset(); b.box.reset();
      for (int q = 0; q < r; q++) (q < h ? a.box : b.box).grow(pbox[b2.prim_order[s + q]]);
      a.start = s; a.range = h; a.l = a.r = -1; b.start = s + h; b.range = r - h; b.l = b.r = -1;
      int ai = (int)T.size(); T.push_back(a); int bi = (int)T.size(); T.push_back(b);
      T[id].l = ai; T[id].r = bi;
      if (a.range > 3) work.push_back(ai);
      if (b.range > 3) work.push_back(bi);
    }
  }
  const int root = resolve(0);

  // 2..4. breadth-first emission
  struct Job { int bnode; uint32_t widx; int depth; };
  std::queue<Job> jobs;
  out.nodes.emplace_back();
  jobs.push({root, 0u, 1});
  while (!jobs.empty()) {
    Job job = jobs.front(); jobs.pop();
    out.max_depth = std::max(out.max_depth, job.depth);
    const BNode& bn = T[job.bnode];
    int kids[8]; int nk = 0;
    if (bn.l < 0 && bn.r < 0) kids[nk++] = job.bnode;            // the root itself is a leaf
    else { kids[nk++] = resolve(bn.l); kids[nk++] = resolve(bn.r); }
    while (nk < 8) {
      int best = -1; double barea = -1;
      for (int i = 0; i < nk; i++) {
        const BNode& c = T[kids[i]];
        if (c.l < 0 && c.r < 0) continue;
        double a = c.box.half_area();
        if (!(a >= 0)) a = 0;
        if (a > barea) { barea = a; best = i; }
      }
      if (best < 0) break;
      int c = kids[best];
      kids[best] = resolve(T[c].l);
      kids[nk++] = resolve(T[c].r);
    }
    // node box = union of children (equals the binary node's box; recomputed so virtual splits are covered)
    Box3 nb; nb.reset();
    for (int i = 0; i < nk; i++) nb.grow(T[kids[i]].box);
    // 3. octant-ordered slots
    int slot_of[8]; bool slot_used[8] = {false}; bool kid_done[8] = {false};
    double cost[8][8];
    for (int i = 0; i < nk; i++) for (int s = 0; s < 8; s++) {
      double c = 0;
      for (int k = 0; k < 3; k++) c += (T[kids[i]].box.centre(k) - nb.centre(k)) * (((s >> k) & 1) ? 1.0 : -1.0);
      cost[i][s] = c;
    }
    for (int round = 0; round < nk; round++) {
      int bi = -1, bs = -1; double bc = -std::numeric_limits<double>::infinity();
      for (int i = 0; i < nk; i++) if (!kid_done[i]) for (int s = 0; s < 8; s++) if (!slot_used[s] && cost[i][s] > bc) { bc = cost[i][s]; bi = i; bs = s; }
      if (bi < 0) {   // NaN costs (degenerate boxes): first free pair
        for (int i = 0; i < nk && bi < 0; i++) if (!kid_done[i]) for (int s = 0; s < 8; s++) if (!slot_used[s]) { bi = i; bs = s; break; }
      }
      slot_of[bi] = bs; kid_done[bi] = true; slot_used[bs] = true;
    }
    int kid_at[8]; for (int s = 0; s < 8; s++) kid_at[s] = -1;
    for (int i = 0; i < nk; i++) kid_at[slot_of[i]] = kids[i];

    // 4. quantisation frame
    WideNode w; std::memset(&w, 0, sizeof(w));
    float org[3]; double scale[3]; uint8_t ebits[3];
    for (int k = 0; k < 3; k++) {
      org[k] = round_down(nb.lo[k]);
      double ext = nb.hi[k] - (double)org[k];
      int e = -126;
      if (ext > 0) { e = (int)std::ceil(std::log2(ext / 255.0)); while (std::ldexp(255.0, e) < ext) e++; }
      e = std::min(127, std::max(-126, e));
      ebits[k] = (uint8_t)(e + 127); scale[k] = std::ldexp(1.0, e);
    }
    w.ox = org[0]; w.oy = org[1]; w.oz = org[2]; w.ex = ebits[0]; w.ey = ebits[1]; w.ez = ebits[2];
    w.prim_base = (uint32_t)out.slot_prim.size();
    int n_internal = 0;
    for (int s = 0; s < 8; s++) if (kid_at[s] >= 0 && !(T[kid_at[s]].l < 0 && T[kid_at[s]].r < 0)) n_internal++;
    w.child_base = (uint32_t)out.nodes.size();
    if (n_internal) out.nodes.resize(out.nodes.size() + n_internal);
    uint32_t next_child = w.child_base; int prim_off = 0;
    for (int s = 0; s < 8; s++) {
      int c = kid_at[s];
      if (c < 0) continue;
      const BNode& cn = T[c];
      uint8_t q[6];
      for (int k = 0; k < 3; k++) {
        double lo = std::floor((cn.box.lo[k] - (double)org[k]) / scale[k]);
        double hi = std::ceil((cn.box.hi[k] - (double)org[k]) / scale[k]);
        q[k] = (uint8_t)std::min(255.0, std::max(0.0, lo));
        q[3 + k] = (uint8_t)std::min(255.0, std::max(0.0, hi));
      }
      w.qlox[s] = q[0]; w.qloy[s] = q[1]; w.qloz[s] = q[2]; w.qhix[s] = q[3]; w.qhiy[s] = q[4]; w.qhiz[s] = q[5];
      if (cn.l < 0 && cn.r < 0) {
        if (cn.range < 1 || cn.range > 3 || prim_off + cn.range > 24) { err = "wide BVH: leaf packing overflow"; return DSRT_ERR_LIMIT; }
        uint8_t unary = cn.range == 1 ? 1 : (cn.range == 2 ? 3 : 7);
        w.meta[s] = (uint8_t)((unary << 5) | prim_off);
        for (int q2 = 0; q2 < cn.range; q2++) out.slot_prim.push_back(b2.prim_order[cn.start + q2]);
        prim_off += cn.range;
      } else {
        w.meta[s] = (uint8_t)((1 << 5) | (24 + s));
        w.imask |= (uint8_t)(1 << s);
        jobs.push({c, next_child++, job.depth + 1});
      }
    }
    out.nodes[job.widx] = w;
  }
  if ((int)out.slot_prim.size() != n_prims) { err = "wide BVH: primitive count mismatch (" + std::to_string(out.slot_prim.size()) + " vs " + std::to_string(n_prims) + ")"; return DSRT_ERR_INVALID; }
  return DSRT_OK;
}

void flatten_records(const dsrt_scene& sc, const WideBVH& wide, std::vector<PrimRecord>& recs, std::vector<ShadeRecord>& shd,
                     std::vector<PrimRecord64>& r64) {
  const size_t n = wide.slot_prim.size();
  recs.assign(n ? n : 1, PrimRecord{}); shd.assign(n ? n : 1, ShadeRecord{}); r64.assign(n ? n : 1, PrimRecord64{});
  for (size_t sl = 0; sl < n; sl++) {
    const int p = wide.slot_prim[sl];
    PrimRecord& r = recs[sl]; ShadeRecord& h = shd[sl]; PrimRecord64& d = r64[sl];
    std::memset(&r, 0, sizeof(r)); std::memset(&h, 0, sizeof(h)); std::memset(&d, 0, sizeof(d));
    r.prim_id = p; r.bsdf = sc.prim_bsdf[p];
    if (sc.prim_type[p] == 1) {
      const double* q = &sc.tri_pos[9 * (size_t)p]; const double* nn = &sc.tri_nrm[9 * (size_t)p];
      r.ax = (float)q[0]; r.ay = (float)q[1]; r.az = (float)q[2]; r.bx = (float)q[3]; r.by = (float)q[4]; r.bz = (float)q[5];
      r.cx = (float)q[6]; r.cy = (float)q[7]; r.cz = (float)q[8]; r.is_tri = 1.0f;
      h.n1x = (float)nn[0]; h.n1y = (float)nn[1]; h.n1z = (float)nn[2]; h.n2x = (float)nn[3]; h.n2y = (float)nn[4]; h.n2z = (float)nn[5];
      h.n3x = (float)nn[6]; h.n3y = (float)nn[7]; h.n3z = (float)nn[8];
      for (int k = 0; k < 9; k++) d.p[k] = q[k];
      d.pad[2] = 1.0;                                // triangle flag
    } else {
      const double* q = &sc.sphere[4 * (size_t)p];
      r.ax = (float)q[0]; r.ay = (float)q[1]; r.az = (float)q[2]; r.bx = (float)q[3]; r.by = (float)(q[3] * q[3]); r.is_tri = 0.0f;
      d.p[0] = q[0]; d.p[1] = q[1]; d.p[2] = q[2]; d.p[3] = q[3]; d.p[4] = q[3] * q[3];   // r2 = r*r, sphere.h:23-24
      d.pad[2] = 0.0;
    }
  }
}

int flatten_lights(int n_lights, const int32_t* light_type, const double* light_param, int ns_area_light, std::vector<Light>& lights) {
  lights.assign((size_t)n_lights, Light{});
  int base = 0;
  for (int i = 0; i < n_lights; i++) {
    Light& L = lights[i]; const double* q = &light_param[28 * (size_t)i];
    std::memset(&L, 0, sizeof(L));
    L.type = light_type[i];
    for (int k = 0; k < 3; k++) { L.radiance[k] = (float)q[k]; L.v0[k] = (float)q[3 + k]; L.dir[k] = (float)q[6 + k]; L.dim_x[k] = (float)q[9 + k]; L.dim_y[k] = (float)q[12 + k]; }
    L.area = (float)q[15];
    for (int k = 0; k < 9; k++) L.s2w[k] = (float)q[16 + k];
    L.is_delta = (L.type == 0 || L.type == 2) ? 1 : 0;     // DirectionalLight / PointLight::is_delta_light, light.h:22,58
    L.n_samples = L.is_delta ? 1 : ns_area_light;          // pathtracer.cpp:474
    L.sample_base = base; base += L.n_samples;
  }
  return base;
}

}  // namespace dsrt
