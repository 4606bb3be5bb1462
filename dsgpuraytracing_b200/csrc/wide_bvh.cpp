// wide_bvh.cpp -- collapse the reference-identical binary SAH tree into the compressed 8-wide BVH the
// sm_100a traversal kernels consume (layout.h), and emit primitives in leaf-contiguous order.
//
//  1. copy the binary tree and refine every leaf down to single primitives with median splits (the reference
//     allows 4 per leaf, and leaves of arbitrary size after a degenerate split, bvh.cpp:143-173); the collapse
//     below then chooses leaf children of 1..3 primitives (the 24-bit primitive part of the hit mask gives each
//     of the 8 children at most 3);
//  2. SAH-optimal collapse (dynamic programme of Ylitie et al. 2017, section 3.1): C(n,i) = cheapest way to turn
//     the binary subtree n into a forest of at most i wide-node children; a subtree with <= 3 primitives may become
//     a leaf child, any subtree may become an internal wide node whose 8 slots are distributed over its two binary
//     children.  This fills the 8 slots far better than greedy largest-area opening (fewer node fetches per ray);
//  3. place children in octant-ordered slots (greedy auction on dot(child centre - node centre, octant
//     direction)) so that slot ^ (7 - ray octant) is a front-to-back visiting priority;
//  4. quantise child boxes to 8 bits per plane, outwards, relative to a float origin rounded down and
//     per-axis power-of-two scales.
// Wide nodes are laid out breadth-first, so the top of the tree is contiguous in memory and the internal
// children of a node are adjacent (one child_base per node).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <future>
#include <unordered_map>

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

// ---- the collapse, as a context object so that the dynamic programme and the emission can run as parallel tasks ----
namespace {

struct Collapse {
  const dsrt_bvh2& b2;
  const std::vector<Box3>& pbox;
  std::vector<BNode> T;
  struct DP { double c[8]; uint8_t split[8]; uint8_t kind1; };      // c[i], i = 1..7 roots; split[i] = roots given to the left child
  std::vector<DP> dp;                                                // kind1: 0 leaf, 1 internal wide node
  std::vector<double> c_int; std::vector<uint8_t> split8;
  static constexpr double c_node = 1.0;
  double c_prim = 1.0;      // cost of one primitive test relative to one node visit (build_wide_bvh argument)
  const std::vector<EndPlane>* end_planes = nullptr;      // axis-aligned area lights (wide_bvh.h)

  Collapse(const dsrt_bvh2& b, const std::vector<Box3>& pb) : b2(b), pbox(pb) {}

  int resolve(int id) const {             // follow single-child chains (bvh.cpp:238-243 does the same at run time)
    while (id >= 0) {
      const BNode& n = T[id];
      if (n.l >= 0 && n.r < 0) id = n.l; else if (n.r >= 0 && n.l < 0) id = n.r; else break;
    }
    return id;
  }

  // 1b. median-split a leaf of r > 1 primitives down to single primitives; its 2(r-1) new nodes go to T[base...]
  void refine_leaf(int id, int base) {
    struct It { int id; };
    int next = base;
    std::vector<int> work; work.push_back(id);
    while (!work.empty()) {
      const int cur = work.back(); work.pop_back();
      const int s = T[cur].start, r = T[cur].range, h = r / 2;
      BNode a, b; a.box.reset(); b.box.reset();
      for (int q = 0; q < r; q++) (q < h ? a.box : b.box).grow(pbox[b2.prim_order[s + q]]);
      a.start = s; a.range = h; a.l = a.r = -1; b.start = s + h; b.range = r - h; b.l = b.r = -1;
      const int ai = next++, bi = next++;
      T[ai] = a; T[bi] = b; T[cur].l = ai; T[cur].r = bi;
      if (a.range > 1) work.push_back(ai);
      if (b.range > 1) work.push_back(bi);
    }
  }

  // 2. one node of the dynamic programme (children already done)
  void dp_node(int n) {
    DP& d = dp[n];
    double area = T[n].box.half_area(); if (!(area >= 0)) area = 0;
    const bool is_leaf = T[n].l < 0;
    const double c_leaf = T[n].range <= 3 ? area * T[n].range * c_prim : std::numeric_limits<double>::infinity();
    if (is_leaf) {
      for (int i = 1; i <= 7; i++) { d.c[i] = c_leaf; d.split[i] = 0; }
      d.kind1 = 0; c_int[n] = std::numeric_limits<double>::infinity();
      return;
    }
    const int L = resolve(T[n].l), R = resolve(T[n].r);
    auto distribute = [&](int j, uint8_t* arg) {                    // best split of j roots between the two children
      double best = std::numeric_limits<double>::infinity(); int bk = 1;
      for (int k = 1; k < j; k++) { double c = dp[L].c[std::min(k, 7)] + dp[R].c[std::min(j - k, 7)]; if (c < best) { best = c; bk = k; } }
      *arg = (uint8_t)bk; return best;
    };
    c_int[n] = distribute(8, &split8[n]) + area * c_node;
    d.c[1] = std::min(c_leaf, c_int[n]); d.kind1 = c_leaf <= c_int[n] ? 0 : 1; d.split[1] = 0;
    for (int i = 2; i <= 7; i++) {
      uint8_t arg = 1; const double cd = distribute(i, &arg);
      if (cd < d.c[i - 1]) { d.c[i] = cd; d.split[i] = arg; } else { d.c[i] = d.c[i - 1]; d.split[i] = 0; }   // 0: use fewer roots
    }
  }
  void dp_sequential(int root) {          // post-order with an explicit stack
    std::vector<int> order, st; st.push_back(root);
    while (!st.empty()) { int n = st.back(); st.pop_back(); order.push_back(n); if (T[n].l >= 0) { st.push_back(resolve(T[n].l)); st.push_back(resolve(T[n].r)); } }
    for (size_t q = order.size(); q-- > 0;) dp_node(order[q]);
  }
  // subtrees with at least `big` primitives fork into tasks; a chain of nodes with one big child is walked iteratively
  void dp_parallel(int n, int big) {
    std::vector<int> chain;
    while (true) {
      if (T[n].l < 0 || T[n].range < big) { dp_sequential(n); break; }
      const int L = resolve(T[n].l), R = resolve(T[n].r);
      const bool bl = T[L].range >= big, br = T[R].range >= big;
      if (bl && br) {
        std::future<void> f = std::async(std::launch::async, [this, L, big] { dp_parallel(L, big); });
        dp_parallel(R, big);
        f.get();
        dp_node(n);
        break;
      }
      if (!bl && !br) { dp_sequential(n); break; }
      dp_sequential(bl ? R : L);
      chain.push_back(n);
      n = bl ? L : R;
    }
    for (size_t q = chain.size(); q-- > 0;) dp_node(chain[q]);
  }

  // children of the wide node rooted at binary node n = leaves of the DP's distribution of 8 slots
  struct Kid { int bnode; bool leaf; };
  void collect(int n, std::vector<Kid>& kids) const {
    kids.clear();
    if (!regrouped.empty()) { const auto it = regrouped.find(n); if (it != regrouped.end()) { kids = it->second; return; } }
    struct Item { int n, i; };
    Item st[64]; int sp = 0;
    if (T[n].l < 0) { kids.push_back({n, true}); return; }           // the root itself is a (<= 3 primitive) leaf
    const int L = resolve(T[n].l), R = resolve(T[n].r);
    st[sp++] = {R, 8 - split8[n]}; st[sp++] = {L, split8[n]};
    while (sp > 0) {
      Item it = st[--sp];
      int i = std::min(it.i, 7);
      while (i > 1 && dp[it.n].split[i] == 0) i--;                    // the DP preferred fewer roots
      if (i == 1) { kids.push_back({it.n, dp[it.n].kind1 == 0}); continue; }
      const int l2 = resolve(T[it.n].l), r2 = resolve(T[it.n].r);
      const int k = dp[it.n].split[i];
      st[sp++] = {r2, i - k}; st[sp++] = {l2, k};
    }
  }

  // 2b. regrouping at the top of the tree.  The dynamic programme can only cut the binary tree it is given, and the reference's
  // binned SAH leaves scene-sized triangles (the Cornell walls) in a subtree whose box is the whole scene: that subtree becomes
  // an internal child every ray has to open.  For a wide node N with children K and an internal child C that covers a large
  // part of N, the SAH cost of any arrangement of the items I = (K \ C) + children(C) into direct children and new internal
  // nodes G_j differs only in sum_j area(G_j) (each item's own term area(item) * cost(item) appears in every arrangement): the
  // current arrangement pays area(C).  Greedy agglomeration (cheapest increase of the summed group area first, <= 8 members per
  // group) until <= 8 slots are left; taken when the groups' summed area is < 0.9 area(C).  Groups become synthetic nodes of T
  // whose children are listed explicitly (`regrouped`); only nodes whose box is >= 1/64 of the root's area are looked at.
  std::unordered_map<int, std::vector<Kid>> regrouped;
  int n_regrouped = 0;
  void regroup_top(int root) {
    const double root_area = T[root].box.half_area();
    if (!(root_area > 0)) return;
    std::deque<int> q; q.push_back(root);
    int budget = 4096;
    std::vector<Kid> kids, sub;
    struct Grp { Box3 box; std::vector<Kid> members; int range; };
    auto area = [](const Box3& b) { const double a = b.half_area(); return a >= 0 ? a : 0.0; };
    while (!q.empty() && budget-- > 0) {
      const int n = q.front(); q.pop_front();
      if (area(T[n].box) < root_area / 64) continue;
      for (int round = 0; round < 8; round++) {
        collect(n, kids);
        int ci = -1; double ca = 0.3 * area(T[n].box);
        for (size_t i = 0; i < kids.size(); i++) if (!kids[i].leaf && area(T[kids[i].bnode].box) > ca) { ca = area(T[kids[i].bnode].box); ci = (int)i; }
        if (ci < 0) break;
        collect(kids[ci].bnode, sub);
        std::vector<Grp> g;
        for (size_t i = 0; i < kids.size(); i++) if ((int)i != ci) g.push_back({T[kids[i].bnode].box, {kids[i]}, T[kids[i].bnode].range});
        for (const Kid& k : sub) g.push_back({T[k.bnode].box, {k}, T[k.bnode].range});
        bool ok = true;
        while (g.size() > 8 && ok) {
          int bi = -1, bj = -1; double bd = std::numeric_limits<double>::infinity();
          for (size_t i = 0; i < g.size(); i++) for (size_t j = i + 1; j < g.size(); j++) {
            if (g[i].members.size() + g[j].members.size() > 8) continue;
            Box3 u = g[i].box; u.grow(g[j].box);
            const double d = area(u) - (g[i].members.size() > 1 ? area(g[i].box) : 0.0) - (g[j].members.size() > 1 ? area(g[j].box) : 0.0);
            if (d < bd) { bd = d; bi = (int)i; bj = (int)j; }
          }
          if (bi < 0) { ok = false; break; }
          g[bi].box.grow(g[bj].box); g[bi].range += g[bj].range;
          g[bi].members.insert(g[bi].members.end(), g[bj].members.begin(), g[bj].members.end());
          g.erase(g.begin() + bj);
        }
        if (!ok) break;
        double cost = 0; for (const Grp& x : g) if (x.members.size() > 1) cost += area(x.box);
        if (!(cost < 0.9 * ca)) break;
        std::vector<Kid> nk;
        for (const Grp& x : g) {
          if (x.members.size() == 1) { nk.push_back(x.members[0]); continue; }
          BNode b; b.box = x.box; b.start = -1; b.range = x.range; b.l = b.r = -1;      // children: regrouped[id], never l / r
          const int id = (int)T.size(); T.push_back(b);
          regrouped[id] = x.members;
          nk.push_back({id, false});
        }
        regrouped[n] = nk; n_regrouped++;
      }
      collect(n, kids);
      for (const Kid& k : kids) if (!k.leaf) q.push_back(k.bnode);
    }
  }

  // is one of these children a flat box (no thickness along `axis`) in the plane of an area light, overlapping its rectangle?
  bool flat_child_on_light(const int* kids, int nk, int axis, double& at, bool& from_low) const {
    for (const EndPlane& P : *end_planes) {
      if (P.axis != axis) continue;
      for (int i = 0; i < nk; i++) {
        const Box3& b = T[kids[i]].box;
        // "flat": the reference's own Cornell quads come out of their COLLADA transform 1e-7 thick
        const double tol = 1e-6 * std::max(1.0, std::fabs(P.coord));
        if (std::fabs(b.lo[axis] - P.coord) > tol || std::fabs(b.hi[axis] - P.coord) > tol) continue;
        bool overlap = true;      // (the reference's Cornell quads are 0.8 x 0.6 under a 0.6 x 0.8 light: overlap, not containment)
        for (int k = 0; k < 3; k++) if (k != axis && (b.hi[k] < P.lo[k] || b.lo[k] > P.hi[k])) overlap = false;
        if (overlap) { at = P.from_low ? b.lo[axis] : b.hi[axis]; from_low = P.from_low; return true; }
      }
    }
    return false;
  }

  // 3..4. one wide node: octant-ordered slots + quantisation.  Its internal children get the node indices
  // child_base, child_base+1, ... (in slot order; their binary nodes are returned in kids_out), the primitives of its
  // leaf children are appended to slot_prim.  Returns the number of internal children or a negative error code.
  int emit_node(int bnode, uint32_t child_base, std::vector<int32_t>& slot_prim, WideNode& w, int* kids_out, std::vector<Kid>& kidv, std::string& err) const {
    collect(bnode, kidv);
    if (kidv.size() > 8) { err = "wide BVH: collapse produced more than 8 children"; return -1; }
    int kids[8]; bool kid_leaf[8]; int nk = 0;
    for (const Kid& k : kidv) { kids[nk] = k.bnode; kid_leaf[nk] = k.leaf; nk++; }
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
    int kid_at[8]; bool leaf_at[8]; for (int s = 0; s < 8; s++) { kid_at[s] = -1; leaf_at[s] = false; }
    for (int i = 0; i < nk; i++) { kid_at[slot_of[i]] = kids[i]; leaf_at[slot_of[i]] = kid_leaf[i]; }

    // 4. quantisation frame.  The device evaluates plane = (1 + q 2^-15) * (2^15 s) + (b - 2^15 s) in float, which is
    // off by up to ~2^-9 of a quantum, and picks near/far planes separately; every child plane therefore gets a
    // guaranteed margin of 1/64 quantum (and a flat box at least one full quantum of thickness): the grid keeps one
    // spare quantum below the node's box, and its pitch never drops below 2^-18 of the coordinate magnitude.
    std::memset(&w, 0, sizeof(w));
    float org[3]; double scale[3]; uint8_t ebits[3];
    for (int k = 0; k < 3; k++) {
      const double ext = nb.hi[k] - nb.lo[k];
      const double mag = std::max(std::max(std::fabs(nb.lo[k]), std::fabs(nb.hi[k])), 1e-30);
      int e = ext > 0 ? (int)std::ceil(std::log2(ext / 252.0)) : -126;
      e = std::max(e, std::ilogb(mag) - 18);
      e = std::min(110, std::max(-120, e));
      while (true) {
        scale[k] = std::ldexp(1.0, e);
        org[k] = round_down(nb.lo[k] - scale[k]);
        if (std::ceil((nb.hi[k] - (double)org[k]) / scale[k] + 1.0 / 64) <= 255.0 || e >= 110) break;
        e++;
      }
      // a child that lies flat in the plane of an area light (EndPlane, wide_bvh.h): lower the origin by a fraction of a
      // quantum so that the plane sits 3/128 quantum past a grid line on the side the shadow rays arrive from (the margin
      // every plane gets is 1/64 = 2/128).  Taken only if the rounded float origin still puts it there and the box still fits.
      double flat_at; bool from_low;
      if (end_planes && flat_child_on_light(kids, nk, k, flat_at, from_low)) {
        const double want = from_low ? 3.0 / 128 : 1.0 - 3.0 / 128;
        const double x0 = (flat_at - (nb.lo[k] - scale[k])) / scale[k];
        double d = want - (x0 - std::floor(x0));
        if (d < 0) d += 1.0;
        const float org_a = round_down(nb.lo[k] - scale[k] * (1.0 + d));
        const double xa = (flat_at - (double)org_a) / scale[k], fa = xa - std::floor(xa);
        if (std::fabs(fa - want) <= 1.0 / 256 && (double)org_a <= nb.lo[k] - scale[k] &&
            std::ceil((nb.hi[k] - (double)org_a) / scale[k] + 1.0 / 64) <= 255.0) org[k] = org_a;
      }
      ebits[k] = (uint8_t)(e + 127 + 15);     // stored with the node test's 2^15 folded in (layout.h)
    }
    w.ox = org[0]; w.oy = org[1]; w.oz = org[2]; w.ex = ebits[0]; w.ey = ebits[1]; w.ez = ebits[2];
#if DSRT_NODE96
    { const uint32_t bx = (uint32_t)ebits[0] << 23, by = (uint32_t)ebits[1] << 23, bz = (uint32_t)ebits[2] << 23;
      std::memcpy(&w.sx, &bx, 4); std::memcpy(&w.sy, &by, 4); std::memcpy(&w.sz, &bz, 4); }
#endif
    w.prim_base = (uint32_t)slot_prim.size();
    w.child_base = child_base;
    int n_internal = 0, prim_off = 0;
    for (int s = 0; s < 8; s++) {
      int c = kid_at[s];
      if (c < 0) continue;
      const BNode& cn = T[c];
      uint8_t q[6];
      for (int k = 0; k < 3; k++) {
        double lo = std::floor((cn.box.lo[k] - (double)org[k]) / scale[k] - 1.0 / 64);
        double hi = std::ceil((cn.box.hi[k] - (double)org[k]) / scale[k] + 1.0 / 64);
        q[k] = (uint8_t)std::min(255.0, std::max(0.0, lo));
        q[3 + k] = (uint8_t)std::min(255.0, std::max(0.0, hi));
      }
      w.qlox[s] = q[0]; w.qloy[s] = q[1]; w.qloz[s] = q[2]; w.qhix[s] = q[3]; w.qhiy[s] = q[4]; w.qhiz[s] = q[5];
      if (leaf_at[s]) {
        if (cn.range < 1 || cn.range > 3 || prim_off + cn.range > 24) { err = "wide BVH: leaf packing overflow"; return -1; }
        w.valid |= ((1u << cn.range) - 1u) << (4 * s);
        for (int q2 = 0; q2 < cn.range; q2++) slot_prim.push_back(b2.prim_order[cn.start + q2]);
        prim_off += cn.range;
      } else {
        w.inner |= 8u << (4 * s);
        w.imask |= (uint8_t)(1 << s);
        kids_out[n_internal++] = c;
      }
    }
    return n_internal;
  }
};

}  // namespace

// area lights (type 3) whose direction is a coordinate axis and whose edges span the other two: light.cpp:71-92
std::vector<EndPlane> light_end_planes(int n_lights, const int32_t* light_type, const double* light_param) {
  std::vector<EndPlane> out;
  for (int i = 0; i < n_lights; i++) {
    if (light_type[i] != 3) continue;
    const double* p = light_param + 28 * (size_t)i;
    const double *pos = p + 3, *dir = p + 6, *dx = p + 9, *dy = p + 12;
    int axis = -1;
    for (int k = 0; k < 3; k++) if (std::fabs(dir[k]) > 0 && dir[(k + 1) % 3] == 0 && dir[(k + 2) % 3] == 0) axis = k;
    if (axis < 0 || dx[axis] != 0 || dy[axis] != 0) continue;
    EndPlane e; e.axis = axis; e.from_low = dir[axis] < 0; e.coord = pos[axis];
    for (int k = 0; k < 3; k++) { const double h = 0.5 * (std::fabs(dx[k]) + std::fabs(dy[k])); e.lo[k] = pos[k] - h; e.hi[k] = pos[k] + h; }
    if (!(std::isfinite(e.coord) && std::isfinite(e.lo[0] + e.lo[1] + e.lo[2] + e.hi[0] + e.hi[1] + e.hi[2]))) continue;
    out.push_back(e);
  }
  return out;
}

int build_wide_bvh(const dsrt_bvh2& b2, const std::vector<Box3>& pbox, int n_prims, WideBVH& out, std::string& err, double prim_cost,
                   const std::vector<EndPlane>* end_planes, bool regroup_top) {
  out.nodes.clear(); out.slot_prim.clear(); out.max_depth = 0;
  if (b2.n_nodes <= 0 || n_prims <= 0) {
    // empty scene: a single node with no children
    WideNode w; std::memset(&w, 0, sizeof(w)); w.ex = w.ey = w.ez = 1;
#if DSRT_NODE96
    { const uint32_t b = 1u << 23; std::memcpy(&w.sx, &b, 4); std::memcpy(&w.sy, &b, 4); std::memcpy(&w.sz, &b, 4); }
#endif
    out.nodes.push_back(w); out.max_depth = 1;
    return DSRT_OK;
  }
  const bool timing = std::getenv("DSRT_BUILD_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  Collapse C(b2, pbox);
  C.c_prim = prim_cost;
  C.end_planes = end_planes && !end_planes->empty() ? end_planes : nullptr;
  std::vector<BNode>& T = C.T;
  // 1. working copy; leaves of more than one primitive are refined (each leaf's new nodes live in its own block of T)
  T.resize((size_t)b2.n_nodes);
  std::atomic<int> bad{-1};
  parallel_for((size_t)b2.n_nodes, (size_t)1 << 16, [&](size_t i0, size_t i1) {
    for (size_t i = i0; i < i1; i++) {
      BNode& n = T[i];
      for (int k = 0; k < 3; k++) { n.box.lo[k] = b2.node_bbox[6 * i + k]; n.box.hi[k] = b2.node_bbox[6 * i + 3 + k]; }
      n.start = b2.node_start[i]; n.range = b2.node_range[i]; n.l = b2.node_left[i]; n.r = b2.node_right[i];
      if (n.start < 0 || n.range < 0 || (int64_t)n.start + (int64_t)n.range > (int64_t)n_prims || n.l >= b2.n_nodes || n.r >= b2.n_nodes) bad.store((int)i);
      // a child always has a larger index than its parent (pre-order numbering of the builder and of the reference's dump):
      // no node can then be its own ancestor, so every walk below terminates; children split the parent's range in two
      else if (n.l >= 0 || n.r >= 0) {
        if (n.l <= (int)i || n.r <= (int)i) bad.store((int)i);
        else {
          const int64_t ls = b2.node_start[n.l], lr = b2.node_range[n.l], rs = b2.node_start[n.r], rr = b2.node_range[n.r];
          if (ls != n.start || lr < 0 || rr < 0 || rs != ls + lr || lr + rr != n.range) bad.store((int)i);
        }
      }
    }
  });
  if (bad.load() >= 0) { err = "dsrt_set_bvh: node " + std::to_string(bad.load()) + " is out of range, is not numbered after its parent, or its children do not partition its primitives"; return DSRT_ERR_INVALID; }
  {
    std::vector<int> leaves; std::vector<size_t> base;
    size_t total = T.size();
    for (int i = 0; i < b2.n_nodes; i++) if (T[i].l < 0 && T[i].r < 0 && T[i].range > 1) { leaves.push_back(i); base.push_back(total); total += 2 * (size_t)(T[i].range - 1); }
    if (total > (size_t)std::numeric_limits<int>::max()) { err = "wide BVH: too many binary nodes"; return DSRT_ERR_LIMIT; }
    T.resize(total);
    parallel_for(leaves.size(), 4096, [&](size_t i0, size_t i1) { for (size_t i = i0; i < i1; i++) C.refine_leaf(leaves[i], (int)base[i]); });
  }
  const int root = C.resolve(0);

  const double t1 = now();
  // 2. dynamic programme over the binary tree (post-order; big subtrees as parallel tasks)
  const size_t NT = T.size();
  C.dp.resize(NT); C.c_int.assign(NT, 0.0); C.split8.assign(NT, 0);
  C.dp_parallel(root, std::max(1 << 14, n_prims / (4 * host_threads())));

  if (regroup_top) C.regroup_top(root);
  const double t2 = now();
  // 3..4. emission.  The top of the tree is emitted breadth-first by one thread until kTopJobs subtrees are pending;
  // every pending subtree is then emitted (breadth-first within itself) as an independent task into its own block and
  // the blocks are concatenated in job order -- the layout depends on kTopJobs only, never on the thread count.
  // Internal children of a node are adjacent (one child_base per node) in both parts.
  constexpr size_t kTopJobs = 1024;
  struct Job { int bnode; uint32_t widx; int depth; };
  std::deque<Job> jobs;
  out.nodes.emplace_back();
  jobs.push_back({root, 0u, 1});
  {
    std::vector<Collapse::Kid> kidv; int kids[8];
    while (!jobs.empty() && jobs.size() < kTopJobs) {
      Job job = jobs.front(); jobs.pop_front();
      out.max_depth = std::max(out.max_depth, job.depth);
      WideNode w;
      const int ni = C.emit_node(job.bnode, (uint32_t)out.nodes.size(), out.slot_prim, w, kids, kidv, err);
      if (ni < 0) return DSRT_ERR_LIMIT;
      for (int i = 0; i < ni; i++) jobs.push_back({kids[i], (uint32_t)(out.nodes.size() + i), job.depth + 1});
      out.nodes.resize(out.nodes.size() + ni);
      out.nodes[job.widx] = w;
    }
  }
  if (!jobs.empty()) {
    struct Block { WideNode root; std::vector<WideNode> nodes; std::vector<int32_t> prims; int depth = 0; int rc = 0; std::string err; };
    std::vector<Job> roots(jobs.begin(), jobs.end());
    std::vector<Block> blocks(roots.size());
    parallel_for(roots.size(), 1, [&](size_t j0, size_t j1) {
      std::vector<Collapse::Kid> kidv; int kids[8];
      for (size_t j = j0; j < j1; j++) {
        Block& B = blocks[j];
        // local indices: the subtree's descendants are B.nodes[0..]; index -1 = the subtree root (lives in the top part)
        struct LJob { int bnode; int lidx; int depth; };
        std::deque<LJob> q; q.push_back({roots[j].bnode, -1, roots[j].depth});
        while (!q.empty()) {
          LJob job = q.front(); q.pop_front();
          B.depth = std::max(B.depth, job.depth);
          WideNode w;
          const int ni = C.emit_node(job.bnode, (uint32_t)B.nodes.size(), B.prims, w, kids, kidv, B.err);
          if (ni < 0) { B.rc = DSRT_ERR_LIMIT; break; }
          for (int i = 0; i < ni; i++) q.push_back({kids[i], (int)(B.nodes.size() + i), job.depth + 1});
          B.nodes.resize(B.nodes.size() + ni);
          if (job.lidx < 0) B.root = w; else B.nodes[job.lidx] = w;
        }
      }
    });
    size_t node_off = out.nodes.size(), prim_off = out.slot_prim.size();
    std::vector<size_t> noff(blocks.size()), poff(blocks.size());
    for (size_t j = 0; j < blocks.size(); j++) {
      if (blocks[j].rc) { err = blocks[j].err; return blocks[j].rc; }
      noff[j] = node_off; poff[j] = prim_off; node_off += blocks[j].nodes.size(); prim_off += blocks[j].prims.size();
      out.max_depth = std::max(out.max_depth, blocks[j].depth);
    }
    if (node_off > 0xffffffffull || prim_off > 0x7fffffffull) { err = "wide BVH: too many nodes"; return DSRT_ERR_LIMIT; }
    out.nodes.resize(node_off); out.slot_prim.resize(prim_off);
    parallel_for(blocks.size(), 1, [&](size_t j0, size_t j1) {
      for (size_t j = j0; j < j1; j++) {
        Block& B = blocks[j];
        auto fix = [&](WideNode w) { w.child_base += (uint32_t)noff[j]; w.prim_base += (uint32_t)poff[j]; return w; };
        out.nodes[roots[j].widx] = fix(B.root);
        for (size_t i = 0; i < B.nodes.size(); i++) out.nodes[noff[j] + i] = fix(B.nodes[i]);
        std::copy(B.prims.begin(), B.prims.end(), out.slot_prim.begin() + (std::ptrdiff_t)poff[j]);
        std::vector<WideNode>().swap(B.nodes); std::vector<int32_t>().swap(B.prims);
      }
    });
  }
  if (timing) std::fprintf(stderr, "build_wide_bvh: %d prims, copy+refine %.2f s, collapse DP %.2f s, emit %.2f s\n", n_prims, t1 - t0, t2 - t1, now() - t2);
  if ((int)out.slot_prim.size() != n_prims) { err = "wide BVH: primitive count mismatch (" + std::to_string(out.slot_prim.size()) + " vs " + std::to_string(n_prims) + ")"; return DSRT_ERR_INVALID; }
  return DSRT_OK;
}

// Sets WideNode::flat (layout.h): leaf slots whose two or three primitives are triangles in one plane.  "In one plane": every
// vertex within 1e-6 of the slot's extent of the plane of the slot's largest triangle (the reference's own wall quads come out
// of their COLLADA transforms 1e-7 off).  A ray leaving one of them then skips its slot mates (drop_source, traverse.cuh).
void mark_flat_slots(const dsrt_scene& sc, WideBVH& wide) {
#if DSRT_NODE96
  parallel_for(wide.nodes.size(), (size_t)1 << 14, [&](size_t n0, size_t n1) {
    for (size_t n = n0; n < n1; n++) {
      WideNode& w = wide.nodes[n];
      w.flat = 0;
      uint32_t rank = 0;
      for (int s = 0; s < 8; s++) {
        const int c = __builtin_popcount((w.valid >> (4 * s)) & 0xfu);
        const uint32_t first = w.prim_base + rank; rank += (uint32_t)c;
        if (c < 2) continue;
        const double* best = nullptr; double best_n2 = 0, nx = 0, ny = 0, nz = 0, ext = 0; bool tris = true;
        for (int i = 0; i < c; i++) {
          const int p = wide.slot_prim[first + i];
          if (sc.prim_type[p] != 1) { tris = false; break; }
          const double* v = sc.tri_pos + 9 * (size_t)p;
          const double ax = v[3] - v[0], ay = v[4] - v[1], az = v[5] - v[2], bx = v[6] - v[0], by = v[7] - v[1], bz = v[8] - v[2];
          const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx, n2 = cx * cx + cy * cy + cz * cz;
          ext = std::max(ext, std::sqrt(std::max(ax * ax + ay * ay + az * az, bx * bx + by * by + bz * bz)));
          if (n2 > best_n2) { best_n2 = n2; best = v; nx = cx; ny = cy; nz = cz; }
        }
        if (!tris || !best || !(best_n2 > 0) || !std::isfinite(best_n2)) continue;
        const double inv = 1.0 / std::sqrt(best_n2);
        double dev = 0;
        for (int i = 0; i < c; i++) {
          const double* v = sc.tri_pos + 9 * (size_t)wide.slot_prim[first + i];
          for (int k = 0; k < 3; k++) {
            dev = std::max(dev, std::fabs(((v[3 * k] - best[0]) * nx + (v[3 * k + 1] - best[1]) * ny + (v[3 * k + 2] - best[2]) * nz) * inv));
            ext = std::max(ext, std::sqrt((v[3 * k] - best[0]) * (v[3 * k] - best[0]) + (v[3 * k + 1] - best[1]) * (v[3 * k + 1] - best[1]) + (v[3 * k + 2] - best[2]) * (v[3 * k + 2] - best[2])));
          }
        }
        if (dev <= 1e-6 * ext) w.flat |= 0xfu << (4 * s);
      }
    }
  });
#else
  (void)sc; (void)wide;
#endif
}

void flatten_records(const dsrt_scene& sc, const WideBVH& wide, std::vector<PrimRecord>& recs, std::vector<ShadeRecord>& shd) {
  const size_t n = wide.slot_prim.size();
  recs.resize(n ? n : 1); shd.resize(n ? n : 1);
  if (!n) { std::memset(&recs[0], 0, sizeof(PrimRecord)); std::memset(&shd[0], 0, sizeof(ShadeRecord)); }
  parallel_for(n, (size_t)1 << 16, [&](size_t s0, size_t s1) {
    for (size_t sl = s0; sl < s1; sl++) {
      const int p = wide.slot_prim[sl];
      PrimRecord& r = recs[sl]; ShadeRecord& h = shd[sl];
      std::memset(&r, 0, sizeof(r)); std::memset(&h, 0, sizeof(h));
      r.prim_id = p; r.bsdf = sc.prim_bsdf[p];
      if (sc.prim_type[p] == 1) {
        const double* q = &sc.tri_pos[9 * (size_t)p]; const double* nn = &sc.tri_nrm[9 * (size_t)p];
        r.ax = (float)q[0]; r.ay = (float)q[1]; r.az = (float)q[2]; r.bx = (float)q[3]; r.by = (float)q[4]; r.bz = (float)q[5];
        r.cx = (float)q[6]; r.cy = (float)q[7]; r.cz = (float)q[8]; r.is_tri = 1.0f;
        h.n1x = (float)nn[0]; h.n1y = (float)nn[1]; h.n1z = (float)nn[2]; h.n2x = (float)nn[3]; h.n2y = (float)nn[4]; h.n2z = (float)nn[5];
        h.n3x = (float)nn[6]; h.n3y = (float)nn[7]; h.n3z = (float)nn[8];
      } else {
        const double* q = &sc.sphere[4 * (size_t)p];
        r.ax = (float)q[0]; r.ay = (float)q[1]; r.az = (float)q[2]; r.bx = (float)q[3]; r.by = (float)(q[3] * q[3]); r.is_tri = 0.0f;
      }
    }
  });
}

void flatten_records64(const dsrt_scene& sc, const WideBVH& wide, std::vector<PrimRecord64>& r64) {
  const size_t n = wide.slot_prim.size();
  r64.resize(n ? n : 1);
  if (!n) std::memset(&r64[0], 0, sizeof(PrimRecord64));
  parallel_for(n, (size_t)1 << 16, [&](size_t s0, size_t s1) {
    for (size_t sl = s0; sl < s1; sl++) {
      const int p = wide.slot_prim[sl];
      PrimRecord64& d = r64[sl];
      std::memset(&d, 0, sizeof(d));
      if (sc.prim_type[p] == 1) {
        const double* q = &sc.tri_pos[9 * (size_t)p];
        for (int k = 0; k < 9; k++) d.p[k] = q[k];
        d.pad[2] = 1.0;                                // triangle flag
      } else {
        const double* q = &sc.sphere[4 * (size_t)p];
        d.p[0] = q[0]; d.p[1] = q[1]; d.p[2] = q[2]; d.p[3] = q[3]; d.p[4] = q[3] * q[3];   // r2 = r*r, sphere.h:23-24
        d.pad[2] = 0.0;
      }
    }
  });
}

int flatten_lights(int n_lights, const int32_t* light_type, const double* light_param, int ns_area_light, bool with_env,
                   std::vector<Light>& lights) {
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
  if (with_env) {   // PathTracer::set_scene appends the environment light after the scene's lights (pathtracer.cpp:88-90)
    Light L; std::memset(&L, 0, sizeof(L));
    L.type = 4; L.is_delta = 0; L.n_samples = ns_area_light; L.sample_base = base; base += L.n_samples;
    lights.push_back(L);
  }
  return base;
}

void build_env_tables(int w, int h, const float* rgb, std::vector<float>& tp, std::vector<float>& t, std::vector<float>& pgt) {
  const double PI = 3.14159265358979323;
  tp.assign((size_t)w * h, 0.f); t.assign((size_t)h, 0.f); pgt.assign((size_t)w * h, 0.f);
  float C = 0;
  for (int y = 0; y < h; y++) {
    const float theta = (float)((y + 0.5) / h * PI);
    const float sin_theta = (float)std::sin((double)theta);
    for (int x = 0; x < w; x++) {
      const float* q = rgb + 3 * ((size_t)x + (size_t)w * y);
      const float illum = 0.2126f * q[0] + 0.7152f * q[1] + 0.0722f * q[2];
      tp[(size_t)y * w + x] = illum * sin_theta;
      C += tp[(size_t)y * w + x];
    }
  }
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) { tp[(size_t)y * w + x] /= C; t[y] += tp[(size_t)y * w + x]; }
    if (t[y] != 0) for (int x = 0; x < w; x++) pgt[(size_t)y * w + x] = tp[(size_t)y * w + x] / t[y];
  }
  for (int y = 0; y < h; y++) {
    if (y > 0) t[y] += t[y - 1];
    for (int x = 1; x < w; x++) pgt[(size_t)y * w + x] += pgt[(size_t)y * w + x - 1];
  }
}

}  // namespace dsrt
