// bvh2_sah.cpp -- host builder for the binary SAH BVH, topology-identical to the reference's
// BVHAccel (src/bvh.cpp:21-202): 32 centroid buckets per axis, cost = halfArea(L)*nL + halfArea(R)*nR over
// the 31 bucket boundaries, best axis = smallest minimum, Hoare partition about the chosen plane with strict
// comparisons, children created only when both sides are non-empty, recursion while a side has more than
// max_leaf (4) primitives.  The bucket index is clamped to [0, 31]: the reference indexes one past the
// bucket array when a centroid lies on the node's upper bound (bvh.cpp:47-50).
//
// Node ids are assigned in preorder (node, left subtree, right subtree); node 0 is the root.
// Subtrees with many primitives are built as parallel tasks (disjoint slices of the order array, node slots from an
// atomic counter); the final preorder renumbering makes the result independent of the schedule, so the topology is
// the reference's whatever the thread count.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <future>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/dsrt.h"
#include "host_util.h"

namespace dsrt {

// Triangle::get_bbox (triangle.cpp:11-23) / Sphere::get_bbox (sphere.h:30-32)
void primitive_boxes(const dsrt_scene* sc, std::vector<Box3>& out) {
  out.resize(sc->n_prims);
  parallel_for((size_t)sc->n_prims, (size_t)1 << 16, [&](size_t i0, size_t i1) {
    for (size_t i = i0; i < i1; i++) {
      Box3& b = out[i];
      if (sc->prim_type[i] == 1) {
        b.reset();
        const double* p = sc->tri_pos + 9 * i;
        b.grow(p); b.grow(p + 3); b.grow(p + 6);
      } else {
        const double* s = sc->sphere + 4 * i;
        for (int k = 0; k < 3; k++) { b.lo[k] = s[k] - s[3]; b.hi[k] = s[k] + s[3]; }
      }
    }
  });
}

namespace {

struct BuildNode { Box3 box; int start, range, left, right; };

struct Builder {
#ifndef DSRT_SAH_BUCKETS
#define DSRT_SAH_BUCKETS 32          // the reference's bucket count (bvh.cpp:21-202); other values are for experiments only (the topology then differs)
#endif
  static constexpr int kBuckets = DSRT_SAH_BUCKETS;
  static constexpr int kMaxLeaf = 4;
  static constexpr int kParallelMin = 1 << 16;     // spawn a task for subtrees at least this large
  static constexpr int kParallelBin = 1 << 21;     // bin nodes at least this large with all threads
  // primitive boxes travel with the order array (one 56-byte item per slot), so binning and partitioning stream through
  // memory instead of gathering boxes through the index array; the swaps are exactly the reference's swaps of `primitives`
  struct Item { Box3 box; int32_t id; };
  std::vector<Item> items;
  std::vector<BuildNode> nodes;                   // preallocated: 2 * n_prims + 1 slots
  std::atomic<int> next_node{0};
  std::atomic<int> spare_threads{0};

  explicit Builder(const std::vector<Box3>& pb) : items(pb.size()) {
    parallel_for(pb.size(), (size_t)1 << 16, [&](size_t i0, size_t i1) { for (size_t i = i0; i < i1; i++) { items[i].box = pb[i]; items[i].id = (int32_t)i; } });
  }

  int alloc2() { return next_node.fetch_add(2); }

  double centroid(int slot, int axis) const {
    const Box3& b = items[slot].box;
    return (b.lo[axis] + b.hi[axis]) * 0.5;
  }

  // returns the bucket boundary (1..31) with the lowest cost on this axis, cost in *best
  int best_split_on_axis(const BuildNode& n, int axis, double* best) const {
    const double lb = n.box.lo[axis], ub = n.box.hi[axis];
    const double interval = (ub - lb) / kBuckets;
    Box3 bbox[kBuckets]; int cnt[kBuckets];
    for (int i = 0; i < kBuckets; i++) { bbox[i].reset(); cnt[i] = 0; }
    auto bin = [&](size_t i0, size_t i1, Box3* bb, int* cc) {
      for (size_t i = i0; i < i1; i++) {
        int b = (int)((centroid(n.start + (int)i, axis) - lb) / interval);
        b = b > kBuckets - 1 ? kBuckets - 1 : (b < 0 ? 0 : b);
        bb[b].grow(items[n.start + i].box);
        cc[b]++;
      }
    };
    if (n.range >= kParallelBin) {
      // the top of the tree: bin slices in parallel and merge (min / max / integer sums: the result does not depend
      // on the slicing)
      std::mutex mu;
      parallel_for((size_t)n.range, (size_t)1 << 18, [&](size_t i0, size_t i1) {
        Box3 lb2[kBuckets]; int lc[kBuckets];
        for (int i = 0; i < kBuckets; i++) { lb2[i].reset(); lc[i] = 0; }
        bin(i0, i1, lb2, lc);
        std::lock_guard<std::mutex> g(mu);
        for (int i = 0; i < kBuckets; i++) { bbox[i].grow(lb2[i]); cnt[i] += lc[i]; }
      });
    } else {
      bin(0, (size_t)n.range, bbox, cnt);
    }
    // suffix unions (the reference's reversed-bucket array, bvh.cpp:54-60) and prefix unions (:64-67)
    Box3 suf[kBuckets]; int sufc[kBuckets];
    for (int i = 0; i < kBuckets; i++) {
      suf[i] = bbox[kBuckets - 1 - i]; sufc[i] = cnt[kBuckets - 1 - i];
      if (i > 0) { suf[i].grow(suf[i - 1]); sufc[i] += sufc[i - 1]; }
    }
    for (int i = 1; i < kBuckets; i++) { bbox[i].grow(bbox[i - 1]); cnt[i] += cnt[i - 1]; }
    int arg = 0;
    for (int i = 0; i < kBuckets - 1; i++) {
      double c = bbox[i].half_area() * cnt[i] + suf[kBuckets - i - 2].half_area() * sufc[kBuckets - i - 2];
      if (c < *best) { *best = c; arg = i + 1; }
    }
    return arg;
  }

  void split(int id) {
    const BuildNode n = nodes[id];
    double best[3]; int plane[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) {
      best[k] = std::numeric_limits<double>::infinity();
      if (n.box.hi[k] == n.box.lo[k]) continue;
      plane[k] = best_split_on_axis(n, k, &best[k]);
    }
    int axis = 0;
    for (int k = 1; k < 3; k++) if (best[k] < best[axis]) axis = k;
    const double lb = n.box.lo[axis], ub = n.box.hi[axis];
    const double cut = lb + (ub - lb) * plane[axis] / kBuckets;
    // Hoare-style partition, bvh.cpp:97-125 (elements equal to the cut stop both scans and are swapped)
    int i = n.start - 1, j = n.start + n.range;
    const int end = n.start + n.range;
    while (i < j) {
      while (true) { i++; if (i >= end) break; if (!(centroid(i, axis) < cut)) break; }
      while (true) { j--; if (j < n.start) break; if (!(centroid(j, axis) > cut)) break; }
      if (i < j) std::swap(items[i], items[j]);
      else break;
    }
    const int nl = i - n.start, nr = n.range - nl;
    int li = -1, ri = -1;
    if (nl != 0 && nr != 0) {
      li = alloc2(); ri = li + 1;
      BuildNode& L = nodes[li]; BuildNode& R = nodes[ri];
      L.box.reset(); R.box.reset();
      if (n.range >= kParallelBin) {
        std::mutex mu;
        parallel_for((size_t)n.range, (size_t)1 << 18, [&](size_t q0, size_t q1) {
          Box3 a, b; a.reset(); b.reset();
          for (size_t q = q0; q < q1; q++) ((int)q < nl ? a : b).grow(items[n.start + q].box);
          std::lock_guard<std::mutex> g(mu);
          L.box.grow(a); R.box.grow(b);
        });
      } else {
        for (int q = 0; q < n.range; q++) (q < nl ? L.box : R.box).grow(items[n.start + q].box);
      }
      L.start = n.start; L.range = nl; L.left = L.right = -1;
      R.start = n.start + nl; R.range = nr; R.left = R.right = -1;
      nodes[id].left = li; nodes[id].right = ri;
    }
    const bool lsmall = nl <= kMaxLeaf, rsmall = nr <= kMaxLeaf;
    if (lsmall && rsmall) return;
    if (lsmall) { if (nl > 0) split(ri); return; }      // nl == 0: everything on one side -> oversized leaf
    if (rsmall) { if (nr > 0) split(li); return; }
    if (std::min(nl, nr) >= kParallelMin && spare_threads.fetch_sub(1) > 0) {
      std::future<void> f = std::async(std::launch::async, [this, li] { split(li); });
      split(ri);
      f.get();
      spare_threads.fetch_add(1);
    } else {
      if (std::min(nl, nr) >= kParallelMin) spare_threads.fetch_add(1);   // undo the failed reservation
      split(li);
      split(ri);
    }
  }
};

}  // namespace
}  // namespace dsrt

extern "C" int dsrt_build_bvh2(const dsrt_scene* sc, double* node_bbox, int32_t* node_start, int32_t* node_range,
                               int32_t* node_left, int32_t* node_right, int32_t* prim_order, int32_t* n_nodes) {
  using namespace dsrt;
  if (!sc || !node_bbox || !node_start || !node_range || !node_left || !node_right || !prim_order || !n_nodes)
    return DSRT_ERR_INVALID;
  if (sc->n_prims < 0) return DSRT_ERR_INVALID;
  const bool timing = std::getenv("DSRT_BUILD_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  std::vector<Box3> pbox;
  primitive_boxes(sc, pbox);
  BuildNode root; root.box.reset();
  for (int i = 0; i < sc->n_prims; i++) root.box.grow(pbox[i]);
  root.start = 0; root.range = sc->n_prims; root.left = root.right = -1;
  Builder b(pbox);
  std::vector<Box3>().swap(pbox);
  b.nodes.resize(2 * (size_t)sc->n_prims + 2);
  b.nodes[0] = root; b.next_node = 1;
  b.spare_threads = host_threads() - 1;
  const double t1 = now();
  b.split(0);   // the reference splits the root unconditionally (bvh.cpp:199-200)
  const double t2 = now();
  b.nodes.resize((size_t)b.next_node.load());
  for (int i = 0; i < sc->n_prims; i++) prim_order[i] = b.items[i].id;
  // renumber in preorder
  std::vector<int> stack; stack.push_back(0);
  std::vector<int> newid(b.nodes.size(), -1);
  int out = 0;
  std::vector<int> visit; visit.reserve(b.nodes.size());
  while (!stack.empty()) {
    int o = stack.back(); stack.pop_back();
    newid[o] = out++; visit.push_back(o);
    if (b.nodes[o].right >= 0) stack.push_back(b.nodes[o].right);
    if (b.nodes[o].left >= 0) stack.push_back(b.nodes[o].left);
  }
  for (int o : visit) {
    const BuildNode& n = b.nodes[o];
    int id = newid[o];
    for (int k = 0; k < 3; k++) { node_bbox[6 * id + k] = n.box.lo[k]; node_bbox[6 * id + 3 + k] = n.box.hi[k]; }
    node_start[id] = n.start; node_range[id] = n.range;
    node_left[id] = n.left >= 0 ? newid[n.left] : -1;
    node_right[id] = n.right >= 0 ? newid[n.right] : -1;
  }
  *n_nodes = out;
  if (timing) std::fprintf(stderr, "dsrt_build_bvh2: %d prims, boxes+items %.2f s, split %.2f s, renumber %.2f s\n", sc->n_prims, t1 - t0, t2 - t1, now() - t2);
  return DSRT_OK;
}
