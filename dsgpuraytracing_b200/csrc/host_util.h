// host_util.h -- small host-side helpers shared by the builder, the wide-BVH flattener and the API.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/dsrt.h"

namespace dsrt {

// vector of trivially copyable elements whose resize() leaves new elements uninitialised (no single-threaded zero fill of
// gigabyte arrays that are about to be overwritten), filled by a parallel copy
template <class T> struct DefaultInitAllocator : std::allocator<T> {
  template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
  template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
template <class T> using PodVec = std::vector<T, DefaultInitAllocator<T>>;

struct Box3 {
  double lo[3], hi[3];
  void reset() {
    for (int k = 0; k < 3; k++) { lo[k] = std::numeric_limits<double>::infinity(); hi[k] = -lo[k]; }
  }
  // std::min / std::max as in BBox::expand (src/bbox.h:69-91); they inline to minsd / maxsd, std::fmin is a libm call
  void grow(const Box3& o) {
    for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], o.lo[k]); hi[k] = std::max(hi[k], o.hi[k]); }
  }
  void grow(const double* p) {
    for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
  }
  // ex*ey + ex*ez + ey*ez exactly as the cost expression is written at bvh.cpp:73-74 (an empty box gives
  // +inf, and inf * 0 primitives = NaN, which never compares less than the running minimum)
  double half_area() const {
    double ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    return ex * ey + ex * ez + ey * ez;
  }
  double centre(int k) const { return 0.5 * (lo[k] + hi[k]); }
};

// sphere (centre xyz, radius) around a box, rounded outwards in float: bounds how far a ray with tmax = inf can reach
inline void bounding_sphere(const Box3& b, int n_prims, float out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0.f;
  if (n_prims <= 0) return;
  double r2 = 0;
  for (int k = 0; k < 3; k++) { out[k] = (float)(0.5 * (b.lo[k] + b.hi[k])); }
  for (int k = 0; k < 3; k++) { const double e = std::max(std::fabs(b.hi[k] - (double)out[k]), std::fabs(b.lo[k] - (double)out[k])); r2 += e * e; }
  out[3] = (float)(std::sqrt(r2) * 1.000001 + 1e-30);
}

// worker threads of the host-side builders; DSRT_HOST_THREADS overrides the core count (the results never depend on it)
inline int host_threads() {
  if (const char* e = std::getenv("DSRT_HOST_THREADS")) { const int n = std::atoi(e); if (n >= 1) return n; }
  return (int)std::max(1u, std::thread::hardware_concurrency());
}

// f(begin, end) over [0, n) in chunks of `grain`, handed out dynamically to up to host_threads() workers.  The chunk
// boundaries do not depend on the thread count, so anything a chunk computes on its own is schedule independent.
template <class F> void parallel_for(size_t n, size_t grain, F f) {
  if (n == 0) return;
  grain = std::max<size_t>(grain, 1);
  const size_t chunks = (n + grain - 1) / grain;
  const int workers = (int)std::min<size_t>((size_t)host_threads(), chunks);
  if (workers <= 1) { f((size_t)0, n); return; }
  std::atomic<size_t> next{0};
  auto body = [&] { for (size_t c; (c = next.fetch_add(1)) < chunks;) f(c * grain, std::min(n, (c + 1) * grain)); };
  std::vector<std::thread> th;
  for (int i = 1; i < workers; i++) th.emplace_back(body);
  body();
  for (std::thread& t : th) t.join();
}

// Triangle::get_bbox (triangle.cpp:11-23) / Sphere::get_bbox (sphere.h:30-32)
void primitive_boxes(const dsrt_scene* sc, std::vector<Box3>& out);

template <class T> void assign_parallel(PodVec<T>& v, const T* src, size_t n) {
  v.resize(n);
  parallel_for(n, (size_t)1 << 22, [&](size_t a, size_t b) { std::memcpy(v.data() + a, src + a, (b - a) * sizeof(T)); });
}

}  // namespace dsrt
