// host_util.h -- small host-side helpers shared by the builder, the wide-BVH flattener and the API.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "../../include/dsrt.h"

namespace dsrt {

struct Box3 {
  double lo[3], hi[3];
  void reset() {
    for (int k = 0; k < 3; k++) { lo[k] = std::numeric_limits<double>::infinity(); hi[k] = -lo[k]; }
  }
  void grow(const Box3& o) {
    for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], o.lo[k]); hi[k] = std::fmax(hi[k], o.hi[k]); }
  }
  void grow(const double* p) {
    for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], p[k]); hi[k] = std::fmax(hi[k], p[k]); }
  }
  // ex*ey + ex*ez + ey*ez exactly as the cost expression is written at bvh.cpp:73-74 (an empty box gives
  // +inf, and inf * 0 primitives = NaN, which never compares less than the running minimum)
  double half_area() const {
    double ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    return ex * ey + ex * ez + ey * ez;
  }
  double centre(int k) const { return 0.5 * (lo[k] + hi[k]); }
};

// Triangle::get_bbox (triangle.cpp:11-23) / Sphere::get_bbox (sphere.h:30-32)
void primitive_boxes(const dsrt_scene* sc, std::vector<Box3>& out);

}  // namespace dsrt
