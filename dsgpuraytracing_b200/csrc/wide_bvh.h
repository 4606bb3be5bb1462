// wide_bvh.h -- host-side flattening of the binary SAH BVH into the compressed 8-wide layout (layout.h).
#pragma once
#include <string>
#include <vector>

#include "host_util.h"
#include "layout.h"

namespace dsrt {

struct WideBVH {
  std::vector<WideNode> nodes;      // node 0 = root, breadth-first
  std::vector<int32_t> slot_prim;   // leaf-contiguous slot -> primitive id
  int max_depth = 0;                // levels of wide nodes (bounds the traversal stack)
};

// primitive / shading records in leaf-contiguous slot order (layout.h)
void flatten_records(const dsrt_scene& sc, const WideBVH& wide, std::vector<PrimRecord>& recs, std::vector<ShadeRecord>& shd);
// marks the leaf slots that hold coplanar triangles (WideNode::flat, layout.h); after build_wide_bvh
void mark_flat_slots(const dsrt_scene& sc, WideBVH& wide);
// fp64 records of the parity kernel (same slot order); built on demand by dsrt_primary_hits(mode 1)
void flatten_records64(const dsrt_scene& sc, const WideBVH& wide, std::vector<PrimRecord64>& r64);
// float light table; returns the number of light samples per path vertex (pathtracer.cpp:474)
int flatten_lights(int n_lights, const int32_t* light_type, const double* light_param, int ns_area_light, bool with_env,
                   std::vector<Light>& out);
// EnvironmentLight constructor (environment_light.cpp:6-53): float tables accumulated in the reference's order
void build_env_tables(int w, int h, const float* rgb, std::vector<float>& pThetaPhi, std::vector<float>& pTheta,
                      std::vector<float>& pPhiGivenTheta);

// A rectangle on which any-hit rays END: an area light whose plane is perpendicular to a coordinate axis.  Shadow rays stop
// 0.1 % short of their light sample (pathtracer.cpp:486-504), so geometry that lies IN the light's plane -- the emissive quad
// every Cornell scene of the reference puts under its area light -- passes the box test of almost every shadow ray unless the
// quantised plane that faces the arriving rays is within 0.001 x distance of the true plane.  A wide node that holds such a
// flat child therefore shifts its quantisation grid by a fraction of a quantum so that this one plane falls just past a grid
// line (wide_bvh.cpp, step 4); every box stays conservative, only where the rounding slack goes changes.
struct EndPlane {
  int axis;             // the light's plane is coord along this axis
  bool from_low;        // rays that carry radiance arrive from the low side (the light faces -axis)
  double coord;
  double lo[3], hi[3];  // extent of the rectangle (lo[axis] == hi[axis] == coord)
};
std::vector<EndPlane> light_end_planes(int n_lights, const int32_t* light_type, const double* light_param);

// regroup_top: children of the top wide nodes are regrouped where that lowers the SAH cost (wide_bvh.cpp, step 2b); fewer node
// visits per ray, but not fewer node steps per WARP on B200 (include/dsrt.h "regroup_top"), hence off by default
// prim_cost: SAH cost of one primitive test relative to one wide-node visit in the collapse (1.0 measured best on B200)
int build_wide_bvh(const dsrt_bvh2& b2, const std::vector<Box3>& pbox, int n_prims, WideBVH& out, std::string& err, double prim_cost = 1.0,
                   const std::vector<EndPlane>* end_planes = nullptr, bool regroup_top = false);

}  // namespace dsrt
