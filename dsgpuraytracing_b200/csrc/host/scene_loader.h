// scene_loader.h -- host-side scene import: COLLADA (.dae) -> flat static scene (include/dsrt.h layout) + camera.
// Restates, headless, what the reference does between main() and PathTracer::set_scene:
//   Collada::ColladaParser::load / parse_*            reference src/collada/collada.cpp:131-936
//   Application::load / init_* / loadCamera           reference src/application.cpp:223-352, 823-853
//   DynamicScene::Mesh, HalfedgeMesh::build, Vertex::computeNormal   src/dynamic_scene/mesh.cpp:16-35,
//                                                     src/halfEdgeMesh.cpp:29-397, src/halfEdgeMesh.h:492-515
//   StaticScene::Mesh / SphereObject::get_primitives  src/static_scene/object.cpp:16-81
//   Camera::configure / place / compute_position      src/camera.cpp:15-111
// so that primitive order (= primitive id), triangle vertex rotation, world positions, vertex normals, BSDF table,
// light table and camera are identical to the reference's.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace dsrt_host {

struct HostCamera {
  // Camera state that Camera::generate_ray reads (camera.h:92-105)
  double pos[3] = {0, 0, 0};
  double c2w[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // column-major
  double targetPos[3] = {0, 0, 0};
  double phi = 0, theta = 0, r = 0, minR = 0, maxR = 0;
  double hFov = 50, vFov = 35, nClip = 0.01, fClip = 100, ar = 1;
  size_t screenW = 600, screenH = 600;
  double screenDist = 0;

  void configure(double hFov_, double vFov_, double nClip_, double fClip_, size_t w, size_t h);   // camera.cpp:15-34
  void place(const double target[3], double phi_, double theta_, double r_, double minR_, double maxR_);  // :36-49
  void compute_position();                                                                           // :88-111
  bool load_info(const std::string& path, std::string& err);   // Application::loadCamera, application.cpp:823-853
};

struct FlatScene {
  std::vector<int32_t> prim_type, prim_bsdf;
  std::vector<double> tri_pos, tri_nrm, sphere;
  std::vector<int32_t> bsdf_type;
  std::vector<float> bsdf_param;
  std::vector<int32_t> light_type;
  std::vector<double> light_param;
  int n_prims() const { return (int)prim_type.size(); }
};

// Loads `path`, sizes the camera for a width x height frame (main.cpp:158-161), places the default orbit camera
// (application.cpp:267-291).  Returns false and fills err on failure (the reference exit()s).
bool load_collada(const std::string& path, size_t width, size_t height, FlatScene& scene, HostCamera& camera, std::string& err);
// Opt-in: meshes the half-edge builder rejects (non-manifold, inconsistently oriented, repeated indices) are imported as plain
// indexed triangles (polygons fan-triangulated) instead of failing.  Off by default: the reference stops on such input.
void set_direct_triangle_fallback(bool on);

}  // namespace dsrt_host
