// oracle/ref_app.h -- TEST INFRASTRUCTURE, not product code.
//
// Headless restatement of what the reference's src/main.cpp:71-188 and Application::{init,load,init_*,set_up_pathtracer,
// loadCamera} (src/application.cpp:49-100, 223-352, 624-633, 823-853) do before a render, minus GL/GLFW: parse the .dae with
// the reference's ColladaParser, build its DynamicScene / StaticScene, place the camera, hand everything to ITS PathTracer.
// Shared by oracle/ref_driver.cpp (the CPU reference arm) and integration/ref_gpu_driver.cpp (the reference's PathTracer
// driving libdsrt.so through the CUDAPathTracer shim).  Compiled against the reference's headers where they lie.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include <chrono>
#include <thread>
#include <unistd.h>

#include "CMU462/CMU462.h"
#include "collada/collada.h"
#include "dynamic_scene/scene.h"
#include "dynamic_scene/mesh.h"
#include "dynamic_scene/sphere.h"
#include "dynamic_scene/ambient_light.h"
#include "dynamic_scene/area_light.h"
#include "dynamic_scene/directional_light.h"
#include "dynamic_scene/point_light.h"
#include "dynamic_scene/spot_light.h"
#include "static_scene/triangle.h"
#include "static_scene/sphere.h"
#include "static_scene/light.h"
#include "static_scene/environment_light.h"
#include "static_scene/object.h"
#include "pathtracer.h"
#include "camera.h"
#include "bvh.h"

using namespace CMU462;
using Collada::CameraInfo;
using Collada::LightInfo;
using Collada::PolymeshInfo;
using Collada::SphereInfo;

struct Args {
  std::string scene, cam, out = ".";
  int w = 1000, h = 1000, spp = 1, nl = 4, depth = 1, threads = 1;
  unsigned seed = 1;
  bool dump_scene = false, ids = false, render = false;
  int env_w = 0, env_h = 0;   // --envmap W H: procedural lat-long environment map (no .exr ships with the reference)
};

// Procedural sky: vertical gradient + a warm "sun" lobe + a faint ground, deterministic, float RGB.
static HDRImageBuffer* make_envmap(int w, int h) {
  HDRImageBuffer* e = new HDRImageBuffer(w, h);
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    double v = (y + 0.5) / h, u = (x + 0.5) / w;
    double sky = v < 0.5 ? 0.35 + 0.9 * (0.5 - v) : 0.08;
    double du = u - 0.3, dv = v - 0.22;
    double sun = 40.0 * std::exp(-(du * du + dv * dv) / 0.0008);
    e->data[x + w * y] = Spectrum((float)(0.6 * sky + sun), (float)(0.75 * sky + 0.9 * sun), (float)(1.0 * sky + 0.7 * sun));
  }
  return e;
}

static void v3(std::vector<double>& v, const Vector3D& a) { v.push_back(a.x); v.push_back(a.y); v.push_back(a.z); }


// what Application owns after load() + set_up_pathtracer()
struct RefApp {
  PathTracer* pt = nullptr;
  Camera* camera = nullptr;
  HDRImageBuffer* envmap = nullptr;
  size_t screenW = 0, screenH = 0;
};

// returns 0 on success (the process exit code otherwise)
static int ref_app_setup(const Args& a, RefApp& app) {
  std::srand(a.seed);                                              // main.cpp:75 (fixed seed instead of time(0))

  Collada::SceneInfo* sceneInfo = new Collada::SceneInfo();        // main.cpp:132-137
  if (Collada::ColladaParser::load(a.scene.c_str(), sceneInfo) < 0) { fprintf(stderr, "cannot load %s\n", a.scene.c_str()); return 3; }

  HDRImageBuffer* envmap = (a.env_w > 0 && a.env_h > 0) ? make_envmap(a.env_w, a.env_h) : NULL;   // main.cpp:99-101 (-e)
  PathTracer* pt = new PathTracer(a.spp, a.depth, a.nl, 1, 1, 1, a.threads, envmap);   // application.cpp:28-40
  pt->useCPU = true;

  // Application::init (application.cpp:87-99): dummy camera, then main.cpp:158-159 sets screenW/H
  Camera& camera = *(app.camera = new Camera());
  {
    CameraInfo ci; ci.hFov = 50; ci.vFov = 35; ci.nClip = 0.01; ci.fClip = 100;
    camera.configure(ci, 600, 600);
  }
  size_t screenW = a.w, screenH = a.h;

  // Application::load (application.cpp:223-299)
  std::vector<DynamicScene::SceneLight*> lights;
  std::vector<DynamicScene::SceneObject*> objects;
  Vector3D c_pos, c_dir;
  for (size_t i = 0; i < sceneInfo->nodes.size(); i++) {
    Collada::Node& node = sceneInfo->nodes[i];
    Collada::Instance* inst = node.instance;
    if (!inst) continue;
    const Matrix4x4& T = node.transform;
    switch (inst->type) {
      case Collada::Instance::CAMERA: {
        CameraInfo* c = static_cast<CameraInfo*>(inst);
        c_pos = (T * Vector4D(c_pos, 1)).to3D();
        c_dir = (T * Vector4D(c->view_dir, 1)).to3D().unit();
        camera.configure(*c, screenW, screenH);                     // init_camera, application.cpp:301-308
        break;
      }
      case Collada::Instance::LIGHT: {
        LightInfo& li = static_cast<LightInfo&>(*inst);              // init_light, application.cpp:314-333
        DynamicScene::SceneLight* L = nullptr;
        switch (li.light_type) {
          case Collada::LightType::AMBIENT: L = new DynamicScene::AmbientLight(li); break;
          case Collada::LightType::DIRECTIONAL: L = new DynamicScene::DirectionalLight(li, T); break;
          case Collada::LightType::AREA: L = new DynamicScene::AreaLight(li, T); break;
          case Collada::LightType::POINT: L = new DynamicScene::PointLight(li, T); break;
          case Collada::LightType::SPOT: L = new DynamicScene::SpotLight(li, T); break;
          default: break;
        }
        lights.push_back(L);
        break;
      }
      case Collada::Instance::SPHERE: {                              // init_sphere, application.cpp:342-347
        SphereInfo& si = static_cast<SphereInfo&>(*inst);
        const Vector3D position = (T * Vector4D(0, 0, 0, 1)).projectTo3D();
        double scale = (T * Vector4D(1, 0, 0, 0)).to3D().norm();
        objects.push_back(new DynamicScene::Sphere(si, position, scale));
        break;
      }
      case Collada::Instance::POLYMESH:                              // init_polymesh, application.cpp:349-352
        objects.push_back(new DynamicScene::Mesh(static_cast<PolymeshInfo&>(*inst), T));
        break;
      default: break;
    }
  }
  DynamicScene::Scene* scene = new DynamicScene::Scene(objects, lights);
  BBox bbox = scene->get_bbox();
  if (!bbox.empty()) {                                               // application.cpp:267-291
    Vector3D target = bbox.centroid();
    double cvd = bbox.extent.norm() / 2 * 1.5;
    camera.place(target, acos(c_dir.y), atan2(c_dir.x, c_dir.z), cvd * 2, cvd / 10.0, cvd * 20.0);
  }

  // Application::set_up_pathtracer (application.cpp:624-633)
  pt->set_camera(&camera);
  pt->set_scene(scene->get_static_scene());
  pt->set_frame_size(screenW, screenH);

  if (!a.cam.empty()) {                                              // Application::loadCamera, application.cpp:823-853
    FILE* pf = fopen(a.cam.c_str(), "r");
    if (!pf) { fprintf(stderr, "cannot open camera file %s\n", a.cam.c_str()); return 3; }
    Camera& cam = camera;
    int n = 0;
    n += fscanf(pf, "%lf %lf %lf", &cam.pos[0], &cam.pos[1], &cam.pos[2]);
    n += fscanf(pf, "%lf %lf %lf", &cam.targetPos[0], &cam.targetPos[1], &cam.targetPos[2]);
    n += fscanf(pf, "%lf", &cam.phi);
    n += fscanf(pf, "%lf", &cam.theta);
    n += fscanf(pf, "%lf", &cam.minR);
    n += fscanf(pf, "%lf", &cam.maxR);
    n += fscanf(pf, "%lf %lf %lf %lf %lf %lf %lf %lf %lf", &cam.c2w(0, 0), &cam.c2w(0, 1), &cam.c2w(0, 2),
                &cam.c2w(1, 0), &cam.c2w(1, 1), &cam.c2w(1, 2), &cam.c2w(2, 0), &cam.c2w(2, 1), &cam.c2w(2, 2));
    fclose(pf);
    if (n != 19) { fprintf(stderr, "bad camera file\n"); return 3; }
  }
  app.pt = pt; app.envmap = envmap; app.screenW = screenW; app.screenH = screenH;
  return 0;
}
