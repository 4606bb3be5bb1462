// integration/cuda_path_tracer_shim.cpp -- see cuda_path_tracer_shim.h.  Every loader below reads the same members of the
// reference's PathTracer that the reference's own cuda_src/setup.cu reads, and hands them to the C ABI of include/dsrt.h.
#include "cuda_path_tracer_shim.h"

#include <cstdio>
#include <cstring>

#include "bvh.h"
#include "bsdf.h"
#include "camera.h"

using namespace CMU462;
using namespace StaticScene;

CUDAPathTracer::CUDAPathTracer(PathTracer* _pathTracer) : pathTracer(_pathTracer) {
  check(dsrt_create(0, &ctx), "dsrt_create");
}

CUDAPathTracer::~CUDAPathTracer() {
  if (ctx) dsrt_destroy(ctx);
}

void CUDAPathTracer::check(int rc, const char* what) {
  if (rc == 0 || status != 0) return;
  status = rc;
  message = std::string(what) + ": " + (ctx ? dsrt_last_error(ctx) : "no context");
  fprintf(stderr, "[CUDAPathTracer] %s\n", message.c_str());       // the reference prints and exit()s; here the caller decides
}

void CUDAPathTracer::init() {                                      // setup.cu:181-201, same order
  loadCamera();
  loadPrimitives();
  loadLights();
  loadBVH();
  createFrameBuffer();
  loadParameters();
  if (ok()) check(dsrt_build_accel(ctx), "dsrt_build_accel");
}

void CUDAPathTracer::createFrameBuffer() {                         // setup.cu:203-219
  screenH = (int)pathTracer->frameBuffer.h;
  screenW = (int)pathTracer->frameBuffer.w;
  frame.assign((size_t)3 * screenW * screenH, 0.f);
}

void CUDAPathTracer::loadCamera() {                                // setup.cu:221-247
  Camera* cam = pathTracer->camera;
  double c2w[9];                                                   // column-major, as include/dsrt.h asks
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) c2w[3 * c + r] = cam->c2w(r, c);
  const double pos[3] = {cam->pos[0], cam->pos[1], cam->pos[2]};
  if (ok()) check(dsrt_set_camera(ctx, pos, c2w, (int32_t)cam->screenW, (int32_t)cam->screenH, cam->screenDist), "dsrt_set_camera");
}

void CUDAPathTracer::loadPrimitives() {                            // setup.cu:249-402
  std::vector<Primitive*>& primitives = pathTracer->primitives;
  const size_t N = primitives.size();
  prim_type.assign(N, 0); prim_bsdf.assign(N, 0);
  tri_pos.assign(9 * N, 0.0); tri_nrm.assign(9 * N, 0.0); sphere.assign(4 * N, 0.0);
  std::map<BSDF*, int> BSDFMap;                                    // index = order of first appearance (setup.cu:260-272)
  std::vector<BSDF*> table;
  for (size_t i = 0; i < N; i++) {
    primMap[primitives[i]] = (int)i;
    prim_type[i] = primitives[i]->getType();                       // triangle.h:74 -> 1, sphere.h:85 -> 0
    BSDF* bsdf = primitives[i]->get_bsdf();
    if (BSDFMap.find(bsdf) == BSDFMap.end()) { const int index = (int)BSDFMap.size(); BSDFMap[bsdf] = index; table.push_back(bsdf); }
    prim_bsdf[i] = BSDFMap[bsdf];
    if (prim_type[i] == 0) {
      Sphere* s = (Sphere*)primitives[i];
      sphere[4 * i] = s->o[0]; sphere[4 * i + 1] = s->o[1]; sphere[4 * i + 2] = s->o[2]; sphere[4 * i + 3] = s->r;
    } else {
      Triangle* t = (Triangle*)primitives[i];
      const Mesh* mesh = t->mesh;
      const size_t v[3] = {t->v1, t->v2, t->v3};
      for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) {    // absolute vertices in double: the kernels make their own layout
        tri_pos[9 * i + 3 * k + c] = mesh->positions[v[k]][c];
        tri_nrm[9 * i + 3 * k + c] = mesh->normals[v[k]][c];
      }
    }
  }
  bsdf_type.assign(table.size(), 0); bsdf_param.assign(8 * table.size(), 0.f);
  for (size_t i = 0; i < table.size(); i++) {                      // setup.cu:323-365; 8 floats {a[3], b[3], ior, 0}
    BSDF* bsdf = table[i]; float* q = &bsdf_param[8 * i];
    bsdf_type[i] = bsdf->getType();
    Spectrum a(0, 0, 0), b(0, 0, 0); float ior = 0.f;
    switch (bsdf_type[i]) {
      case 0: a = ((DiffuseBSDF*)bsdf)->albedo; break;
      case 1: a = ((MirrorBSDF*)bsdf)->reflectance; break;
      case 2: b = ((RefractionBSDF*)bsdf)->transmittance; ior = ((RefractionBSDF*)bsdf)->ior; break;
      case 3: a = ((GlassBSDF*)bsdf)->reflectance; b = ((GlassBSDF*)bsdf)->transmittance; ior = ((GlassBSDF*)bsdf)->ior; break;
      case 4: a = ((EmissionBSDF*)bsdf)->radiance; break;
      default: break;
    }
    q[0] = a.r; q[1] = a.g; q[2] = a.b; q[3] = b.r; q[4] = b.g; q[5] = b.b; q[6] = ior;
  }
  have_prims = true;
  commitScene();
}

void CUDAPathTracer::loadLights() {                                // setup.cu:689-774 (toGPULight)
  const std::vector<SceneLight*>& lights = pathTracer->scene->lights;
  light_type.clear(); light_param.clear();
  for (SceneLight* L : lights) {
    if (dynamic_cast<EnvironmentLight*>(L)) continue;              // reaches the GPU through dsrt_set_envmap, not the light table
    double q[28]; std::memset(q, 0, sizeof(q));
    const int type = L->getType();                                 // light.h:24-99: 0 directional, 1 hemisphere, 2 point, 3 area
    switch (type) {
      case 0: { DirectionalLight* d = (DirectionalLight*)L; q[0] = d->radiance.r; q[1] = d->radiance.g; q[2] = d->radiance.b;
                q[3] = d->dirToLight.x; q[4] = d->dirToLight.y; q[5] = d->dirToLight.z; break; }
      case 1: { InfiniteHemisphereLight* h = (InfiniteHemisphereLight*)L; q[0] = h->radiance.r; q[1] = h->radiance.g; q[2] = h->radiance.b;
                for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) q[16 + 3 * c + r] = h->sampleToWorld(r, c); break; }
      case 2: { PointLight* p = (PointLight*)L; q[0] = p->radiance.r; q[1] = p->radiance.g; q[2] = p->radiance.b;
                q[3] = p->position.x; q[4] = p->position.y; q[5] = p->position.z; break; }
      case 3: { AreaLight* a = (AreaLight*)L; q[0] = a->radiance.r; q[1] = a->radiance.g; q[2] = a->radiance.b;
                q[3] = a->position.x; q[4] = a->position.y; q[5] = a->position.z;
                q[6] = a->direction.x; q[7] = a->direction.y; q[8] = a->direction.z;
                q[9] = a->dim_x.x; q[10] = a->dim_x.y; q[11] = a->dim_x.z;
                q[12] = a->dim_y.x; q[13] = a->dim_y.y; q[14] = a->dim_y.z; q[15] = a->area; break; }
      default: continue;                                           // spot / sphere / mesh lights are empty stubs in the reference (light.cpp:61-115)
    }
    light_type.push_back(type);
    light_param.insert(light_param.end(), q, q + 28);
  }
  if (pathTracer->envLight) {                                      // PathTracer's envmap constructor argument (pathtracer.cpp:41-45)
    const HDRImageBuffer* e = pathTracer->envLight->envMap;
    std::vector<float> rgb((size_t)e->w * e->h * 3);
    for (size_t i = 0; i < (size_t)e->w * e->h; i++) { rgb[3 * i] = e->data[i].r; rgb[3 * i + 1] = e->data[i].g; rgb[3 * i + 2] = e->data[i].b; }
    if (ok()) check(dsrt_set_envmap(ctx, (int32_t)e->w, (int32_t)e->h, rgb.data()), "dsrt_set_envmap");
  }
  have_lights = true;
  commitScene();
}

void CUDAPathTracer::commitScene() {
  if (!have_prims || !have_lights || !ok()) return;
  dsrt_scene sc;
  std::memset(&sc, 0, sizeof(sc));
  sc.n_prims = (int32_t)prim_type.size(); sc.prim_type = prim_type.data(); sc.prim_bsdf = prim_bsdf.data();
  sc.tri_pos = tri_pos.data(); sc.tri_nrm = tri_nrm.data(); sc.sphere = sphere.data();
  sc.n_bsdf = (int32_t)bsdf_type.size(); sc.bsdf_type = bsdf_type.data(); sc.bsdf_param = bsdf_param.data();
  sc.n_lights = (int32_t)light_type.size(); sc.light_type = light_type.data(); sc.light_param = light_param.data();
  check(dsrt_set_scene(ctx, &sc), "dsrt_set_scene");
}

void CUDAPathTracer::loadBVH() {                                   // setup.cu:415-476: the reference's own SAH tree, flattened
  std::vector<double> box; std::vector<int32_t> start, range, left, right;
  struct Item { BVHNode* n; int parent; bool isLeft; };
  std::vector<Item> stack;
  if (pathTracer->bvh && pathTracer->bvh->root) stack.push_back({pathTracer->bvh->root, -1, false});
  while (!stack.empty()) {                                         // pre-order: node, left subtree, right subtree
    const Item it = stack.back(); stack.pop_back();
    const int id = (int)start.size();
    if (it.parent >= 0) (it.isLeft ? left : right)[it.parent] = id;
    BVHNode* n = it.n;
    for (int c = 0; c < 3; c++) box.push_back(n->bb.min[c]);
    for (int c = 0; c < 3; c++) box.push_back(n->bb.max[c]);
    start.push_back((int32_t)n->start); range.push_back((int32_t)n->range); left.push_back(-1); right.push_back(-1);
    if (n->r) stack.push_back({n->r, id, false});
    if (n->l) stack.push_back({n->l, id, true});
  }
  const std::vector<Primitive*>& ordered = pathTracer->bvh->primitives;          // BVHAccel's reordered primitive list
  std::vector<int32_t> order(ordered.size());
  for (size_t i = 0; i < ordered.size(); i++) order[i] = primMap[ordered[i]];
  dsrt_bvh2 b;
  std::memset(&b, 0, sizeof(b));
  b.n_nodes = (int32_t)start.size(); b.node_bbox = box.data(); b.node_start = start.data(); b.node_range = range.data();
  b.node_left = left.data(); b.node_right = right.data(); b.prim_order = order.data();
  if (ok()) check(dsrt_set_bvh(ctx, &b), "dsrt_set_bvh");
}

void CUDAPathTracer::loadParameters() {                            // setup.cu:777-811
  if (ok()) check(dsrt_set_params(ctx, (int32_t)pathTracer->ns_aa, (int32_t)pathTracer->ns_area_light, (int32_t)pathTracer->max_ray_depth, seed),
                  "dsrt_set_params");
}

void CUDAPathTracer::startRayTracingPT() {                         // setup.cu:147-179
  if (ok()) check(dsrt_render(ctx, 0, (int32_t)pathTracer->ns_aa, 1, frame.data(), &last_stats), "dsrt_render");
}

void CUDAPathTracer::updateHostSampleBuffer() {                    // setup.cu:813-827
  if (ok()) pathTracer->updateBufferFromGPU(frame.data());
}

// PathTracer::updateBufferFromGPU is declared in src/pathtracer.h:200 but defined in cuda_src/setup.cu:829-843, the file this
// shim replaces: sampleBuffer <- frame, then the tone-mapped frameBuffer.
void CMU462::PathTracer::updateBufferFromGPU(float* gpuBuffer) {
  const size_t w = sampleBuffer.w, h = sampleBuffer.h;
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) {
      const float* p = gpuBuffer + 3 * (y * w + x);
      sampleBuffer.update_pixel(Spectrum(p[0], p[1], p[2]), x, y);
    }
  sampleBuffer.toColor(frameBuffer, 0, 0, w, h);
}
