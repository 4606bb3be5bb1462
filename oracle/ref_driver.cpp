// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Headless driver that links against the reference's OWN CPU translation units
// (compiled in place from /root/reference by oracle/build_ref.sh; nothing from the
// reference is copied into this repository).  It restates what the reference's
// src/main.cpp:71-188 and Application::{init,load,init_*,set_up_pathtracer,loadCamera}
// (src/application.cpp:49-100, 223-352, 624-633, 823-853) do for a headless "-c" run,
// minus GL/GLFW/CUDA, and dumps everything the parity tests need as .npy files:
//
//   --dump-scene : flattened primitives / BSDF table / lights / camera / SAH BVH topology
//   --ids        : primary-ray closest-hit primitive ids + t at pixel centres
//                  (BVHAccel::intersect, src/bvh.cpp:343-363)
//   --render     : the reference CPU render (PathTracer::start_raytracing, 1 thread,
//                  fixed srand seed) as a raw linear float RGB buffer + segment counters
//
// Built with -fno-access-control so private members of reference classes can be read.
#include "ref_app.h"

// segment counters, incremented by the sed-inserted hooks in the patched copy of bvh.cpp
long long g_ref_closest_calls = 0;
long long g_ref_any_calls = 0;

// declared in pathtracer.h:200 but only defined in cuda_src/setup.cu:829
namespace CMU462 { void PathTracer::updateBufferFromGPU(float*) {} }

// ---------------------------------------------------------------- npy writer
static void npy_write(const std::string& path, const char* descr, size_t elem,
                      const std::vector<size_t>& shape, const void* data) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
  std::string sh = "(";
  size_t n = 1;
  for (size_t i = 0; i < shape.size(); i++) { sh += std::to_string(shape[i]) + ","; n *= shape[i]; }
  sh += ")";
  std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': " + sh + ", }";
  size_t total = 10 + hdr.size() + 1;
  size_t pad = (64 - total % 64) % 64;
  hdr += std::string(pad, ' ') + "\n";
  unsigned char magic[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0,
                             (unsigned char)(hdr.size() & 0xff), (unsigned char)(hdr.size() >> 8)};
  fwrite(magic, 1, 10, f);
  fwrite(hdr.data(), 1, hdr.size(), f);
  if (n) fwrite(data, elem, n, f);
  fclose(f);
}

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc; i++) {
    std::string s = argv[i];
    auto next = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", s.c_str()); exit(2); } return argv[++i]; };
    if (s == "-s") a.spp = atoi(next());
    else if (s == "-l") a.nl = atoi(next());
    else if (s == "-t") a.threads = atoi(next());
    else if (s == "-m") a.depth = atoi(next());
    else if (s == "-w") a.w = atoi(next());
    else if (s == "-h") a.h = atoi(next());
    else if (s == "-f") a.cam = next();
    else if (s == "--seed") a.seed = (unsigned)strtoul(next(), 0, 10);
    else if (s == "--out") a.out = next();
    else if (s == "--dump-scene") a.dump_scene = true;
    else if (s == "--ids") a.ids = true;
    else if (s == "--render") a.render = true;
    else if (s == "--envmap") { a.env_w = atoi(next()); a.env_h = atoi(next()); }
    else a.scene = s;
  }
  if (a.scene.empty()) { fprintf(stderr, "usage: ref_driver [-s -l -t -m -w -h -f] [--seed N] [--out DIR] [--dump-scene] [--ids] [--render] scene.dae\n"); return 2; }

  RefApp app;
  if (int rc = ref_app_setup(a, app)) return rc;
  PathTracer* pt = app.pt; Camera& camera = *app.camera; HDRImageBuffer* envmap = app.envmap;
  const size_t screenW = app.screenW, screenH = app.screenH;

  const std::vector<StaticScene::Primitive*>& prims = pt->primitives;
  const size_t N = prims.size();
  std::map<const StaticScene::Primitive*, int> primIndex;
  for (size_t i = 0; i < N; i++) primIndex[prims[i]] = (int)i;

  if (a.dump_scene) {
    std::vector<int32_t> ptype(N), pbsdf(N);
    std::vector<double> tpos(N * 9, 0.0), tnrm(N * 9, 0.0), sph(N * 4, 0.0);
    std::map<BSDF*, int> bsdfIndex;
    std::vector<BSDF*> bsdfs;
    for (size_t i = 0; i < N; i++) {
      StaticScene::Primitive* p = prims[i];
      BSDF* b = p->get_bsdf();
      if (!bsdfIndex.count(b)) { bsdfIndex[b] = (int)bsdfs.size(); bsdfs.push_back(b); }
      pbsdf[i] = bsdfIndex[b];
      ptype[i] = p->getType();                                      // triangle.h:74 -> 1, sphere.h:85 -> 0
      if (ptype[i] == 1) {
        StaticScene::Triangle* t = static_cast<StaticScene::Triangle*>(p);
        const Vector3D* P = t->mesh->positions; const Vector3D* Nn = t->mesh->normals;
        size_t idx[3] = {t->v1, t->v2, t->v3};
        for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) {
          tpos[i * 9 + k * 3 + c] = P[idx[k]][c];
          tnrm[i * 9 + k * 3 + c] = Nn[idx[k]][c];
        }
      } else {
        StaticScene::Sphere* s = static_cast<StaticScene::Sphere*>(p);
        sph[i * 4 + 0] = s->o.x; sph[i * 4 + 1] = s->o.y; sph[i * 4 + 2] = s->o.z; sph[i * 4 + 3] = s->r;
      }
    }
    npy_write(a.out + "/prim_type.npy", "<i4", 4, {N}, ptype.data());
    npy_write(a.out + "/prim_bsdf.npy", "<i4", 4, {N}, pbsdf.data());
    npy_write(a.out + "/tri_pos.npy", "<f8", 8, {N, 9}, tpos.data());
    npy_write(a.out + "/tri_nrm.npy", "<f8", 8, {N, 9}, tnrm.data());
    npy_write(a.out + "/sphere.npy", "<f8", 8, {N, 4}, sph.data());

    // BSDF table: type, a[3] (albedo | reflectance | radiance), b[3] (transmittance), ior
    size_t nb = bsdfs.size();
    std::vector<int32_t> btype(nb);
    std::vector<float> bpar(nb * 8, 0.f);
    for (size_t i = 0; i < nb; i++) {
      BSDF* b = bsdfs[i];
      btype[i] = b->getType();
      float* q = &bpar[i * 8];
      switch (btype[i]) {
        case 0: { Spectrum s = static_cast<DiffuseBSDF*>(b)->albedo; q[0] = s.r; q[1] = s.g; q[2] = s.b; break; }
        case 1: { Spectrum s = static_cast<MirrorBSDF*>(b)->reflectance; q[0] = s.r; q[1] = s.g; q[2] = s.b; break; }
        case 2: { RefractionBSDF* r = static_cast<RefractionBSDF*>(b); q[3] = r->transmittance.r; q[4] = r->transmittance.g; q[5] = r->transmittance.b; q[6] = r->ior; break; }
        case 3: { GlassBSDF* g = static_cast<GlassBSDF*>(b); q[0] = g->reflectance.r; q[1] = g->reflectance.g; q[2] = g->reflectance.b;
                  q[3] = g->transmittance.r; q[4] = g->transmittance.g; q[5] = g->transmittance.b; q[6] = g->ior; break; }
        case 4: { Spectrum s = static_cast<EmissionBSDF*>(b)->radiance; q[0] = s.r; q[1] = s.g; q[2] = s.b; break; }
      }
    }
    npy_write(a.out + "/bsdf_type.npy", "<i4", 4, {nb}, btype.data());
    npy_write(a.out + "/bsdf_param.npy", "<f4", 4, {nb, 8}, bpar.data());

    // lights: type + 24 doubles {radiance[3], dirToLight|position[3], direction[3], dim_x[3], dim_y[3], area, sampleToWorld[9] col-major}
    size_t nl = pt->scene->lights.size();
    std::vector<int32_t> ltype(nl);
    std::vector<double> lpar(nl * 28, 0.0);
    for (size_t i = 0; i < nl; i++) {
      StaticScene::SceneLight* L = pt->scene->lights[i];
      ltype[i] = L->getType();
      // EnvironmentLight::getType() also returns 1 (environment_light.h:47): flattened as type 4
      if (dynamic_cast<StaticScene::EnvironmentLight*>(L)) { ltype[i] = 4; continue; }
      double* q = &lpar[i * 28];
      switch (ltype[i]) {
        case 0: { auto* d = static_cast<StaticScene::DirectionalLight*>(L); q[0] = d->radiance.r; q[1] = d->radiance.g; q[2] = d->radiance.b;
                  q[3] = d->dirToLight.x; q[4] = d->dirToLight.y; q[5] = d->dirToLight.z; break; }
        case 1: { auto* h = static_cast<StaticScene::InfiniteHemisphereLight*>(L); q[0] = h->radiance.r; q[1] = h->radiance.g; q[2] = h->radiance.b;
                  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) q[16 + c * 3 + r] = h->sampleToWorld(r, c); break; }
        case 2: { auto* p = static_cast<StaticScene::PointLight*>(L); q[0] = p->radiance.r; q[1] = p->radiance.g; q[2] = p->radiance.b;
                  q[3] = p->position.x; q[4] = p->position.y; q[5] = p->position.z; break; }
        case 3: { auto* ar = static_cast<StaticScene::AreaLight*>(L); q[0] = ar->radiance.r; q[1] = ar->radiance.g; q[2] = ar->radiance.b;
                  q[3] = ar->position.x; q[4] = ar->position.y; q[5] = ar->position.z;
                  q[6] = ar->direction.x; q[7] = ar->direction.y; q[8] = ar->direction.z;
                  q[9] = ar->dim_x.x; q[10] = ar->dim_x.y; q[11] = ar->dim_x.z;
                  q[12] = ar->dim_y.x; q[13] = ar->dim_y.y; q[14] = ar->dim_y.z;
                  q[15] = ar->area; break; }
        default: break;
      }
    }
    if (envmap) npy_write(a.out + "/env_rgb.npy", "<f4", 4, {(size_t)envmap->h, (size_t)envmap->w, 3}, envmap->data.data());
    npy_write(a.out + "/light_type.npy", "<i4", 4, {nl}, ltype.data());
    npy_write(a.out + "/light_param.npy", "<f8", 8, {nl, 28}, lpar.data());

    // camera: pos[3], c2w[9] column-major, screenW, screenH, screenDist, hFov, vFov
    std::vector<double> cam;
    v3(cam, camera.pos);
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) cam.push_back(camera.c2w(r, c));
    cam.push_back((double)camera.screenW); cam.push_back((double)camera.screenH); cam.push_back(camera.screenDist);
    cam.push_back(camera.hFov); cam.push_back(camera.vFov);
    npy_write(a.out + "/camera.npy", "<f8", 8, {cam.size()}, cam.data());

    // SAH BVH topology, preorder (node, left subtree, right subtree)
    std::vector<double> nb_box; std::vector<int32_t> nstart, nrange, nleft, nright;
    struct Item { StaticScene::BVHNode* n; int parent; bool isLeft; };
    std::vector<Item> stack; stack.push_back({pt->bvh->root, -1, false});
    while (!stack.empty()) {
      Item it = stack.back(); stack.pop_back();
      int id = (int)nstart.size();
      if (it.parent >= 0) { if (it.isLeft) nleft[it.parent] = id; else nright[it.parent] = id; }
      StaticScene::BVHNode* n = it.n;
      for (int c = 0; c < 3; c++) nb_box.push_back(n->bb.min[c]);
      for (int c = 0; c < 3; c++) nb_box.push_back(n->bb.max[c]);
      nstart.push_back((int)n->start); nrange.push_back((int)n->range); nleft.push_back(-1); nright.push_back(-1);
      if (n->r) stack.push_back({n->r, id, false});
      if (n->l) stack.push_back({n->l, id, true});
    }
    size_t M = nstart.size();
    npy_write(a.out + "/node_bbox.npy", "<f8", 8, {M, 6}, nb_box.data());
    npy_write(a.out + "/node_start.npy", "<i4", 4, {M}, nstart.data());
    npy_write(a.out + "/node_range.npy", "<i4", 4, {M}, nrange.data());
    npy_write(a.out + "/node_left.npy", "<i4", 4, {M}, nleft.data());
    npy_write(a.out + "/node_right.npy", "<i4", 4, {M}, nright.data());
    std::vector<int32_t> order(N);
    for (size_t i = 0; i < N; i++) order[i] = primIndex[pt->bvh->primitives[i]];
    npy_write(a.out + "/prim_order.npy", "<i4", 4, {N}, order.data());
    fprintf(stderr, "[ref_driver] scene: %zu prims, %zu bsdfs, %zu lights, %zu bvh nodes\n", N, nb, nl, M);
  }

  if (a.ids) {
    std::vector<int32_t> ids(screenW * screenH);
    std::vector<double> ts(screenW * screenH);
    for (size_t y = 0; y < screenH; y++) for (size_t x = 0; x < screenW; x++) {
      Ray r = camera.generate_ray((x + 0.5) / screenW, (y + 0.5) / screenH);
      StaticScene::Intersection isect;
      bool hit = pt->bvh->intersect(r, &isect);
      ids[y * screenW + x] = hit ? primIndex[isect.primitive] : -1;
      ts[y * screenW + x] = hit ? isect.t : INF_D;
    }
    npy_write(a.out + "/hit_id.npy", "<i4", 4, {screenH, screenW}, ids.data());
    npy_write(a.out + "/hit_t.npy", "<f8", 8, {screenH, screenW}, ts.data());
  }

  if (a.render) {
    g_ref_closest_calls = g_ref_any_calls = 0;
    auto t0 = std::chrono::steady_clock::now();
    pt->start_raytracing();                                          // pathtracer.cpp:192-221
    while (1) {                                                      // main.cpp:174-181
      pt->m.lock(); int st = pt->state; pt->m.unlock();
      if (st == PathTracer::DONE) break;
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    auto t1 = std::chrono::steady_clock::now();
    double sec = std::chrono::duration<double>(t1 - t0).count();
    npy_write(a.out + "/rgb.npy", "<f4", 4, {screenH, screenW, 3}, pt->sampleBuffer.data.data());
    double cnt[3] = {(double)g_ref_closest_calls, (double)g_ref_any_calls, sec};
    npy_write(a.out + "/counters.npy", "<f8", 8, {3}, cnt);
    fprintf(stderr, "[ref_driver] render %.3f s, closest=%lld any=%lld (%.3f Mseg/s)\n", sec, g_ref_closest_calls,
            g_ref_any_calls, (g_ref_closest_calls + g_ref_any_calls) / sec / 1e6);
  }
  fflush(stdout); fflush(stderr);
  _exit(0);                                                          // main.cpp:204-208 also skips destructors
}
