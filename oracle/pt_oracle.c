/* oracle/pt_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, or called by, the product).
 *
 * Plain-C, double-precision CPU restatement of the reference's PathTracer render path
 * (Khrylx/DSGPURayTracing, CPU branch).  Every function cites the reference file:line it
 * follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
 * may load this library, and only as the checker.
 *
 * PARITY PIN: the reference ships no golden vectors for this path, so the
 * pin is the reference's own CPU code compiled here (oracle/_ref/ref_driver, built by
 * oracle/build_ref.sh).  tests/test_oracle_vs_reference.py checks this restatement against it:
 * SAH BVH topology node-for-node, primary-hit ids bit-exact (ties included), and -- because this
 * file draws glibc rand() in exactly the reference's call order in RNG mode 0 -- the rendered
 * radiance buffer for the same srand() seed.
 *
 * Compile WITHOUT -march=native / -ffast-math (see oracle/Makefile): operation order and the
 * absence of FMA contraction are part of the contract.
 *
 * Flat scene layout (identical to what oracle/ref_driver.cpp dumps and to include/dsrt.h):
 *   prim_type[i]  1 = triangle, 0 = sphere          (triangle.h:74, sphere.h:85)
 *   prim_bsdf[i]  index into the BSDF table
 *   tri_pos[9i..] p1,p2,p3 world space (rotated polygon order, object.cpp:36-41)
 *   tri_nrm[9i..] vertex normals n1,n2,n3
 *   sphere[4i..]  centre xyz, radius
 *   bsdf_type     0 diffuse 1 mirror 2 refraction 3 glass 4 emission   (bsdf.h:123-236)
 *   bsdf_param[8] a[3] (albedo|reflectance|radiance), b[3] (transmittance), ior, pad
 *   light_type    0 directional 1 hemisphere 2 point 3 area            (light.h:24-99)
 *                 4 environment map (EnvironmentLight; the reference reuses id 1 for it, environment_light.h:47)
 *   light_param[28] radiance[3], dirToLight|position[3], direction[3], dim_x[3], dim_y[3], area,
 *                 sampleToWorld[9] column-major (at offset 16)
 *   cam[17]       pos[3], c2w[9] column-major, screenW, screenH, screenDist, hFov, vFov
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323          /* CMU462/misc.h:10 */
#define ORC_EPS_D 0.00000000001             /* misc.h:13 */
#define ORC_EPS_N 0.005                     /* misc.h: EPS_N 5e-3 */

typedef struct { double x, y, z; } v3;
typedef struct { float r, g, b; } spec;

static inline v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline v3 vmulr(v3 a, double c) { return V(a.x * c, a.y * c, a.z * c); }   /* v * c,  vector3D.h:74 */
static inline v3 vmull(double c, v3 a) { return V(c * a.x, c * a.y, c * a.z); }   /* c * v,  vector3D.h:136 */
static inline v3 vdiv(v3 a, double c) { double rc = 1.0 / c; return V(rc * a.x, rc * a.y, rc * a.z); } /* :79 */
static inline double vdot(v3 u, v3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }
static inline v3 vcross(v3 u, v3 v) { return V(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x); }
static inline double vnorm(v3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
static inline double vnorm2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline v3 vunit(v3 a) { double rn = 1. / sqrt(a.x * a.x + a.y * a.y + a.z * a.z); return V(rn * a.x, rn * a.y, rn * a.z); } /* :121 */
static inline v3 vnormalize(v3 a) { double c = 1. / vnorm(a); return V(a.x * c, a.y * c, a.z * c); } /* :129,:100,:95 */
static inline double vget(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* column-major 3x3, M*x = x0*c0 + x1*c1 + x2*c2 (matrix3x3.cpp:138-142) */
typedef struct { v3 c[3]; } m3;
static inline v3 m3mul(const m3* M, v3 x) {
  return vadd(vadd(vmull(x.x, M->c[0]), vmull(x.y, M->c[1])), vmull(x.z, M->c[2]));
}
static inline m3 m3T(const m3* A) {
  m3 B;
  B.c[0] = V(A->c[0].x, A->c[1].x, A->c[2].x);
  B.c[1] = V(A->c[0].y, A->c[1].y, A->c[2].y);
  B.c[2] = V(A->c[0].z, A->c[1].z, A->c[2].z);
  return B;
}

static inline spec S(float r, float g, float b) { spec s = {r, g, b}; return s; }
static inline spec sadd(spec a, spec b) { return S(a.r + b.r, a.g + b.g, a.b + b.b); }
static inline spec smul(spec a, spec b) { return S(a.r * b.r, a.g * b.g, a.b * b.b); }
static inline spec sscale(spec a, float s) { return S(a.r * s, a.g * s, a.b * s); }
static inline float sillum(spec a) { return 0.2126f * a.r + 0.7152f * a.g + 0.0722f * a.b; } /* spectrum.h:94-96 */

/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int n_prims;
  const int32_t* prim_type;
  const int32_t* prim_bsdf;
  const double* tri_pos;
  const double* tri_nrm;
  const double* sphere;
  int n_bsdf;
  const int32_t* bsdf_type;
  const float* bsdf_param;
  int n_lights;
  const int32_t* light_type;
  const double* light_param;
  const double* cam;
  /* optional lat-long environment map (EnvironmentLight, environment_light.cpp): light_type 4 refers to it */
  int env_w, env_h;
  const float* env_rgb;        /* env_h * env_w * 3 */
  const float* env_pThetaPhi;  /* tables built by orc_env_build */
  const float* env_pTheta;
  const float* env_pPhiGivenTheta;
} orc_scene;

typedef struct {
  int n_nodes;
  const double* node_bbox;   /* n_nodes x 6: min xyz, max xyz */
  const int32_t* node_start;
  const int32_t* node_range;
  const int32_t* node_left;  /* -1 if none */
  const int32_t* node_right;
  const int32_t* prim_order; /* BVH slot -> primitive id */
} orc_bvh;

typedef struct {
  v3 o, d;
  double min_t, max_t; /* max_t is 'mutable' in the reference (ray.h:18-19) */
  int depth;
} oray;

typedef struct {
  double t;
  int prim; /* primitive id, -1 none */
  v3 n;
  int bsdf;
} oisect;

/* ------------------------------------------------------------------ primitive bounding boxes */
typedef struct { v3 mn, mx; } obox;
static obox box_empty(void) { obox b; b.mn = V(INFINITY, INFINITY, INFINITY); b.mx = V(-INFINITY, -INFINITY, -INFINITY); return b; }
static void box_expand_p(obox* b, v3 p) { /* bbox.h:69-77 */
  b->mn.x = fmin(b->mn.x, p.x); b->mn.y = fmin(b->mn.y, p.y); b->mn.z = fmin(b->mn.z, p.z);
  b->mx.x = fmax(b->mx.x, p.x); b->mx.y = fmax(b->mx.y, p.y); b->mx.z = fmax(b->mx.z, p.z);
}
static void box_expand_b(obox* b, const obox* o) { /* bbox.h:59-67 */
  b->mn.x = fmin(b->mn.x, o->mn.x); b->mn.y = fmin(b->mn.y, o->mn.y); b->mn.z = fmin(b->mn.z, o->mn.z);
  b->mx.x = fmax(b->mx.x, o->mx.x); b->mx.y = fmax(b->mx.y, o->mx.y); b->mx.z = fmax(b->mx.z, o->mx.z);
}
static inline v3 tri_p(const orc_scene* s, int prim, int k) { const double* p = s->tri_pos + 9 * (size_t)prim + 3 * k; return V(p[0], p[1], p[2]); }
static inline v3 tri_n(const orc_scene* s, int prim, int k) { const double* p = s->tri_nrm + 9 * (size_t)prim + 3 * k; return V(p[0], p[1], p[2]); }

static obox prim_bbox(const orc_scene* s, int prim) {
  obox b;
  if (s->prim_type[prim] == 1) { /* Triangle::get_bbox, triangle.cpp:11-23 */
    b = box_empty();
    box_expand_p(&b, tri_p(s, prim, 0)); box_expand_p(&b, tri_p(s, prim, 1)); box_expand_p(&b, tri_p(s, prim, 2));
  } else { /* Sphere::get_bbox, sphere.h:30-32 */
    const double* q = s->sphere + 4 * (size_t)prim;
    v3 o = V(q[0], q[1], q[2]); v3 r = V(q[3], q[3], q[3]);
    b.mn = vsub(o, r); b.mx = vadd(o, r);
  }
  return b;
}

/* ------------------------------------------------------------------ SAH BVH build (bvh.cpp:21-202) */
typedef struct {
  int cap, n;
  double* bbox; int32_t *start, *range, *left, *right;
} bnodes;
static int bn_new(bnodes* B, const obox* bb, int start, int range) {
  if (B->n == B->cap) {
    B->cap = B->cap ? B->cap * 2 : 1024;
    B->bbox = (double*)realloc(B->bbox, sizeof(double) * 6 * B->cap);
    B->start = (int32_t*)realloc(B->start, 4 * B->cap); B->range = (int32_t*)realloc(B->range, 4 * B->cap);
    B->left = (int32_t*)realloc(B->left, 4 * B->cap); B->right = (int32_t*)realloc(B->right, 4 * B->cap);
  }
  int id = B->n++;
  double* q = B->bbox + 6 * id;
  q[0] = bb->mn.x; q[1] = bb->mn.y; q[2] = bb->mn.z; q[3] = bb->mx.x; q[4] = bb->mx.y; q[5] = bb->mx.z;
  B->start[id] = start; B->range[id] = range; B->left[id] = -1; B->right[id] = -1;
  return id;
}
typedef struct { obox bb; int cnt; } bucket;
static inline double half_area(const obox* b) { /* extent products as written at bvh.cpp:73-74 */
  double ex = b->mx.x - b->mn.x, ey = b->mx.y - b->mn.y, ez = b->mx.z - b->mn.z;
  return ex * ey + ex * ez + ey * ez;
}

static void build_rec(const orc_scene* s, const obox* pb, int32_t* order, bnodes* B, int node, int bucketNum, int max_leaf) {
  int start = B->start[node], range = B->range[node];
  double nmin[3] = {B->bbox[6 * node], B->bbox[6 * node + 1], B->bbox[6 * node + 2]};
  double nmax[3] = {B->bbox[6 * node + 3], B->bbox[6 * node + 4], B->bbox[6 * node + 5]};
  double minC[3] = {INFINITY, INFINITY, INFINITY};
  int minB[3] = {0, 0, 0}; /* uninitialised in the reference (bvh.cpp:30) when no split is finite */
  bucket* Bk = (bucket*)malloc(sizeof(bucket) * bucketNum);
  bucket* rBk = (bucket*)malloc(sizeof(bucket) * bucketNum);
  for (int k = 0; k < 3; k++) {
    double ub = nmax[k], lb = nmin[k];
    if (ub == lb) continue;                                   /* bvh.cpp:35-37 */
    double interval = (ub - lb) / bucketNum;
    for (int i = 0; i < bucketNum; i++) { Bk[i].bb = box_empty(); Bk[i].cnt = 0; }
    for (int i = 0; i < range; i++) {                         /* bvh.cpp:43-51 */
      const obox* p = &pb[order[start + i]];
      double c = (vget(p->mn, k) + vget(p->mx, k)) * 0.5;
      int bi = (int)((c - lb) / interval);
      if (bi > bucketNum - 1) bi = bucketNum - 1;             /* the F3 clamp (see oracle/build_ref.sh) */
      if (bi < 0) bi = 0;
      box_expand_b(&Bk[bi].bb, p); Bk[bi].cnt++;
    }
    for (int i = 0; i < bucketNum; i++) {                     /* bvh.cpp:54-60 */
      rBk[i] = Bk[bucketNum - i - 1];
      if (i > 0) { box_expand_b(&rBk[i].bb, &rBk[i - 1].bb); rBk[i].cnt += rBk[i - 1].cnt; }
    }
    for (int i = 1; i < bucketNum; i++) { box_expand_b(&Bk[i].bb, &Bk[i - 1].bb); Bk[i].cnt += Bk[i - 1].cnt; } /* :64-67 */
    for (int i = 0; i < bucketNum - 1; i++) {                 /* bvh.cpp:70-79 */
      bucket* b1 = &Bk[i]; bucket* b2 = &rBk[bucketNum - i - 2];
      double C = half_area(&b1->bb) * b1->cnt + half_area(&b2->bb) * b2->cnt;
      if (C < minC[k]) { minC[k] = C; minB[k] = i + 1; }
    }
  }
  free(Bk); free(rBk);
  int axis = 0; double cost = minC[0];                        /* bvh.cpp:83-90 */
  for (int i = 1; i < 3; i++) if (minC[i] < cost) { axis = i; cost = minC[i]; }
  double ub = nmax[axis], lb = nmin[axis];
  double pLine = lb + (ub - lb) * minB[axis] / bucketNum;     /* bvh.cpp:95 */
  int i = start - 1, j = start + range;
  while (i < j) {                                             /* bvh.cpp:101-125 */
    double c1 = 0, c2 = 0;
    do { i++; if (i >= start + range) break; const obox* p = &pb[order[i]]; c1 = (vget(p->mn, axis) + vget(p->mx, axis)) * 0.5; } while (c1 < pLine);
    do { j--; if (j < start) break; const obox* p = &pb[order[j]]; c2 = (vget(p->mn, axis) + vget(p->mx, axis)) * 0.5; } while (c2 > pLine);
    if (i < j) { int32_t t = order[i]; order[i] = order[j]; order[j] = t; } else break;
  }
  int lR = i - start, rR = range - lR;                        /* bvh.cpp:128-139 */
  obox lbb = box_empty(), rbb = box_empty();
  for (int q = 0; q < range; q++) { const obox* p = &pb[order[start + q]]; if (q < lR) box_expand_b(&lbb, p); else box_expand_b(&rbb, p); }
  if (!(lR == 0 || rR == 0)) {                                /* bvh.cpp:143-144 */
    int l = bn_new(B, &lbb, start, lR); B->left[node] = l;
    int r = bn_new(B, &rbb, start + lR, rR); B->right[node] = r;
  }
  int L = B->left[node], R = B->right[node];
  if (lR <= max_leaf && rR <= max_leaf) return;               /* bvh.cpp:147-173 */
  else if (lR <= max_leaf) { if (lR > 0) build_rec(s, pb, order, B, R, bucketNum, max_leaf); }
  else if (rR <= max_leaf) { if (rR > 0) build_rec(s, pb, order, B, L, bucketNum, max_leaf); }
  else { build_rec(s, pb, order, B, L, bucketNum, max_leaf); build_rec(s, pb, order, B, R, bucketNum, max_leaf); }
}

/* Builds the reference SAH BVH (BVHAccel::BVHAccel, bvh.cpp:181-202: 32 buckets, leaf 4) and returns it
 * renumbered in preorder (node, left subtree, right subtree).  Output arrays must hold 2*n_prims entries. */
int orc_build_bvh(const orc_scene* s, double* node_bbox, int32_t* node_start, int32_t* node_range,
                  int32_t* node_left, int32_t* node_right, int32_t* prim_order) {
  int n = s->n_prims;
  obox* pb = (obox*)malloc(sizeof(obox) * (n > 0 ? n : 1));
  obox root = box_empty();
  for (int i = 0; i < n; i++) { pb[i] = prim_bbox(s, i); prim_order[i] = i; box_expand_b(&root, &pb[i]); }
  bnodes B; memset(&B, 0, sizeof(B));
  bn_new(&B, &root, 0, n);
  build_rec(s, pb, prim_order, &B, 0, 32, 4);
  /* preorder renumbering */
  int* stack = (int*)malloc(sizeof(int) * (B.n + 1)); int* parent = (int*)malloc(sizeof(int) * (B.n + 1));
  char* isl = (char*)malloc(B.n + 1);
  int sp = 0, out = 0; stack[sp] = 0; parent[sp] = -1; isl[sp] = 0; sp++;
  while (sp) {
    sp--; int o = stack[sp], par = parent[sp]; char il = isl[sp];
    int id = out++;
    if (par >= 0) { if (il) node_left[par] = id; else node_right[par] = id; }
    memcpy(node_bbox + 6 * id, B.bbox + 6 * o, 48);
    node_start[id] = B.start[o]; node_range[id] = B.range[o]; node_left[id] = -1; node_right[id] = -1;
    if (B.right[o] >= 0) { stack[sp] = B.right[o]; parent[sp] = id; isl[sp] = 0; sp++; }
    if (B.left[o] >= 0) { stack[sp] = B.left[o]; parent[sp] = id; isl[sp] = 1; sp++; }
  }
  free(stack); free(parent); free(isl); free(pb);
  free(B.bbox); free(B.start); free(B.range); free(B.left); free(B.right);
  return out;
}

/* ------------------------------------------------------------------ BBox::intersect (bbox.cpp:10-30) */
int orc_bbox_intersect(const double* bb, const double* o, const double* d, double* t0, double* t1) {
  for (int i = 0; i < 3; i++) {
    if (d[i] != 0.0) {
      double tx1 = (bb[i] - o[i]) / d[i];
      double tx2 = (bb[3 + i] - o[i]) / d[i];
      *t0 = fmax(*t0, fmin(tx1, tx2));
      *t1 = fmin(*t1, fmax(tx1, tx2));
    }
  }
  return *t0 <= *t1;
}

/* ------------------------------------------------------------------ Triangle::intersect (triangle.cpp:25-104) */
static int tri_intersect(const orc_scene* s, int prim, oray* r, oisect* i) {
  v3 p1 = tri_p(s, prim, 0), p2 = tri_p(s, prim, 1), p3 = tri_p(s, prim, 2);
  v3 e1 = vsub(p2, p1), e2 = vsub(p3, p1), sv = vsub(r->o, p1);
  double f = vdot(vcross(e1, r->d), e2);
  if (f == 0) return 0;
  double u = vdot(vcross(sv, r->d), e2) / f;
  double v = vdot(vcross(e1, r->d), sv) / f;
  double t = vdot(vcross(e1, vneg(sv)), e2) / f;
  if (i) {
    if (!(u >= 0 && v >= 0 && u + v <= 1 && t > r->min_t && t < r->max_t && t < i->t)) return 0;
    r->max_t = t;
    i->bsdf = s->prim_bsdf[prim]; i->t = t; i->prim = prim;
    v3 n = vadd(vadd(vmull(1 - u - v, tri_n(s, prim, 0)), vmull(u, tri_n(s, prim, 1))), vmull(v, tri_n(s, prim, 2)));
    if (vdot(r->d, n) > 0) n = vneg(n);
    i->n = n;
    return 1;
  }
  return (u >= 0 && v >= 0 && u + v <= 1 && t > r->min_t && t < r->max_t);
}

/* ------------------------------------------------------------------ Sphere::test/intersect (sphere.cpp:10-77) */
static int sph_intersect(const orc_scene* s, int prim, oray* r, oisect* i) {
  const double* q = s->sphere + 4 * (size_t)prim;
  v3 o = V(q[0], q[1], q[2]); double r2 = q[3] * q[3];
  v3 m = vsub(o, r->o);
  double b = vdot(m, r->d);
  double c = vdot(m, m) - r2;
  double delta = b * b - c;
  if (delta < 0) return 0;
  double t1 = b - sqrt(delta), t2 = b + sqrt(delta);
  if (!i) {
    /* Sphere::intersect(r) passes the SAME variable for t1 and t2 (sphere.cpp:42-44), so both names
       alias the last value written, t2: the any-hit test is  !(t2 >= max_t || t2 <= min_t). */
    if (t2 >= r->max_t || t2 <= r->min_t) return 0;
    return 1;
  }
  if (t1 >= r->max_t || t2 <= r->min_t) return 0;
  i->bsdf = s->prim_bsdf[prim]; i->prim = prim;
  double t = t1;
  if (t1 <= r->min_t) t = t2;                      /* no t < i->t test, sphere.cpp:64-72 */
  v3 n = vsub(vadd(r->o, vmulr(r->d, t)), o);
  n = vnormalize(n);
  i->n = n; i->t = t; r->max_t = t;
  return 1;
}
static inline int prim_intersect(const orc_scene* s, int prim, oray* r, oisect* i) {
  return s->prim_type[prim] == 1 ? tri_intersect(s, prim, r, i) : sph_intersect(s, prim, r, i);
}

/* ------------------------------------------------------------------ BVH traversal (bvh.cpp:227-329) */
typedef struct { long long closest, any, box_tests, prim_tests; } orc_counters;
static orc_counters g_cnt;

static int node_closest(const orc_scene* s, const orc_bvh* b, int node, oray* ray, oisect* i) {
  int l = b->node_left[node], r = b->node_right[node];
  if (l < 0 && r < 0) {
    int hit = 0;
    for (int j = 0; j < b->node_range[node]; j++) {
      g_cnt.prim_tests++;
      int res = prim_intersect(s, b->prim_order[j + b->node_start[node]], ray, i);
      hit = hit || res;
    }
    return hit;
  }
  if (l < 0) return node_closest(s, b, r, ray, i);
  if (r < 0) return node_closest(s, b, l, ray, i);
  double tminl = -INFINITY, tminr = -INFINITY, tmaxl = INFINITY, tmaxr = INFINITY;
  v3 nd = V(ray->d.x + ORC_EPS_D, ray->d.y + ORC_EPS_D, ray->d.z + ORC_EPS_D);   /* bvh.cpp:250-252 */
  nd = vnormalize(nd);
  double o[3] = {ray->o.x, ray->o.y, ray->o.z}, d[3] = {nd.x, nd.y, nd.z};
  g_cnt.box_tests += 2;
  int hitl = orc_bbox_intersect(b->node_bbox + 6 * l, o, d, &tminl, &tmaxl);
  int hitr = orc_bbox_intersect(b->node_bbox + 6 * r, o, d, &tminr, &tmaxr);
  if (hitl && hitr) {
    int first = (tminl <= tminr) ? l : r, second = (tminl <= tminr) ? r : l;
    hitl = node_closest(s, b, first, ray, i);
    if (!hitl || i->t > fmax(tminl, tminr)) hitr = node_closest(s, b, second, ray, i);
    return hitl || hitr;
  } else if (hitl) return node_closest(s, b, l, ray, i);
  else if (hitr) return node_closest(s, b, r, ray, i);
  return 0;
}
static int node_any(const orc_scene* s, const orc_bvh* b, int node, oray* ray) {
  int l = b->node_left[node], r = b->node_right[node];
  if (l < 0 && r < 0) {
    for (int j = 0; j < b->node_range[node]; j++) {
      g_cnt.prim_tests++;
      if (prim_intersect(s, b->prim_order[j + b->node_start[node]], ray, NULL)) return 1;
    }
    return 0;
  }
  if (l < 0) return node_any(s, b, r, ray);
  if (r < 0) return node_any(s, b, l, ray);
  double tminl = -INFINITY, tminr = -INFINITY, tmaxl = INFINITY, tmaxr = INFINITY;
  v3 nd = V(ray->d.x + ORC_EPS_D, ray->d.y + ORC_EPS_D, ray->d.z + ORC_EPS_D);
  nd = vnormalize(nd);
  double o[3] = {ray->o.x, ray->o.y, ray->o.z}, d[3] = {nd.x, nd.y, nd.z};
  g_cnt.box_tests += 2;
  int hitl = orc_bbox_intersect(b->node_bbox + 6 * l, o, d, &tminl, &tmaxl);
  int hitr = orc_bbox_intersect(b->node_bbox + 6 * r, o, d, &tminr, &tmaxr);
  if (hitl && hitr) {
    int first = (tminl <= tminr) ? l : r, second = (tminl <= tminr) ? r : l;
    return node_any(s, b, first, ray) || node_any(s, b, second, ray);
  } else if (hitl) return node_any(s, b, l, ray);
  else if (hitr) return node_any(s, b, r, ray);
  return 0;
}
static int bvh_closest(const orc_scene* s, const orc_bvh* b, oray* r, oisect* i) { g_cnt.closest++; return node_closest(s, b, 0, r, i); }
static int bvh_any(const orc_scene* s, const orc_bvh* b, oray* r) { g_cnt.any++; return node_any(s, b, 0, r); }

/* ------------------------------------------------------------------ Camera::generate_ray (camera.cpp:113-129) */
void orc_generate_ray(const double* cam, double x, double y, double* o_out, double* d_out) {
  v3 pos = V(cam[0], cam[1], cam[2]);
  m3 c2w; for (int c = 0; c < 3; c++) c2w.c[c] = V(cam[3 + 3 * c], cam[4 + 3 * c], cam[5 + 3 * c]);
  double W = cam[12], H = cam[13], dist = cam[14];
  v3 sp = V(-(x - 0.5) * W / dist, -(y - 0.5) * H / dist, 1);
  v3 dir = vneg(sp);
  v3 wsp = vadd(m3mul(&c2w, sp), pos);
  v3 wd = m3mul(&c2w, dir);
  wd = vnormalize(wd);
  o_out[0] = wsp.x; o_out[1] = wsp.y; o_out[2] = wsp.z; d_out[0] = wd.x; d_out[1] = wd.y; d_out[2] = wd.z;
}

/* ------------------------------------------------------------------ make_coord_space (bsdf.cpp:13-30) */
static void make_coord_space(m3* o2w, v3 n) {
  v3 z = n, h = z;
  if (fabs(h.x) <= fabs(h.y) && fabs(h.x) <= fabs(h.z)) h.x = 1.0;
  else if (fabs(h.y) <= fabs(h.x) && fabs(h.y) <= fabs(h.z)) h.y = 1.0;
  else h.z = 1.0;
  z = vnormalize(z);
  v3 y = vcross(h, z); y = vnormalize(y);
  v3 x = vcross(z, y); x = vnormalize(x);
  o2w->c[0] = x; o2w->c[1] = y; o2w->c[2] = z;
}
void orc_make_coord_space(const double* n, double* out9) {
  m3 m; make_coord_space(&m, V(n[0], n[1], n[2]));
  for (int c = 0; c < 3; c++) { out9[3 * c] = m.c[c].x; out9[3 * c + 1] = m.c[c].y; out9[3 * c + 2] = m.c[c].z; }
}

/* ------------------------------------------------------------------ random numbers
 * mode 0: glibc rand() in the reference's call order  (sampler.cpp:14, pathtracer.cpp:539, bsdf.cpp:147)
 * mode 1: Philox4x32-10 keyed by (seed), counter (pixel, sample, depth, block) -- the SAME streams the
 *         CUDA path uses (dsgpuraytracing_b200/csrc/rng.cuh), so oracle and GPU walk the same paths.
 *         u = (x >> 8) * 2^-24  (exactly representable in float)
 */
typedef struct { int mode; uint32_t seed, pixel, sample; int arg_order_rtl; } orng;

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
void orc_philox(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t depth, uint32_t block, uint32_t* out4) {
  uint32_t c[4] = {pixel, sample, depth, block};
  philox4x32_10(c, seed, 0x5EEDu);
  memcpy(out4, c, 16);
}
/* raw generator for the Random123 known-answer vectors (tests/test_oracle_vs_reference.py) */
void orc_philox_raw(const uint32_t* ctr4, uint32_t k0, uint32_t k1, uint32_t* out4) {
  uint32_t c[4] = {ctr4[0], ctr4[1], ctr4[2], ctr4[3]};
  philox4x32_10(c, k0, k1);
  memcpy(out4, c, 16);
}
static inline double u24(uint32_t x) { return (double)(x >> 8) * (1.0 / 16777216.0); }
static void rng_block(const orng* g, int depth, int block, double u[4]) {
  uint32_t c[4]; orc_philox(g->seed, g->pixel, g->sample, (uint32_t)depth, (uint32_t)block, c);
  for (int i = 0; i < 4; i++) u[i] = u24(c[i]);
}
static inline double rnd01(void) { return rand() / (double)RAND_MAX; }
/* Vector2D(rand()/.., rand()/..) -- g++ evaluates the two constructor arguments right-to-left on x86-64
   (checked against oracle/_ref in tests); sampler.cpp:14 */
static void rng_grid2(const orng* g, int depth, int block, int half, double* x, double* y) {
  if (g->mode == 0) {
    double a = rnd01(), b = rnd01();
    if (g->arg_order_rtl) { *y = a; *x = b; } else { *x = a; *y = b; }
  } else { double u[4]; rng_block(g, depth, block, u); *x = u[2 * half]; *y = u[2 * half + 1]; }
}
/* two sequential draws r1, r2 (sampler.cpp:24-25, 45-46) */
static void rng_seq2(const orng* g, int depth, int block, int half, double* r1, double* r2) {
  if (g->mode == 0) { *r1 = rnd01(); *r2 = rnd01(); }
  else { double u[4]; rng_block(g, depth, block, u); *r1 = u[2 * half]; *r2 = u[2 * half + 1]; }
}
static double rng_one(const orng* g, int depth, int block, int idx) {
  if (g->mode == 0) return rnd01();
  double u[4]; rng_block(g, depth, block, u); return u[idx];
}
/* Philox block ids (shared with the CUDA path): 0 camera jitter; 1 bsdf {dir u0,u1; glass choice u2; roulette u3};
   2 + j/2 light sample j (flat over lights), halves {u0,u1} / {u2,u3}. */
#define BLK_CAM 0
#define BLK_BSDF 1
#define BLK_LIGHT0 2

/* ------------------------------------------------------------------ samplers (sampler.cpp:20-55) */
static v3 cosine_hemisphere(double r1, double r2, float* pdf) {
  double theta = acos(1 - 2 * r1) / 2;
  double phi = 2 * ORC_PI * r2;
  double sin_theta = sin(theta), cos_theta = cos(theta);
  *pdf = (float)(cos_theta / ORC_PI);
  return V(sin_theta * cos(phi), sin_theta * sin(phi), cos_theta);
}
static v3 uniform_hemisphere(double r1, double r2) {
  double sin_theta = sqrt(1 - r1 * r1);
  double phi = 2 * ORC_PI * r2;
  return V(sin_theta * cos(phi), sin_theta * sin(phi), r1);
}

/* ------------------------------------------------------------------ BSDFs (bsdf.cpp:34-202) */
static inline spec bs_a(const orc_scene* s, int b) { const float* q = s->bsdf_param + 8 * b; return S(q[0], q[1], q[2]); }
static inline spec bs_b(const orc_scene* s, int b) { const float* q = s->bsdf_param + 8 * b; return S(q[3], q[4], q[5]); }
static inline float bs_ior(const orc_scene* s, int b) { return s->bsdf_param[8 * b + 6]; }

static spec bsdf_f(const orc_scene* s, int b) { /* only Diffuse is non-zero: bsdf.cpp:34-36, 50-59, 88, 117, 195 */
  if (s->bsdf_type[b] == 0) return sscale(bs_a(s, b), (float)(1.0 / ORC_PI));
  return S(0, 0, 0);
}
static spec bsdf_emission(const orc_scene* s, int b) { return s->bsdf_type[b] == 4 ? bs_a(s, b) : S(0, 0, 0); }
static int bsdf_is_delta(const orc_scene* s, int b) { int t = s->bsdf_type[b]; return t == 1 || t == 2 || t == 3; }

static int refract(v3 wo, v3* wi, float ior) { /* bsdf.cpp:167-191, float arithmetic as written */
  int sign = 1;
  float ratio = ior;
  if (wo.z > 0) { sign = -1; ratio = 1 / ratio; }
  float cos2_wi = (float)(1 - ratio * ratio * (1 - wo.z * wo.z));
  if (cos2_wi < 0) { *wi = V(-wo.x, -wo.y, wo.z); return 0; }
  *wi = vunit(V(-wo.x * ratio, -wo.y * ratio, sign * sqrt(cos2_wi)));
  return 1;
}

static spec bsdf_sample_f(const orc_scene* s, int b, v3 wo, v3* wi, float* pdf, const orng* g, int depth) {
  switch (s->bsdf_type[b]) {
    case 0: { double r1, r2; rng_seq2(g, depth, BLK_BSDF, 0, &r1, &r2); *wi = cosine_hemisphere(r1, r2, pdf);
              return sscale(bs_a(s, b), (float)(1.0 / ORC_PI)); }
    case 1: { *wi = V(-wo.x, -wo.y, wo.z); *pdf = 1; return sscale(bs_a(s, b), (float)(1 / fmax(wo.z, 1e-8))); }
    case 2: {
      *pdf = 1;
      if (!refract(wo, wi, bs_ior(s, b))) return S(0, 0, 0);
      double ni = bs_ior(s, b), no = 1;
      if (wo.z < 0) { double t = ni; ni = no; no = t; }
      double ratio = no / ni;
      /* transmittance * ratio*ratio * (1/max(|wi.z|,1e-8)): left-to-right, each scalar converted to float
         by Spectrum::operator*(float) (bsdf.cpp:110) */
      return sscale(sscale(sscale(bs_b(s, b), (float)ratio), (float)ratio), (float)(1 / fmax(fabs(wi->z), 1e-8)));
    }
    case 3: {
      *pdf = 1;
      if (!refract(wo, wi, bs_ior(s, b))) return sscale(bs_b(s, b), (float)(1 / fmax(fabs(wi->z), 1e-8))); /* :128-131 */
      double ni = bs_ior(s, b), no = 1;
      double cos_i = fabs(wi->z), cos_o = fabs(wo.z);
      if (wo.z < 0) { double t = ni; ni = no; no = t; }
      double r1 = (no * cos_i - ni * cos_o) / (no * cos_i + ni * cos_o);
      double r2 = (ni * cos_i - no * cos_o) / (ni * cos_i + no * cos_o);
      double Fr = 0.5 * (r1 * r1 + r2 * r2);
      if (rng_one(g, depth, BLK_BSDF, 2) <= Fr) {
        *wi = V(-wo.x, -wo.y, wo.z);
        return sscale(bs_a(s, b), (float)(1 / fmax(fabs(wi->z), 1e-8)));
      } else {
        double ratio = no / ni;
        return sscale(sscale(sscale(bs_b(s, b), (float)ratio), (float)ratio), (float)(1 / fmax(fabs(wi->z), 1e-8)));
      }
    }
    case 4: { double r1, r2; rng_seq2(g, depth, BLK_BSDF, 0, &r1, &r2); *wi = cosine_hemisphere(r1, r2, pdf); return S(0, 0, 0); }
  }
  *pdf = 1; *wi = V(0, 0, 1); return S(0, 0, 0);
}

/* ------------------------------------------------------------------ EnvironmentLight (environment_light.cpp:6-201)
 * Tables are float and are accumulated in the reference's order (constructor, :6-53). */
void orc_env_build(int w, int h, const float* rgb, float* pThetaPhi, float* pTheta, float* pPhiGivenTheta) {
  float C = 0;
  for (int y = 0; y < h; y++) {
    float theta = (float)((y + 0.5) / h * ORC_PI);
    float sin_theta = (float)sin(theta);
    for (int x = 0; x < w; x++) {
      const float* q = rgb + 3 * (x + w * y);
      pThetaPhi[y * w + x] = sillum(S(q[0], q[1], q[2])) * sin_theta;
      C += pThetaPhi[y * w + x];
    }
  }
  for (int y = 0; y < h; y++) {
    pTheta[y] = 0;
    for (int x = 0; x < w; x++) { pThetaPhi[y * w + x] /= C; pTheta[y] += pThetaPhi[y * w + x]; pPhiGivenTheta[y * w + x] = 0; }
    if (pTheta[y] != 0) for (int x = 0; x < w; x++) pPhiGivenTheta[y * w + x] = pThetaPhi[y * w + x] / pTheta[y];
  }
  for (int y = 0; y < h; y++) {
    if (y > 0) pTheta[y] += pTheta[y - 1];
    for (int x = 1; x < w; x++) pPhiGivenTheta[y * w + x] += pPhiGivenTheta[y * w + x - 1];
  }
}
static int lower_bound_f(const float* a, int n, float v) {   /* std::lower_bound: first index with a[i] >= v, n if none */
  int lo = 0, hi = n;
  while (lo < hi) { int mid = lo + (hi - lo) / 2; if (a[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}
/* EnvironmentLight::sample_dir (:129-199): bilinear lookup with wrap-around */
static spec env_sample_dir(const orc_scene* s, v3 d) {
  const int w = s->env_w, h = s->env_h;
  double theta = acos(d.y);
  double sin_theta = sqrt(1 - d.y * d.y);
  double cl = d.z / sin_theta; cl = fmin(fmax(cl, -1.0), 1.0);
  double phi = sin_theta == 0 ? ORC_PI : acos(cl);
  if (d.x > 0) phi = 2 * ORC_PI - phi;
  double u = phi / (2 * ORC_PI), v = theta / ORC_PI;
  float tu = (float)(u * w - 0.5), tv = (float)(v * h - 0.5);
  int su = (int)tu, sv = (int)tv;
  float a, b; int px1, px2, py1, py2;
  if (tu < 0) { a = tu + 1; px1 = w - 1; px2 = 0; } else if (tu >= w - 1) { a = tu - w + 1; px1 = w - 1; px2 = 0; } else { a = tu - su; px1 = su; px2 = su + 1; }
  if (tv < 0) { b = tv + 1; py1 = h - 1; py2 = 0; } else if (tv >= h - 1) { b = tv - h + 1; py1 = h - 1; py2 = 0; } else { b = tv - sv; py1 = sv; py2 = sv + 1; }
  const float* E = s->env_rgb;
  #define ENVPX(x, y) S(E[3 * ((x) + w * (y))], E[3 * ((x) + w * (y)) + 1], E[3 * ((x) + w * (y)) + 2])
  spec z11 = ENVPX(px1, py1), z21 = ENVPX(px2, py1), z12 = ENVPX(px1, py2), z22 = ENVPX(px2, py2);
  #undef ENVPX
  spec zy1 = sadd(sscale(z11, 1 - a), sscale(z21, a));
  spec zy2 = sadd(sscale(z12, 1 - a), sscale(z22, a));
  return sadd(sscale(zy1, 1 - b), sscale(zy2, b));
}
/* EnvironmentLight::importanceSampling (:71-113); r1, r2 are the two uniforms as FLOATS */
static void env_importance(const orc_scene* s, float r1, float r2, v3* wi, float* pdf) {
  const int w = s->env_w, h = s->env_h;
  const float* pTheta = s->env_pTheta;
  r1 *= pTheta[h - 1];
  int t = lower_bound_f(pTheta, h, r1);
  if (t >= h) t = h - 1;                       /* (the reference would read past the end) */
  float prev = t > 0 ? pTheta[t - 1] : 0;
  float y = t + (r1 - prev) / (pTheta[t] - prev);
  float theta = (float)(fminf(y / h, 1.f) * ORC_PI);
  const float* row = s->env_pPhiGivenTheta + (size_t)t * w;
  r2 *= row[w - 1];
  int q = lower_bound_f(row, w, r2);
  if (q >= w) q = w - 1;
  prev = q > 0 ? row[q - 1] : 0;
  float x = q + (r2 - prev) / (row[q] - prev);
  float phi = (float)(fminf(x / w, 1.f) * 2 * ORC_PI);
  double sin_theta = sin(theta), cos_theta = cos(theta);
  float p = s->env_pThetaPhi[(size_t)t * w + q];
  p = (float)(p / (sin_theta * (2 * ORC_PI / w) * (ORC_PI / h)));
  *pdf = p;
  *wi = V(-sin_theta * sin(phi), cos_theta, sin_theta * cos(phi));
}

/* ------------------------------------------------------------------ lights (light.cpp:17-92) */
static int light_is_delta(int type) { return type == 0 || type == 2; }
static spec light_sample_L(const orc_scene* s, int l, v3 p, v3* wi, float* dist, float* pdf, const orng* g, int depth, int j) {
  const double* q = s->light_param + 28 * l;
  spec rad = S((float)q[0], (float)q[1], (float)q[2]);
  switch (s->light_type[l]) {
    case 0: *wi = V(q[3], q[4], q[5]); *dist = INFINITY; *pdf = 1.0f; return rad;
    case 1: {
      double r1, r2; rng_seq2(g, depth, BLK_LIGHT0 + j / 2, j & 1, &r1, &r2);
      v3 dir = uniform_hemisphere(r1, r2);
      m3 M; for (int c = 0; c < 3; c++) M.c[c] = V(q[16 + 3 * c], q[17 + 3 * c], q[18 + 3 * c]);
      *wi = m3mul(&M, dir); *dist = INFINITY; *pdf = (float)(1.0 / (2.0 * M_PI)); return rad;
    }
    case 2: { v3 d = vsub(V(q[3], q[4], q[5]), p); *wi = vunit(d); *dist = (float)vnorm(d); *pdf = 1.0f; return rad; }
    case 3: {
      double sx, sy; rng_grid2(g, depth, BLK_LIGHT0 + j / 2, j & 1, &sx, &sy);
      sx -= 0.5f; sy -= 0.5f;
      v3 pos = V(q[3], q[4], q[5]), dir = V(q[6], q[7], q[8]), dx = V(q[9], q[10], q[11]), dy = V(q[12], q[13], q[14]);
      float area = (float)q[15];
      v3 d = vsub(vadd(vadd(pos, vmull(sx, dx)), vmull(sy, dy)), p);
      float cosTheta = (float)vdot(d, dir);
      float sqDist = (float)vnorm2(d);
      float dst = sqrtf(sqDist);
      *wi = vdiv(d, dst);
      *dist = dst;
      *pdf = sqDist / (area * fabsf(cosTheta));   /* all-float: std::fabs(float) overload */
      return cosTheta < 0 ? rad : S(0, 0, 0);
    }
  }
  if (s->light_type[l] == 4) {      /* EnvironmentLight::sample_L (:115-127) */
    float r1, r2;
    if (g->mode == 0) { r1 = rand() / (float)RAND_MAX; r2 = rand() / (float)RAND_MAX; }
    else { double u[4]; rng_block(g, depth, BLK_LIGHT0 + j / 2, u); r1 = (float)u[2 * (j & 1)]; r2 = (float)u[2 * (j & 1) + 1]; }
    env_importance(s, r1, r2, wi, pdf);
    *dist = INFINITY;
    return env_sample_dir(s, *wi);
  }
  *wi = V(0, 1, 0); *dist = INFINITY; *pdf = 1; return S(0, 0, 0);
}

/* ------------------------------------------------------------------ PathTracer::trace_ray (pathtracer.cpp:407-553) */
typedef struct { const orc_scene* s; const orc_bvh* b; int ns_area_light, max_ray_depth; orng g; } octx;

static spec trace_ray(octx* c, oray* r, int includeLe) {
  const orc_scene* s = c->s;
  oisect isect; isect.t = INFINITY; isect.prim = -1; isect.bsdf = -1; isect.n = V(0, 0, 0);
  if (!bvh_closest(s, c->b, r, &isect)) {                    /* pathtracer.cpp:411-427 */
    if (s->env_w > 0 && includeLe) return env_sample_dir(s, r->d);
    return S(0, 0, 0);
  }
  spec L_out = includeLe ? bsdf_emission(s, isect.bsdf) : S(0, 0, 0);
  v3 hit_p = vadd(r->o, vmulr(r->d, isect.t));
  m3 o2w; make_coord_space(&o2w, isect.n);
  m3 w2o = m3T(&o2w);
  v3 w_out = m3mul(&w2o, vsub(r->o, hit_p));
  w_out = vnormalize(w_out);
  v3 dir_to_light; float dist_to_light, pdf;
  int jbase = 0;
  for (int l = 0; l < s->n_lights; l++) {
    spec L = S(0, 0, 0);
    int delta = light_is_delta(s->light_type[l]);
    int ns = delta ? 1 : c->ns_area_light;
    double scale = 1.0 / ns;
    for (int i = 0; i < ns; i++) {
      spec light_L = light_sample_L(s, l, hit_p, &dir_to_light, &dist_to_light, &pdf, &c->g, r->depth, jbase + i);
      double eps = delta ? ORC_EPS_N : 0;
      oray sR; sR.o = vadd(vadd(hit_p, vmull(eps, isect.n)), vmull(ORC_EPS_D, dir_to_light)); sR.d = dir_to_light;
      sR.min_t = 0.0; sR.max_t = dist_to_light * 0.999; sR.depth = 0;
      if (bvh_any(s, c->b, &sR)) continue;
      v3 w_in = m3mul(&w2o, dir_to_light);
      w_in = vnormalize(w_in);
      double cos_theta = fmax(0.0, w_in.z);
      spec f = bsdf_f(s, isect.bsdf);
      L = sadd(L, smul(sscale(light_L, (float)(cos_theta / pdf)), f));
    }
    L_out = sadd(L_out, sscale(L, (float)scale));
    jbase += ns;
  }
  if (r->depth >= c->max_ray_depth) return L_out;
  v3 w_in;
  spec f = bsdf_sample_f(s, isect.bsdf, w_out, &w_in, &pdf, &c->g, r->depth);
  double cos_theta = fabs(w_in.z);
  float tp = fmaxf(1 - sillum(f), 0.f);
  double terminateProbability = tp;
  if (rng_one(&c->g, r->depth, BLK_BSDF, 3) < terminateProbability) return L_out;
  v3 v = m3mul(&o2w, w_in);
  v = vnormalize(v);
  oray refR; refR.o = vadd(hit_p, vmull(ORC_EPS_D, v)); refR.d = v; refR.min_t = 0.0; refR.max_t = INFINITY; refR.depth = r->depth + 1;
  spec indirL = trace_ray(c, &refR, bsdf_is_delta(s, isect.bsdf));
  return sadd(L_out, smul(sscale(indirL, (float)(cos_theta / (pdf * (1 - terminateProbability)))), f));
}

/* ------------------------------------------------------------------ render (pathtracer.cpp:192-221, 555-637)
 * rng_mode 0: srand(seed) then glibc rand() in the reference's order (tiles 32x32 row-major, rows, pixels, samples)
 * rng_mode 1: Philox streams; samples [spp_begin, spp_begin+spp_count) of every pixel.
 * rgb_out: H*W*3 floats, row 0 = bottom (image.h:113-117); counters[4] = closest, any, box tests, prim tests. */
void orc_render(const orc_scene* s, const orc_bvh* b, int W, int H, int spp_begin, int spp_count, int spp_total,
                int ns_area_light, int max_depth, int rng_mode, uint32_t seed, int arg_order_rtl,
                float* rgb_out, double* counters) {
  memset(&g_cnt, 0, sizeof(g_cnt));
  octx c; c.s = s; c.b = b; c.ns_area_light = ns_area_light; c.max_ray_depth = max_depth;
  c.g.mode = rng_mode; c.g.seed = seed; c.g.arg_order_rtl = arg_order_rtl;
  if (rng_mode == 0) srand(seed);
  const int T = 32;
  for (int ty = 0; ty < H; ty += T) for (int tx = 0; tx < W; tx += T) {
    int ex = tx + T < W ? tx + T : W, ey = ty + T < H ? ty + T : H;
    for (int y = ty; y < ey; y++) for (int x = tx; x < ex; x++) {
      spec acc = S(0, 0, 0);
      c.g.pixel = (uint32_t)(y * W + x);
      for (int i = 0; i < spp_count; i++) {                       /* raytrace_pixel, pathtracer.cpp:571-579 */
        c.g.sample = (uint32_t)(spp_begin + i);
        double rx, ry; rng_grid2(&c.g, 0, BLK_CAM, 0, &rx, &ry);
        double px = (x + rx) / W, py = (y + ry) / H;
        oray r; double o[3], d[3];
        orc_generate_ray(s->cam, px, py, o, d);
        r.o = V(o[0], o[1], o[2]); r.d = V(d[0], d[1], d[2]); r.min_t = 0.0; r.max_t = INFINITY; r.depth = 0;
        acc = sadd(acc, trace_ray(&c, &r, 1));
      }
      acc = sscale(acc, (float)(1.0 / spp_total));
      float* q = rgb_out + 3 * ((size_t)y * W + x);
      q[0] = acc.r; q[1] = acc.g; q[2] = acc.b;
    }
  }
  if (counters) { counters[0] = (double)g_cnt.closest; counters[1] = (double)g_cnt.any; counters[2] = (double)g_cnt.box_tests; counters[3] = (double)g_cnt.prim_tests; }
}

/* Primary closest hits at pixel centres (SURVEY 8d gate 1).  tie[i] = 1 when another primitive reports
 * exactly the same t as the winner (exact tie -> excluded from the bit-exact comparison). */
void orc_primary_hits(const orc_scene* s, const orc_bvh* b, int W, int H, int32_t* ids, double* ts, uint8_t* tie) {
  for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
    double o[3], d[3];
    orc_generate_ray(s->cam, (x + 0.5) / W, (y + 0.5) / H, o, d);
    oray r; r.o = V(o[0], o[1], o[2]); r.d = V(d[0], d[1], d[2]); r.min_t = 0.0; r.max_t = INFINITY; r.depth = 0;
    oisect is; is.t = INFINITY; is.prim = -1; is.bsdf = -1; is.n = V(0, 0, 0);
    int hit = bvh_closest(s, b, &r, &is);
    size_t k = (size_t)y * W + x;
    ids[k] = hit ? is.prim : -1; ts[k] = hit ? is.t : INFINITY;
    if (tie) {
      tie[k] = 0;
      if (hit) {
        for (int p = 0; p < s->n_prims && !tie[k]; p++) {
          if (p == is.prim) continue;
          oray r2 = r; r2.max_t = INFINITY; oisect i2; i2.t = INFINITY; i2.prim = -1;
          if (prim_intersect(s, p, &r2, &i2) && i2.t == is.t) tie[k] = 1;
        }
      }
    }
  }
}

/* Single-ray entry points for known-answer tests */
int orc_closest_hit(const orc_scene* s, const orc_bvh* b, const double* o, const double* d, double max_t, double* t, double* n) {
  oray r; r.o = V(o[0], o[1], o[2]); r.d = V(d[0], d[1], d[2]); r.min_t = 0.0; r.max_t = max_t; r.depth = 0;
  oisect is; is.t = INFINITY; is.prim = -1; is.bsdf = -1; is.n = V(0, 0, 0);
  int hit = bvh_closest(s, b, &r, &is);
  if (t) *t = is.t;
  if (n) { n[0] = is.n.x; n[1] = is.n.y; n[2] = is.n.z; }
  return hit ? is.prim : -1;
}
int orc_any_hit(const orc_scene* s, const orc_bvh* b, const double* o, const double* d, double max_t) {
  oray r; r.o = V(o[0], o[1], o[2]); r.d = V(d[0], d[1], d[2]); r.min_t = 0.0; r.max_t = max_t; r.depth = 0;
  return bvh_any(s, b, &r);
}
/* brute-force closest hit over all primitives in id order (no BVH) */
int orc_closest_hit_brute(const orc_scene* s, const double* o, const double* d, double* t) {
  oray r; r.o = V(o[0], o[1], o[2]); r.d = V(d[0], d[1], d[2]); r.min_t = 0.0; r.max_t = INFINITY; r.depth = 0;
  oisect is; is.t = INFINITY; is.prim = -1;
  for (int p = 0; p < s->n_prims; p++) prim_intersect(s, p, &r, &is);
  if (t) *t = is.t;
  return is.prim;
}

/* HDRImageBuffer::toColor + ImageBuffer::update_pixel (image.h:174-189, 49-58): RGBA8 packed as 0xAABBGGRR */
void orc_to_color(const float* rgb, int n_pixels, uint32_t* out) {
  float gamma = 2.2f, level = 1.0f;
  float one_over_gamma = 1.0f / gamma;
  float exposure = (float)sqrt(pow(2, level));
  for (int i = 0; i < n_pixels; i++) {
    float c[3];
    for (int k = 0; k < 3; k++) {
      float v = powf(rgb[3 * i + k] * exposure, one_over_gamma);
      c[k] = (v < 1.f) ? v : 1.f;   /* clamp(0.f, 1.f, c) == min(max(0,1), c): misc.h:70-72 with swapped arguments */
    }
    uint32_t R = (uint32_t)(uint8_t)(c[0] * 255), G = (uint32_t)(uint8_t)(c[1] * 255), B = (uint32_t)(uint8_t)(c[2] * 255);
    out[i] = R | (G << 8) | (B << 16) | (255u << 24);
  }
}
