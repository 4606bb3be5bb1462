// shade.cuh -- float restatement of the reference's shading math for the wavefront shade stage.
//   make_coord_space                      src/bsdf.cpp:13-30
//   Diffuse/Mirror/Refraction/Glass/Emission BSDF::f, sample_f, reflect, refract   src/bsdf.cpp:34-202
//   Directional/InfiniteHemisphere/Point/Area light sample_L                        src/static_scene/light.cpp:17-92
//   Camera::generate_ray                  src/camera.cpp:113-129
//   cosine / uniform hemisphere samplers  src/sampler.cpp:20-55
// Quirks of the reference estimator are kept on purpose: point lights have no 1/r^2
// falloff, glass under total internal reflection returns the TRANSMITTANCE, Fresnel choice is not weighted,
// shadow rays stop at 0.999*dist, delta lights offset the shadow origin by EPS_N along the (unnormalised)
// shading normal, Russian roulette terminates with probability max(1 - illum(f), 0).
#pragma once
#include "hd.h"
#include "layout.h"
#include "rng.cuh"
#include "traverse.cuh"   // source_code()

namespace dsrt {

constexpr float kPi = 3.14159265358979323f;
constexpr float kEpsN = 5e-3f;   // CMU462/misc.h:12

struct V3 { float x, y, z; };
DSRT_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
DSRT_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
DSRT_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
DSRT_HD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
DSRT_HD V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
DSRT_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DSRT_HD V3 cross(V3 u, V3 v) { return v3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x); }
DSRT_HD V3 normalize(V3 a) { const float r = hd_rsqrt(dot(a, a)); return r * a; }

struct Frame { V3 x, y, z; };   // columns of o2w
// make_coord_space, bsdf.cpp:13-30
DSRT_HD Frame make_coord_space(V3 n) {
  V3 h = n;
  const float ax = fabsf(h.x), ay = fabsf(h.y), az = fabsf(h.z);
  if (ax <= ay && ax <= az) h.x = 1.0f;
  else if (ay <= ax && ay <= az) h.y = 1.0f;
  else h.z = 1.0f;
  Frame f;
  f.z = normalize(n);
  f.y = normalize(cross(h, f.z));
  f.x = normalize(cross(f.z, f.y));
  return f;
}
DSRT_HD V3 to_world(const Frame& f, V3 w) { return w.x * f.x + (w.y * f.y + w.z * f.z); }
DSRT_HD V3 to_local(const Frame& f, V3 w) { return v3(dot(f.x, w), dot(f.y, w), dot(f.z, w)); }

DSRT_HD float illum(V3 s) { return 0.2126f * s.x + 0.7152f * s.y + 0.0722f * s.z; }  // spectrum.h:94-96

// CosineWeightedHemisphereSampler3D::get_sample(pdf), sampler.cpp:44-55:
// theta = acos(1-2 r1)/2  =>  cos(theta) = sqrt(1-r1), sin(theta) = sqrt(r1)  (half-angle identities)
DSRT_HD V3 cosine_hemisphere(float r1, float r2, float* pdf) {
  const float st = sqrtf(r1), ct = sqrtf(1.0f - r1);
  float sp, cp;
  hd_sincospi(2.0f * r2, &sp, &cp);
  *pdf = ct * (1.0f / kPi);
  return v3(st * cp, st * sp, ct);
}
// UniformHemisphereSampler3D::get_sample, sampler.cpp:20-31
DSRT_HD V3 uniform_hemisphere(float r1, float r2) {
  const float st = sqrtf(fmaxf(0.0f, 1.0f - r1 * r1));
  float sp, cp;
  hd_sincospi(2.0f * r2, &sp, &cp);
  return v3(st * cp, st * sp, r1);
}

// BSDF::refract, bsdf.cpp:167-191
DSRT_HD bool refract(V3 wo, V3* wi, float ior) {
  float sign = 1.0f, ratio = ior;
  if (wo.z > 0.0f) { sign = -1.0f; ratio = 1.0f / ratio; }
  const float cos2 = 1.0f - ratio * ratio * (1.0f - wo.z * wo.z);
  if (cos2 < 0.0f) { *wi = v3(-wo.x, -wo.y, wo.z); return false; }
  *wi = normalize(v3(-wo.x * ratio, -wo.y * ratio, sign * sqrtf(cos2)));
  return true;
}

// BSDF::f -- only the diffuse lobe is non-zero (bsdf.cpp:34-36; mirror/refraction/glass/emission return 0)
DSRT_HD V3 bsdf_f(const Bsdf& b) {
  if (b.type == 0) return (1.0f / kPi) * v3(b.a[0], b.a[1], b.a[2]);
  return v3(0.f, 0.f, 0.f);
}

// BSDF::sample_f for the five reference BSDFs; u = (dir0, dir1, glass choice, -)
DSRT_HD V3 bsdf_sample_f(const Bsdf& b, V3 wo, V3* wi, float* pdf, float4 u) {
  const V3 A = v3(b.a[0], b.a[1], b.a[2]), B = v3(b.b[0], b.b[1], b.b[2]);
  switch (b.type) {
    case 0: *wi = cosine_hemisphere(u.x, u.y, pdf); return (1.0f / kPi) * A;                       // bsdf.cpp:38-44
    case 1: *wi = v3(-wo.x, -wo.y, wo.z); *pdf = 1.0f; return (1.0f / fmaxf(wo.z, 1e-8f)) * A;      // bsdf.cpp:61-69
    case 2: {                                                                                      // bsdf.cpp:90-111
      *pdf = 1.0f;
      if (!refract(wo, wi, b.ior)) return v3(0.f, 0.f, 0.f);
      float ni = b.ior, no = 1.0f;
      if (wo.z < 0.0f) { const float t = ni; ni = no; no = t; }
      const float ratio = no / ni;
      return (ratio * ratio / fmaxf(fabsf(wi->z), 1e-8f)) * B;
    }
    case 3: {                                                                                      // bsdf.cpp:119-158
      *pdf = 1.0f;
      if (!refract(wo, wi, b.ior)) return (1.0f / fmaxf(fabsf(wi->z), 1e-8f)) * B;                 // TIR quirk, :128-131
      float ni = b.ior, no = 1.0f;
      const float cos_i = fabsf(wi->z), cos_o = fabsf(wo.z);
      if (wo.z < 0.0f) { const float t = ni; ni = no; no = t; }
      const float r1 = (no * cos_i - ni * cos_o) / (no * cos_i + ni * cos_o);
      const float r2 = (ni * cos_i - no * cos_o) / (ni * cos_i + no * cos_o);
      const float Fr = 0.5f * (r1 * r1 + r2 * r2);
      if (u.z <= Fr) { *wi = v3(-wo.x, -wo.y, wo.z); return (1.0f / fmaxf(fabsf(wi->z), 1e-8f)) * A; }
      const float ratio = no / ni;
      return (ratio * ratio / fmaxf(fabsf(wi->z), 1e-8f)) * B;
    }
    default: *wi = cosine_hemisphere(u.x, u.y, pdf); return v3(0.f, 0.f, 0.f);                      // emission, bsdf.cpp:199-202
  }
}

// SceneLight::sample_L; (u0,u1) = this sample's pair.  Returns radiance, fills wi / dist / pdf.
DSRT_HD V3 light_sample_L(const Light& L, V3 p, float u0, float u1, V3* wi, float* dist, float* pdf) {
  const V3 rad = v3(L.radiance[0], L.radiance[1], L.radiance[2]);
  switch (L.type) {
    case 0: *wi = v3(L.v0[0], L.v0[1], L.v0[2]); *dist = kInfF; *pdf = 1.0f; return rad;            // light.cpp:17-23
    case 1: {                                                                                      // light.cpp:34-42
      const V3 d = uniform_hemisphere(u0, u1);
      *wi = d.x * v3(L.s2w[0], L.s2w[1], L.s2w[2]) + (d.y * v3(L.s2w[3], L.s2w[4], L.s2w[5]) + d.z * v3(L.s2w[6], L.s2w[7], L.s2w[8]));
      *dist = kInfF; *pdf = 1.0f / (2.0f * kPi); return rad;
    }
    case 2: {                                                                                      // light.cpp:49-57
      const V3 d = v3(L.v0[0], L.v0[1], L.v0[2]) - p;
      const float n2 = dot(d, d);
      *wi = hd_rsqrt(n2) * d; *dist = sqrtf(n2); *pdf = 1.0f; return rad;
    }
    default: {                                                                                     // light.cpp:80-92
      const float sx = u0 - 0.5f, sy = u1 - 0.5f;
      const V3 d = v3(L.v0[0], L.v0[1], L.v0[2]) + sx * v3(L.dim_x[0], L.dim_x[1], L.dim_x[2]) +
                   sy * v3(L.dim_y[0], L.dim_y[1], L.dim_y[2]) - p;
      const float cosTheta = dot(d, v3(L.dir[0], L.dir[1], L.dir[2]));
      const float sqDist = dot(d, d);
      const float dst = sqrtf(sqDist);
      *wi = (1.0f / dst) * d; *dist = dst;
      *pdf = sqDist / (L.area * fabsf(cosTheta));
      return cosTheta < 0.0f ? rad : v3(0.f, 0.f, 0.f);
    }
  }
}

// ---- EnvironmentLight (environment_light.cpp:71-199), float restatement -------------------------------------------
DSRT_HD int env_lower_bound(const float* a, int n, float v) {      // std::lower_bound
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (a[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}
// sample_dir: bilinear lookup with wrap-around (:129-199)
DSRT_HD V3 env_sample_dir(const EnvMap& e, V3 d) {
  const int w = e.w, h = e.h;
  const float dy = fminf(fmaxf(d.y, -1.0f), 1.0f);
  const float theta = acosf(dy);
  const float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - dy * dy));
  float phi = sin_theta == 0.0f ? kPi : acosf(fminf(fmaxf(d.z / sin_theta, -1.0f), 1.0f));
  if (d.x > 0.0f) phi = 2.0f * kPi - phi;
  const float tu = phi / (2.0f * kPi) * (float)w - 0.5f, tv = theta / kPi * (float)h - 0.5f;
  const int su = (int)tu, sv = (int)tv;
  float a, b; int px1, px2, py1, py2;
  if (tu < 0.0f) { a = tu + 1.0f; px1 = w - 1; px2 = 0; } else if (tu >= (float)(w - 1)) { a = tu - (float)w + 1.0f; px1 = w - 1; px2 = 0; } else { a = tu - (float)su; px1 = su; px2 = su + 1; }
  if (tv < 0.0f) { b = tv + 1.0f; py1 = h - 1; py2 = 0; } else if (tv >= (float)(h - 1)) { b = tv - (float)h + 1.0f; py1 = h - 1; py2 = 0; } else { b = tv - (float)sv; py1 = sv; py2 = sv + 1; }
  const float* p11 = e.rgb + 3 * (px1 + w * py1); const float* p21 = e.rgb + 3 * (px2 + w * py1);
  const float* p12 = e.rgb + 3 * (px1 + w * py2); const float* p22 = e.rgb + 3 * (px2 + w * py2);
  const V3 zy1 = (1.0f - a) * v3(p11[0], p11[1], p11[2]) + a * v3(p21[0], p21[1], p21[2]);
  const V3 zy2 = (1.0f - a) * v3(p12[0], p12[1], p12[2]) + a * v3(p22[0], p22[1], p22[2]);
  return (1.0f - b) * zy1 + b * zy2;
}
// importanceSampling (:71-113): inverse-CDF over theta rows, then over phi within the row, piecewise linear
DSRT_HD void env_importance(const EnvMap& e, float r1, float r2, V3* wi, float* pdf) {
  const int w = e.w, h = e.h;
  r1 *= e.pTheta[h - 1];
  int t = env_lower_bound(e.pTheta, h, r1); if (t >= h) t = h - 1;
  float prev = t > 0 ? e.pTheta[t - 1] : 0.0f;
  const float y = (float)t + (r1 - prev) / (e.pTheta[t] - prev);
  const float theta = fminf(y / (float)h, 1.0f) * kPi;
  const float* row = e.pPhiGivenTheta + (size_t)t * w;
  r2 *= row[w - 1];
  int q = env_lower_bound(row, w, r2); if (q >= w) q = w - 1;
  prev = q > 0 ? row[q - 1] : 0.0f;
  const float x = (float)q + (r2 - prev) / (row[q] - prev);
  const float phi = fminf(x / (float)w, 1.0f) * 2.0f * kPi;
  const float st = sinf(theta), ct = cosf(theta);
  *pdf = e.pThetaPhi[(size_t)t * w + q] / (st * (2.0f * kPi / (float)w) * (kPi / (float)h));
  *wi = v3(-st * sinf(phi), ct, st * cosf(phi));
}

// Camera::generate_ray, camera.cpp:113-129 (x,y in [0,1])
DSRT_HD void generate_ray(const Camera& c, float x, float y, V3* o, V3* d) {
  const V3 sp = v3(-(x - 0.5f) * c.w_over_dist, -(y - 0.5f) * c.h_over_dist, 1.0f);
  const V3 cx = v3(c.c2w[0], c.c2w[1], c.c2w[2]), cy = v3(c.c2w[3], c.c2w[4], c.c2w[5]), cz = v3(c.c2w[6], c.c2w[7], c.c2w[8]);
  const V3 w = sp.x * cx + (sp.y * cy + sp.z * cz);
  *o = w + v3(c.pos[0], c.pos[1], c.pos[2]);
  *d = normalize(v3(-w.x, -w.y, -w.z));
}

// ---- one path vertex ---------------------------------------------------------------------------------------
struct SceneDev {
  const Bsdf* bsdf;
  const Light* lights;
  const float4* shade;   // 3 float4 per slot: vertex normals
  int n_lights;
  int n_light_samples;   // sum over lights of samples per vertex
  EnvMap env;            // env.w == 0 when there is no EnvironmentLight
};
struct PathIn { float4 ray_o, ray_d, thr, hit; uint32_t pix, smp; };
struct PathOut {
  V3 emission; bool has_emission;       // throughput * Le, to be added to the pixel
  bool cont; float4 new_o, new_d, new_thr;
};

// PathTracer::trace_ray after a successful closest-hit query (pathtracer.cpp:435-552) for ONE path vertex:
// emission (if includeLe), one shadow ray per light sample handed to `sink.shadow(j, a, b, c)`, then
// BSDF::sample_f + Russian roulette for the continuation ray.  Shared by k_shade and the CPU harness in tests/.
template <class Sink>
DSRT_HD void shade_path(const PathIn& in, const float4* __restrict__ prims, const SceneDev& sc, uint32_t seed, int max_depth,
                        int depth, PathOut& out, Sink& sink) {
  const int slot = hd_f2i(in.hit.w);
  const V3 rd = v3(in.ray_d.x, in.ray_d.y, in.ray_d.z);
  const V3 thr = v3(in.thr.x, in.thr.y, in.thr.z);
  const int include_le = (hd_f2i(in.thr.w) >> 8) & 1;
  const float4 a = hd_ldg(prims + 3 * (size_t)slot), b = hd_ldg(prims + 3 * (size_t)slot + 1), c = hd_ldg(prims + 3 * (size_t)slot + 2);
  const Bsdf bs = sc.bsdf[hd_f2i(c.w)];
  const int src_code = source_code(slot, b.w != 0.0f);      // travels with every ray that leaves this primitive (traverse.cuh)
  V3 hit_p, n_sh;
  if (b.w != 0.0f) {
    // hit point from the barycentrics (stays on the triangle's plane to float precision);
    // n = (1-u-v) n1 + u n2 + v n3, unnormalised, flipped against the ray (triangle.cpp:94-99)
    const V3 p1 = v3(a.x, a.y, a.z), p2 = v3(b.x, b.y, b.z), p3 = v3(c.x, c.y, c.z);
    hit_p = p1 + (in.hit.y * (p2 - p1) + in.hit.z * (p3 - p1));
    const float4 n1 = hd_ldg(sc.shade + 3 * (size_t)slot), n2 = hd_ldg(sc.shade + 3 * (size_t)slot + 1), n3 = hd_ldg(sc.shade + 3 * (size_t)slot + 2);
    n_sh = (1.0f - in.hit.y - in.hit.z) * v3(n1.x, n1.y, n1.z) + (in.hit.y * v3(n2.x, n2.y, n2.z) + in.hit.z * v3(n3.x, n3.y, n3.z));
    if (dot(rd, n_sh) > 0.0f) n_sh = v3(-n_sh.x, -n_sh.y, -n_sh.z);
  } else {
    // outward unit normal, not flipped (sphere.cpp:69-72); the hit point is re-projected onto the sphere
    const V3 cc = v3(a.x, a.y, a.z);
    n_sh = normalize(v3(in.ray_o.x, in.ray_o.y, in.ray_o.z) + in.hit.x * rd - cc);
    hit_p = cc + b.x * n_sh;
  }
  out.has_emission = include_le && bs.type == 4;                                                  // pathtracer.cpp:435
  out.emission = out.has_emission ? thr * v3(bs.a[0], bs.a[1], bs.a[2]) : v3(0.f, 0.f, 0.f);
  const Frame fr = make_coord_space(n_sh);
  const V3 w_out = normalize(to_local(fr, v3(-rd.x, -rd.y, -rd.z)));                              // pathtracer.cpp:455-456

  // direct lighting: one shadow ray per light sample (pathtracer.cpp:469-523)
  const V3 f_direct = bsdf_f(bs);
  for (int l = 0; l < sc.n_lights; l++) {
    const Light& L = sc.lights[l];
    const float scale = 1.0f / (float)L.n_samples;
    for (int i = 0; i < L.n_samples; i++) {
      const int j = L.sample_base + i;
      float u0 = 0.f, u1 = 0.f;
      if (L.type == 1 || L.type >= 3) {
        const float4 u = rng_block(seed, in.pix, in.smp, (uint32_t)depth, kBlockLight0 + (uint32_t)(j >> 1));
        u0 = (j & 1) ? u.z : u.x; u1 = (j & 1) ? u.w : u.y;
      }
      V3 wi; float dist, pdf; V3 light_L;
      if (L.type == 4) {                          // EnvironmentLight::sample_L, environment_light.cpp:115-127
        env_importance(sc.env, u0, u1, &wi, &pdf); dist = kInfF;
        light_L = env_sample_dir(sc.env, wi);
      } else {
        light_L = light_sample_L(L, hit_p, u0, u1, &wi, &dist, &pdf);
      }
      const V3 w_in = normalize(to_local(fr, wi));
      const float cos_theta = fmaxf(0.0f, w_in.z);
      const V3 contrib = (cos_theta / pdf * scale) * (thr * (light_L * f_direct));
      const V3 so = L.is_delta ? hit_p + kEpsN * n_sh : hit_p;                                    // pathtracer.cpp:497-500
      sink.shadow(j, make_float4(so.x, so.y, so.z, dist * 0.999f), make_float4(wi.x, wi.y, wi.z, hd_i2f(src_code)),
                  make_float4(contrib.x, contrib.y, contrib.z, hd_u2f(in.pix)));
    }
  }

  // indirect: BSDF sample + Russian roulette (pathtracer.cpp:527-552)
  out.cont = false;
  if (depth < max_depth) {
    const float4 u = rng_block(seed, in.pix, in.smp, (uint32_t)depth, kBlockBsdf);
    V3 w_in; float pdf = 1.f;
    const V3 f = bsdf_sample_f(bs, w_out, &w_in, &pdf, u);
    const float cos_theta = fabsf(w_in.z);
    const float p_term = fmaxf(1.0f - illum(f), 0.0f);
    if (!(u.w < p_term)) {
      const V3 nd = normalize(to_world(fr, w_in));
      const V3 nthr = (cos_theta / (pdf * (1.0f - p_term))) * (thr * f);
      const int next_le = (bs.type == 1 || bs.type == 2 || bs.type == 3) ? 1 : 0;                 // BSDF::is_delta
      out.cont = true;
      out.new_o = make_float4(hit_p.x, hit_p.y, hit_p.z, kInfF);
      out.new_d = make_float4(nd.x, nd.y, nd.z, hd_i2f(src_code));
      out.new_thr = make_float4(nthr.x, nthr.y, nthr.z, hd_i2f((depth + 1) | (next_le << 8)));
    }
  }
}

}  // namespace dsrt
