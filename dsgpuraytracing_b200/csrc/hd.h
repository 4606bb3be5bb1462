// hd.h -- host/device portability shims.  The device functions of rng.cuh / traverse.cuh / shade.cuh are
// __host__ __device__ so that the CPU harness under tests/ (tests/cpu_walk.cu: a checker, never part of the
// product path) can walk the same wide BVH with the same code on a box without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define DSRT_HD __host__ __device__ __forceinline__
#else
#error "compile with nvcc"
#endif

namespace dsrt {

DSRT_HD float hd_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
// fma with the result clamped to [0, 1] (NaN -> 0): one FFMA.SAT on the device
DSRT_HD float hd_fma_sat(float a, float b, float c) {
#ifdef __CUDA_ARCH__
  return __saturatef(__fmaf_rn(a, b, c));
#else
  const float r = fmaf(a, b, c);
  return r > 0.0f ? (r < 1.0f ? r : 1.0f) : 0.0f;      // NaN compares false: 0
#endif
}
DSRT_HD float hd_mul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
DSRT_HD float hd_sub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
DSRT_HD double hd_dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
DSRT_HD double hd_dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}
DSRT_HD double hd_dsub(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  volatile double r = a - b; return r;
#endif
}
DSRT_HD double hd_ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
DSRT_HD double hd_dsqrt(double a) {
#ifdef __CUDA_ARCH__
  return __dsqrt_rn(a);
#else
  return sqrt(a);
#endif
}
DSRT_HD float hd_d2f_ru(double a) {
#ifdef __CUDA_ARCH__
  return __double2float_ru(a);
#else
  float f = (float)a; if ((double)f < a) f = nextafterf(f, INFINITY); return f;
#endif
}
DSRT_HD int hd_clz(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
DSRT_HD int hd_popc(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
DSRT_HD uint32_t hd_umulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
DSRT_HD float hd_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
DSRT_HD uint32_t hd_f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
DSRT_HD float hd_i2f(int i) { return hd_u2f((uint32_t)i); }
DSRT_HD int hd_f2i(float f) { return (int)hd_f2u(f); }
DSRT_HD float hd_rcp(float x) {
#ifdef __CUDA_ARCH__
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}
DSRT_HD float hd_rsqrt(float x) {
#ifdef __CUDA_ARCH__
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
DSRT_HD void hd_sincospi(float x, float* s, float* c) {
#ifdef __CUDA_ARCH__
  sincospif(x, s, c);
#else
  *s = (float)sin(3.14159265358979323846 * (double)x); *c = (float)cos(3.14159265358979323846 * (double)x);
#endif
}
template <typename T> DSRT_HD T hd_ldg(const T* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

}  // namespace dsrt
