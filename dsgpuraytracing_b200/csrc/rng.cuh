// rng.cuh -- counter-based Philox4x32-10 streams for the wavefront path tracer.
//
// key     = (seed, 0x5EED)
// counter = (pixel index y*W+x, camera-sample index, path depth, block)
// block 0: camera jitter (u0,u1)                          -- UniformGridSampler2D, sampler.cpp:7-16
// block 1: BSDF {direction u0,u1; glass reflect/refract u2; Russian roulette u3}
//                                                          -- sampler.cpp:44-55, bsdf.cpp:147, pathtracer.cpp:539
// block 2 + j/2: light sample j (flat index over lights), halves {u0,u1} / {u2,u3}
//                                                          -- light.cpp:37, 83
// u = (x >> 8) * 2^-24 in [0,1): exactly representable in float, so the CPU oracle (oracle/pt_oracle.c,
// orc_philox) feeds bit-identical uniforms to its double-precision restatement of the reference.
// Because the stream is a pure function of (seed, pixel, sample, depth, block), a frame is reproducible
// across batch sizes and across the multi-GPU sample split.
#pragma once
#include <stdint.h>

#include "hd.h"

namespace dsrt {

constexpr uint32_t kBlockCamera = 0;
constexpr uint32_t kBlockBsdf = 1;
constexpr uint32_t kBlockLight0 = 2;

DSRT_HD uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = hd_umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = hd_umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

DSRT_HD float u24(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

DSRT_HD float4 rng_block(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t depth,
                                            uint32_t block) {
  const uint4 r = philox4x32_10(make_uint4(pixel, sample, depth, block), seed, 0x5EEDu);
  return make_float4(u24(r.x), u24(r.y), u24(r.z), u24(r.w));
}

}  // namespace dsrt
