// layout.h -- HBM data layout shared by the host flattening code and the sm_100a kernels.
//
// Wide node: 8-wide BVH node, 80 bytes of information (compressed wide-BVH layout after
// Ylitie, Karras, Laine 2017).  Child boxes are quantised to 8 bits per plane relative to the node's
// own box: lo = origin + qlo * 2^(e-127), hi = origin + qhi * 2^(e-127) (the exponent bytes ex, ey, ez hold e + 15: the
// node test multiplies by 2^15, traverse.cuh byte_unit), rounded OUTWARDS so every
// quantised box encloses the exact double-precision box of the reference's binary SAH tree.
// Internal children are stored contiguously from child_base in slot order; the primitives of all
// leaf children are stored contiguously from prim_base (leaf-contiguous primitive order, ascending slot).
//
// The node test (traverse.cuh) answers with one NIBBLE per slot (slot s -> bits 4s..4s+3, all four set when the ray meets
// the child's box), so that a hit costs one predicated OR with a constant.  Two 32-bit words of the node turn that into the
// two things the traversal needs:
//   valid  nibble s = the low c bits set when slot s is a leaf child with c primitives (c = 1..3), 0 otherwise;
//          (hits & valid) is the mask of primitives to test, and primitive bit k is record
//          prim_base + popc(valid & ((1 << k) - 1))
//   inner  bit 4s+3 set when slot s is an internal child; (hits & inner) are the children to open, (inner >> 3) is the
//          slot-ordered internal-child mask used to index them from child_base
// imask (bit s = slot s is internal) is kept for host-side consumers.
#pragma once
#include <stdint.h>

namespace dsrt {

constexpr float kInfF = __builtin_huge_valf();

// DSRT_NODE96 (default): the node is padded to 96 bytes = three 256-bit loads (sm_100 LDG.256) on 32-byte boundaries.  An
// 80-byte node already occupies three 32-byte sectors wherever it lies, so the padding costs no cache or DRAM traffic; it
// takes two load instructions (and two passes through the LSU data pipe) off every node visit, and the spare words hold the
// three plane scales 2^15 * 2^(e-127) as ready-made floats (no exponent decoding in the node test).
#ifndef DSRT_NODE96
#define DSRT_NODE96 1
#endif
struct alignas(DSRT_NODE96 ? 32 : 16) WideNode {
  float ox, oy, oz;
  uint8_t ex, ey, ez, imask;
#if DSRT_NODE96
  float sx, sy, sz;              // 2^15 * 2^(e-127) per axis: the floats the exponent bytes encode
  uint32_t flat;                 // nibble s = 0xf when leaf slot s holds two or three COPLANAR triangles (the halves of a wall
                                 // quad): a ray that starts on one of them cannot hit the others, so drop_source (traverse.cuh)
                                 // takes the whole slot out of the hit mask; set by mark_flat_slots (wide_bvh.cpp), 0 otherwise
#endif
  uint32_t prim_base;            // (prim_base, valid) are re-read as one 64-bit word when a lane turns hit bits into records
  uint32_t valid;
  uint32_t child_base;
  uint32_t inner;
  uint8_t qlox[8], qloy[8];
  uint8_t qloz[8], qhix[8];
  uint8_t qhiy[8], qhiz[8];
};
static_assert(sizeof(WideNode) == (DSRT_NODE96 ? 96 : 80), "WideNode must be 80 (96) bytes");
constexpr int kNodeQuads = (int)(sizeof(WideNode) / 16);       // 128-bit words per node
constexpr int kNodeInfoBytes = 80;                             // bytes of information per node (what the roofline counts)

// Primitive record in leaf-contiguous order, 48 bytes = three 128-bit loads.
//   triangle: a = (p1.xyz, prim_id bits)  b = (p2.xyz, 1.0f)      c = (p3.xyz, bsdf bits)
//   sphere:   a = (centre.xyz, prim_id)   b = (r, r*r, 0, 0.0f)   c = (0,0,0, bsdf bits)
struct alignas(16) PrimRecord {
  float ax, ay, az; int32_t prim_id;
  float bx, by, bz; float is_tri;
  float cx, cy, cz; int32_t bsdf;
};
static_assert(sizeof(PrimRecord) == 48, "PrimRecord must be 48 bytes");

// Shading record indexed by SLOT (same order as PrimRecord): vertex normals, 48 bytes.
struct alignas(16) ShadeRecord {
  float n1x, n1y, n1z, pad0;
  float n2x, n2y, n2z, pad1;
  float n3x, n3y, n3z, pad2;
};
static_assert(sizeof(ShadeRecord) == 48, "ShadeRecord must be 48 bytes");

// fp64 primitive record for the parity kernel (reference arithmetic), by slot: 9 doubles + 3 spare
struct alignas(16) PrimRecord64 {
  double p[9];      // triangle: p1,p2,p3   sphere: centre, r, r2(computed as r*r in double), 0...
  double pad[3];
};
static_assert(sizeof(PrimRecord64) == 96, "PrimRecord64 must be 96 bytes");

struct Bsdf {       // 32 bytes
  float a[3];       // albedo | reflectance | radiance
  float b[3];       // transmittance
  float ior;
  int32_t type;     // 0 diffuse 1 mirror 2 refraction 3 glass 4 emission
};

struct Light {      // float restatement of light_param
  float radiance[3];
  float v0[3];      // dirToLight | position
  float dir[3];
  float dim_x[3];
  float dim_y[3];
  float area;
  float s2w[9];     // column-major
  int32_t type;     // 0 directional 1 hemisphere 2 point 3 area
  int32_t is_delta;
  int32_t sample_base;  // first flat light-sample index j of this light
  int32_t n_samples;
};

// Lat-long environment map + the importance-sampling tables of EnvironmentLight (environment_light.cpp:6-53)
struct EnvMap {
  const float* rgb;             // h*w*3
  const float* pThetaPhi;       // h*w   joint pdf (normalised illum * sin theta)
  const float* pTheta;          // h     marginal CDF
  const float* pPhiGivenTheta;  // h*w   conditional CDFs
  int w, h;                     // w == 0: no environment light
};

struct Camera {
  float pos[3];
  float c2w[9];     // column-major
  float w_over_dist, h_over_dist;
  double pos64[3], c2w64[9], W64, H64, dist64;
  int32_t width, height;
};

}  // namespace dsrt
