// rb_common.cuh — vector math, RNG and records shared by every kernel.
//
// All per-pixel logic is written as __host__ __device__ functions so that the
// same source can also be compiled for the host by the test-only emulation
// harness (tests/emu): that harness is never linked into the product library.
//
// Arithmetic contract: no implicit FMA contraction (nvcc -fmad=false), IEEE
// division and square root (nvcc defaults), glm operation order. Explicit fused
// operations are spelled fmaf_() and appear only where the oracle spells the
// same (ray/triangle test) or where exact parity is not required (box culling).
#ifndef RB_COMMON_CUH_
#define RB_COMMON_CUH_

#include <float.h>
#include <math.h>
#include <stdint.h>

#include "../../include/restir_b200.h"
#include "det_math.h"

#if defined(__CUDACC__)
#define RB_HD __host__ __device__ __forceinline__
#define RB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define RB_HD inline
#define RB_HD_NOINLINE inline
#endif

namespace rb {

struct V3 {
  float x, y, z;
};
struct alignas(16) F4 {
  float x, y, z, w;
};
struct alignas(16) U4 {
  uint32_t x, y, z, w;
};
struct alignas(8) U2 {
  uint32_t x, y;
};

RB_HD float fmaf_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return __builtin_fmaf(a, b, c);
#endif
}
RB_HD float sqrtf_(float x) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(x);
#else
  return __builtin_sqrtf(x);
#endif
}
RB_HD float fdiv_(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
RB_HD float frcp_(float x) {  // 1.0f / x, correctly rounded (same bits as the division, shorter on the device)
#if defined(__CUDA_ARCH__)
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}
RB_HD float fabsf_(float x) { return dm::u2f(dm::f2u(x) & 0x7FFFFFFFu); }
RB_HD uint32_t f2u(float x) { return dm::f2u(x); }
RB_HD float u2f(uint32_t x) { return dm::u2f(x); }

RB_HD V3 v3(float a) { return {a, a, a}; }
RB_HD V3 v3(float x, float y, float z) { return {x, y, z}; }
RB_HD V3 xyz(const F4& a) { return {a.x, a.y, a.z}; }
RB_HD F4 f4(const V3& a, float w) { return {a.x, a.y, a.z, w}; }
RB_HD V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
RB_HD V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
RB_HD V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
RB_HD V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
RB_HD V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
RB_HD V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
// glm::dot / cross / length / normalize / reflect, P/glm/detail/func_geometric.inl:48-110
RB_HD float dot(V3 a, V3 b) {
  V3 t = a * b;
  return t.x + t.y + t.z;
}
RB_HD V3 cross(V3 x, V3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
RB_HD float length(V3 v) { return sqrtf_(dot(v, v)); }
RB_HD V3 normalize(V3 v) { return v * frcp_(sqrtf_(dot(v, v))); }
RB_HD V3 reflect(V3 I, V3 N) { return I - N * dot(N, I) * 2.0f; }
RB_HD float gmax(float a, float b) { return (a < b) ? b : a; }  // glm::max
RB_HD float gmin(float a, float b) { return (b < a) ? b : a; }  // glm::min
RB_HD float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
RB_HD int imin(int a, int b) { return a < b ? a : b; }
RB_HD int lowest_bit(uint32_t m) {  // index of the lowest set bit, m != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)m) - 1;
#else
  return __builtin_ctz(m);
#endif
}
RB_HD int imax(int a, int b) { return a > b ? a : b; }

#define RB_PI 3.14159265358979323846264338327950288f
#define RB_TWO_PI 6.28318530717958647692528676655900576f
#define RB_ROOT_PI 1.772453850905516027f
#define RB_ONE_OVER_PI 0.318309886183790671537767526745028724f
#define RB_ONE_OVER_TWO_PI 0.159154943091895335768883763372514362f
#define RB_NEG_FLT_MAX (-FLT_MAX)

// ---- counter RNG (SURVEY §8c seam 2; same construction as the oracle's) ---------
enum Pass : uint32_t { PASS_GBUF = 0, PASS_INITIAL = 1, PASS_TEMPORAL = 2, PASS_SPATIAL = 3, PASS_MIS = 4 };

RB_HD uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
// frame_key folds (seed, frame, pass, iter); computed once per launch on the host
RB_HD uint32_t rng_frame_key(uint32_t seed, uint32_t frame, uint32_t pass, uint32_t iter) {
  uint32_t h = fmix32(seed ^ 0x9E3779B9u);
  h = fmix32(h ^ (frame * 0x85EBCA77u + 0x165667B1u));
  h = fmix32(h ^ ((pass * 64u + iter) * 0xC2B2AE3Du + 0x27D4EB2Fu));
  return h;
}
RB_HD uint32_t rng_pixel_key(uint32_t frame_key, uint32_t pixel) { return fmix32(frame_key ^ (pixel * 0x9E3779B1u)); }
RB_HD uint32_t rng_bits(uint32_t key, uint32_t slot) { return fmix32(key + slot * 0x9E3779B9u); }
RB_HD float bits_to_unit(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }
// Utils::getRandomValue(a,b), P/utils.cpp:199-202
RB_HD float rng_value(uint32_t key, uint32_t slot, float a, float b) { return a + (b - a) * bits_to_unit(rng_bits(key, slot)); }

// ---- records ----------------------------------------------------------------------
// LightSample / Reservoir, P/Reservoir.h:6-60 (+ lightIdx instrumentation)
struct LightSample {
  V3 samplePoint, sampleNormal, L_i;
  int lightIdx;
};
RB_HD LightSample invalid_sample() {
  LightSample s;
  s.samplePoint = v3(RB_NEG_FLT_MAX);
  s.sampleNormal = v3(RB_NEG_FLT_MAX);
  s.L_i = v3(RB_NEG_FLT_MAX);
  s.lightIdx = -1;
  return s;
}
RB_HD bool sample_valid(const LightSample& s) {  // LightSample::isValid, :11-17
  bool pointOk = s.samplePoint.x != RB_NEG_FLT_MAX && s.samplePoint.y != RB_NEG_FLT_MAX && s.samplePoint.z != RB_NEG_FLT_MAX;
  bool normalOk = s.sampleNormal.x != RB_NEG_FLT_MAX && s.sampleNormal.y != RB_NEG_FLT_MAX && s.sampleNormal.z != RB_NEG_FLT_MAX;
  bool L_iOk = s.L_i.x > 0 || s.L_i.y > 0 || s.L_i.z > 0;
  return pointOk && normalOk && L_iOk;
}
struct Reservoir {
  LightSample bestSample;
  float w_sum, W;
  int confidence;
};
RB_HD Reservoir empty_reservoir() {
  Reservoir r;
  r.bestSample = invalid_sample();
  r.w_sum = 0;
  r.W = 0;
  r.confidence = 0;
  return r;
}
// Reservoir::addSample, :33-47
RB_HD bool add_sample(Reservoir& r, const LightSample& s, float w, int c, uint32_t key, uint32_t slot) {
  r.w_sum += w;
  r.confidence += c;
  if (w == 0 && r.w_sum == 0) return false;
  if (rng_value(key, slot, 0, 1) < w / r.w_sum) {
    r.bestSample = s;
    return true;
  }
  return false;
}

// GBufferElement, P/GBufferElement.h:6-23 (+ cached 1/I_M and the emissive flag)
struct GElem {
  V3 pos, normal, diffuse, specular, emission;
  float shininess, depth, invIM;
  uint32_t matType;
  bool isEmissive;
};
RB_HD bool emissive3(const V3& e) { return e.x > 0 || e.y > 0 || e.z > 0; }

// SoA planes in HBM, 16-byte records for 128-bit coalesced access (DESIGN.md "Data layout")
struct GBufPlanes {
  F4* pos_depth;    // {pos.xyz, depth}
  F4* normal_shin;  // {normal.xyz, shininess}
  F4* diffuse_iim;  // {diffuse.rgb, 1/I_M}
  F4* spec_type;    // {specular.rgb, bits: matType | emissive<<8}
  F4* emission;     // {emission.rgb, 0}   (read only where the emissive bit is set)
  U2* hit_ids;      // {geomID, primID}
};
struct ResPlanes {
  F4* point_wsum;  // {samplePoint.xyz, w_sum}
  F4* normal_W;    // {sampleNormal.xyz, W}
  F4* Li_conf;     // {L_i.rgb, bits(confidence)}
  int* light_idx;
};

struct CamState {  // what GBuffer keeps of the camera, P/GBufferElement.h:136-139
  V3 pos;
  float focal;
  float viewMat[16];
  float invViewMat[16];
};

RB_HD F4 ldg4(const F4* p) {
#if defined(__CUDA_ARCH__)
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return {v.x, v.y, v.z, v.w};
#else
  return *p;
#endif
}
RB_HD U2 ldg2(const U2* p) {
#if defined(__CUDA_ARCH__)
  uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  return {v.x, v.y};
#else
  return *p;
#endif
}
RB_HD F4 ld4(const F4* p) {
#if defined(__CUDA_ARCH__)
  float4 v = *reinterpret_cast<const float4*>(p);
  return {v.x, v.y, v.z, v.w};
#else
  return *p;
#endif
}
RB_HD void st4(F4* p, const F4& v) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<float4*>(p) = make_float4(v.x, v.y, v.z, v.w);
#else
  *p = v;
#endif
}

RB_HD Reservoir load_reservoir(const ResPlanes& R, size_t i) {
  F4 a = ld4(R.point_wsum + i), b = ld4(R.normal_W + i), c = ld4(R.Li_conf + i);
  Reservoir r;
  r.bestSample.samplePoint = xyz(a);
  r.w_sum = a.w;
  r.bestSample.sampleNormal = xyz(b);
  r.W = b.w;
  r.bestSample.L_i = xyz(c);
  r.confidence = (int)f2u(c.w);
  r.bestSample.lightIdx = R.light_idx[i];
  return r;
}
RB_HD void store_reservoir(const ResPlanes& R, size_t i, const Reservoir& r) {
  st4(R.point_wsum + i, f4(r.bestSample.samplePoint, r.w_sum));
  st4(R.normal_W + i, f4(r.bestSample.sampleNormal, r.W));
  st4(R.Li_conf + i, f4(r.bestSample.L_i, u2f((uint32_t)r.confidence)));
  R.light_idx[i] = r.bestSample.lightIdx;
}
RB_HD GElem load_gelem(const GBufPlanes& G, size_t i) {
  F4 a = ld4(G.pos_depth + i), b = ld4(G.normal_shin + i), c = ld4(G.diffuse_iim + i), d = ld4(G.spec_type + i);
  GElem e;
  e.pos = xyz(a);
  e.depth = a.w;
  e.normal = xyz(b);
  e.shininess = b.w;
  e.diffuse = xyz(c);
  e.invIM = c.w;
  e.specular = xyz(d);
  uint32_t bits = f2u(d.w);
  e.matType = bits & 0xFFu;
  e.isEmissive = (bits & 0x100u) != 0;
  e.emission = v3(0);
  if (e.isEmissive) e.emission = xyz(ld4(G.emission + i));
  return e;
}
RB_HD void store_gelem(const GBufPlanes& G, size_t i, const GElem& e, uint32_t geomID, uint32_t primID) {
  st4(G.pos_depth + i, f4(e.pos, e.depth));
  st4(G.normal_shin + i, f4(e.normal, e.shininess));
  st4(G.diffuse_iim + i, f4(e.diffuse, e.invIM));
  st4(G.spec_type + i, f4(e.specular, u2f((e.matType & 0xFFu) | (e.isEmissive ? 0x100u : 0u))));
  st4(G.emission + i, f4(e.emission, 0.0f));
  G.hit_ids[i] = U2{geomID, primID};
}

}  // namespace rb
#endif
