// restir_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain C++17 restatement of the reference's ReSTIR DI frame loop
// (Tonz24/restir-embree; P/ = template/src/pg/pg1_embree/), function by
// function, each citing the reference file:line it follows.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product (restir_embree_b200/csrc) never does.
//
// PARITY PIN STATUS: the reference ships no tests or golden vectors
//  and its Embree binary is Windows-only, so end-to-end parity is
// "unpinned by the reference".  What IS pinned: the leaf functions restated
// here (triangle sampling, disk sampling, light CDF, cosine / cosine-lobe
// distributions, Phong evalBRDF / evalPdf / sampleBRDF / calc_I_M incl. the
// Boost incomplete beta, sanitize, the mt19937 stream) are checked against the
// reference's OWN sources compiled in place (oracle/_ref, see oracle/Makefile
// and tests/test_ref_pin.py, fixtures in tests/golden/).
//
// Seams, all selectable at run time:
//   rng_mode   0 counter : stateless hash keyed (seed, frame, pass, iter, pixel, slot)
//              1 legacy  : serial std::mt19937{123} + uniform_real_distribution<float>
//                          in y-outer/x-inner order, as the reference's _DEBUG build
//                          (P/utils.cpp:175-176,199-202)
//   math_mode  0 det     : restir_embree_b200/csrc/det_math.h (bit-reproducible on GPU)
//              1 libm    : std::pow/cos/sin/lgamma/exp as the reference calls them
//   tracer     0 brute   : Möller–Trumbore over all triangles (ground truth)
//              1 bvh2    : CPU BVH2 with conservative culling (same hit set)
//   light sampler via RbParams.lightSampler (cdf = reference / alias).
//
// Ray/triangle arithmetic contract (Embree 3.13.5 is a closed binary — module
// named in DESIGN.md; this is the published Möller–Trumbore algorithm with a
// fixed operation order, see tri_test below). Closest hit = smallest t, ties
// broken by smaller (geomID, primID).

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/restir_b200.h"
#include "../restir_embree_b200/csrc/det_math.h"

#ifdef ORACLE_HAVE_BOOST
#include <boost/math/special_functions/beta.hpp>
#endif

namespace orc {

// ----------------------------------------------------------------------------
// glm-order vector math (P/glm/detail/func_geometric.inl:48-110,
// type_mat3x3.inl:468-474, type_mat4x4.inl:561-571)
// ----------------------------------------------------------------------------
struct V3 {
  float x, y, z;
};
static inline V3 v3(float a) { return {a, a, a}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline float dot(V3 a, V3 b) {
  V3 t = a * b;
  return t.x + t.y + t.z;
}
static inline V3 cross(V3 x, V3 y) {
  return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}
static inline float length(V3 v) { return std::sqrt(dot(v, v)); }
static inline V3 normalize(V3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }
static inline V3 reflect(V3 I, V3 N) { return I - N * dot(N, I) * 2.0f; }
static inline float gmax(float a, float b) { return (a < b) ? b : a; }  // glm::max
static inline float gmin(float a, float b) { return (b < a) ? b : a; }  // glm::min
static inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }

static const float kPi = 3.14159265358979323846264338327950288f;
static const float kTwoPi = 6.28318530717958647692528676655900576f;
static const float kRootPi = 1.772453850905516027f;
static const float kOneOverPi = 0.318309886183790671537767526745028724f;
static const float kOneOverTwoPi = 0.159154943091895335768883763372514362f;

// ----------------------------------------------------------------------------
// math seam
// ----------------------------------------------------------------------------
struct Math {
  int mode = 0;
  float pow(float x, float y) const { return mode ? std::pow(x, y) : dm::powf_(x, y); }
  float cos(float x) const { return mode ? std::cos(x) : dm::cosf_(x); }
  float sin(float x) const { return mode ? std::sin(x) : dm::sinf_(x); }
  float exp(float x) const { return mode ? std::exp(x) : dm::expf_(x); }
  float lgamma(float x) const { return mode ? std::lgamma(x) : dm::lgammaf_(x); }
  float atan2(float y, float x) const { return mode ? ::atan2f(y, x) : dm::atan2f_(y, x); }  // P/SphericalMap.cpp:11
  float acos(float x) const { return mode ? ::acosf(x) : dm::acosf_(x); }  // :12, the float overload (MSVC's global acos(float))
  float ibeta(float x, float a, float b) const {  // MaterialPhong::ibeta(x,a,b), P/MaterialPhong.cpp:246-248
#ifdef ORACLE_HAVE_BOOST
    if (mode) return boost::math::beta(a, b, x);
#endif
    return dm::ibetaf_(a, b, x);
  }
};

// ----------------------------------------------------------------------------
// RNG seam
// ----------------------------------------------------------------------------
enum Pass { PASS_GBUF = 0, PASS_INITIAL = 1, PASS_TEMPORAL = 2, PASS_SPATIAL = 3, PASS_MIS = 4 };

static inline uint32_t fmix32(uint32_t h) {  // murmur3 finaliser
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
// key of one pixel in one pass of one frame
static inline uint32_t rng_key(uint32_t seed, uint32_t frame, uint32_t pass, uint32_t iter, uint32_t pixel) {
  uint32_t h = fmix32(seed ^ 0x9E3779B9u);
  h = fmix32(h ^ (frame * 0x85EBCA77u + 0x165667B1u));
  h = fmix32(h ^ ((pass * 64u + iter) * 0xC2B2AE3Du + 0x27D4EB2Fu));
  h = fmix32(h ^ (pixel * 0x9E3779B1u));
  return h;
}
static inline uint32_t rng_bits(uint32_t key, uint32_t slot) { return fmix32(key + slot * 0x9E3779B9u); }
static inline float bits_to_unit(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

struct Rng {
  int mode = 0;  // 0 counter, 1 legacy
  uint32_t key = 0;
  std::mt19937* gen = nullptr;
  std::uniform_real_distribution<float>* dist = nullptr;
  // Utils::getRandomValue(0,1) value, P/utils.cpp:199-202
  float unit(uint32_t slot) { return mode ? (*dist)(*gen) : bits_to_unit(rng_bits(key, slot)); }
  uint32_t raw(uint32_t slot) { return mode ? (*gen)() : rng_bits(key, slot); }
  float value(uint32_t slot, float a, float b) { return a + (b - a) * unit(slot); }
};

// ----------------------------------------------------------------------------
// Scene (what ModelLoader::loadScene produces, P/ModelLoader.cpp:218-321)
// ----------------------------------------------------------------------------
struct Tri {
  V3 p0, p1, p2;
  V3 n0, n1, n2;
  V3 e1, e2;
  float uv[3][2] = {{0, 0}, {0, 0}, {0, 0}};  // attribute slot 1 (P/ModelLoader.cpp:282-283)
  uint32_t geom, prim, material;
  int emissive_id;  // running id over emissive triangles (P/ModelLoader.cpp:255-257,301-306), -1 if none
  float area;       // Triangle ctor, P/triangle.cpp:13-16
};
struct Box {
  V3 lo, hi;
};
struct BvhNode {
  Box box;
  int left, right;  // internal: children; leaf: left = -1
  int first, count;
};

struct Hit {
  float t = FLT_MAX, u = 0, v = 0;
  int tri = -1;
};

// Möller–Trumbore with a fixed operation order (fused multiply-adds spelled
// out so that g++ and nvcc produce the same bits). Accepts tnear < t < tfar.
static inline bool tri_test(const V3& o, const V3& d, const Tri& T, float tnear, float tfar, float* t, float* u,
                            float* v) {
  const V3& e1 = T.e1;
  const V3& e2 = T.e2;
  float px = std::fmaf(d.y, e2.z, -(d.z * e2.y));
  float py = std::fmaf(d.z, e2.x, -(d.x * e2.z));
  float pz = std::fmaf(d.x, e2.y, -(d.y * e2.x));
  float det = std::fmaf(e1.z, pz, std::fmaf(e1.y, py, e1.x * px));
  if (det == 0.0f) return false;
  float inv = 1.0f / det;
  float tx = o.x - T.p0.x, ty = o.y - T.p0.y, tz = o.z - T.p0.z;
  float uu = std::fmaf(tz, pz, std::fmaf(ty, py, tx * px)) * inv;
  if (!(uu >= 0.0f && uu <= 1.0f)) return false;
  float qx = std::fmaf(ty, e1.z, -(tz * e1.y));
  float qy = std::fmaf(tz, e1.x, -(tx * e1.z));
  float qz = std::fmaf(tx, e1.y, -(ty * e1.x));
  float vv = std::fmaf(d.z, qz, std::fmaf(d.y, qy, d.x * qx)) * inv;
  if (!(vv >= 0.0f && uu + vv <= 1.0f)) return false;
  float tt = std::fmaf(e2.z, qz, std::fmaf(e2.y, qy, e2.x * qx)) * inv;
  if (!(tt > tnear && tt < tfar)) return false;
  *t = tt;
  *u = uu;
  *v = vv;
  return true;
}

struct Scene {
  std::vector<Tri> tris;
  std::vector<RbMaterial> mats;
  // textured materials (rb_set_textures seam): Texture's members (P/Texture.h:44-48) + Material::textures_ slots
  struct Tex {
    int width = 0, height = 0, scan_width = 0, pixel_size = 0;
    std::vector<unsigned char> data;
    bool clamp = false;  // TextureClamp: REPEAT (materials, ModelLoader::TextureProxy) / CLAMP_TO_EDGE (Texture's default: the sky)
  };
  std::vector<Tex> textures;
  Tex sky;  // SphericalMap::texture (rb_set_sky seam); width == 0 = none
  std::vector<RbMaterialTextures> mat_tex;  // empty = untextured
  std::vector<V3> tangents;  // [3 * n_tris] attribute slot 3 (P/ModelLoader.cpp:286-287), read by normal maps only; empty = none
  std::vector<int> emissive;  // TriangleCDF::tris (indices into tris)
  // TriangleCDF, P/TriangleCDF.cpp:8-34
  float totalSurface = 0;
  std::vector<float> cdf;
  // alias seam
  std::vector<float> alias_prob;
  std::vector<uint32_t> alias_idx;
  // tracer
  int tracer = 0;
  std::vector<BvhNode> nodes;
  std::vector<int> order;  // leaf triangle indices
  float pad = 0;

  bool lightsValid() const { return !emissive.empty(); }  // TriangleCDF::isValid, P/TriangleCDF.h:19-21

  void buildLights() {
    totalSurface = 0;
    for (int id : emissive) totalSurface += tris[id].area;
    cdf.clear();
    for (size_t i = 0; i < emissive.size(); ++i) {
      float normArea = tris[emissive[i]].area / totalSurface;
      float pred = i == 0 ? 0.0f : cdf[i - 1];
      cdf.push_back(pred + normArea);
    }
    // (the reference's std::sort of the cdf is a no-op on a non-decreasing sequence)
    // Vose alias table, fixed processing order (DESIGN.md "alias table")
    size_t N = emissive.size();
    alias_prob.assign(N, 1.0f);
    alias_idx.resize(N);
    std::vector<float> q(N);
    std::vector<uint32_t> small, large;
    for (size_t i = 0; i < N; ++i) {
      alias_idx[i] = (uint32_t)i;
      q[i] = (tris[emissive[i]].area / totalSurface) * (float)N;
      if (q[i] < 1.0f)
        small.push_back((uint32_t)i);
      else
        large.push_back((uint32_t)i);
    }
    while (!small.empty() && !large.empty()) {
      uint32_t s = small.back();
      small.pop_back();
      uint32_t l = large.back();
      large.pop_back();
      alias_prob[s] = q[s];
      alias_idx[s] = l;
      q[l] = (q[l] + q[s]) - 1.0f;
      if (q[l] < 1.0f)
        small.push_back(l);
      else
        large.push_back(l);
    }
  }

  // ---- BVH2 (median split), only a culling structure: hit set == brute force
  int buildNode(int first, int count, std::vector<V3>& cent) {
    BvhNode n;
    n.first = first;
    n.count = count;
    n.left = n.right = -1;
    Box b{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
    Box cb = b;
    for (int i = first; i < first + count; ++i) {
      const Tri& T = tris[order[i]];
      const V3* ps[3] = {&T.p0, &T.p1, &T.p2};
      for (auto p : ps) {
        b.lo = {std::min(b.lo.x, p->x), std::min(b.lo.y, p->y), std::min(b.lo.z, p->z)};
        b.hi = {std::max(b.hi.x, p->x), std::max(b.hi.y, p->y), std::max(b.hi.z, p->z)};
      }
      const V3& c = cent[order[i]];
      cb.lo = {std::min(cb.lo.x, c.x), std::min(cb.lo.y, c.y), std::min(cb.lo.z, c.z)};
      cb.hi = {std::max(cb.hi.x, c.x), std::max(cb.hi.y, c.y), std::max(cb.hi.z, c.z)};
    }
    b.lo = b.lo - v3(pad);
    b.hi = b.hi + v3(pad);
    n.box = b;
    int idx = (int)nodes.size();
    nodes.push_back(n);
    if (count > 4) {
      V3 ext = cb.hi - cb.lo;
      int axis = ext.x >= ext.y ? (ext.x >= ext.z ? 0 : 2) : (ext.y >= ext.z ? 1 : 2);
      auto key = [&](int t) { return axis == 0 ? cent[t].x : axis == 1 ? cent[t].y : cent[t].z; };
      int mid = first + count / 2;
      std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                       [&](int a, int b2) { return key(a) < key(b2); });
      int l = buildNode(first, mid - first, cent);
      int r = buildNode(mid, first + count - mid, cent);
      nodes[idx].left = l;
      nodes[idx].right = r;
    }
    return idx;
  }
  void buildBvh() {
    nodes.clear();
    order.resize(tris.size());
    std::vector<V3> cent(tris.size());
    float maxabs = 0;
    for (size_t i = 0; i < tris.size(); ++i) {
      order[i] = (int)i;
      const Tri& T = tris[i];
      cent[i] = (T.p0 + T.p1 + T.p2) * (1.0f / 3.0f);
      const V3* ps[3] = {&T.p0, &T.p1, &T.p2};
      for (auto p : ps) maxabs = std::max(maxabs, std::max(std::fabs(p->x), std::max(std::fabs(p->y), std::fabs(p->z))));
    }
    pad = maxabs * 1e-5f + 1e-30f;
    if (!tris.empty()) {
      nodes.reserve(tris.size());
      buildNode(0, (int)tris.size(), cent);
    }
  }
  static inline bool slab(const Box& b, const V3& o, const V3& d, float tnear, float tfar) {
    float t0 = tnear, t1 = tfar;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    const float lo[3] = {b.lo.x, b.lo.y, b.lo.z}, hi[3] = {b.hi.x, b.hi.y, b.hi.z};
    for (int a = 0; a < 3; ++a) {
      if (dd[a] == 0.0f) {
        if (oo[a] < lo[a] || oo[a] > hi[a]) return false;
        continue;
      }
      if (dd[a] != dd[a]) return true;  // NaN direction: let the triangle tests decide
      double inv = 1.0 / (double)dd[a];
      double a0 = ((double)lo[a] - oo[a]) * inv, a1 = ((double)hi[a] - oo[a]) * inv;
      if (a0 > a1) std::swap(a0, a1);
      // widen by a relative epsilon: culling must never drop a triangle the brute force hits
      a0 -= std::fabs(a0) * 1e-6;
      a1 += std::fabs(a1) * 1e-6;
      if (a0 > t0) t0 = (float)std::min(a0, (double)FLT_MAX);
      if (a1 < t1) t1 = (float)std::max(a1, -(double)FLT_MAX);
      if ((double)t0 > (double)t1 * (1.0 + 1e-6) + 1e-30) return false;
    }
    return true;
  }

  Hit closest(const V3& o, const V3& d, float tnear, float tfar) const {
    Hit best;
    best.t = tfar;
    auto consider = [&](int ti) {
      float t, u, v;
      // accept t == best.t for the tie-break, so test against an open upper bound just above
      if (tri_test(o, d, tris[ti], tnear, tfar, &t, &u, &v)) {
        if (best.tri < 0 || t < best.t || (t == best.t && ti < best.tri)) {
          best.t = t;
          best.u = u;
          best.v = v;
          best.tri = ti;
        }
      }
    };
    if (tracer == 0 || nodes.empty()) {
      for (int i = 0; i < (int)tris.size(); ++i) consider(i);
    } else {
      int stack[128];
      int sp = 0;
      stack[sp++] = 0;
      while (sp) {
        const BvhNode& n = nodes[stack[--sp]];
        float cull_far = best.tri >= 0 ? best.t : tfar;
        // keep equal-t candidates reachable: the slab test is already widened
        if (!slab(n.box, o, d, tnear, cull_far)) continue;
        if (n.left < 0) {
          for (int i = n.first; i < n.first + n.count; ++i) consider(order[i]);
        } else {
          stack[sp++] = n.left;
          stack[sp++] = n.right;
        }
      }
    }
    if (best.tri < 0) best.t = FLT_MAX;
    return best;
  }
  bool occluded(const V3& o, const V3& d, float tnear, float tfar) const {
    float t, u, v;
    if (tracer == 0 || nodes.empty()) {
      for (int i = 0; i < (int)tris.size(); ++i)
        if (tri_test(o, d, tris[i], tnear, tfar, &t, &u, &v)) return true;
      return false;
    }
    int stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
      const BvhNode& n = nodes[stack[--sp]];
      if (!slab(n.box, o, d, tnear, tfar)) continue;
      if (n.left < 0) {
        for (int i = n.first; i < n.first + n.count; ++i)
          if (tri_test(o, d, tris[order[i]], tnear, tfar, &t, &u, &v)) return true;
      } else {
        stack[sp++] = n.left;
        stack[sp++] = n.right;
      }
    }
    return false;
  }
};

// ----------------------------------------------------------------------------
// Reservoir / G-buffer records (P/Reservoir.h:6-60, P/GBufferElement.h:6-23)
// ----------------------------------------------------------------------------
struct LightSample {
  V3 samplePoint{-FLT_MAX, -FLT_MAX, -FLT_MAX}, sampleNormal{-FLT_MAX, -FLT_MAX, -FLT_MAX};
  V3 L_i{-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int lightIdx = -1;  // instrumentation only: emissive-triangle id of the sample
  bool isValid() const {
    bool pointOk = samplePoint.x != -FLT_MAX && samplePoint.y != -FLT_MAX && samplePoint.z != -FLT_MAX;
    bool normalOk = sampleNormal.x != -FLT_MAX && sampleNormal.y != -FLT_MAX && sampleNormal.z != -FLT_MAX;
    bool L_iOk = L_i.x > 0 || L_i.y > 0 || L_i.z > 0;
    return pointOk && normalOk && L_iOk;
  }
};
struct InitialCandidateSample {
  LightSample sample{};
  float W{0};
  float misWeight{};
};
struct Reservoir {
  LightSample bestSample{};
  float w_sum{0};
  float W{0};
  int confidence{0};
  // Reservoir::addSample, P/Reservoir.h:33-47 (draw only when not the 0/0 case)
  bool addSample(const LightSample& s, float w, int c, Rng& rng, uint32_t slot) {
    w_sum += w;
    confidence += c;
    if (w == 0 && w_sum == 0) return false;
    if (rng.value(slot, 0, 1) < w / w_sum) {
      bestSample = s;
      return true;
    }
    return false;
  }
  bool hasSample() const { return w_sum > 0.0f; }
  void capConfidence(int cap) { confidence = std::min(confidence, cap); }
};
struct GBufferElement {
  V3 worldSpacePos{0, 0, 0}, worldSpaceNormal{0, 0, 0};
  V3 diffuseColor{0, 0, 0}, specularColor{0, 0, 0}, emission{0, 0, 0};
  float shininess{0}, depth{0};
  uint8_t materialType{0};
  // not in the reference record: cached 1/I_M (pure function of this element and its frame's camera)
  float invIM{0};
  uint32_t geomID{0xFFFFFFFFu}, primID{0xFFFFFFFFu};
};
static inline bool emissive(const V3& e) { return e.x > 0 || e.y > 0 || e.z > 0; }

struct GBuffer {
  std::vector<GBufferElement> px;
  V3 cameraPosWS{0, 0, 0};
  float viewMat[16]{}, invViewMat[16]{};
  float focalLength{0};
};

struct alignas(64) Counters {
  uint64_t closest = 0, any_written = 0, any_traced = 0;
  // temporal pass outcome per pixel: [0] backward reprojection failed (:644), [1] depth test
  // at the reprojected pixel failed (:660), [2] forward reprojection failed (:671), [3] depth test at the
  // forward-reprojected pixel failed (:686), [4] merged with the previous frame's reservoir
  uint64_t temporal[5] = {0, 0, 0, 0, 0};
};
static inline int thread_id() {
#ifdef _OPENMP
  return omp_get_thread_num() & 255;
#else
  return 0;
#endif
}

struct Oracle {
  int width = 0, height = 0;
  int band_y0 = 0, band_y1 = 0;
  std::vector<int> rows;  // bench sampling only (orc_set_row_segments): render these rows instead of [band_y0, band_y1)
  uint32_t seed = 123;
  Math math;
  int rng_mode = 0;
  int cache_iim = 0;
  std::mt19937 legacy_gen{123};
  std::uniform_real_distribution<float> legacy_dist{0.0f, 1.0f};
  Scene scene;
  bool have_scene = false;
  RbParams P;
  GBuffer gBuffer, gBufferLastFrame;
  std::vector<Reservoir> res[3];
  int readIdx = 0, writeIdx = 1, lastIdx = 2;
  std::vector<float> frame;
  std::vector<float> accumulator, display;  // N1: P/simpleguidx11.h:153,156
  uint32_t frameIdx = 0;
  Counters ctrs[256];  // one slot per OpenMP thread
  std::string err;

  Reservoir& R_read(int x, int y) { return res[readIdx][(size_t)y * width + x]; }
  Reservoir& R_write(int x, int y) { return res[writeIdx][(size_t)y * width + x]; }
  Reservoir& R_last(int x, int y) { return res[lastIdx][(size_t)y * width + x]; }
  void swapReservoirBuffers() { std::swap(readIdx, writeIdx); }  // P/simpleguidx11.h:116

  Rng rngFor(uint32_t pass, uint32_t iter, int x, int y) {
    Rng r;
    r.mode = rng_mode;
    r.gen = &legacy_gen;
    r.dist = &legacy_dist;
    r.key = rng_key(seed, frameIdx, pass, iter, (uint32_t)(y * width + x));
    return r;
  }

  // ---- Phong (P/MaterialPhong.cpp:122-248) ----------------------------------
  float gamma_quot(float a, float b) const { return math.exp(math.lgamma(a) - math.lgamma(b)); }  // :224-226
  float calc_I_M(float nDotV, float n) const {                                                      // :228-244
    float costerm = nDotV;
    float sinterm_sq = 1.0f - costerm * costerm;
    float halfn = 0.5f * n;
    float negterm = costerm;
    sinterm_sq = gclamp(sinterm_sq, 0.0f, 1.0f);
    if (n >= 1e-18f) negterm *= halfn * math.ibeta(sinterm_sq, halfn, 0.5f);
    return (kTwoPi * costerm + kRootPi * gamma_quot(halfn + 0.5f, halfn + 1.0f) * (math.pow(sinterm_sq, halfn) - negterm)) /
           (n + 2.0f);
  }
  float inv_I_M(const GBufferElement& e, const V3& cameraPos) const {
    const V3 V = normalize(cameraPos - e.worldSpacePos);
    float nDotV = dot(V, e.worldSpaceNormal);
    return 1.0f / calc_I_M(nDotV, e.shininess);
  }
  V3 phong_evalBRDF(const GBufferElement& e, const V3& cameraPos, const V3& omega_i) const {  // :122-148
    const V3 V = normalize(cameraPos - e.worldSpacePos);
    V3 f_r = e.diffuseColor * kOneOverPi;
    float i_m;
    if (cache_iim)
      i_m = e.invIM;
    else {
      float nDotV = dot(V, e.worldSpaceNormal);
      i_m = 1.0f / calc_I_M(nDotV, e.shininess);
    }
    const V3 omega_r = normalize(reflect(-V, e.worldSpaceNormal));
    f_r = f_r + e.specularColor * i_m * math.pow(gmax(dot(omega_i, omega_r), 0.0f), e.shininess);
    return f_r;
  }
  static float cosw_pdf(const V3& normal, const V3& omega_i) {  // CosineWeightedDistribution::getPdf, P/Distribution.h:33-35
    return gmax(dot(normal, omega_i), 0.0f) * kOneOverPi;
  }
  float lobe_pdf(const V3& omega_i, const V3& omega_r, float gamma) const {  // CosineLobeDistribution::getPdf, :65-67
    return (gamma + 1.0f) * kOneOverTwoPi * math.pow(gmax(0.0f, dot(omega_i, omega_r)), gamma);
  }
  static float maxComponent(const V3& v) { return gmax(gmax(v.x, v.y), v.z); }  // P/utils.h:61-63
  float phong_evalPdf(const GBufferElement& e, const V3& cameraPos, const V3& omega_i) const {  // :150-172
    float maxDiffuse = maxComponent(e.diffuseColor);
    float maxSpecular = maxComponent(e.specularColor);
    float pdfFactor = maxDiffuse / (maxDiffuse + maxSpecular);
    float pdf = cosw_pdf(e.worldSpaceNormal, omega_i) * pdfFactor;
    const V3 omega_o = normalize(e.worldSpacePos - cameraPos);
    const V3 omega_r = normalize(reflect(omega_o, e.worldSpaceNormal));
    pdf += lobe_pdf(omega_i, omega_r, e.shininess) * (1.0f - pdfFactor);
    return pdf;
  }
  static V3 orthogonal(const V3& v) {  // Utils::orthogonal, P/utils.cpp:204-207
    return std::fabs(v.x) > std::fabs(v.z) ? V3{v.y, -v.x, 0.0f} : V3{0.0f, v.z, -v.y};
  }
  static V3 toWorld(const V3& o1, const V3& o2, const V3& n, const V3& s) {  // glm::mat3{o1,o2,n} * s
    return {o1.x * s.x + o2.x * s.y + n.x * s.z, o1.y * s.x + o2.y * s.y + n.y * s.z, o1.z * s.x + o2.z * s.y + n.z * s.z};
  }
  V3 cosw_sample(const V3& normal, Rng& rng, uint32_t s1, uint32_t s2) const {  // P/Distribution.h:10-31
    float r1 = rng.value(s1, 0, 1);
    float r2 = rng.value(s2, 0, 1);
    float x = math.cos(kPi * 2.0f * r1) * std::sqrt(1.0f - r2);
    float y = math.sin(kPi * 2.0f * r1) * std::sqrt(1.0f - r2);
    float z = std::sqrt(r2);
    V3 sample = normalize(V3{x, y, z});
    V3 o2 = normalize(orthogonal(normal));
    V3 o1 = normalize(cross(normal, o2));
    o2 = normalize(cross(o1, normal));
    return toWorld(o1, o2, normal, sample);
  }
  V3 lobe_sample(const V3& omega_r, float gamma, Rng& rng, uint32_t s1, uint32_t s2) const {  // :43-63
    float r1 = rng.value(s1, 0, 1);
    float r2 = rng.value(s2, 0, 1);
    float x = math.cos(2.0f * kPi * r1) * std::sqrt(1.0f - math.pow(r2, 2.0f / (gamma + 1.0f)));
    float y = math.sin(2.0f * kPi * r1) * std::sqrt(1.0f - math.pow(r2, 2.0f / (gamma + 1.0f)));
    float z = math.pow(r2, 1.0f / (gamma + 1.0f));
    V3 sample = normalize(V3{x, y, z});
    V3 o2 = normalize(orthogonal(omega_r));
    V3 o1 = normalize(cross(omega_r, o2));
    o2 = normalize(cross(o1, omega_r));
    return toWorld(o1, o2, omega_r, sample);
  }
  struct PTInfoGI {
    V3 omega_i;
    float pdf;
  };
  // MaterialPhong::sampleBRDF, P/MaterialPhong.cpp:174-222 (f_r is unused by the ReSTIR caller)
  PTInfoGI phong_sampleBRDF(const GBufferElement& e, const V3& cameraPos, Rng& rng, uint32_t base) const {
    V3 omega_o = normalize(e.worldSpacePos - cameraPos);
    float maxDiffuse = maxComponent(e.diffuseColor);
    float maxSpecular = maxComponent(e.specularColor);
    float r0 = rng.value(base + 0, 0.0f, maxDiffuse + maxSpecular);
    float pdfFactor = maxDiffuse / (maxDiffuse + maxSpecular);
    V3 omega_i;
    const V3 omega_r = normalize(reflect(omega_o, e.worldSpaceNormal));
    if (r0 < maxDiffuse)
      omega_i = cosw_sample(e.worldSpaceNormal, rng, base + 1, base + 2);
    else
      omega_i = lobe_sample(omega_r, e.shininess, rng, base + 1, base + 2);
    float pdfDiffuse = cosw_pdf(e.worldSpaceNormal, omega_i) * pdfFactor;
    float pdfSpecular = lobe_pdf(omega_i, omega_r, e.shininess) * (1.0f - pdfFactor);
    return {omega_i, pdfDiffuse + pdfSpecular};
  }
  // MaterialLambert::sampleBRDF, P/MaterialLambert.cpp:43-53
  PTInfoGI lambert_sampleBRDF(const GBufferElement& e, Rng& rng, uint32_t base) const {
    V3 omega_i = cosw_sample(e.worldSpaceNormal, rng, base + 1, base + 2);
    return {omega_i, cosw_pdf(e.worldSpaceNormal, omega_i)};
  }
  // dispatchers, P/ReSTIRIntegrator.h:32-59
  V3 evalBRDF(const GBufferElement& e, const V3& cameraPos, const V3& omega_i) const {
    if (e.materialType == RB_MAT_PHONG || e.materialType == RB_MAT_DIELECTRIC) return phong_evalBRDF(e, cameraPos, omega_i);
    return e.diffuseColor * kOneOverPi;  // MaterialLambert::evalBRDF, P/MaterialLambert.cpp:33-41
  }
  PTInfoGI sampleBRDF(const GBufferElement& e, const V3& cameraPos, Rng& rng, uint32_t base) const {
    if (e.materialType == RB_MAT_LAMBERT) return lambert_sampleBRDF(e, rng, base);
    return phong_sampleBRDF(e, cameraPos, rng, base);
  }

  // ---- ray queries (P/Intersection.h) ---------------------------------------
  // Intersection::testOcclusion, :43-60
  bool testOcclusion(const V3& from, const V3& to) {
    const float dist = length(to - from);
    const V3 dir = normalize(to - from);
    Counters& ctr = ctrs[thread_id()];
    ctr.any_written++;
    ctr.any_traced++;
    return scene.occluded(from, dir, FLT_MIN + P.tnearOffset, dist - P.tfarOffset);
  }
  struct HitInfo {
    bool didHit = false;
    V3 normal{0, 0, 0}, hitPoint{0, 0, 0};
    float dst = FLT_MAX;
    uint32_t hitTriId = 0;
    int tri = -1;
    float uv[2] = {0, 0};
  };
  // Intersection::intersectEmbree + getGeometryAttributes, :8-41, 85-113
  HitInfo intersect(const V3& org, const V3& dir, float tnear, float tfar) {
    HitInfo h;
    ctrs[thread_id()].closest++;
    Hit hit = scene.closest(org, dir, tnear, tfar);
    if (hit.tri >= 0) {
      const Tri& T = scene.tris[hit.tri];
      // rtcInterpolate0 of the normal attribute: w*n0 + u*n1 + v*n2, w = 1-u-v
      float w = 1.0f - hit.u - hit.v;
      V3 n = T.n0 * w + T.n1 * hit.u + T.n2 * hit.v;
      n = normalize(n);
      if (dot(-dir, n) <= 0.0f) n = n * -1.0f;
      h.normal = n;
      // rtcInterpolate0 of attribute slot 1 (P/Intersection.h:99-100), same contract as the normal
      h.uv[0] = T.uv[0][0] * w + T.uv[1][0] * hit.u + T.uv[2][0] * hit.v;
      h.uv[1] = T.uv[0][1] * w + T.uv[1][1] * hit.u + T.uv[2][1] * hit.v;
      // normal map (:25-39): TBN from the interpolated tangent (slot 3) and the flipped normal, texel * 2 - 1; the
      // result is neither re-normalised nor flipped again
      if (!scene.mat_tex.empty() && scene.mat_tex[T.material].normal >= 0 && !scene.tangents.empty()) {
        const V3* tn = scene.tangents.data() + 3 * (size_t)hit.tri;
        V3 tangent = tn[0] * w + tn[1] * hit.u + tn[2] * hit.v;
        V3 Tt = tangent - dot(tangent, h.normal) * h.normal;
        Tt = normalize(Tt);
        V3 B = normalize(cross(h.normal, Tt));
        V3 N = texSample(scene.textures[scene.mat_tex[T.material].normal], h.uv) * 2.0f - V3{1.0f, 1.0f, 1.0f};
        const V3 n0 = h.normal;  // glm mat3 * vec3: (m0 * x + m1 * y) + m2 * z per component
        h.normal = {Tt.x * N.x + B.x * N.y + n0.x * N.z, Tt.y * N.x + B.y * N.y + n0.y * N.z, Tt.z * N.x + B.z * N.y + n0.z * N.z};
      }
      h.didHit = true;
      h.hitPoint = org + dir * hit.t;
      h.dst = hit.t;
      h.hitTriId = T.emissive_id >= 0 ? (uint32_t)T.emissive_id : 0u;
      h.tri = hit.tri;
    }
    return h;
  }

  // ---- camera (P/camera.cpp:20-42) ------------------------------------------
  void generateRay(int x, int y, V3* org, V3* dir, Rng& rng) {
    V3 d_c{(float)x - (float)width / 2.0f, (float)height / 2.0f - (float)y, -gBuffer.focalLength};
    const float* m = gBuffer.invViewMat;  // mat3(invViewMat) * d_c
    V3 d_w{m[0] * d_c.x + m[4] * d_c.y + m[8] * d_c.z, m[1] * d_c.x + m[5] * d_c.y + m[9] * d_c.z,
           m[2] * d_c.x + m[6] * d_c.y + m[10] * d_c.z};
    d_w = normalize(d_w);
    if (rng.mode == 1) {  // the discarded Sampling::sampleDiskUniform(apertureSize), :30
      (void)rng.unit(0);
      (void)rng.unit(1);
    }
    *org = gBuffer.cameraPosWS;
    *dir = d_w;
  }

  // ---- evaluateF / evaluatePHat (P/ReSTIRIntegrator.cpp:180-211) -------------
  V3 evaluateF(const LightSample& s, const V3& cameraPos, const GBufferElement& g, bool testVisibility) {
    if (!s.isValid() || g.emission.x > 0 || g.emission.y > 0 || g.emission.z > 0) return v3(0);
    V3 lightDir = s.samplePoint - g.worldSpacePos;
    float r_sqr = dot(lightDir, lightDir);
    lightDir = normalize(lightDir);
    float cosThetaI = gmax(dot(lightDir, g.worldSpaceNormal), 0.0f);
    float cosThetaY = std::fabs(dot(-lightDir, s.sampleNormal));
    float G = cosThetaI * cosThetaY / r_sqr;
    V3 f_r = evalBRDF(g, cameraPos, lightDir);
    bool V = true;
    if (testVisibility) V = !testOcclusion(g.worldSpacePos, s.samplePoint);
    return s.L_i * f_r * G * (float)V;
  }
  float evaluatePHat(const LightSample& s, const V3& cameraPos, const GBufferElement& g, bool testVisibility) {
    return length(evaluateF(s, cameraPos, g, testVisibility));
  }

  // ---- light sampling -------------------------------------------------------
  struct TriPick {
    int emissiveIdx;
    float pdf;
  };
  // TriangleCDF::getTriangle, P/TriangleCDF.cpp:36-54 (its private uniform_int draw never
  // touches the shared stream) / alias seam
  TriPick pickTriangle(Rng& rng, uint32_t slot) {
    size_t N = scene.emissive.size();
    if (P.lightSampler == RB_LS_ALIAS) {
      uint32_t h = rng.raw(slot);
      uint32_t i = (uint32_t)(((uint64_t)h * (uint64_t)N) >> 32);
      float frac = rng.mode ? rng.unit(0) : bits_to_unit(rng_bits(rng.key, slot | 0x40000000u));
      uint32_t idx = (frac < scene.alias_prob[i]) ? i : scene.alias_idx[i];
      return {(int)idx, scene.tris[scene.emissive[idx]].area / scene.totalSurface};
    }
    float ksi = rng.value(slot, 0.0f, 1.0f);
    size_t index = std::lower_bound(scene.cdf.begin(), scene.cdf.end(), ksi) - scene.cdf.begin();
    if (index >= N) index = N - 1;
    if (index == 0) return {0, scene.cdf[0]};
    return {(int)index, scene.cdf[index] - scene.cdf[index - 1]};
  }
  float getPDFForTriangle(const Tri& t) const {  // P/TriangleCDF.h:25-31
    float pdf = t.area / scene.totalSurface;
    pdf *= 1.0f / t.area;
    return pdf;
  }
  float m_area(float pdfArea, float pdfBrdf) const {  // P/ReSTIRIntegrator.h:62-67
    if (pdfArea == 0.0f && pdfBrdf == 0.0f) return 0.0f;
    return pdfArea / ((float)P.M_Area * pdfArea + (float)P.M_Brdf * pdfBrdf);
  }
  float m_brdf(float pdfBrdf, float pdfArea) const {  // :69-74
    if (pdfArea == 0.0f && pdfBrdf == 0.0f) return 0.0f;
    return pdfBrdf / ((float)P.M_Area * pdfArea + (float)P.M_Brdf * pdfBrdf);
  }
  // ReSTIRIntegrator::areaSampleLight, P/ReSTIRIntegrator.cpp:89-124
  InitialCandidateSample areaSampleLight(const GBufferElement& elem, Rng& rng, uint32_t base) {
    TriPick pick = pickTriangle(rng, base + 0);
    const Tri& T = scene.tris[scene.emissive[pick.emissiveIdx]];
    // Sampling::sampleTriangle, P/Sampling.cpp:63-76
    float r1 = rng.value(base + 1, 0, 1);
    float r2 = rng.value(base + 2, 0, 1);
    float x = 1.0f - std::sqrt(r1);
    float y = std::sqrt(r1) * (1.0f - r2);
    float z = std::sqrt(r1) * r2;
    V3 samplePoint = T.p0 * x + T.p1 * y + T.p2 * z;
    V3 normal = normalize(T.n0 * x + T.n1 * y + T.n2 * z);
    float triPointPdf = 1.0f / T.area;

    float pdf_area = pick.pdf * triPointPdf;
    V3 lightDir = samplePoint - elem.worldSpacePos;
    float r_sqr = dot(lightDir, lightDir);
    lightDir = normalize(lightDir);
    float cosThetaY = gmax(dot(-lightDir, normal), 0.0f);
    float areaMeasureFactor = cosThetaY / r_sqr;
    float pdfAsIfBrdf = phong_evalPdf(elem, gBuffer.cameraPosWS, lightDir);  // getMaterialPDFEvalFunc: always Phong
    float pdfAsIfBrdfAreaMeasure = pdfAsIfBrdf * areaMeasureFactor;
    LightSample s;
    s.samplePoint = samplePoint;
    s.sampleNormal = normal;
    const RbMaterial& m = scene.mats[T.material];
    s.L_i = {m.emission[0], m.emission[1], m.emission[2]};
    s.lightIdx = pick.emissiveIdx;
    float misWeight = m_area(pdf_area, pdfAsIfBrdfAreaMeasure);
    return {s, 1.0f / pdf_area, misWeight};
  }
  // ReSTIRIntegrator::brdfSampleLight, :126-177
  InitialCandidateSample brdfSampleLight(const GBufferElement& elem, Rng& rng, uint32_t base) {
    PTInfoGI payloadGI = sampleBRDF(elem, gBuffer.cameraPosWS, rng, base);
    V3 org = elem.worldSpacePos + P.normalOffset * elem.worldSpaceNormal;
    HitInfo hi = intersect(org, payloadGI.omega_i, FLT_MIN + P.tnearOffset, FLT_MAX);
    LightSample s;
    float W = 0, misWeight = 0;
    if (hi.didHit && scene.tris[hi.tri].emissive_id >= 0) {
      V3 lightDir = hi.hitPoint - elem.worldSpacePos;
      float r_sqr = dot(lightDir, lightDir);
      lightDir = normalize(lightDir);
      float cosThetaY = gmax(dot(-lightDir, hi.normal), 0.0f);
      float areaMeasureFactor = cosThetaY / r_sqr;
      const Tri& T = scene.tris[scene.emissive[hi.hitTriId]];
      float brdfPdf = payloadGI.pdf;
      float pdf_area = getPDFForTriangle(T);
      float brdfPdfAreaMeasure = brdfPdf * areaMeasureFactor;
      s.samplePoint = hi.hitPoint;
      s.sampleNormal = hi.normal;
      const RbMaterial& m = scene.mats[scene.tris[hi.tri].material];
      s.L_i = {m.emission[0], m.emission[1], m.emission[2]};
      s.lightIdx = (int)hi.hitTriId;
      W = 1.0f / brdfPdfAreaMeasure;
      misWeight = m_brdf(brdfPdfAreaMeasure, pdf_area);
    }
    return {s, W, misWeight};
  }

  // ---- passes ---------------------------------------------------------------
  // Texture::get_texel(x, y), REPEAT, P/Texture.cpp:72-107
  static V3 texel(const Scene::Tex& T, int x, int y) {
    int c_x = std::abs(x % T.width), c_y = std::abs(y % T.height);
    if (T.clamp) {  // CLAMP_TO_EDGE: glm::clamp(x, 0, width_ - 1)
      c_x = std::min(std::max(x, 0), T.width - 1);
      c_y = std::min(std::max(y, 0), T.height - 1);
    }
    const int offset = c_y * T.scan_width + c_x * T.pixel_size;
    if (T.pixel_size > 4) {
      float f[3];
      memcpy(f, T.data.data() + offset, 12);
      return {f[0], f[1], f[2]};
    }
    const float b = T.data[offset] / 255.0f, g = T.data[offset + 1] / 255.0f, r = T.data[offset + 2] / 255.0f;
    return {r, g, b};
  }
  static V3 mix3(const V3& x, const V3& y, float a) { return x * (1.0f - a) + y * a; }  // glm::mix
  // Texture::getTexelBilinear, :170-194
  static V3 texSample(const Scene::Tex& T, const float* uv) {
    const float pcx = uv[0] * T.width, pcy = (1 - uv[1]) * T.height;
    const float flx = std::floor(pcx), fly = std::floor(pcy);
    const float tx = pcx - flx, ty = pcy - fly;
    V3 x0y0 = texel(T, flx, fly), x1y0 = texel(T, flx + 1, fly), x0y1 = texel(T, flx, fly + 1), x1y1 = texel(T, flx + 1, fly + 1);
    V3 x1 = mix3(x0y0, x1y0, tx);
    V3 x2 = mix3(x0y1, x1y1, tx);
    return mix3(x1, x2, ty);
  }

  // SphericalMap::getTexel, P/SphericalMap.cpp:10-14. INVPI is the double 1.0 / M_PI: the float products are promoted,
  // the sums are formed in double and rounded once when they are stored into `const float x, y`.
  V3 skyTexel(const V3& dir) const {
    const double INVPI = 1.0 / 3.14159265358979323846;
    const float x = (float)(0.5f + (double)(0.5f * math.atan2(dir.y, dir.x)) * INVPI);
    const float y = (float)(1.0f - (double)math.acos(dir.z) * INVPI);
    const float uv[2] = {x, y};
    return texSample(scene.sky, uv);
  }

  // ReSTIRIntegrator::gBufferFillPass, :213-234
  void gBufferFillPass(int x, int y) {
    Rng rng = rngFor(PASS_GBUF, 0, x, y);
    V3 org, dir;
    generateRay(x, y, &org, &dir, rng);
    HitInfo hi = intersect(org, dir, FLT_MIN + 0.01f, FLT_MAX);  // Ray ctor defaults, P/Ray.h:8
    GBufferElement e;
    if (hi.didHit) {
      const Tri& T = scene.tris[hi.tri];
      const RbMaterial& m = scene.mats[T.material];
      e.worldSpacePos = hi.hitPoint;
      e.worldSpaceNormal = hi.normal;
      e.depth = length(hi.hitPoint - org);
      e.materialType = (uint8_t)m.type;
      e.diffuseColor = {m.diffuse[0], m.diffuse[1], m.diffuse[2]};
      e.specularColor = {m.specular[0], m.specular[1], m.specular[2]};
      e.emission = {m.emission[0], m.emission[1], m.emission[2]};
      e.shininess = m.shininess;
      if (!scene.mat_tex.empty()) {  // Material::getDiffuseColor / getSpecularColor / getShininess, P/material.cpp:105-134
        const RbMaterialTextures& sl = scene.mat_tex[T.material];
        if (sl.diffuse >= 0) e.diffuseColor = texSample(scene.textures[sl.diffuse], hi.uv);
        if (sl.specular >= 0) e.specularColor = texSample(scene.textures[sl.specular], hi.uv);
        if (sl.shininess >= 0) {
          V3 tx = texSample(scene.textures[sl.shininess], hi.uv);
          e.shininess = 2.0f / (tx.x * tx.x) - 2.0f;
        }
      }
      e.geomID = T.geom;
      e.primID = T.prim;
      if (cache_iim && !emissive(e.emission) && (e.materialType == RB_MAT_PHONG || e.materialType == RB_MAT_DIELECTRIC))
        e.invIM = inv_I_M(e, gBuffer.cameraPosWS);
    } else {
      e.emission = P.useSkybox ? skyTexel(dir) : V3{P.bgColor[0], P.bgColor[1], P.bgColor[2]};  // :231
    }
    gBuffer.px[(size_t)y * width + x] = e;
  }

  // ReSTIRIntegrator::initialRenderPass, :236-298
  void initialRenderPass(int x, int y) {
    const GBufferElement& g = gBuffer.px[(size_t)y * width + x];
    if (emissive(g.emission) || !scene.lightsValid()) {
      R_write(x, y) = Reservoir{};
      return;
    }
    Rng rng = rngFor(PASS_INITIAL, 0, x, y);
    Reservoir r{};
    const V3 cam = gBuffer.cameraPosWS;
    if (P.M_Area > 0) {
      float inv_MArea = 1.0f / (float)P.M_Area;
      for (int i = 0; i < P.M_Area; ++i) {
        uint32_t base = 4u * (uint32_t)i;
        InitialCandidateSample c = areaSampleLight(g, rng, base);
        float p_hat = evaluatePHat(c.sample, cam, g, !P.doVisibilityPass);
        float w;
        if (P.M_Brdf > 0)
          w = c.misWeight * p_hat * c.W;
        else
          w = inv_MArea * p_hat * c.W;
        r.addSample(c.sample, w, 1, rng, base + 3);
      }
    }
    if (P.M_Brdf > 0) {
      float inv_MBrdf = 1.0f / (float)P.M_Brdf;
      for (int i = 0; i < P.M_Brdf; ++i) {
        uint32_t base = 4u * (uint32_t)(P.M_Area + i);
        InitialCandidateSample c = brdfSampleLight(g, rng, base);
        float p_hat = evaluatePHat(c.sample, cam, g, !P.doVisibilityPass);
        float w;
        if (P.M_Area > 0)
          w = c.misWeight * p_hat * c.W;
        else
          w = inv_MBrdf * p_hat * c.W;
        r.addSample(c.sample, w, 1, rng, base + 3);
      }
    }
    float p_hat = evaluatePHat(r.bestSample, cam, g, !P.doVisibilityPass);
    r.W = p_hat > 0.0f ? 1.0f / p_hat * r.w_sum : 0.0f;
    r.capConfidence(P.confidenceCap);
    R_write(x, y) = r;
  }

  // ReSTIRIntegrator::visibilityPass, :302-312
  void visibilityPass(int x, int y) {
    const V3 samplePoint = R_write(x, y).bestSample.samplePoint;
    const V3 shadingPoint = gBuffer.px[(size_t)y * width + x].worldSpacePos;
    bool V = !testOcclusion(shadingPoint, samplePoint);
    if (!V) R_write(x, y).W = 0;
  }

  // reprojectBackward / reprojectForward, :544-587
  bool reproject(const GBuffer& gb, const V3& wsPos, int* sx, int* sy) const {
    const float* m = gb.viewMat;  // glm mat4 * vec4 order: (m0*x + m1*y) + (m2*z + m3*w)
    float vx = (m[0] * wsPos.x + m[4] * wsPos.y) + (m[8] * wsPos.z + m[12] * 1.0f);
    float vy = (m[1] * wsPos.x + m[5] * wsPos.y) + (m[9] * wsPos.z + m[13] * 1.0f);
    float vz = (m[2] * wsPos.x + m[6] * wsPos.y) + (m[10] * wsPos.z + m[14] * 1.0f);
    if (vz >= 0) return false;
    float focal = gb.focalLength;
    float fx = std::round((-vx / vz) * focal + (float)width / 2.0f);
    float fy = std::round((vy / vz) * focal + (float)height / 2.0f);
    // int conversion of out-of-range floats is UB in C++; clamp first (same on GPU)
    if (!(fx >= -1.0f)) return false;
    if (!(fy >= -1.0f)) return false;
    if (fx > (float)width || fy > (float)height) return false;
    int screenX = (int)fx, screenY = (int)fy;
    if (screenX < 0 || screenX > width - 1 || screenY < 0 || screenY > height - 1) return false;
    *sx = screenX;
    *sy = screenY;
    return true;
  }

  // ReSTIRIntegrator::temporalReusePass, :625-732
  void temporalReusePass(int x, int y) {
    Rng rng = rngFor(PASS_TEMPORAL, 0, x, y);
    Reservoir resultReservoir{};
    const GBufferElement currentElem = gBuffer.px[(size_t)y * width + x];
    const V3 currentCamPos = gBuffer.cameraPosWS;
    int px, py;
    bool ok = reproject(gBufferLastFrame, currentElem.worldSpacePos, &px, &py);
    const Reservoir currentReservoir = R_read(x, y);
    const LightSample& currentSample = currentReservoir.bestSample;
    // same pixel, not the reprojected one (:641); temporalFetchReprojected = the repaired variant (no reference behaviour)
    const Reservoir prevReservoir = (P.temporalFetchReprojected && ok) ? R_last(px, py) : R_last(x, y);
    const LightSample& prevSample = prevReservoir.bestSample;
    uint64_t* tstat = ctrs[thread_id()].temporal;
    if (!ok) {
      tstat[0]++;
      R_write(x, y) = currentReservoir;
      return;
    }
    const GBufferElement prevElem = gBufferLastFrame.px[(size_t)py * width + px];
    const V3 prevCamPos = gBufferLastFrame.cameraPosWS;
    float currentDepth = length(currentElem.worldSpacePos - currentCamPos);
    float prevDepth = length(prevElem.worldSpacePos - prevCamPos);
    float depthRatio = currentDepth > prevDepth ? prevDepth / currentDepth : currentDepth / prevDepth;
    if (depthRatio < 0.9f) {
      tstat[1]++;
      R_write(x, y) = currentReservoir;
      return;
    }
    const GBufferElement prevElemAtCurrent = gBufferLastFrame.px[(size_t)y * width + x];
    int fx, fy;
    if (!reproject(gBuffer, prevElemAtCurrent.worldSpacePos, &fx, &fy)) {
      tstat[2]++;
      R_write(x, y) = currentReservoir;
      return;
    }
    const GBufferElement fwReprojected = gBuffer.px[(size_t)fy * width + fx];
    float currentDepthP = length(prevElemAtCurrent.worldSpacePos - prevCamPos);
    float prevDepthP = length(fwReprojected.worldSpacePos - currentCamPos);
    float depthRatioP = currentDepthP > prevDepthP ? prevDepthP / currentDepthP : currentDepthP / prevDepthP;
    if (depthRatioP < 0.9f) {
      tstat[3]++;
      R_write(x, y) = currentReservoir;
      return;
    }
    tstat[4]++;
    float p_cur = evaluatePHat(currentSample, currentCamPos, currentElem, true);
    float p_prev = evaluatePHat(currentSample, prevCamPos, prevElem, true);
    float m_cur = p_cur * (float)currentReservoir.confidence /
                  (p_cur * (float)currentReservoir.confidence + p_prev * (float)prevReservoir.confidence);
    if (!(m_cur > 0)) m_cur = 0.0f;
    float p_hat_cur = evaluatePHat(currentSample, currentCamPos, currentElem, true);
    float w_cur = m_cur * p_hat_cur * currentReservoir.W;
    resultReservoir.addSample(currentSample, w_cur, currentReservoir.confidence, rng, 0);

    p_cur = evaluatePHat(prevSample, currentCamPos, currentElem, true);
    p_prev = evaluatePHat(prevSample, prevCamPos, prevElem, true);
    float m_prev = p_prev * (float)prevReservoir.confidence /
                   (p_cur * (float)currentReservoir.confidence + p_prev * (float)prevReservoir.confidence);
    if (!(m_prev > 0)) m_prev = 0.0f;
    float p_hat_prev = evaluatePHat(prevSample, currentCamPos, currentElem, true);
    float w_prev = m_prev * p_hat_prev * prevReservoir.W;
    resultReservoir.addSample(prevSample, w_prev, prevReservoir.confidence, rng, 1);
    resultReservoir.capConfidence(P.confidenceCap);
    float final_p_hat = evaluatePHat(resultReservoir.bestSample, currentCamPos, currentElem, true);
    resultReservoir.W = final_p_hat > 0.0f ? resultReservoir.w_sum / final_p_hat : 0.0f;
    R_write(x, y) = resultReservoir;
  }

  // ReSTIRIntegrator::spatialReusePass, :316-542
  void spatialReusePass(int x, int y, int iter) {
    const GBufferElement thisElem = gBuffer.px[(size_t)y * width + x];
    if (emissive(thisElem.emission)) {
      R_write(x, y) = R_read(x, y);
      return;
    }
    Rng rng = rngFor(PASS_SPATIAL, (uint32_t)iter, x, y);
    const V3 cam = gBuffer.cameraPosWS;
    const int k = P.spatialReuseNeighborCount;
    std::vector<std::pair<int, int>> nb;
    nb.push_back({x, y});
    int M = 1;
    for (int i = 0; i < k; ++i) {
      // Sampling::sampleDiskUniform, P/Sampling.cpp:78-87, truncated to int on assignment (:338)
      float theta = rng.value(2u * i, 0, 2.0f) * kPi;
      float r = std::sqrt(rng.value(2u * i + 1, 0, P.spatialReuseRadius));
      float ox = r * math.cos(theta);
      float oy = r * math.sin(theta);
      int nx = x + (int)ox, ny = y + (int)oy;
      nx = std::min(std::max(nx, 0), width - 1);  // glm::clamp to screen, :340
      ny = std::min(std::max(ny, 0), height - 1);
      const GBufferElement& ne = gBuffer.px[(size_t)ny * width + nx];
      if (emissive(ne.emission)) continue;
      if (P.rejectDissimilarNeighbors) {
        float normalSimilarity = dot(ne.worldSpaceNormal, thisElem.worldSpaceNormal);
        if (normalSimilarity < P.minNormalSimilarity) continue;
        float depthRatio = 0;
        if (ne.depth > 0) depthRatio = thisElem.depth / ne.depth;
        float halfDepthDiff = P.maxDepthDifference * 0.5f;
        float lo = 1.0f - halfDepthDiff, hi = 1.0f + halfDepthDiff;
        if (depthRatio < lo || depthRatio > hi) continue;
      }
      nb.push_back({nx, ny});
      M += 1;
    }
    auto G = [&](int i) -> const GBufferElement& { return gBuffer.px[(size_t)nb[i].second * width + nb[i].first]; };
    auto RR = [&](int i) -> const Reservoir& { return R_read(nb[i].first, nb[i].second); };
    Reservoir resultReservoir{};
    int confidenceSum = 0, confidenceSumNonCanonical = 0;
    for (size_t i = 0; i < nb.size(); ++i) {
      int c = RR((int)i).confidence;
      confidenceSum += c;
      if (i != 0) confidenceSumNonCanonical += c;
    }
    int selectedSampleIndex = 0;
    float rcpM = M > 0 ? 1.0f / (float)M : 0.0f;
    const int mode = P.spatialWeightCalc;
    const int n = (int)nb.size();
    for (int i = 0; i < n; ++i) {
      const Reservoir reservoir_i = RR(i);
      const LightSample& sample_i = reservoir_i.bestSample;
      float misNom = 0, misDenom = 0, misWeight = rcpM;
      if (mode == RB_SW_BALANCE_HEURISTIC) {
        misWeight = 0.0f;
        for (int j = 0; j < n; ++j) {
          float p_hat = evaluatePHat(sample_i, cam, G(j), true);
          misDenom += p_hat * RR(j).confidence;
          if (i == j) misNom = p_hat * reservoir_i.confidence;
        }
        if (misDenom > 0) misWeight = misNom / misDenom;
      }
      if (mode == RB_SW_PAIRWISE_MIS) {
        misWeight = 0.0f;
        if (i == 0) {
          float sum = 0.0f;
          float p_hat_c = evaluatePHat(sample_i, cam, G(i), true) * (float)reservoir_i.confidence;
          for (int j = 1; j < n; ++j) {
            float p_hat_j = evaluatePHat(sample_i, cam, G(j), true);
            float denom = p_hat_c + p_hat_j * (float)confidenceSumNonCanonical;
            if (denom > 0) {
              float confFract = (float)RR(j).confidence / (float)confidenceSum;
              sum += confFract * (p_hat_c / denom);
            }
          }
          misWeight = ((float)reservoir_i.confidence / (float)confidenceSum) + sum;
        } else {
          float p_hat_i = evaluatePHat(sample_i, cam, G(i), true);
          float p_hat_c = evaluatePHat(sample_i, cam, G(0), true);
          p_hat_i *= (float)confidenceSumNonCanonical;
          float denom = p_hat_i + p_hat_c * (float)RR(0).confidence;
          if (denom > 0 && confidenceSum > 0)
            misWeight = ((float)reservoir_i.confidence / (float)confidenceSum) * (p_hat_i / denom);
        }
      }
      float resamplingPhat = evaluatePHat(sample_i, cam, thisElem, true);
      float resamplingWeight = misWeight * resamplingPhat * reservoir_i.W;
      if (resultReservoir.addSample(sample_i, resamplingWeight, reservoir_i.confidence, rng, 2u * k + i))
        selectedSampleIndex = i;
    }
    float final_p_hat = evaluatePHat(resultReservoir.bestSample, cam, thisElem, true);
    if (mode == RB_SW_CONSTANT || mode == RB_SW_BALANCE_HEURISTIC || mode == RB_SW_PAIRWISE_MIS) {
      resultReservoir.W = final_p_hat > 0.0f ? resultReservoir.w_sum / final_p_hat : 0.0f;
    } else if (mode == RB_SW_CONSTANT_DEBIAS_Z_TERM) {
      int Z = 0;
      float correctionFactor = 1.0f;
      for (int i = 0; i < n; ++i)
        if (!testOcclusion(G(i).worldSpacePos, resultReservoir.bestSample.samplePoint)) Z += 1;
      if (Z > 0 && M > 0) correctionFactor = (1.0f / (float)Z) / rcpM;
      resultReservoir.W = final_p_hat > 0.0f ? correctionFactor * resultReservoir.w_sum / final_p_hat : 0.0f;
    } else if (mode == RB_SW_CONSTANT_DEBIAS_CONTRIB) {
      const LightSample selectedSample = RR(selectedSampleIndex).bestSample;
      float misNom = 0, misDenom = 0, contribWeight = 0, correctionFactor = 0;
      for (int i = 0; i < n; ++i) {
        float p_hat = evaluatePHat(selectedSample, cam, G(i), true);
        misDenom += p_hat * (float)RR(i).confidence;
        if (i == selectedSampleIndex) misNom = p_hat * (float)RR(i).confidence;
      }
      if (misDenom > 0) contribWeight = misNom / misDenom;
      if (M > 0) correctionFactor = contribWeight / rcpM;
      resultReservoir.W = final_p_hat > 0.0f ? correctionFactor * resultReservoir.w_sum / final_p_hat : 0.0f;
    }
    resultReservoir.capConfidence(P.confidenceCap);
    R_write(x, y) = resultReservoir;
  }

  // final shading loop, P/simpleguidx11.cpp:452-472 + Integrator::sanitize, P/Integrator.cpp:6-23
  void shadePixel(int x, int y) {
    V3 pixel;
    const Reservoir& r = R_read(x, y);
    const GBufferElement& g = gBuffer.px[(size_t)y * width + x];
    if (r.hasSample()) {
      V3 f = evaluateF(r.bestSample, gBuffer.cameraPosWS, g, true);
      pixel = f * r.W;
    } else
      pixel = g.emission;
    if (std::isnan(pixel.x) || std::isnan(pixel.y) || std::isnan(pixel.z)) pixel = v3(0);
    if (pixel.x < 0 || pixel.y < 0 || pixel.z < 0) pixel = v3(0);
    float* o = &frame[((size_t)y * width + x) * 3];
    o[0] = pixel.x;
    o[1] = pixel.y;
    o[2] = pixel.z;
  }

  // ---- N2: the ground-truth estimator ------------------------------------------------------------------
  // NEEPathIntegrator::integrateImpl2 with calcDI, without calcGI (P/NEEPathIntegrator.cpp:76-131) around
  // DirectMISIntegrator::calculateDirectLighting (P/DirectMISIntegrator.cpp:18-144). The material virtuals
  // (MaterialPhong::evaluateLightingGI / evaluateBRDF / getPdfForSample, P/MaterialPhong.cpp:18-119; MaterialLambert's,
  // P/MaterialLambert.cpp:10-31) take the primary ray's direction. Draw order as in the reference: lobe select, two
  // direction draws, light pick, two triangle draws (counter slots 0, 1-2, 4, 5-6).
  static float powerHeuristic(float pdf, float pdfOther) {  // :10-15
    const float pdf_sqr = pdf * pdf;
    const float pdfOther_sqr = pdfOther * pdfOther;
    return pdf_sqr / (pdfOther_sqr + pdf_sqr);
  }
  static V3 vdiv(const V3& v, float s) { return {v.x / s, v.y / s, v.z / s}; }  // glm vec3 / scalar: true divisions
  V3 phongBRDF_ray(const GBufferElement& e, const V3& rayDir, const V3& omega_i) const {  // MaterialPhong::evaluateBRDF, :69-92
    V3 f_r = e.diffuseColor * kOneOverPi;
    float nDotV = dot(-rayDir, e.worldSpaceNormal);
    float i_m = 1.0f / calc_I_M(nDotV, e.shininess);
    const V3 omega_r = normalize(reflect(rayDir, e.worldSpaceNormal));
    f_r = f_r + e.specularColor * i_m * math.pow(gmax(dot(omega_i, omega_r), 0.0f), e.shininess);
    return f_r;
  }
  float phongPdf_ray(const GBufferElement& e, const V3& rayDir, const V3& omega_i) const {  // ::getPdfForSample, :94-119
    float maxDiffuse = maxComponent(e.diffuseColor);
    float maxSpecular = maxComponent(e.specularColor);
    float pdfFactor = maxDiffuse / (maxDiffuse + maxSpecular);
    float pdf = cosw_pdf(e.worldSpaceNormal, omega_i) * pdfFactor;
    const V3 omega_r = normalize(reflect(rayDir, e.worldSpaceNormal));
    pdf += lobe_pdf(omega_i, omega_r, e.shininess) * (1.0f - pdfFactor);
    return pdf;
  }
  struct GIPayload {
    V3 omega_i, f_r;
    float pdf;
  };
  GIPayload evaluateLightingGI(const GBufferElement& e, const V3& rayDir, Rng& rng) const {
    if (e.materialType == RB_MAT_LAMBERT) {  // P/MaterialLambert.cpp:10-18
      V3 omega_i = cosw_sample(e.worldSpaceNormal, rng, 1, 2);
      float pdf = cosw_pdf(e.worldSpaceNormal, omega_i);
      return {omega_i, vdiv(e.diffuseColor, kPi), pdf};
    }
    // P/MaterialPhong.cpp:18-67
    float maxDiffuse = maxComponent(e.diffuseColor);
    float maxSpecular = maxComponent(e.specularColor);
    float r0 = rng.value(0, 0.0f, maxDiffuse + maxSpecular);
    float pdfFactor = maxDiffuse / (maxDiffuse + maxSpecular);
    V3 omega_i, f_r;
    const V3 omega_r = normalize(reflect(rayDir, e.worldSpaceNormal));
    if (r0 < maxDiffuse) {
      omega_i = cosw_sample(e.worldSpaceNormal, rng, 1, 2);
      f_r = e.diffuseColor * kOneOverPi;
    } else {
      omega_i = lobe_sample(omega_r, e.shininess, rng, 1, 2);
      float nDotV = dot(-rayDir, e.worldSpaceNormal);
      float i_m = 1.0f / calc_I_M(nDotV, e.shininess);
      f_r = e.specularColor * i_m * math.pow(gmax(dot(omega_i, omega_r), 0.0f), e.shininess);
    }
    float pdfDiffuse = cosw_pdf(e.worldSpaceNormal, omega_i) * pdfFactor;
    float pdfSpecular = lobe_pdf(omega_i, omega_r, e.shininess) * (1.0f - pdfFactor);
    float pdf = pdfDiffuse + pdfSpecular;
    if (dot(e.worldSpaceNormal, omega_i) < 0) return {omega_i, v3(0), pdf};
    return {omega_i, f_r, pdf};
  }
  V3 misEvaluateBRDFSample(const GBufferElement& e, const V3& rayDir, Rng& rng) {  // P/DirectMISIntegrator.cpp:92-144
    V3 L_direct = v3(0);
    GIPayload payloadGI = evaluateLightingGI(e, rayDir, rng);
    V3 org = e.worldSpacePos + P.normalOffset * e.worldSpaceNormal;
    HitInfo hi = intersect(org, payloadGI.omega_i, FLT_MIN + P.tnearOffset, FLT_MAX);
    if (hi.didHit) {
      const RbMaterial& m = scene.mats[scene.tris[hi.tri].material];
      if (m.emission[0] + m.emission[1] + m.emission[2] > 0) {  // Material::isEmissive, P/material.h:135-137
        V3 L_i{m.emission[0], m.emission[1], m.emission[2]};
        V3 lightDir = hi.hitPoint - e.worldSpacePos;
        float r_sqr = dot(lightDir, lightDir);
        lightDir = normalize(lightDir);
        float cosThetaI = gmax(dot(lightDir, e.worldSpaceNormal), 0.0f);
        float cosThetaY = gmax(dot(-lightDir, hi.normal), 0.0f);
        float areaMeasureFactor = cosThetaY / r_sqr;
        const Tri& tri = scene.tris[scene.emissive[hi.hitTriId]];
        float brdfPdf = payloadGI.pdf;
        float pdfAsIfLight = getPDFForTriangle(tri);
        float brdfPdfAreaMeasure = brdfPdf * areaMeasureFactor;
        float misWeight = powerHeuristic(brdfPdfAreaMeasure, pdfAsIfLight);
        L_direct = vdiv(misWeight * L_i * payloadGI.f_r * cosThetaI, brdfPdf);
      }
    }
    return L_direct;
  }
  V3 misEvaluateLightSample(const GBufferElement& e, const V3& rayDir, Rng& rng) {  // :38-90
    V3 L_direct = v3(0);
    if (!scene.lightsValid()) return L_direct;
    TriPick pick = pickTriangle(rng, 4);
    const Tri& T = scene.tris[scene.emissive[pick.emissiveIdx]];
    float r1 = rng.value(5, 0, 1);  // Sampling::sampleTriangle, P/Sampling.cpp:63-76
    float r2 = rng.value(6, 0, 1);
    float bx = 1.0f - std::sqrt(r1);
    float by = std::sqrt(r1) * (1.0f - r2);
    float bz = std::sqrt(r1) * r2;
    V3 samplePoint = T.p0 * bx + T.p1 * by + T.p2 * bz;
    V3 normal = normalize(T.n0 * bx + T.n1 * by + T.n2 * bz);
    float lightPdf = pick.pdf * (1.0f / T.area);
    if (lightPdf == 0) return v3(0);
    V3 lightDir = samplePoint - e.worldSpacePos;
    float r_sqr = dot(lightDir, lightDir);
    lightDir = normalize(lightDir);
    if (r_sqr == 0) return v3(0);
    float cosThetaI = gmax(dot(lightDir, e.worldSpaceNormal), 0.0f);
    float cosThetaY = gmax(dot(-lightDir, normal), 0.0f);
    float areaMeasureFactor = cosThetaY / r_sqr;
    if (cosThetaI > 0 && cosThetaY > 0 && !testOcclusion(e.worldSpacePos, samplePoint)) {
      const bool lambert = e.materialType == RB_MAT_LAMBERT;
      float pdfAsIfBrdf = lambert ? cosw_pdf(e.worldSpaceNormal, lightDir) : phongPdf_ray(e, rayDir, lightDir);
      float pdfAsIfBrdfAreaMeasure = pdfAsIfBrdf * areaMeasureFactor;
      const RbMaterial& m = scene.mats[T.material];
      V3 L_i{m.emission[0], m.emission[1], m.emission[2]};
      float misWeight = powerHeuristic(lightPdf, pdfAsIfBrdfAreaMeasure);
      if (misWeight > 0.0f) {
        float G = cosThetaI * cosThetaY / r_sqr;
        V3 f_r = lambert ? e.diffuseColor * kOneOverPi : phongBRDF_ray(e, rayDir, lightDir);
        L_direct = vdiv(misWeight * L_i * f_r * G, lightPdf);
      }
    }
    return L_direct;
  }
  static void sanitize(V3& l) {  // Integrator::sanitize, P/Integrator.cpp:6-23
    if (std::isnan(l.x) || std::isnan(l.y) || std::isnan(l.z)) l = v3(0);
    if (l.x < 0 || l.y < 0 || l.z < 0) l = v3(0);
  }
  // Raytracer::get_pixel (P/raytracer.cpp:40-46) for every pixel, serially in the legacy-RNG mode
  void produceMis(const RbCamera& cam, uint32_t frame_idx, uint32_t techniques) {
    frameIdx = frame_idx;
    memcpy(gBuffer.viewMat, cam.viewMat, sizeof(float) * 16);
    memcpy(gBuffer.invViewMat, cam.invViewMat, sizeof(float) * 16);
    gBuffer.cameraPosWS = {cam.pos[0], cam.pos[1], cam.pos[2]};
    gBuffer.focalLength = cam.focal_px;
    forPixels([&](int x, int y) {
      gBufferFillPass(x, y);  // Camera::GenerateRay (its two discarded draws in legacy mode) + intersectEmbree
      const GBufferElement& e = gBuffer.px[(size_t)y * width + x];
      V3 out;
      if (e.geomID == 0xFFFFFFFFu || emissive(e.emission)) {
        out = e.emission;  // background colour / emitter seen from the camera vertex
      } else {
        Rng crng = rngFor(PASS_GBUF, 0, x, y);
        crng.mode = 0;  // (direction only: the draws were consumed by gBufferFillPass)
        V3 org, dir;
        generateRay(x, y, &org, &dir, crng);
        Rng rng = rngFor(PASS_MIS, 0, x, y);
        V3 L_direct = v3(0);
        if (techniques & 1u) L_direct = L_direct + misEvaluateBRDFSample(e, dir, rng);
        if (techniques & 2u) L_direct = L_direct + misEvaluateLightSample(e, dir, rng);
        sanitize(L_direct);
        out = v3(0) + L_direct;
      }
      float* o = &frame[3 * ((size_t)y * width + x)];
      o[0] = out.x, o[1] = out.y, o[2] = out.z;
    });
  }

  template <class F>
  void forPixels(F&& f) {
    if (rng_mode == 1) {  // _DEBUG build: serial, y outer, x inner
      for (int y = band_y0; y < band_y1; ++y)
        for (int x = 0; x < width; ++x) f(x, y);
    } else if (!rows.empty()) {
      const int n = (int)rows.size();
#pragma omp parallel for schedule(dynamic, 1)
      for (int i = 0; i < n; ++i)
        for (int x = 0; x < width; ++x) f(x, rows[i]);
    } else {
#pragma omp parallel for schedule(dynamic, 1)
      for (int y = band_y0; y < band_y1; ++y)
        for (int x = 0; x < width; ++x) f(x, y);
    }
  }

  // SimpleGuiDX11::produceRestir, P/simpleguidx11.cpp:359-487
  void produceRestir(const RbCamera& cam, uint32_t frame_idx, double* pass_s) {
    using clk = std::chrono::steady_clock;
    auto sec = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    frameIdx = frame_idx;
    auto t0 = clk::now();
    memcpy(gBuffer.viewMat, cam.viewMat, sizeof(float) * 16);
    memcpy(gBuffer.invViewMat, cam.invViewMat, sizeof(float) * 16);
    gBuffer.cameraPosWS = {cam.pos[0], cam.pos[1], cam.pos[2]};
    gBuffer.focalLength = cam.focal_px;
    auto run = [&](auto&& body) { forPixels(body); };
    run([&](int x, int y) { gBufferFillPass(x, y); });
    auto t1 = clk::now();
    run([&](int x, int y) { initialRenderPass(x, y); });
    auto t2 = clk::now();
    if (P.doVisibilityPass) run([&](int x, int y) { visibilityPass(x, y); });
    auto t3 = clk::now();
    if (P.doTemporalReuse && frame_idx > 0) {
      swapReservoirBuffers();
      run([&](int x, int y) { temporalReusePass(x, y); });
    }
    auto t4 = clk::now();
    if (P.doSpatialReuse) {
      for (int i = 0; i < P.spatialPassCount; ++i) {
        swapReservoirBuffers();
        run([&](int x, int y) { spatialReusePass(x, y, i); });
      }
    }
    auto t5 = clk::now();
    swapReservoirBuffers();
    run([&](int x, int y) { shadePixel(x, y); });
    auto t6 = clk::now();
    // memcpy(reservoirsLastFrame, ...) and gBufferLastFrame.setDataFrom(gBuffer), :478-481 — by rotation
    std::swap(lastIdx, readIdx);  // last' = read; old last becomes a ping-pong buffer
    std::swap(gBufferLastFrame.px, gBuffer.px);
    gBufferLastFrame.cameraPosWS = gBuffer.cameraPosWS;
    memcpy(gBufferLastFrame.viewMat, gBuffer.viewMat, sizeof(float) * 16);
    memcpy(gBufferLastFrame.invViewMat, gBuffer.invViewMat, sizeof(float) * 16);
    gBufferLastFrame.focalLength = gBuffer.focalLength;
    auto t7 = clk::now();
    if (pass_s) {
      pass_s[0] = sec(t0, t1);
      pass_s[1] = sec(t1, t2);
      pass_s[2] = sec(t2, t3);
      pass_s[3] = sec(t3, t4);
      pass_s[4] = sec(t4, t5);
      pass_s[5] = sec(t5, t6);
      pass_s[6] = sec(t6, t7);
      pass_s[7] = sec(t0, t7);
    }
  }
};

}  // namespace orc

// ============================================================================
// C interface (ctypes)
// ============================================================================
using namespace orc;

extern "C" {

void* orc_create(int width, int height, uint32_t seed, int rng_mode, int math_mode, int tracer_mode, int cache_iim) {
  Oracle* o = new Oracle();
  o->width = width;
  o->height = height;
  o->band_y0 = 0;
  o->band_y1 = height;
  o->seed = seed;
  o->rng_mode = rng_mode;
  o->math.mode = math_mode;
  o->scene.tracer = tracer_mode;
  o->cache_iim = cache_iim;
  {
    RbParams& p = o->P;
    memset(&p, 0, sizeof(p));
    p.M_Area = 1;
    p.M_Brdf = 1;
    p.spatialReuseNeighborCount = 5;
    p.spatialPassCount = 1;
    p.confidenceCap = 20;
    p.spatialReuseRadius = 30;
    p.minNormalSimilarity = 0.85f;
    p.maxDepthDifference = 0.2f;
    p.tnearOffset = 0.01f;
    p.tfarOffset = 0.001f;
    p.normalOffset = 0.001f;
    p.bgColor[0] = p.bgColor[1] = p.bgColor[2] = 0.5f;
  }
  size_t n = (size_t)width * height;
  o->gBuffer.px.assign(n, GBufferElement{});
  o->gBufferLastFrame.px.assign(n, GBufferElement{});
  for (auto& r : o->res) r.assign(n, Reservoir{});
  o->frame.assign(n * 3, 0.0f);
  return o;
}
void orc_destroy(void* h) { delete (Oracle*)h; }

int orc_set_band(void* h, int y0, int y1) {
  Oracle* o = (Oracle*)h;
  if (y0 < 0 || y1 > o->height || y0 > y1) return -1;
  o->band_y0 = y0;
  o->band_y1 = y1;
  o->rows.clear();
  return 0;
}
// bench.py's bounded sample of a frame: several row segments spread over the image (frame cost varies along y)
int orc_set_row_segments(void* h, const int* y0y1, int n_segments) {
  Oracle* o = (Oracle*)h;
  o->rows.clear();
  for (int s = 0; s < n_segments; ++s) {
    if (y0y1[2 * s] < 0 || y0y1[2 * s + 1] > o->height || y0y1[2 * s] > y0y1[2 * s + 1]) return -1;
    for (int y = y0y1[2 * s]; y < y0y1[2 * s + 1]; ++y) o->rows.push_back(y);
  }
  return 0;
}
// torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU arm sets its thread count explicitly
void orc_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}
int orc_max_threads(void) { return omp_get_max_threads(); }

int orc_upload_scene(void* h, const RbSceneDesc* sd) {
  Oracle* o = (Oracle*)h;
  Scene& S = o->scene;
  S.tris.clear();
  S.emissive.clear();
  S.tangents.clear();
  bool any_tan = false;
  for (uint32_t s = 0; s < sd->n_surfaces; ++s) any_tan = any_tan || (sd->surfaces[s].tangent && sd->surfaces[s].n_tris);
  S.mats.assign(sd->materials, sd->materials + sd->n_materials);
  S.textures.clear();
  S.mat_tex.clear();
  int triIdCtr = 0;
  for (uint32_t s = 0; s < sd->n_surfaces; ++s) {
    const RbSurface& sf = sd->surfaces[s];
    if (sf.material >= sd->n_materials) return -1;
    const RbMaterial& m = S.mats[sf.material];
    bool isEmissive = m.emission[0] + m.emission[1] + m.emission[2] > 0;  // Material::isEmissive, P/material.h:135-137
    for (uint32_t i = 0; i < sf.n_tris; ++i) {
      Tri T;
      const float* p = sf.pos + 9 * (size_t)i;
      const float* n = sf.normal + 9 * (size_t)i;
      T.p0 = {p[0], p[1], p[2]};
      T.p1 = {p[3], p[4], p[5]};
      T.p2 = {p[6], p[7], p[8]};
      T.n0 = {n[0], n[1], n[2]};
      T.n1 = {n[3], n[4], n[5]};
      T.n2 = {n[6], n[7], n[8]};
      if (sf.uv) {
        const float* w = sf.uv + 6 * (size_t)i;
        for (int k = 0; k < 3; ++k) T.uv[k][0] = w[2 * k], T.uv[k][1] = w[2 * k + 1];
      }
      if (any_tan) {
        const float* w = sf.tangent ? sf.tangent + 9 * (size_t)i : nullptr;
        for (int k = 0; k < 3; ++k) S.tangents.push_back(w ? V3{w[3 * k], w[3 * k + 1], w[3 * k + 2]} : V3{0, 0, 0});
      }
      T.e1 = T.p1 - T.p0;
      T.e2 = T.p2 - T.p0;
      T.geom = s;
      T.prim = i;
      T.material = sf.material;
      T.area = 0.5f * length(cross(T.e1, T.e2));
      T.emissive_id = -1;
      if (isEmissive) {
        T.emissive_id = triIdCtr++;
        S.emissive.push_back((int)S.tris.size());
      }
      S.tris.push_back(T);
    }
  }
  S.buildLights();
  if (S.tracer == 1) S.buildBvh();
  o->have_scene = true;
  return 0;
}

int orc_set_textures(void* h, const RbTexture* textures, uint32_t n_textures, const RbMaterialTextures* per_material, uint32_t n_materials) {
  Oracle* o = (Oracle*)h;
  Scene& S = o->scene;
  if (n_materials != S.mats.size()) return -1;
  S.textures.resize(n_textures);
  for (uint32_t t = 0; t < n_textures; ++t) {
    const RbTexture& T = textures[t];
    S.textures[t].width = T.width, S.textures[t].height = T.height;
    S.textures[t].scan_width = T.scan_width, S.textures[t].pixel_size = T.pixel_size;
    const unsigned char* d = (const unsigned char*)T.data;
    S.textures[t].data.assign(d, d + (size_t)T.scan_width * T.height);
  }
  S.mat_tex.assign(per_material, per_material + n_materials);
  return 0;
}

int orc_set_params(void* h, const RbParams* p) {
  Oracle* o = (Oracle*)h;
  if (p->useSkybox && o->scene.sky.width == 0) return -4;  // the reference would dereference a null skybox
  o->P = *p;
  return 0;
}

// Scene::setSkybox (P/Scene.cpp:47-50): SphericalMap over a Texture with the constructor's defaults (BILINEAR, CLAMP_TO_EDGE)
int orc_set_sky(void* h, const RbTexture* sky) {
  Oracle* o = (Oracle*)h;
  Scene::Tex& T = o->scene.sky;
  T = Scene::Tex{};
  if (!sky) {
    o->P.useSkybox = 0;
    return 0;
  }
  T.width = sky->width, T.height = sky->height, T.scan_width = sky->scan_width, T.pixel_size = sky->pixel_size;
  const unsigned char* d = (const unsigned char*)sky->data;
  T.data.assign(d, d + (size_t)sky->scan_width * sky->height);
  T.clamp = true;
  return 0;
}

int orc_render_frame(void* h, const RbCamera* cam, uint32_t frame_idx, float* rgb_out, double* pass_seconds) {
  Oracle* o = (Oracle*)h;
  if (!o->have_scene) return -3;
  o->produceRestir(*cam, frame_idx, pass_seconds);
  if (rgb_out) memcpy(rgb_out, o->frame.data(), o->frame.size() * sizeof(float));
  return 0;
}

int orc_render_mis_frame(void* h, const RbCamera* cam, uint32_t frame_idx, uint32_t techniques, float* rgb_out) {
  Oracle* o = (Oracle*)h;
  if (!o->have_scene) return -3;
  o->produceMis(*cam, frame_idx, techniques);
  if (rgb_out) memcpy(rgb_out, o->frame.data(), o->frame.size() * sizeof(float));
  return 0;
}

// ---- after the path: the Producer loop's accumulate / tonemap / statistics ----------------------------
// Utils::aces, P/utils.cpp:190-197 (float constants a..e; glm vec3 arithmetic is component-wise; glm::clamp)
static void ref_aces(float* x3, const Math&) {
  const float a = 2.51, b = 0.03, c = 2.43, d = 0.59, e = 0.14;
  for (int i = 0; i < 3; ++i) {
    const float x = x3[i];
    const float v = (x * (a * x + b)) / (x * (c * x + d) + e);
    x3[i] = std::min(std::max(v, 0.0f), 1.0f);  // glm::clamp = min(max(x, minVal), maxVal)
  }
}
// Utils::compress(float&), P/utils.cpp:220-230 — note the double literal in "u <= 0.0031308"
static void ref_compress(float& u, const Math& m) {
  if (u <= 0.0f)
    u = 0.0f;
  else if (u >= 1.0f)
    u = 1.0f;
  else if (u <= 0.0031308)
    u *= 12.92f;
  else
    u = 1.055f * m.pow(u, 1.0f / 2.4f) - 0.055f;
}
void orc_aces(float* rgb) { ref_aces(rgb, Math{}); }
float orc_compress(int math_mode, float u) {
  Math m;
  m.mode = math_mode;
  ref_compress(u, m);
  return u;
}
// test hook: put an arbitrary image into frame_data
int orc_set_frame(void* h, const float* rgb) {
  Oracle* o = (Oracle*)h;
  o->frame.assign(rgb, rgb + (size_t)o->width * o->height * 3);
  return 0;
}
// P/simpleguidx11.cpp:246-253 (accumulate), :262-295 (display), :308-326 (mean / variance), over the band rows
int orc_accumulate_display(void* h, uint32_t acc_frame_ctr, int tonemap, int gamma_correct, float* display_rgba_out, double* stats4) {
  Oracle* o = (Oracle*)h;
  const size_t n = (size_t)o->width * o->height;
  if (o->accumulator.size() != n * 3) o->accumulator.assign(n * 3, 0.0f);
  if (o->display.size() != n * 4) o->display.assign(n * 4, 0.0f);
  if (o->frame.size() != n * 3) return -1;
  const float a = 1.0f / static_cast<float>(acc_frame_ctr + 1);
  double pixelSum = 0, pixelSqrSum = 0;
  for (int y = o->band_y0; y < o->band_y1; ++y)
    for (int x = 0; x < o->width; ++x) {
      const size_t offset = (size_t)y * o->width + x;
      float* acc = &o->accumulator[3 * offset];
      const float* fr = &o->frame[3 * offset];
      for (int c = 0; c < 3; ++c) acc[c] = acc[c] * (1.0f - a) + fr[c] * a;  // glm::mix: x * (1 - a) + y * a
      float pix[3] = {acc[0], acc[1], acc[2]};
      if (tonemap) ref_aces(pix, o->math);
      if (gamma_correct)
        for (int c = 0; c < 3; ++c) ref_compress(pix[c], o->math);
      float* dsp = &o->display[4 * offset];
      dsp[0] = pix[0], dsp[1] = pix[1], dsp[2] = pix[2], dsp[3] = 1.0f;
      const float pixelMean = (acc[0] + acc[1] + acc[2]) / 3.0f;
      pixelSum += pixelMean;
      pixelSqrSum += pixelMean * pixelMean;
    }
  const double count = (double)(o->band_y1 - o->band_y0) * o->width;
  if (stats4) {
    stats4[0] = pixelSum, stats4[1] = pixelSqrSum;
    stats4[2] = pixelSum / count;
    stats4[3] = pixelSqrSum / count - stats4[2] * stats4[2];
  }
  if (display_rgba_out) memcpy(display_rgba_out, o->display.data(), n * 16);
  return 0;
}

// same ids and packed layouts as rb_readback (include/restir_b200.h RbBufferId)
int orc_readback(void* h, int id, void* dst, size_t bytes) {
  Oracle* o = (Oracle*)h;
  size_t n = (size_t)o->width * o->height;
  // after produceRestir the frame's G-buffer lives in gBufferLastFrame and the final reservoirs in res[lastIdx]
  const std::vector<GBufferElement>& G = o->gBufferLastFrame.px;
  const std::vector<Reservoir>& R = o->res[o->lastIdx];
  auto need = [&](size_t b) { return bytes >= b; };
  float* f = (float*)dst;
  uint32_t* u = (uint32_t*)dst;
  int32_t* ii = (int32_t*)dst;
  switch (id) {
    case RB_BUF_GBUF_POS_DEPTH:
      if (!need(n * 16)) return -1;
      for (size_t i = 0; i < n; ++i) {
        f[4 * i] = G[i].worldSpacePos.x, f[4 * i + 1] = G[i].worldSpacePos.y, f[4 * i + 2] = G[i].worldSpacePos.z;
        f[4 * i + 3] = G[i].depth;
      }
      return 0;
    case RB_BUF_GBUF_NORMAL_SHIN:
      if (!need(n * 16)) return -1;
      for (size_t i = 0; i < n; ++i) {
        f[4 * i] = G[i].worldSpaceNormal.x, f[4 * i + 1] = G[i].worldSpaceNormal.y, f[4 * i + 2] = G[i].worldSpaceNormal.z;
        f[4 * i + 3] = G[i].shininess;
      }
      return 0;
    case RB_BUF_GBUF_DIFFUSE_IIM:
      if (!need(n * 16)) return -1;
      for (size_t i = 0; i < n; ++i) {
        f[4 * i] = G[i].diffuseColor.x, f[4 * i + 1] = G[i].diffuseColor.y, f[4 * i + 2] = G[i].diffuseColor.z;
        f[4 * i + 3] = G[i].invIM;
      }
      return 0;
    case RB_BUF_GBUF_SPEC_TYPE:
      if (!need(n * 16)) return -1;
      for (size_t i = 0; i < n; ++i) {
        f[4 * i] = G[i].specularColor.x, f[4 * i + 1] = G[i].specularColor.y, f[4 * i + 2] = G[i].specularColor.z;
        u[4 * i + 3] = (uint32_t)G[i].materialType | (emissive(G[i].emission) ? 0x100u : 0u);
      }
      return 0;
    case RB_BUF_GBUF_EMISSION:
      if (!need(n * 16)) return -1;
      for (size_t i = 0; i < n; ++i) {
        f[4 * i] = G[i].emission.x, f[4 * i + 1] = G[i].emission.y, f[4 * i + 2] = G[i].emission.z;
        f[4 * i + 3] = 0;
      }
      return 0;
    case RB_BUF_HIT_IDS:
      if (!need(n * 8)) return -1;
      for (size_t i = 0; i < n; ++i) u[2 * i] = G[i].geomID, u[2 * i + 1] = G[i].primID;
      return 0;
    case RB_BUF_RES_POINT_WSUM:
      if (!need(n * 16)) return -1;
      for (size_t i = 0; i < n; ++i) {
        f[4 * i] = R[i].bestSample.samplePoint.x, f[4 * i + 1] = R[i].bestSample.samplePoint.y,
              f[4 * i + 2] = R[i].bestSample.samplePoint.z;
        f[4 * i + 3] = R[i].w_sum;
      }
      return 0;
    case RB_BUF_RES_NORMAL_W:
      if (!need(n * 16)) return -1;
      for (size_t i = 0; i < n; ++i) {
        f[4 * i] = R[i].bestSample.sampleNormal.x, f[4 * i + 1] = R[i].bestSample.sampleNormal.y,
              f[4 * i + 2] = R[i].bestSample.sampleNormal.z;
        f[4 * i + 3] = R[i].W;
      }
      return 0;
    case RB_BUF_RES_LI_CONF:
      if (!need(n * 16)) return -1;
      for (size_t i = 0; i < n; ++i) {
        f[4 * i] = R[i].bestSample.L_i.x, f[4 * i + 1] = R[i].bestSample.L_i.y, f[4 * i + 2] = R[i].bestSample.L_i.z;
        ii[4 * i + 3] = R[i].confidence;
      }
      return 0;
    case RB_BUF_RES_LIGHT_IDX:
      if (!need(n * 4)) return -1;
      for (size_t i = 0; i < n; ++i) ii[i] = R[i].bestSample.lightIdx;
      return 0;
    case RB_BUF_FRAME_RGB:
      if (!need(n * 12)) return -1;
      memcpy(dst, o->frame.data(), n * 12);
      return 0;
    case RB_BUF_ACCUMULATOR:
      if (!need(n * 12) || o->accumulator.size() != n * 3) return -1;
      memcpy(dst, o->accumulator.data(), n * 12);
      return 0;
    case RB_BUF_DISPLAY:
      if (!need(n * 16) || o->display.size() != n * 4) return -1;
      memcpy(dst, o->display.data(), n * 16);
      return 0;
    case RB_BUF_ALIAS_PROB:
      if (!need(o->scene.alias_prob.size() * 4)) return -1;
      memcpy(dst, o->scene.alias_prob.data(), o->scene.alias_prob.size() * 4);
      return 0;
    case RB_BUF_ALIAS_IDX:
      if (!need(o->scene.alias_idx.size() * 4)) return -1;
      memcpy(dst, o->scene.alias_idx.data(), o->scene.alias_idx.size() * 4);
      return 0;
    case RB_BUF_LIGHT_CDF:
      if (!need(o->scene.cdf.size() * 4)) return -1;
      memcpy(dst, o->scene.cdf.data(), o->scene.cdf.size() * 4);
      return 0;
  }
  return -1;
}

void orc_counters(void* h, uint64_t* out3, int reset) {
  Oracle* o = (Oracle*)h;
  out3[0] = out3[1] = out3[2] = 0;
  for (auto& c : o->ctrs) {
    out3[0] += c.closest;
    out3[1] += c.any_written;
    out3[2] += c.any_traced;
    if (reset) c = Counters{};
  }
}
void orc_temporal_stats(void* h, uint64_t* out5, int reset) {
  Oracle* o = (Oracle*)h;
  for (int k = 0; k < 5; ++k) out5[k] = 0;
  for (auto& c : o->ctrs)
    for (int k = 0; k < 5; ++k) {
      out5[k] += c.temporal[k];
      if (reset) c.temporal[k] = 0;
    }
}
uint32_t orc_num_emissive(void* h) { return (uint32_t)((Oracle*)h)->scene.emissive.size(); }
uint32_t orc_num_triangles(void* h) { return (uint32_t)((Oracle*)h)->scene.tris.size(); }
float orc_total_emissive_area(void* h) { return ((Oracle*)h)->scene.totalSurface; }

int orc_trace_closest(void* h, const RbRay* rays, RbHit* hits, uint32_t n) {
  Oracle* o = (Oracle*)h;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const RbRay& r = rays[i];
    Hit hit = o->scene.closest({r.org_x, r.org_y, r.org_z}, {r.dir_x, r.dir_y, r.dir_z}, r.tnear, r.tfar);
    RbHit& out = hits[i];
    if (hit.tri >= 0) {
      out.t = hit.t, out.u = hit.u, out.v = hit.v;
      out.geomID = o->scene.tris[hit.tri].geom;
      out.primID = o->scene.tris[hit.tri].prim;
    } else {
      out.t = r.tfar, out.u = out.v = 0;
      out.geomID = out.primID = 0xFFFFFFFFu;
    }
  }
  return 0;
}
int orc_trace_occluded(void* h, const RbRay* rays, uint8_t* occ, uint32_t n) {
  Oracle* o = (Oracle*)h;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const RbRay& r = rays[i];
    occ[i] = o->scene.occluded({r.org_x, r.org_y, r.org_z}, {r.dir_x, r.dir_y, r.dir_z}, r.tnear, r.tfar) ? 1 : 0;
  }
  return 0;
}

// ---- leaf functions exposed for unit tests and for the oracle/_ref pin -------
float orc_dm_pow(float x, float y) { return dm::powf_(x, y); }
float orc_dm_sin(float x) { return dm::sinf_(x); }
float orc_dm_cos(float x) { return dm::cosf_(x); }
float orc_dm_exp(float x) { return dm::expf_(x); }
float orc_dm_lgamma(float x) { return dm::lgammaf_(x); }
float orc_dm_ibeta(float a, float b, float x) { return dm::ibetaf_(a, b, x); }
float orc_dm_atan2(float y, float x) { return dm::atan2f_(y, x); }
float orc_dm_acos(float x) { return dm::acosf_(x); }
uint32_t orc_rng_bits(uint32_t seed, uint32_t frame, uint32_t pass, uint32_t iter, uint32_t pixel, uint32_t slot) {
  return rng_bits(rng_key(seed, frame, pass, iter, pixel), slot);
}

static Oracle make_leaf(int math_mode) {
  Oracle o;
  o.math.mode = math_mode;
  return o;
}
float orc_calc_I_M(int math_mode, float nDotV, float n) { return make_leaf(math_mode).calc_I_M(nDotV, n); }
static GBufferElement elem_from(const float* e) {
  GBufferElement g;
  g.worldSpacePos = {e[0], e[1], e[2]};
  g.worldSpaceNormal = {e[3], e[4], e[5]};
  g.diffuseColor = {e[6], e[7], e[8]};
  g.specularColor = {e[9], e[10], e[11]};
  g.shininess = e[12];
  g.materialType = RB_MAT_PHONG;
  return g;
}
// elem = {pos3, normal3, diffuse3, specular3, shininess}
void orc_phong_evalBRDF(int math_mode, const float* elem, const float* cam, const float* wi, float* out3) {
  Oracle o = make_leaf(math_mode);
  V3 r = o.phong_evalBRDF(elem_from(elem), {cam[0], cam[1], cam[2]}, {wi[0], wi[1], wi[2]});
  out3[0] = r.x, out3[1] = r.y, out3[2] = r.z;
}
float orc_phong_evalPdf(int math_mode, const float* elem, const float* cam, const float* wi) {
  Oracle o = make_leaf(math_mode);
  return o.phong_evalPdf(elem_from(elem), {cam[0], cam[1], cam[2]}, {wi[0], wi[1], wi[2]});
}
// legacy-stream leaf samplers: run with a caller-seeded mt19937 so they can be compared
// draw-for-draw with the reference sources compiled in oracle/_ref
// n calls on one stream; elems = n x 13 floats, cams = n x 3 floats
void orc_legacy_phong_sampleBRDF(int math_mode, uint32_t mt_seed, int n, const float* elems, const float* cams, float* out4n) {
  Oracle o = make_leaf(math_mode);
  o.rng_mode = 1;
  o.legacy_gen.seed(mt_seed);
  Rng rng = o.rngFor(0, 0, 0, 0);
  for (int i = 0; i < n; ++i) {
    GBufferElement g = elem_from(elems + 13 * i);
    const float* cam = cams + 3 * i;
    auto s = o.phong_sampleBRDF(g, {cam[0], cam[1], cam[2]}, rng, 0);
    out4n[4 * i] = s.omega_i.x, out4n[4 * i + 1] = s.omega_i.y, out4n[4 * i + 2] = s.omega_i.z, out4n[4 * i + 3] = s.pdf;
  }
}
void orc_legacy_sampleDiskUniform(int math_mode, uint32_t mt_seed, int n, float radius, float* out2n) {
  Oracle o = make_leaf(math_mode);
  o.rng_mode = 1;
  o.legacy_gen.seed(mt_seed);
  Rng rng = o.rngFor(0, 0, 0, 0);
  for (int i = 0; i < n; ++i) {
    float theta = rng.value(0, 0, 2.0f) * kPi;
    float r = std::sqrt(rng.value(1, 0, radius));
    out2n[2 * i] = r * o.math.cos(theta);
    out2n[2 * i + 1] = r * o.math.sin(theta);
  }
}
// tri = {p0,p1,p2,n0,n1,n2} 18 floats; out = {point3, normal3, pdf} per sample
void orc_legacy_sampleTriangle(uint32_t mt_seed, int n, const float* tri, float* out7n) {
  Oracle o;
  o.rng_mode = 1;
  o.legacy_gen.seed(mt_seed);
  Rng rng = o.rngFor(0, 0, 0, 0);
  V3 p0{tri[0], tri[1], tri[2]}, p1{tri[3], tri[4], tri[5]}, p2{tri[6], tri[7], tri[8]};
  V3 n0{tri[9], tri[10], tri[11]}, n1{tri[12], tri[13], tri[14]}, n2{tri[15], tri[16], tri[17]};
  float area = 0.5f * length(cross(p1 - p0, p2 - p0));
  for (int i = 0; i < n; ++i) {
    float r1 = rng.value(0, 0, 1);
    float r2 = rng.value(1, 0, 1);
    float x = 1.0f - std::sqrt(r1);
    float y = std::sqrt(r1) * (1.0f - r2);
    float z = std::sqrt(r1) * r2;
    V3 s = p0 * x + p1 * y + p2 * z;
    V3 nn = normalize(n0 * x + n1 * y + n2 * z);
    float* q = out7n + 7 * i;
    q[0] = s.x, q[1] = s.y, q[2] = s.z, q[3] = nn.x, q[4] = nn.y, q[5] = nn.z, q[6] = 1.0f / area;
  }
}
// light CDF pick with the legacy stream: out = {index, pdf} per draw (scene must be uploaded)
void orc_legacy_cdf_pick(void* h, uint32_t mt_seed, int n, float* out2n) {
  Oracle* o = (Oracle*)h;
  int save_mode = o->rng_mode, save_ls = o->P.lightSampler;
  o->rng_mode = 1;
  o->P.lightSampler = RB_LS_CDF;
  o->legacy_gen.seed(mt_seed);
  Rng rng = o->rngFor(0, 0, 0, 0);
  for (int i = 0; i < n; ++i) {
    auto p = o->pickTriangle(rng, 0);
    out2n[2 * i] = (float)p.emissiveIdx;
    out2n[2 * i + 1] = p.pdf;
  }
  o->rng_mode = save_mode;
  o->P.lightSampler = save_ls;
}
void orc_legacy_floats(uint32_t mt_seed, int n, float* out) {
  std::mt19937 g(mt_seed);
  std::uniform_real_distribution<float> d(0.0f, 1.0f);
  for (int i = 0; i < n; ++i) out[i] = d(g);
}
int orc_have_boost(void) {
#ifdef ORACLE_HAVE_BOOST
  return 1;
#else
  return 0;
#endif
}
}
