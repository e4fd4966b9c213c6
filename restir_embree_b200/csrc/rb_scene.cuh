// rb_scene.cuh — device scene records and the 8-wide quantised BVH traversal.
//
// Replaces what the reference keeps inside Embree (rtcIntersect1 / rtcOccluded1,
// P/Intersection.h:43-83) plus the attribute fetch of getGeometryAttributes
// (:85-113). B200 has no RT cores: this is ordinary SM code.
//
// Node8 (80 B = 5 x 16 B, AoSoA, 16-byte aligned):
//   n0 = { origin.xyz, bits: ex | ey<<8 | ez<<16 | imask<<24 }   ex.. = biased exponents of the grid step
//   n1 = { child_base, tri_base, meta[0..3], meta[4..7] }
//   n2 = { qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7] }
//   n3 = { qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7] }
//   n4 = { qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7] }
// Child boxes are 8-bit grid coordinates relative to the node origin, rounded
// outwards. meta[s] of slot s: 0 = empty; internal child: (1<<5) | (24+s);
// leaf: (unary triangle count in bits 5..7) | triangle offset (0..23).
// Children sit in slots chosen at build time so that (slot ^ ray octant) orders
// them front to back (compressed-wide-BVH scheme of Ylitie, Karras & Laine 2017).
#ifndef RB_SCENE_CUH_
#define RB_SCENE_CUH_

#include "rb_common.cuh"

namespace rb {

#define RB_LEAF_MAX 3
#define RB_STACK_MAX 40

struct SceneDev {
  // geometry
  const F4* node8;        // [5 * n_nodes]
  const F4* tri_isect;    // [3 * n_leaf_tris] leaf order: {v0.xyz,e1.x} {e1.y,e1.z,e2.x,e2.y} {e2.z,bits(tri id),0,0}
  const F4* tri_normals;  // [3 * n_tris] scene order: {n0.xyz,n1.x} {n1.y,n1.z,n2.x,n2.y} {n2.z,0,0,0}
  const U4* tri_info;     // [n_tris] scene order: {geomID, primID, material, emissive id (int, -1 none)}
  const F4* mat;          // [3 * n_mat]: {diffuse.rgb, shininess} {specular.rgb, bits(type)} {emission.rgb, ior}
  // emissive triangles (TriangleCDF::tris order, P/ModelLoader.cpp:301-306)
  const F4* light;  // [6 * n_lights]: {p0,area} {p1,area/total} {p2,1/area} {n0,Le.r} {n1,Le.g} {n2,Le.b}
  const float* cdf;
  const float* alias_prob;
  const uint32_t* alias_idx;
  uint32_t n_lights;
  uint32_t n_tris;
  uint32_t n_nodes;
  float total_area;
};

struct HitRec {
  float t, u, v;
  uint32_t tri;  // scene-order triangle index, 0xFFFFFFFF = miss
};

// Möller–Trumbore with the fixed operation order of the parity contract
// (DESIGN.md "ray/triangle arithmetic"; the oracle spells the same sequence).
RB_HD bool tri_test(const V3& o, const V3& d, const F4& a, const F4& b, const F4& c, float tnear, float tfar, float* t,
                    float* u, float* v) {
  const float e1x = a.w, e1y = b.x, e1z = b.y, e2x = b.z, e2y = b.w, e2z = c.x;
  float px = fmaf_(d.y, e2z, -(d.z * e2y));
  float py = fmaf_(d.z, e2x, -(d.x * e2z));
  float pz = fmaf_(d.x, e2y, -(d.y * e2x));
  float det = fmaf_(e1z, pz, fmaf_(e1y, py, e1x * px));
  if (det == 0.0f) return false;
  float inv = fdiv_(1.0f, det);
  float tx = o.x - a.x, ty = o.y - a.y, tz = o.z - a.z;
  float uu = fmaf_(tz, pz, fmaf_(ty, py, tx * px)) * inv;
  if (!(uu >= 0.0f && uu <= 1.0f)) return false;
  float qx = fmaf_(ty, e1z, -(tz * e1y));
  float qy = fmaf_(tz, e1x, -(tx * e1z));
  float qz = fmaf_(tx, e1y, -(ty * e1x));
  float vv = fmaf_(d.z, qz, fmaf_(d.y, qy, d.x * qx)) * inv;
  if (!(vv >= 0.0f && uu + vv <= 1.0f)) return false;
  float tt = fmaf_(e2z, qz, fmaf_(e2y, qy, e2x * qx)) * inv;
  if (!(tt > tnear && tt < tfar)) return false;
  *t = tt;
  *u = uu;
  *v = vv;
  return true;
}

RB_HD uint32_t byte_of(uint32_t w, int i) { return (w >> (8 * i)) & 0xFFu; }
RB_HD int bfind(uint32_t x) {  // index of the highest set bit, x != 0
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)x);
#else
  return 31 - __builtin_clz(x);
#endif
}
RB_HD int popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// One ray against the BVH. ANY = true: stop at the first hit with tnear < t < tfar
// (rtcOccluded1); ANY = false: closest hit, ties by smaller scene-order index.
template <bool ANY>
RB_HD bool trace8(const SceneDev& sc, const V3& o, const V3& d, float tnear, float tfar, HitRec* out) {
  HitRec best;
  best.t = tfar;
  best.u = best.v = 0;
  best.tri = 0xFFFFFFFFu;
  if (sc.n_nodes == 0) {
    if (out) *out = best;
    return false;
  }
  // safe reciprocal: a zero component becomes a huge finite slope (keeps the slab test conservative, no NaN)
  const float tiny = 1e-30f;
  float dx = fabsf_(d.x) < tiny ? (dm::f2u(d.x) >> 31 ? -tiny : tiny) : d.x;
  float dy = fabsf_(d.y) < tiny ? (dm::f2u(d.y) >> 31 ? -tiny : tiny) : d.y;
  float dz = fabsf_(d.z) < tiny ? (dm::f2u(d.z) >> 31 ? -tiny : tiny) : d.z;
  const float idx = fdiv_(1.0f, dx), idy = fdiv_(1.0f, dy), idz = fdiv_(1.0f, dz);
  const uint32_t oct_inv = (dx < 0 ? 0u : 1u) | (dy < 0 ? 0u : 2u) | (dz < 0 ? 0u : 4u);  // 7 - octant

  if (!(d.x == d.x && d.y == d.y && d.z == d.z)) {  // NaN direction (from == to): nothing can be hit
    if (out) *out = best;
    return false;
  }

  U2 stack[RB_STACK_MAX];
  int sp = 0;
  // root: one pending internal child at bit 31 with imask 0 -> node index 0
  U2 ngroup = U2{0u, 0x80000000u};

  while (true) {
    // ---- descend into the nearest pending child of the current node group -------------
    const uint32_t hits = ngroup.y;
    const int bit = bfind(hits);
    ngroup.y &= ~(1u << bit);
    if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
    const uint32_t slot = ((uint32_t)(bit - 24)) ^ oct_inv;
    const uint32_t node_index = ngroup.x + popc((hits & 0xFFu) & ~(0xFFFFFFFFu << slot));

    const F4* np = sc.node8 + 5 * (size_t)node_index;
    const F4 n0 = ldg4(np + 0), n1 = ldg4(np + 1), n2 = ldg4(np + 2), n3 = ldg4(np + 3), n4 = ldg4(np + 4);
    const uint32_t ebits = f2u(n0.w);
    const uint32_t imask = ebits >> 24;
    const float ax = u2f(byte_of(ebits, 0) << 23) * idx, ay = u2f(byte_of(ebits, 1) << 23) * idy,
                az = u2f(byte_of(ebits, 2) << 23) * idz;
    const float bx = (n0.x - o.x) * idx, by = (n0.y - o.y) * idy, bz = (n0.z - o.z) * idz;
    const float tcull = ANY ? tfar : best.t * 1.000001f;
    uint32_t hitmask = 0;
    const uint32_t metaw[2] = {f2u(n1.z), f2u(n1.w)};
    const uint32_t qlx[2] = {f2u(n2.x), f2u(n2.y)}, qly[2] = {f2u(n2.z), f2u(n2.w)}, qlz[2] = {f2u(n3.x), f2u(n3.y)};
    const uint32_t qhx[2] = {f2u(n3.z), f2u(n3.w)}, qhy[2] = {f2u(n4.x), f2u(n4.y)}, qhz[2] = {f2u(n4.z), f2u(n4.w)};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t meta = byte_of(metaw[i >> 2], i & 3);
      if (meta == 0) continue;
      const float lx = (float)byte_of(qlx[i >> 2], i & 3), hx = (float)byte_of(qhx[i >> 2], i & 3);
      const float ly = (float)byte_of(qly[i >> 2], i & 3), hy = (float)byte_of(qhy[i >> 2], i & 3);
      const float lz = (float)byte_of(qlz[i >> 2], i & 3), hz = (float)byte_of(qhz[i >> 2], i & 3);
      const float t0x = fmaf_(dx < 0 ? hx : lx, ax, bx), t1x = fmaf_(dx < 0 ? lx : hx, ax, bx);
      const float t0y = fmaf_(dy < 0 ? hy : ly, ay, by), t1y = fmaf_(dy < 0 ? ly : hy, ay, by);
      const float t0z = fmaf_(dz < 0 ? hz : lz, az, bz), t1z = fmaf_(dz < 0 ? lz : hz, az, bz);
      const float tmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tnear));
      const float tmax = fminf(fminf(t1x, t1y), fminf(t1z, tcull));
      if (tmin <= tmax) {
        const uint32_t internal = (imask >> i) & 1u;
        const uint32_t shift = (meta & 31u) ^ (internal ? oct_inv : 0u);
        hitmask |= (meta >> 5) << shift;
      }
    }
    ngroup.x = f2u(n1.x);
    ngroup.y = (hitmask & 0xFF000000u) | imask;
    const uint32_t tbase = f2u(n1.y);
    uint32_t tbits = hitmask & 0x00FFFFFFu;

    // ---- leaf triangles of this node ---------------------------------------------------
    while (tbits != 0) {
      const int ti = bfind(tbits);
      tbits &= ~(1u << ti);
      const F4* tp = sc.tri_isect + 3 * (size_t)(tbase + (uint32_t)ti);
      const F4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
      float t, u, v;
      if (tri_test(o, d, a, b, c, tnear, tfar, &t, &u, &v)) {
        const uint32_t id = f2u(c.y);
        if (ANY) {
          if (out) {
            out->t = t, out->u = u, out->v = v, out->tri = id;
          }
          return true;
        }
        if (best.tri == 0xFFFFFFFFu || t < best.t || (t == best.t && id < best.tri)) {
          best.t = t, best.u = u, best.v = v, best.tri = id;
        }
      }
    }

    if (ngroup.y <= 0x00FFFFFFu) {
      if (sp == 0) break;
      ngroup = stack[--sp];
    }
  }
  if (out) *out = best;
  return best.tri != 0xFFFFFFFFu;
}

// Intersection::testOcclusion, P/Intersection.h:43-60 (no normal offset; tnear = FLT_MIN + tnearOffset;
// tfar = dist - tfarOffset)
RB_HD bool test_occlusion(const SceneDev& sc, const V3& from, const V3& to, float tnearOffset, float tfarOffset) {
  const float dist = length(to - from);
  const V3 dir = normalize(to - from);
  return trace8<true>(sc, from, dir, FLT_MIN + tnearOffset, dist - tfarOffset, nullptr);
}

// Intersection::intersectEmbree + getGeometryAttributes, :8-41, 85-113, for untextured materials:
// interpolated normalised shading normal flipped to face the ray, hit point = org + dir * t.
struct SurfaceHit {
  bool didHit;
  V3 normal, hitPoint;
  float t;
  uint32_t tri, geomID, primID, material;
  int emissiveId;
};
RB_HD SurfaceHit intersect_surface(const SceneDev& sc, const V3& org, const V3& dir, float tnear, float tfar) {
  SurfaceHit h;
  h.didHit = false;
  h.normal = v3(0);
  h.hitPoint = v3(0);
  h.t = FLT_MAX;
  h.tri = h.geomID = h.primID = 0xFFFFFFFFu;
  h.material = 0;
  h.emissiveId = -1;
  HitRec r;
  if (!trace8<false>(sc, org, dir, tnear, tfar, &r)) return h;
  const F4* np = sc.tri_normals + 3 * (size_t)r.tri;
  const F4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
  const V3 n0 = xyz(a), n1 = v3(a.w, b.x, b.y), n2 = v3(b.z, b.w, c.x);
  const float w = 1.0f - r.u - r.v;
  V3 n = n0 * w + n1 * r.u + n2 * r.v;  // rtcInterpolate0 of attribute slot 0
  n = normalize(n);
  if (dot(-dir, n) <= 0.0f) n = n * -1.0f;
  const U4 info = sc.tri_info[r.tri];
  h.didHit = true;
  h.normal = n;
  h.hitPoint = org + dir * r.t;
  h.t = r.t;
  h.tri = r.tri;
  h.geomID = info.x;
  h.primID = info.y;
  h.material = info.z;
  h.emissiveId = (int)info.w;
  return h;
}

}  // namespace rb
#endif
