// rb_scene.cuh — device scene records and the 8-wide quantised BVH traversal.
//
// Replaces what the reference keeps inside Embree (rtcIntersect1 / rtcOccluded1,
// P/Intersection.h:43-83) plus the attribute fetch of getGeometryAttributes
// (:85-113). B200 has no RT cores: this is ordinary SM code.
//
// Node8 (96 B = three 32-byte records, each fetched with ONE 256-bit load — LDG.E.ENL2.256, new on sm_100):
//   r0 = { origin'.xyz, bits: ex | ey<<8 | ez<<16 | imask<<24 | child_base, tri_base, tri_shift, 0 }
//   r1 = { qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7] | qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7] }
//   r2 = { qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7] | hit halves: word k = half(slot k) | half(slot k+4) << 16 }
// ex.. = biased exponents of the grid step. Child boxes are 7-bit grid coordinates q (one per byte), rounded
// outwards; a plane sits at origin' + (128 + q) * step and the float 128 + q is assembled with ONE byte permute
// (0x43000000 | q << 16) — no integer-to-float conversion on the quarter-rate pipe.
// The 16-bit hit half of a slot is what a box hit ORs into the node's hit accumulator (one LOP3 per child):
// 0 = empty slot; internal child in slot s: 1 << (12 + (s & 3)); leaf: (unary triangle count) << offset, the
// offset counted within the slot's group of four (0..9), so that
//   triangle bits = (acc & 0xFFF) | ((acc >> 16 & 0xFFF) << tri_shift),   tri_shift = triangles in slots 0..3
//   internal hits = (acc >> 12 & 0xF) | (acc >> 28) << 4                  (slot order)
// Children sit in slots chosen at build time so that (slot ^ ray octant) orders them front to back (compressed-
// wide-BVH scheme of Ylitie, Karras & Laine 2017); the internal hit byte is put into that order with three
// conditional bit swaps (perm8) — no table, no memory access.
// Why this shape: the traversal kernels are bound by the L1 data pipe (wavefronts = load instructions x distinct
// lines, profiles/), so a node visit is three loads, and by the half-rate ALU pipe, so the hit mask costs ~2 ops
// per child instead of 7.
#ifndef RB_SCENE_CUH_
#define RB_SCENE_CUH_

#include "rb_common.cuh"

namespace rb {

#ifndef RB_LEAF_MAX
#define RB_LEAF_MAX 3
#endif
#define RB_NODE_F4 6  // 16-byte records per node (96 B)
#define RB_STACK_MAX 48  // >= 2 * tree depth + 2: a node visit pushes at most a node group and a triangle group

struct I4 {
  int x, y, z, w;
};
// Texture's members, P/Texture.h:44-50
struct TexDev {
  const unsigned char* data;
  int width, height, scan_width, pixel_size;
  int clamp;  // TextureClamp: 0 = REPEAT (material maps, ModelLoader::TextureProxy), 1 = CLAMP_TO_EDGE (the sky)
};
// per-material constants of MaterialPhong::calc_I_M (rb_passes.cuh: make_mat_const)
struct MatConst {
  double lbeta;     // log B(shininess / 2, 1/2)
  float gq;         // gamma(shininess / 2 + 1/2) / gamma(shininess / 2 + 1)
  float shininess;  // the value both were computed for (bit compare: a textured shininess falls back to computing them)
};
struct SceneDev {
  // geometry
  const F4* node8;        // [RB_NODE_F4 * n_nodes]
  const F4* tri_isect;    // [3 * n_leaf_tris] leaf order: {v0.xyz,e1.x} {e1.y,e1.z,e2.x,e2.y} {e2.z,bits(tri id),0,0}
  const F4* tri_normals;  // [3 * n_tris] scene order: {n0.xyz,n1.x} {n1.y,n1.z,n2.x,n2.y} {n2.z,0,0,0}
  const U4* tri_info;     // [n_tris] scene order: {geomID, primID, material, emissive id (int, -1 none)}
  const F4* mat;          // [3 * n_mat]: {diffuse.rgb, shininess} {specular.rgb, bits(type)} {emission.rgb, ior}
  const MatConst* mat_const;  // [n_mat] or null
  // emissive triangles (TriangleCDF::tris order, P/ModelLoader.cpp:301-306)
  const F4* light;  // [6 * n_lights]: {p0,area} {p1,area/total} {p2,1/area} {n0,Le.r} {n1,Le.g} {n2,Le.b}
  const float* cdf;
  const float* alias_prob;
  const uint32_t* alias_idx;
  const U2* alias_pair;  // {bits(alias_prob[i]), alias_idx[i]}
  const F4* light_cull;  // [n_lights]: {bounding sphere centre, radius | +inf = not eligible} (rb_host_scene.h); may be null
  float maxabs;          // largest |coordinate| of any vertex of the scene
  uint32_t n_lights;
  uint32_t n_tris;
  uint32_t n_nodes;
  float total_area;
  uint32_t q7_base;  // 0x43000000 (see q7f)
  // second BVH over the emissive triangles only (leaf ids are scene triangle ids): BRDF-sampled rays only care
  // whether their closest hit is an emitter, so they are traced against this small tree first (rb_passes.cuh)
  const F4* em_node8;
  const F4* em_tri_isect;
  uint32_t em_n_nodes;
  // textured materials (rb_set_textures): per-triangle texture coordinates, texture table, per-material slots
  const F4* tri_uv;     // [2 * n_tris] scene order: {u0,v0,u1,v1} {u2,v2,0,0}; null = no surface carries uv
  const struct TexDev* tex;  // [n_tex]
  const I4* mat_tex;    // [n_mat]: {diffuse, specular, shininess, normal} texture index or -1; null = untextured scene
  // normal maps (Material::kNormalMapSlot): per-vertex tangents, attribute slot 3 of the reference's Embree geometry
  // (P/ModelLoader.cpp:286-287). Non-null ONLY while some material of the scene has a normal map, so that every other
  // scene pays one uniform pointer test per hit.
  const F4* tri_tan;    // [3 * n_tris] scene order, packed like tri_normals
  // sky (rb_set_sky): SphericalMap::texture, P/SphericalMap.h:16; data == nullptr = none
  TexDev sky;
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
  float inv = frcp_(det);
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
// x permuted so that bit (s ^ o) of the result is bit s of x (8-bit x, 3-bit o)
RB_HD uint32_t perm8(uint32_t x, uint32_t o) {
  if (o & 4u) x = ((x << 4) | (x >> 4)) & 0xFFu;
  if (o & 2u) x = ((x & 0x33u) << 2) | ((x >> 2) & 0x33u);
  if (o & 1u) x = ((x & 0x55u) << 1) | ((x >> 1) & 0x55u);
  return x;
}
struct alignas(32) F8 {
  F4 a, b;
};
RB_HD F8 ldg8(const F4* p) {  // 32-byte aligned
#if defined(__CUDA_ARCH__)
  return *reinterpret_cast<const F8*>(__builtin_assume_aligned(p, 32));
#else
  return F8{p[0], p[1]};
#endif
}
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
// 128 + (byte k of w) as a float, for 7-bit bytes. c43 is SceneDev::q7_base = 0x43000000, read from the
// kernel parameters: PRMT takes one immediate, and it should be the selector — with a literal constant the compiler
// keeps the constant as the immediate and re-materialises the selector into a register before every permute
// (~50 extra moves per node visit).
template <int K>
RB_HD float q7f(uint32_t w, uint32_t c43) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(__byte_perm(w, c43, 0x7044u | (K << 8)));
#else
  return u2f(c43 | (((w >> (8 * K)) & 0xFFu) << 16));
#endif
}

// Traversal state of one ray (registers + a per-thread stack).
struct Trav {
  V3 o, d;
  float dx, dy, dz;     // direction with zero components replaced by +-1e-30 (finite slopes)
  float idx, idy, idz;  // 1 / (dx,dy,dz)
  float tnear, tfar;
  uint32_t oct_inv;
  U2 ngroup;
  int sp;
  HitRec best;
  bool hit_any;
  uint32_t tie_id;  // TIE mode only: a hit at exactly tfar counts when its triangle id is smaller than this
#ifdef RB_TRAV_STATS
  uint32_t n_tri_tests;
#endif
};

// returns false when the ray cannot hit anything (empty scene, NaN direction)
RB_HD bool trav_init(Trav& T, const SceneDev& sc, const V3& o, const V3& d, float tnear, float tfar) {
  T.o = o;
  T.d = d;
  T.tnear = tnear;
  T.tfar = tfar;
  T.best.t = tfar;
  T.best.u = T.best.v = 0;
  T.best.tri = 0xFFFFFFFFu;
  T.hit_any = false;
  T.tie_id = 0u;
#ifdef RB_TRAV_STATS
  T.n_tri_tests = 0;
#endif
  T.sp = 0;
  T.ngroup = U2{0u, 0x80000000u};  // root: one pending internal child at bit 31 with imask 0 -> node index 0
  if (sc.n_nodes == 0) return false;
  if (!(d.x == d.x && d.y == d.y && d.z == d.z)) return false;  // NaN direction (from == to)
  // safe reciprocal: a zero component becomes a huge finite slope (keeps the slab test conservative, no NaN)
  const float tiny = 1e-30f;
  T.dx = fabsf_(d.x) < tiny ? (f2u(d.x) >> 31 ? -tiny : tiny) : d.x;
  T.dy = fabsf_(d.y) < tiny ? (f2u(d.y) >> 31 ? -tiny : tiny) : d.y;
  T.dz = fabsf_(d.z) < tiny ? (f2u(d.z) >> 31 ? -tiny : tiny) : d.z;
  T.idx = frcp_(T.dx), T.idy = frcp_(T.dy), T.idz = frcp_(T.dz);
  T.oct_inv = (T.dx < 0 ? 0u : 1u) | (T.dy < 0 ? 0u : 2u) | (T.dz < 0 ? 0u : 4u);  // 7 - octant
  return true;
}

// Box test of child I, result = d##I = tmax - tmin (a subtraction on the FMA pipe): the box is hit iff d is not negative.
// The traversal kernels are bound by the ALU pipe (compare / select / logic / permute / min-max, profiles/), so the
// hit masks are formed without compares: ONE byte permute in sign-replicate mode turns the sign bits of the two
// children that share a hit word into a 16 + 16-bit MISS mask, one LOP3 applies it — 8 ALU-pipe instructions per node
// where compare + select + mask + or took 29. (A difference is -0 only for x - x with x = -0... never: x - x = +0 in
// round-to-nearest; an invalid inf - inf gives the positive canonical NaN on the device and min() drops NaN operands,
// i.e. "hit": culling stays conservative, and hit SETS do not depend on culling.)
#define RB_CHILD_TEST(I, WX0, WX1, WY0, WY1, WZ0, WZ1)                                                \
  float d##I;                                                                                         \
  {                                                                                                   \
    const float t0x = fmaf_(q7f<(I)&3>(WX0, c43), ax, bx), t1x = fmaf_(q7f<(I)&3>(WX1, c43), ax, bx); \
    const float t0y = fmaf_(q7f<(I)&3>(WY0, c43), ay, by), t1y = fmaf_(q7f<(I)&3>(WY1, c43), ay, by); \
    const float t0z = fmaf_(q7f<(I)&3>(WZ0, c43), az, bz), t1z = fmaf_(q7f<(I)&3>(WZ1, c43), az, bz); \
    /* [lo, hi] must meet [tnear, tcull]: hi - lo, tcull - lo and hi - tnear all non-negative. Three subtractions  \
       (FMA pipe) and one three-input minimum instead of two more min / max on the ALU pipe. */        \
    const float lo = fmaxf(fmaxf(t0x, t0y), t0z), hi = fminf(fminf(t1x, t1y), t1z);                   \
    d##I = fminf(fminf(hi - lo, tcull - lo), hi - T.tnear);                                           \
  }
// hit-word bits of the slot pair (k, k + 4) that survive: low half unless d_lo < 0, high half unless d_hi < 0
RB_HD uint32_t hit_halves(uint32_t word, float d_lo, float d_hi) {
#if defined(__CUDA_ARCH__)
  uint32_t miss;  // bytes 0,1 = sign of d_lo replicated, bytes 2,3 = sign of d_hi replicated
  asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(miss) : "r"(__float_as_uint(d_lo)), "r"(__float_as_uint(d_hi)));
  return word & ~miss;
#else
  return word & ((d_lo >= 0.0f || d_lo != d_lo ? 0x0000FFFFu : 0u) | (d_hi >= 0.0f || d_hi != d_hi ? 0xFFFF0000u : 0u));
#endif
}

// ---- traversal primitives ---------------------------------------------------------------------------
// A node group is {child_base, hit bits << 24 | imask}; a triangle group is {tri_base, triangle hit bits}.
// The result of a traversal does not depend on the order in which groups are processed (any-hit: some hit
// exists; closest: minimum with an index tie-break), which is what lets the device kernel reschedule them.
RB_HD bool has_node_work(const Trav& T) { return T.ngroup.y > 0x00FFFFFFu; }

// Visit the nearest pending child of the current node group: test its eight child boxes. Leaves the hit
// internal children in T.ngroup (pushing the remainder of the old group), returns the hit leaf triangles.
// Stack: U2* (per-thread array in local memory) or SmemStack (k_trace_queue: entry i of this thread sits at
// base[i * stride] in shared memory — conflict-free, no local-memory round trip for the push / pop of every visit)
struct SmemStack {
  U2* base;
  int stride;
  RB_HD U2& operator[](int i) const { return base[i * stride]; }
};
template <bool ANY, class Stack>
RB_HD U2 trav_node_step(Trav& T, Stack stack, const SceneDev& sc) {
  const uint32_t hits = T.ngroup.y;
  const int bit = bfind(hits);
  T.ngroup.y &= ~(1u << bit);
  if (T.ngroup.y > 0x00FFFFFFu) stack[T.sp++] = T.ngroup;
#ifndef RB_ANY_ORDERED
  // any-hit rays: the visiting order cannot change the answer, and 85 % of the frame's shadow rays are unoccluded (every
  // hit child is visited whatever the order) — slot order saves the octant permutation of the hit byte
  const uint32_t slot = ANY ? (uint32_t)(bit - 24) : ((uint32_t)(bit - 24)) ^ T.oct_inv;
#else
  const uint32_t slot = ((uint32_t)(bit - 24)) ^ T.oct_inv;
#endif
  const uint32_t node_index = T.ngroup.x + popc((hits & 0xFFu) & ~(0xFFFFFFFFu << slot));

  const F4* np = sc.node8 + RB_NODE_F4 * (size_t)node_index;
  const F8 r0 = ldg8(np), r1 = ldg8(np + 2), r2 = ldg8(np + 4);
  const F4 &n0 = r0.a, &n1 = r0.b, &n2 = r1.a, &n3 = r1.b, &n4 = r2.a, &n5 = r2.b;
  const uint32_t ebits = f2u(n0.w);
  const uint32_t imask = ebits >> 24;
  const float ax = u2f(byte_of(ebits, 0) << 23) * T.idx, ay = u2f(byte_of(ebits, 1) << 23) * T.idy,
              az = u2f(byte_of(ebits, 2) << 23) * T.idz;
  const float bx = (n0.x - T.o.x) * T.idx, by = (n0.y - T.o.y) * T.idy, bz = (n0.z - T.o.z) * T.idz;
  const float tcull = ANY ? T.tfar : T.best.t * 1.000001f;
  // near / far plane words per axis, chosen once per node by the ray direction sign
  const bool nx = T.dx < 0, ny = T.dy < 0, nz = T.dz < 0;
  const uint32_t x0a = f2u(nx ? n3.z : n2.x), x0b = f2u(nx ? n3.w : n2.y), x1a = f2u(nx ? n2.x : n3.z), x1b = f2u(nx ? n2.y : n3.w);
  const uint32_t y0a = f2u(ny ? n4.x : n2.z), y0b = f2u(ny ? n4.y : n2.w), y1a = f2u(ny ? n2.z : n4.x), y1b = f2u(ny ? n2.w : n4.y);
  const uint32_t z0a = f2u(nz ? n4.z : n3.x), z0b = f2u(nz ? n4.w : n3.y), z1a = f2u(nz ? n3.x : n4.z), z1b = f2u(nz ? n3.y : n4.w);
  const uint32_t c43 = sc.q7_base;
  RB_CHILD_TEST(0, x0a, x1a, y0a, y1a, z0a, z1a)
  RB_CHILD_TEST(1, x0a, x1a, y0a, y1a, z0a, z1a)
  RB_CHILD_TEST(2, x0a, x1a, y0a, y1a, z0a, z1a)
  RB_CHILD_TEST(3, x0a, x1a, y0a, y1a, z0a, z1a)
  RB_CHILD_TEST(4, x0b, x1b, y0b, y1b, z0b, z1b)
  RB_CHILD_TEST(5, x0b, x1b, y0b, y1b, z0b, z1b)
  RB_CHILD_TEST(6, x0b, x1b, y0b, y1b, z0b, z1b)
  RB_CHILD_TEST(7, x0b, x1b, y0b, y1b, z0b, z1b)
  const uint32_t acc = hit_halves(f2u(n5.x), d0, d4) | hit_halves(f2u(n5.y), d1, d5) | hit_halves(f2u(n5.z), d2, d6) |
                       hit_halves(f2u(n5.w), d3, d7);
  const uint32_t tri_bits = (acc & 0xFFFu) | (((acc >> 16) & 0xFFFu) << (f2u(n1.z) & 31u));
#ifndef RB_ANY_ORDERED
  const uint32_t ihits_slot = ((acc >> 12) & 0xFu) | ((acc >> 28) << 4);
  const uint32_t ihits = ANY ? ihits_slot : perm8(ihits_slot, T.oct_inv);  // closest-hit: front to back
#else
  const uint32_t ihits = perm8(((acc >> 12) & 0xFu) | ((acc >> 28) << 4), T.oct_inv);  // front to back
#endif
  T.ngroup.x = f2u(n1.x);
  T.ngroup.y = (ihits << 24) | imask;
  return U2{f2u(n1.y), tri_bits};
}

// Test ONE triangle of a triangle group (clears its bit). Returns true when the ray is finished by it (ANY hit).
// TIE (any-hit only): the ray asks "does anything precede the hit (tfar, tie_id)?" under the closest-hit order of
// the parity contract (smaller t, then smaller triangle id) — a hit at exactly tfar counts iff its id is smaller.
template <bool ANY, bool TIE = false>
RB_HD bool trav_tri_one(Trav& T, U2& tgroup, const SceneDev& sc) {
  const int ti = bfind(tgroup.y);
  tgroup.y &= ~(1u << ti);
  const F4* tp = sc.tri_isect + 3 * (size_t)(tgroup.x + (uint32_t)ti);
  const F4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
  float t, u, v;
#ifdef RB_TRAV_STATS
  T.n_tri_tests++;
#endif
  const float tfar_test = TIE ? u2f(f2u(T.tfar) + 1u) : T.tfar;  // next float above tfar (tfar > 0, finite)
  if (tri_test(T.o, T.d, a, b, c, T.tnear, tfar_test, &t, &u, &v)) {
    const uint32_t id = f2u(c.y);
    if (TIE && !(t < T.tfar || id < T.tie_id)) return false;
    if (ANY) {
      T.best.t = t, T.best.u = u, T.best.v = v, T.best.tri = id;
      T.hit_any = true;
      return true;
    }
    if (T.best.tri == 0xFFFFFFFFu || t < T.best.t || (t == T.best.t && id < T.best.tri)) {
      T.best.t = t, T.best.u = u, T.best.v = v, T.best.tri = id;
    }
  }
  return false;
}

// Sequential step (one ray per thread, no rescheduling): node visit, then its triangles, then the next group.
// Returns false when the ray is finished.
template <bool ANY, bool TIE = false>
RB_HD bool trav_step(Trav& T, U2* stack, const SceneDev& sc) {
  U2 tgroup = trav_node_step<ANY, U2*>(T, stack, sc);
  while (tgroup.y != 0)
    if (trav_tri_one<ANY, TIE>(T, tgroup, sc)) return false;
  if (!has_node_work(T)) {
    if (T.sp == 0) return false;
    T.ngroup = stack[--T.sp];
  }
  return true;
}

// One ray to completion.
template <bool ANY>
RB_HD bool trace8(const SceneDev& sc, const V3& o, const V3& d, float tnear, float tfar, HitRec* out) {
  Trav T;
  U2 stack[RB_STACK_MAX];
  if (trav_init(T, sc, o, d, tnear, tfar)) {
    while (trav_step<ANY>(T, stack, sc)) {
    }
  }
  if (out) *out = T.best;
  return T.best.tri != 0xFFFFFFFFu;
}
// does any triangle precede the hit (tfar, tie_id) along the ray? (see trav_tri_one)
RB_HD bool trace8_precedes(const SceneDev& sc, const V3& o, const V3& d, float tnear, float tfar, uint32_t tie_id) {
  Trav T;
  U2 stack[RB_STACK_MAX];
  if (trav_init(T, sc, o, d, tnear, tfar)) {
    T.tie_id = tie_id;
    while (trav_step<true, true>(T, stack, sc)) {
    }
  }
  return T.best.tri != 0xFFFFFFFFu;
}
// the same scene seen through its emissive-only BVH
RB_HD SceneDev emissive_view(const SceneDev& sc) {
  SceneDev e = sc;
  e.node8 = sc.em_node8;
  e.tri_isect = sc.em_tri_isect;
  e.n_nodes = sc.em_n_nodes;
  return e;
}

// Shadow ray of Intersection::testOcclusion, P/Intersection.h:43-60 (no normal offset; tnear = FLT_MIN + tnearOffset;
// tfar = dist - tfarOffset)
RB_HD void shadow_ray(const V3& from, const V3& to, float tfarOffset, V3* dir, float* tfar) {
  *tfar = length(to - from) - tfarOffset;
  *dir = normalize(to - from);
}
RB_HD bool test_occlusion(const SceneDev& sc, const V3& from, const V3& to, float tnearOffset, float tfarOffset) {
  V3 dir;
  float tfar;
  shadow_ray(from, to, tfarOffset, &dir, &tfar);
  return trace8<true>(sc, from, dir, FLT_MIN + tnearOffset, tfar, nullptr);
}

// Intersection::intersectEmbree + getGeometryAttributes, :8-41, 85-113: interpolated normalised shading normal
// flipped to face the ray, hit point = org + dir * t, texture coordinates, normal map (TBN) where the material has one.
// Texture::get_texel(x, y) (P/Texture.cpp:72-107): REPEAT abs(x % w) / CLAMP_TO_EDGE clamp(x, 0, w - 1); 8-bit B,G,R / 255
// or float R,G,B
RB_HD V3 tex_texel(const TexDev& T, int x, int y) {
  int cx = x % T.width, cy = y % T.height;
  cx = cx < 0 ? -cx : cx;
  cy = cy < 0 ? -cy : cy;
  if (T.clamp) {
    cx = x < 0 ? 0 : (x > T.width - 1 ? T.width - 1 : x);
    cy = y < 0 ? 0 : (y > T.height - 1 ? T.height - 1 : y);
  }
  const unsigned char* p = T.data + (size_t)cy * T.scan_width + (size_t)cx * T.pixel_size;
  if (T.pixel_size > 4) {
    const float* f = reinterpret_cast<const float*>(p);
    return v3(f[0], f[1], f[2]);
  }
  const float b = (float)p[0] / 255.0f, g = (float)p[1] / 255.0f, r = (float)p[2] / 255.0f;
  return v3(r, g, b);
}
// Texture::getTexelBilinear (:170-194): pixel = (u * w, (1 - v) * h); glm::mix(a, b, t) = a * (1 - t) + b * t
RB_HD V3 tex_sample(const TexDev& T, float u, float v) {
  const float px = u * (float)T.width, py = (1.0f - v) * (float)T.height;
  const float fx = floorf(px), fy = floorf(py);
  const float tx = px - fx, ty = py - fy;
  const V3 x0y0 = tex_texel(T, (int)fx, (int)fy);
  const V3 x1y0 = tex_texel(T, (int)(fx + 1.0f), (int)fy);
  const V3 x0y1 = tex_texel(T, (int)fx, (int)(fy + 1.0f));
  const V3 x1y1 = tex_texel(T, (int)(fx + 1.0f), (int)(fy + 1.0f));
  const V3 x1 = x0y0 * (1.0f - tx) + x1y0 * tx;
  const V3 x2 = x0y1 * (1.0f - tx) + x1y1 * tx;
  return x1 * (1.0f - ty) + x2 * ty;
}

// SphericalMap::getTexel (P/SphericalMap.cpp:10-14): x = 0.5f + 0.5f * atan2f(d.y, d.x) * INVPI, y = 1.0f - acos(d.z) * INVPI
// with the DOUBLE constant INVPI = 1.0 / M_PI — float products promoted, sums formed in double, rounded once on the store
// to float; atan2f / acosf are det_math's. Out of line (one call per primary-ray miss); everything by VALUE: a
// reference into the kernel's parameter block would make the compiler copy the whole FrameCtx to local memory.
RB_HD_NOINLINE V3 sky_texel(TexDev sky, V3 dir) {
  const double INVPI = 1.0 / 3.14159265358979323846;
  const float x = (float)(0.5f + (double)(0.5f * dm::atan2f_(dir.y, dir.x)) * INVPI);
  const float y = (float)(1.0f - (double)dm::acosf_(dir.z) * INVPI);
  return tex_sample(sky, x, y);
}

// Normal map, Intersection::intersectEmbree :25-39 (glm operation order: dot = (x+y)+z, normalize = v * inversesqrt(dot),
// cross = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y), mat3 * vec3 = (m0*x + m1*y) + m2*z):
//   T = normalize(tangent - dot(tangent, n) * n);  B = normalize(cross(n, T));  n' = mat3(T, B, n) * (texel * 2 - 1)
// `n` is the interpolated normal AFTER the flip towards the ray; the tangent is the raw interpolation of slot 3; the
// result is neither re-normalised nor flipped again. Out of line: the callers' register budget is that of the scenes
// without normal maps; arguments by value / plain device pointers (see sky_texel).
RB_HD_NOINLINE V3 apply_normal_map(const F4* tp, const TexDev* map, float w, float u, float v, float tex_u, float tex_v, V3 n) {
  const F4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
  const V3 t0 = xyz(a), t1 = v3(a.w, b.x, b.y), t2 = v3(b.z, b.w, c.x);
  const V3 tangent = t0 * w + t1 * u + t2 * v;  // rtcInterpolate0 of attribute slot 3 (P/Intersection.h:106-107)
  V3 T = tangent - n * dot(tangent, n);
  T = normalize(T);
  const V3 B = normalize(cross(n, T));
  const V3 N = tex_sample(*map, tex_u, tex_v) * 2.0f - v3(1.0f);
  return v3(T.x * N.x + B.x * N.y + n.x * N.z, T.y * N.x + B.y * N.y + n.y * N.z, T.z * N.x + B.z * N.y + n.z * N.z);
}

struct SurfaceHit {
  bool didHit;
  V3 normal, hitPoint;
  float tex_u, tex_v;  // rtcInterpolate0 of attribute slot 1 (P/Intersection.h:99-100); 0 when the scene has no uv
  float t;
  uint32_t tri, geomID, primID, material;
  int emissiveId;
};
// NMAP = false: an instantiation without the normal-map branch, for a kernel whose register budget has no room for it
// (k_initial_resolve); the host launches it only while sc.tri_tan == nullptr, where the two are the same function.
template <bool NMAP = true>
RB_HD SurfaceHit surface_from_hit(const SceneDev& sc, const V3& org, const V3& dir, const HitRec& r) {
  SurfaceHit h;
  h.didHit = false;
  h.normal = v3(0);
  h.hitPoint = v3(0);
  h.t = FLT_MAX;
  h.tri = h.geomID = h.primID = 0xFFFFFFFFu;
  h.material = 0;
  h.emissiveId = -1;
  h.tex_u = h.tex_v = 0.0f;
  if (r.tri == 0xFFFFFFFFu) return h;
  const F4* np = sc.tri_normals + 3 * (size_t)r.tri;
  const F4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
  const V3 n0 = xyz(a), n1 = v3(a.w, b.x, b.y), n2 = v3(b.z, b.w, c.x);
  const float w = 1.0f - r.u - r.v;
  V3 n = n0 * w + n1 * r.u + n2 * r.v;  // rtcInterpolate0 of attribute slot 0
  n = normalize(n);
  if (dot(-dir, n) <= 0.0f) n = n * -1.0f;
  if (sc.tri_uv != nullptr) {  // same interpolation contract as the normal: w * a0 + u * a1 + v * a2
    const F4 q0 = ldg4(sc.tri_uv + 2 * (size_t)r.tri), q1 = ldg4(sc.tri_uv + 2 * (size_t)r.tri + 1);
    h.tex_u = q0.x * w + q0.z * r.u + q1.x * r.v;
    h.tex_v = q0.y * w + q0.w * r.u + q1.y * r.v;
  }
  const U4 info = sc.tri_info[r.tri];
  if (NMAP && sc.tri_tan != nullptr) {  // some material of the scene has a normal map
    const int slot = sc.mat_tex[info.z].w;
    if (slot >= 0) n = apply_normal_map(sc.tri_tan + 3 * (size_t)r.tri, sc.tex + slot, w, r.u, r.v, h.tex_u, h.tex_v, n);
  }
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
RB_HD SurfaceHit intersect_surface(const SceneDev& sc, const V3& org, const V3& dir, float tnear, float tfar) {
  HitRec r;
  trace8<false>(sc, org, dir, tnear, tfar, &r);
  return surface_from_hit(sc, org, dir, r);
}

}  // namespace rb
#endif
