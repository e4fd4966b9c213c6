// rb_build.cuh — GPU BVH construction (replaces rtcCommitScene, P/Scene.cpp:15).
//
//   1. bounds_body   : triangle boxes (padded) + scene bounds (atomic min/max)
//   2. morton_body   : 63-bit Morton code of the box centre
//   3. radix sort of (code, triangle) — device: cub::DeviceRadixSort; emulation: std::stable_sort
//   4. karras_body   : binary radix tree over the sorted codes (Karras 2012), one thread per internal node
//   5. fit_body      : bottom-up box fit, one thread per leaf, second arrival at a node proceeds; the same sweep
//                      fills the collapse table C(node, i) = cheapest SAH cost of the subtree represented as at
//                      most i roots (wide nodes or leaves), i = 1..7 (the dynamic program of Ylitie, Karras &
//                      Laine 2017, section 3.1)
//   6. collapse_body : top-down, level-synchronous collapse of the binary tree into 8-wide nodes following the
//                      table's decisions (which subtrees become the <= 8 children of a wide node; subtrees with
//                      <= RB_LEAF_MAX triangles may become leaf children); choose child slots for octant
//                      ordering; quantise child boxes to 7 bits outward; emit leaf triangles in node order
// Every body is a __host__ __device__ function of a thread index so the emulation harness
// in tests/emu can run the identical code sequentially on the CPU.
#ifndef RB_BUILD_CUH_
#define RB_BUILD_CUH_

#include "rb_scene.cuh"

namespace rb {

RB_HD int atomic_add_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  int o = *p;
  *p += v;
  return o;
#endif
}
RB_HD void atomic_min_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}
RB_HD void atomic_max_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
  atomicMax(p, v);
#else
  if (v > *p) *p = v;
#endif
}
RB_HD void fence_() {
#if defined(__CUDA_ARCH__)
  __threadfence();
#endif
}
// monotone float <-> int map for atomic min/max
RB_HD int f2ord(float f) {
  int k = (int)f2u(f);
  return k < 0 ? (k ^ 0x7FFFFFFF) : k;
}
RB_HD float ord2f(int k) { return u2f((uint32_t)(k < 0 ? (k ^ 0x7FFFFFFF) : k)); }

RB_HD int clz32_(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
RB_HD int clz64_(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}

// Cost of a leaf triangle relative to a node visit in the collapse table. A triangle test is a third of a node visit in
// instructions, but the traversal runs its triangle phase at ~11 of 32 lanes against ~27 in the node phase, so what a
// triangle costs the warp is close to a node visit: measured on the bench frame 0.15 / 0.3 / 0.5 / 0.8 / 1.0 / 1.2 ->
// 146.2 / 148.1 / 150.0 / 150.6 / 150.5 / 150.5 fps (results do not depend on the tree).
#ifndef RB_COLLAPSE_C_PRIM
#define RB_COLLAPSE_C_PRIM 0.8f
#endif

struct BuildCtx {
  uint32_t n;            // triangles
  const float* tri_pos;  // [n][9] scene order (or the subset being built)
  const uint32_t* id_map;  // optional: triangle index in tri_pos -> scene triangle id written into the leaves
  float pad;             // box padding (absolute)
  int* scene_bounds;     // [6] ordered ints: lo.xyz (min), hi.xyz (max)
  // per triangle, scene order
  F4* tbox_lo;
  F4* tbox_hi;
  // sorted
  uint64_t* morton;
  uint32_t* order;  // sorted position -> scene triangle
  // binary radix tree: internal nodes [0, n-2]; child refs: >= 0 internal, < 0 leaf ~sortedIndex
  int* left;
  int* right;
  int* parent;       // [n-1] parent of internal (root: -1)
  int* leaf_parent;  // [n]
  int* range_lo;     // [n-1] first sorted leaf covered
  int* range_hi;     // [n-1] last sorted leaf covered
  F4* ibox_lo;       // [n-1]
  F4* ibox_hi;
  int* visit;  // [n-1] arrival counters
  // optimal collapse (dynamic program over the binary tree, filled bottom-up by fit_body)
  float* dp_cost;     // [7 * (n-1)]: C(node, i), i = 1..7 — SAH cost of the subtree as a forest of at most i roots
  uint32_t* dp_dec;   // [2 * (n-1)]: 8 decision bytes, byte j-1 for j = 1..8 (see dp_node)
  float c_prim;       // cost of a triangle test relative to a node visit
  // collapse
  F4* node8;       // [RB_NODE_F4 * max_nodes]
  F4* tri_isect;   // [3 * n]
  int* counters;   // [0] nodes allocated, [1] leaf triangles emitted, [2] next-queue length
  const int* q_in;  // pairs (binary node, out node8 index)
  int* q_out;
  int q_in_len;
};

RB_HD void bounds_body(const BuildCtx& c, uint32_t i) {
  const float* p = c.tri_pos + 9 * (size_t)i;
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = fminf(p[a], fminf(p[3 + a], p[6 + a])) - c.pad;
    hi[a] = fmaxf(p[a], fmaxf(p[3 + a], p[6 + a])) + c.pad;
  }
  c.tbox_lo[i] = F4{lo[0], lo[1], lo[2], 0};
  c.tbox_hi[i] = F4{hi[0], hi[1], hi[2], 0};
  for (int a = 0; a < 3; ++a) {
    atomic_min_i(c.scene_bounds + a, f2ord(lo[a]));
    atomic_max_i(c.scene_bounds + 3 + a, f2ord(hi[a]));
  }
}

RB_HD uint64_t spread21(uint32_t v) {  // 21 bits -> every third bit
  uint64_t x = v & 0x1FFFFFu;
  x = (x | x << 32) & 0x1F00000000FFFFull;
  x = (x | x << 16) & 0x1F0000FF0000FFull;
  x = (x | x << 8) & 0x100F00F00F00F00Full;
  x = (x | x << 4) & 0x10C30C30C30C30C3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
RB_HD void morton_body(const BuildCtx& c, uint32_t i) {
  const float blo[3] = {ord2f(c.scene_bounds[0]), ord2f(c.scene_bounds[1]), ord2f(c.scene_bounds[2])};
  const float bhi[3] = {ord2f(c.scene_bounds[3]), ord2f(c.scene_bounds[4]), ord2f(c.scene_bounds[5])};
  const F4 lo = c.tbox_lo[i], hi = c.tbox_hi[i];
  const float ctr[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
  uint32_t q[3];
  for (int a = 0; a < 3; ++a) {
    float ext = bhi[a] - blo[a];
    float u = ext > 0 ? (ctr[a] - blo[a]) / ext : 0.0f;
    u = fminf(fmaxf(u, 0.0f), 1.0f);
    float s = u * 2097151.0f;
    q[a] = (uint32_t)s;
  }
  c.morton[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
  c.order[i] = i;
}

RB_HD int karras_delta(const BuildCtx& c, int i, int j) {
  if (j < 0 || j >= (int)c.n) return -1;
  const uint64_t a = c.morton[i], b = c.morton[j];
  if (a == b) return 64 + clz32_((uint32_t)i ^ (uint32_t)j);
  return clz64_(a ^ b);
}
RB_HD void karras_body(const BuildCtx& c, uint32_t ii) {
  const int i = (int)ii;
  const int d = (karras_delta(c, i, i + 1) - karras_delta(c, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = karras_delta(c, i, i - d);
  int lmax = 2;
  while (karras_delta(c, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (karras_delta(c, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = karras_delta(c, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (karras_delta(c, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + (d < 0 ? d : 0);
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  int L, R;
  if (lo == gamma) {
    L = ~gamma;
    c.leaf_parent[gamma] = i;
  } else {
    L = gamma;
    c.parent[gamma] = i;
  }
  if (hi == gamma + 1) {
    R = ~(gamma + 1);
    c.leaf_parent[gamma + 1] = i;
  } else {
    R = gamma + 1;
    c.parent[gamma + 1] = i;
  }
  c.left[i] = L;
  c.right[i] = R;
  c.range_lo[i] = lo;
  c.range_hi[i] = hi;
  if (i == 0) c.parent[0] = -1;
}

RB_HD void child_box(const BuildCtx& c, int ref, F4* lo, F4* hi) {
  if (ref < 0) {
    const uint32_t t = c.order[~ref];
    *lo = c.tbox_lo[t];
    *hi = c.tbox_hi[t];
  } else {
    *lo = c.ibox_lo[ref];
    *hi = c.ibox_hi[ref];
  }
}
RB_HD float box_area(const F4& lo, const F4& hi) {
  const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}
// Collapse table of one child: a single triangle costs area * c_prim however many roots it may use.
RB_HD void dp_load(const BuildCtx& c, int ref, const F4& lo, const F4& hi, float* C /*[8], 1..7 used*/) {
  if (ref < 0) {
    const float v = box_area(lo, hi) * c.c_prim;
    for (int i = 1; i <= 7; ++i) C[i] = v;
    return;
  }
  for (int i = 1; i <= 7; ++i) {
#if defined(__CUDA_ARCH__)
    C[i] = __ldcg(c.dp_cost + 7 * (size_t)ref + (i - 1));  // written by another thread: bypass L1
#else
    C[i] = c.dp_cost[7 * (size_t)ref + (i - 1)];
#endif
  }
}
// C(n,1) = min(leaf, internal) with leaf = A*P*c_prim (P <= RB_LEAF_MAX) and internal = distribute(n,8) + A;
// C(n,i) = min(distribute(n,i), C(n,i-1)); distribute(n,j) = min_k C(left,k) + C(right,j-k).
// Decision byte j-1 (j = 2..8): low 3 bits = the k of distribute(n,j), bit 7 = "C(n,j) is C(n,j-1)".
// Decision byte 0: 1 = the subtree is a leaf when it is a single root.
RB_HD void dp_node(const BuildCtx& c, int node, const F4& alo, const F4& ahi, const F4& blo, const F4& bhi, const F4& nlo,
                   const F4& nhi) {
  float CL[8], CR[8];
  dp_load(c, c.left[node], alo, ahi, CL);
  dp_load(c, c.right[node], blo, bhi, CR);
  float dist[9];
  uint32_t dec[8];
  for (int j = 2; j <= 8; ++j) {
    float best = FLT_MAX;
    int bk = 1;
    for (int k = 1; k < j; ++k) {
      if (k > 7 || j - k > 7) continue;
      const float v = CL[k] + CR[j - k];
      if (v < best) best = v, bk = k;
    }
    dist[j] = best;
    dec[j - 1] = (uint32_t)bk;
  }
  const float A = box_area(nlo, nhi);
  const int P = c.range_hi[node] - c.range_lo[node] + 1;
  float C[8];
  const float internal = dist[8] + A;
  const float leaf = P <= RB_LEAF_MAX ? A * (float)P * c.c_prim : FLT_MAX;
  C[1] = leaf <= internal ? leaf : internal;
  dec[0] = leaf <= internal ? 1u : 0u;
  for (int i = 2; i <= 7; ++i) {
    if (dist[i] < C[i - 1]) {
      C[i] = dist[i];
    } else {
      C[i] = C[i - 1];
      dec[i - 1] |= 0x80u;
    }
  }
  for (int i = 1; i <= 7; ++i) c.dp_cost[7 * (size_t)node + (i - 1)] = C[i];
  c.dp_dec[2 * (size_t)node + 0] = dec[0] | (dec[1] << 8) | (dec[2] << 16) | (dec[3] << 24);
  c.dp_dec[2 * (size_t)node + 1] = dec[4] | (dec[5] << 8) | (dec[6] << 16) | (dec[7] << 24);
}
RB_HD uint32_t dp_decision(const BuildCtx& c, int node, int j) {  // decision byte of "j roots", j = 1..8
  return (c.dp_dec[2 * (size_t)node + ((j - 1) >> 2)] >> (8 * ((j - 1) & 3))) & 0xFFu;
}

RB_HD void fit_body(const BuildCtx& c, uint32_t leaf) {
  int node = c.leaf_parent[leaf];
  while (node >= 0) {
    fence_();
    if (atomic_add_i(c.visit + node, 1) == 0) return;  // first arrival waits for the sibling subtree
    fence_();
    F4 alo, ahi, blo, bhi;
#if defined(__CUDA_ARCH__)
    // children boxes were written by other threads: bypass L1
    const int L = c.left[node], R = c.right[node];
    if (L < 0) {
      const uint32_t t = c.order[~L];
      alo = c.tbox_lo[t], ahi = c.tbox_hi[t];
    } else {
      float4 x = __ldcg(reinterpret_cast<const float4*>(c.ibox_lo + L)), y = __ldcg(reinterpret_cast<const float4*>(c.ibox_hi + L));
      alo = F4{x.x, x.y, x.z, x.w}, ahi = F4{y.x, y.y, y.z, y.w};
    }
    if (R < 0) {
      const uint32_t t = c.order[~R];
      blo = c.tbox_lo[t], bhi = c.tbox_hi[t];
    } else {
      float4 x = __ldcg(reinterpret_cast<const float4*>(c.ibox_lo + R)), y = __ldcg(reinterpret_cast<const float4*>(c.ibox_hi + R));
      blo = F4{x.x, x.y, x.z, x.w}, bhi = F4{y.x, y.y, y.z, y.w};
    }
#else
    child_box(c, c.left[node], &alo, &ahi);
    child_box(c, c.right[node], &blo, &bhi);
#endif
    const F4 nlo = F4{fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0};
    const F4 nhi = F4{fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0};
    c.ibox_lo[node] = nlo;
    c.ibox_hi[node] = nhi;
    dp_node(c, node, alo, ahi, blo, bhi, nlo, nhi);
    node = c.parent[node];
  }
}

RB_HD int ref_count(const BuildCtx& c, int ref) { return ref < 0 ? 1 : (c.range_hi[ref] - c.range_lo[ref] + 1); }
RB_HD int ref_first(const BuildCtx& c, int ref) { return ref < 0 ? ~ref : c.range_lo[ref]; }
// Child boxes are stored as 7-bit grid coordinates q in [0,127]; the traversal decodes a plane as
// origin' + (128 + q) * step, where 128 + q is built directly as the float 0x43000000 | q << 16 (one byte
// permute, no integer-to-float conversion). origin' = node_lo - 128 * step.
#define RB_QMAX 127.0f
// biased exponent e of the grid step 2^(e-127) such that 126 steps cover the extent (one step of slack for the
// rounding of origin')
RB_HD uint32_t grid_exponent(float extent) {
  if (!(extent > 0.0f)) return 1u;
  const float step = extent / 126.0f;
  uint32_t bits = f2u(step);
  uint32_t e = (bits >> 23) & 0xFFu;
  if (bits & 0x7FFFFFu) e += 1;  // round the step up to a power of two
  if (e < 1u) e = 1u;
  if (e > 254u) e = 254u;
  while (e < 254u && u2f(e << 23) * 126.0f < extent) e += 1;
  return e;
}
RB_HD float float_prev(float x) {  // next representable float below x
  if (x == 0.0f) return -1.401298464e-45f;
  uint32_t b = f2u(x);
  return u2f(x > 0.0f ? b - 1u : b + 1u);
}
// origin' for one axis: the largest float with origin' + 128*step <= lo
RB_HD float grid_origin(float lo, float step) {
  float o = lo - 128.0f * step;
  while (o + 128.0f * step > lo) o = float_prev(o);
  return o;
}
RB_HD void write_tri(const BuildCtx& c, uint32_t dst, uint32_t tri) {
  const float* p = c.tri_pos + 9 * (size_t)tri;
  const float e1x = p[3] - p[0], e1y = p[4] - p[1], e1z = p[5] - p[2];
  const float e2x = p[6] - p[0], e2y = p[7] - p[1], e2z = p[8] - p[2];
  F4* o = c.tri_isect + 3 * (size_t)dst;
  o[0] = F4{p[0], p[1], p[2], e1x};
  o[1] = F4{e1y, e1z, e2x, e2y};
  o[2] = F4{e2z, u2f(c.id_map ? c.id_map[tri] : tri), 0, 0};
}

RB_HD uint32_t pack4(const uint32_t* b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); }

// Emit one Node8 at out_index from the children `refs` (binary-tree refs), whose union box is lo..hi.
RB_HD void emit_node8(const BuildCtx& c, int out_index, const int* refs, int n_items, const F4& nlo, const F4& nhi) {
  F4 clo[8], chi[8];
  bool internal[8];
  int n_int = 0, n_leaf_tris = 0;
  for (int k = 0; k < n_items; ++k) {
    child_box(c, refs[k], &clo[k], &chi[k]);
    internal[k] = refs[k] >= 0 && !(dp_decision(c, refs[k], 1) & 1u);
    if (internal[k])
      n_int++;
    else
      n_leaf_tris += ref_count(c, refs[k]);
  }
  // slot assignment: greedily give each child the free slot whose octant direction best matches
  // its offset from the node centre, so that (slot ^ ray octant) sorts children front to back
  const float cx = 0.5f * (nlo.x + nhi.x), cy = 0.5f * (nlo.y + nhi.y), cz = 0.5f * (nlo.z + nhi.z);
  int slot_of[8];
  int child_in_slot[8];
  for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
  for (int k = 0; k < n_items; ++k) slot_of[k] = -1;
  for (int round = 0; round < n_items; ++round) {
    float bestc = -FLT_MAX;
    int bk = -1, bs = -1;
    for (int k = 0; k < n_items; ++k) {
      if (slot_of[k] >= 0) continue;
      const float ox = 0.5f * (clo[k].x + chi[k].x) - cx, oy = 0.5f * (clo[k].y + chi[k].y) - cy,
                  oz = 0.5f * (clo[k].z + chi[k].z) - cz;
      for (int s = 0; s < 8; ++s) {
        if (child_in_slot[s] >= 0) continue;
        const float cost = ((s & 1) ? ox : -ox) + ((s & 2) ? oy : -oy) + ((s & 4) ? oz : -oz);
        if (cost > bestc) {
          bestc = cost;
          bk = k;
          bs = s;
        }
      }
    }
    slot_of[bk] = bs;
    child_in_slot[bs] = bk;
  }
  const int child_base = n_int ? atomic_add_i(c.counters + 0, n_int) : 0;
  const int tri_base = n_leaf_tris ? atomic_add_i(c.counters + 1, n_leaf_tris) : 0;
  const uint32_t ex = grid_exponent(nhi.x - nlo.x), ey = grid_exponent(nhi.y - nlo.y), ez = grid_exponent(nhi.z - nlo.z);
  const float sx = u2f(ex << 23), sy = u2f(ey << 23), sz = u2f(ez << 23);
  const float gox = grid_origin(nlo.x, sx), goy = grid_origin(nlo.y, sy), goz = grid_origin(nlo.z, sz);
  uint32_t imask = 0;
  uint32_t half[8], qlo[3][8], qhi[3][8];
  int tri_shift = 0;  // triangles in slots 0..3
  int int_cursor = 0, tri_cursor = 0;
  int q_base = 0;
  if (n_int) q_base = atomic_add_i(c.counters + 2, n_int);
  for (int s = 0; s < 8; ++s) {
    half[s] = 0;
    if (s == 4) tri_shift = tri_cursor;
    for (int a = 0; a < 3; ++a) qlo[a][s] = qhi[a][s] = 0;
    const int k = child_in_slot[s];
    if (k < 0) continue;
    const float l3[3] = {clo[k].x, clo[k].y, clo[k].z}, h3[3] = {chi[k].x, chi[k].y, chi[k].z};
    const float o3[3] = {gox, goy, goz}, s3[3] = {sx, sy, sz};
    for (int a = 0; a < 3; ++a) {
      float fl = floorf((l3[a] - o3[a]) / s3[a]) - 128.0f;
      float fh = ceilf((h3[a] - o3[a]) / s3[a]) - 128.0f;
      fl = fminf(fmaxf(fl, 0.0f), RB_QMAX);
      fh = fminf(fmaxf(fh, 0.0f), RB_QMAX);
      // outward rounding must survive the rounding of the float expressions themselves
      while (fl > 0.0f && o3[a] + (128.0f + fl) * s3[a] > l3[a]) fl -= 1.0f;
      while (fh < RB_QMAX && o3[a] + (128.0f + fh) * s3[a] < h3[a]) fh += 1.0f;
      qlo[a][s] = (uint32_t)fl;
      qhi[a][s] = (uint32_t)fh;
    }
    if (internal[k]) {
      imask |= 1u << s;
      half[s] = 1u << (12 + (s & 3));
      // children of this node occupy consecutive node8 indices in slot order
      c.q_out[2 * (q_base + int_cursor) + 0] = refs[k];
      c.q_out[2 * (q_base + int_cursor) + 1] = child_base + int_cursor;
      int_cursor++;
    } else {
      const int cnt = ref_count(c, refs[k]), first = ref_first(c, refs[k]);
      half[s] = ((1u << cnt) - 1u) << (tri_cursor - (s >= 4 ? tri_shift : 0));
      for (int t = 0; t < cnt; ++t) write_tri(c, (uint32_t)(tri_base + tri_cursor + t), c.order[first + t]);
      tri_cursor += cnt;
    }
  }
  F4* o = c.node8 + RB_NODE_F4 * (size_t)out_index;
  o[0] = F4{gox, goy, goz, u2f(ex | (ey << 8) | (ez << 16) | (imask << 24))};
  o[1] = F4{u2f((uint32_t)child_base), u2f((uint32_t)tri_base), u2f((uint32_t)tri_shift), u2f(0u)};
  o[2] = F4{u2f(pack4(qlo[0])), u2f(pack4(qlo[0] + 4)), u2f(pack4(qlo[1])), u2f(pack4(qlo[1] + 4))};
  o[3] = F4{u2f(pack4(qlo[2])), u2f(pack4(qlo[2] + 4)), u2f(pack4(qhi[0])), u2f(pack4(qhi[0] + 4))};
  o[4] = F4{u2f(pack4(qhi[1])), u2f(pack4(qhi[1] + 4)), u2f(pack4(qhi[2])), u2f(pack4(qhi[2] + 4))};
  o[5] = F4{u2f(half[0] | (half[4] << 16)), u2f(half[1] | (half[5] << 16)), u2f(half[2] | (half[6] << 16)),
            u2f(half[3] | (half[7] << 16))};
}

RB_HD void collapse_body(const BuildCtx& c, uint32_t w) {
  const int bnode = c.q_in[2 * w], out_index = c.q_in[2 * w + 1];
  // the children of this wide node: follow distribute(bnode, 8) down to single roots
  int refs[8];
  int n_items = 0;
  int st_ref[16], st_budget[16];
  int sp = 0;
  {
    const int k = (int)(dp_decision(c, bnode, 8) & 7u);
    st_ref[sp] = c.right[bnode], st_budget[sp++] = 8 - k;
    st_ref[sp] = c.left[bnode], st_budget[sp++] = k;
  }
  while (sp > 0) {
    const int ref = st_ref[--sp];
    int budget = st_budget[sp];
    if (ref < 0) {
      refs[n_items++] = ref;
      continue;
    }
    uint32_t d = 0;
    while (budget > 1 && ((d = dp_decision(c, ref, budget)) & 0x80u)) budget--;
    if (budget == 1) {
      refs[n_items++] = ref;
      continue;
    }
    const int k = (int)(d & 7u);
    st_ref[sp] = c.right[ref], st_budget[sp++] = budget - k;
    st_ref[sp] = c.left[ref], st_budget[sp++] = k;
  }
  emit_node8(c, out_index, refs, n_items, c.ibox_lo[bnode], c.ibox_hi[bnode]);
}

// Scenes with <= RB_LEAF_MAX triangles: one node, one leaf child holding them all (sorted order).
RB_HD void tiny_root_body(const BuildCtx& c) {
  F4 lo = c.tbox_lo[c.order[0]], hi = c.tbox_hi[c.order[0]];
  for (uint32_t i = 1; i < c.n; ++i) {
    const F4 a = c.tbox_lo[c.order[i]], b = c.tbox_hi[c.order[i]];
    lo = F4{fminf(lo.x, a.x), fminf(lo.y, a.y), fminf(lo.z, a.z), 0};
    hi = F4{fmaxf(hi.x, b.x), fmaxf(hi.y, b.y), fmaxf(hi.z, b.z), 0};
  }
  const uint32_t ex = grid_exponent(hi.x - lo.x), ey = grid_exponent(hi.y - lo.y), ez = grid_exponent(hi.z - lo.z);
  for (uint32_t i = 0; i < c.n; ++i) write_tri(c, i, c.order[i]);
  F4* o = c.node8;
  o[0] = F4{grid_origin(lo.x, u2f(ex << 23)), grid_origin(lo.y, u2f(ey << 23)), grid_origin(lo.z, u2f(ez << 23)),
            u2f(ex | (ey << 8) | (ez << 16))};
  o[1] = F4{u2f(0u), u2f(0u), u2f(c.n), u2f(0u)};
  o[2] = F4{u2f(0u), u2f(0u), u2f(0u), u2f(0u)};
  o[3] = F4{u2f(0u), u2f(0u), u2f(127u), u2f(0u)};
  o[4] = F4{u2f(127u), u2f(0u), u2f(127u), u2f(0u)};
  o[5] = F4{u2f((1u << c.n) - 1u), u2f(0u), u2f(0u), u2f(0u)};
  c.counters[0] = 1;
  c.counters[1] = (int)c.n;
}

}  // namespace rb
#endif
