// emu.cpp — TEST-ONLY host emulation of the CUDA kernel bodies.
//
// The product's per-thread kernel bodies (rb_build.cuh, rb_scene.cuh, rb_passes.cuh) are
// __host__ __device__ functions; this harness compiles the SAME source with g++ and runs
// each "kernel" as a plain loop over thread indices, so the CPU-only test tier
// (pytest -m "not gpu") can check the kernel logic — BVH build, traversal, all passes —
// bit for bit against the oracle before any GPU time is spent.
//
// It is NOT part of the product: nothing under restir_embree_b200/ loads it, bench.py never
// touches it, and the shipped library (librestir_b200.so) has no CPU path at all.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <functional>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../restir_embree_b200/csrc/rb_build.cuh"
#include "../../restir_embree_b200/csrc/rb_host_scene.h"
#include "../../restir_embree_b200/csrc/rb_passes.cuh"

using namespace rb;

struct Emu {
  int width = 0, height = 0, y0 = 0, y1 = 0;
  uint32_t seed = 123;
  RbParams P{};
  HostScene hs;
  std::vector<F4> node8, tri_isect, em_node8, em_tri_isect;
  uint32_t em_n_nodes = 0;
  uint32_t n_nodes = 0, depth = 0;
  SceneDev sc{};
  bool haveScene = false, havePrev = false;
  // frame state
  struct GStore {
    std::vector<F4> a, b, c, d, e;
    std::vector<U2> ids;
  } gs[2];
  struct RStore {
    std::vector<F4> a, b, c;
    std::vector<int> li;
  } rs[3];
  int gCur = 0, rRead = 0, rWrite = 1, rLast = 2;
  std::vector<float> frame, accumulator;
  std::vector<F4> display;
  unsigned long long counters[8] = {0};
  CamState prevCam{};
  std::string err;
  // in-flight frame (phases) and band bookkeeping
  FrameCtx fc{};
  RbParams Pf{};
  bool open = false, wave = false, wave_spatial = false, wave_spatial_staged = false;
  uint32_t frame_idx = 0, qcount = 0;
  int prevGy0 = 0, prevGy1 = 0;
  std::vector<RayQ> rays;
  std::vector<uint8_t> occ;
  std::vector<HitRec> hits;
  std::vector<F4> brdf_dir;
  std::vector<MatConst> mat_const;
  std::vector<U4> cand;
  bool shaded = false;
  std::vector<TexDev> tex_tab;  // rb_set_textures mirror
  std::vector<std::vector<unsigned char>> tex_data;
  std::vector<I4> tex_slots;
  std::vector<unsigned char> sky_data;  // rb_set_sky mirror
  std::vector<RayQ> chain;
  uint32_t chain_count[2] = {0, 0};
  std::vector<uint32_t> deferred;
  uint32_t deferred_count = 0;
  unsigned long long n_deferred_total = 0;  // over all frames: lets a test see that the deferral path ran

  GBufPlanes gp(int i) { return GBufPlanes{gs[i].a.data(), gs[i].b.data(), gs[i].c.data(), gs[i].d.data(), gs[i].e.data(), gs[i].ids.data()}; }
  ResPlanes rp(int i) { return ResPlanes{rs[i].a.data(), rs[i].b.data(), rs[i].c.data(), rs[i].li.data()}; }
};

static int emu_build_bvh_impl(const float* tri_pos, const uint32_t* id_map, uint32_t n, float maxabs, std::vector<F4>& node8_out,
                              std::vector<F4>& tri_out, uint32_t* n_nodes_out, uint32_t* depth_out) {
  BuildCtx c{};
  c.n = n;
  c.tri_pos = tri_pos;
  c.id_map = id_map;
  c.pad = maxabs * (1.0f / 262144.0f) + 1e-30f;
  int bounds[6] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, (int)0x80000000, (int)0x80000000, (int)0x80000000};
  c.scene_bounds = bounds;
  std::vector<F4> tlo(n), thi(n), ilo(n), ihi(n);
  std::vector<uint64_t> morton(n), morton_sorted(n);
  std::vector<uint32_t> order(n), order_sorted(n);
  std::vector<int> left(n), right(n), parent(n), leaf_parent(n), rlo(n), rhi(n), visit(n, 0);
  c.tbox_lo = tlo.data(), c.tbox_hi = thi.data();
  c.morton = morton.data(), c.order = order.data();
  c.left = left.data(), c.right = right.data(), c.parent = parent.data(), c.leaf_parent = leaf_parent.data();
  c.range_lo = rlo.data(), c.range_hi = rhi.data(), c.ibox_lo = ilo.data(), c.ibox_hi = ihi.data(), c.visit = visit.data();
  std::vector<float> dp_cost(7 * (size_t)n);
  std::vector<uint32_t> dp_dec(2 * (size_t)n);
  c.dp_cost = dp_cost.data(), c.dp_dec = dp_dec.data();
  c.c_prim = getenv("EMU_CPRIM") ? (float)atof(getenv("EMU_CPRIM")) : RB_COLLAPSE_C_PRIM;
  int counters[4] = {1, 0, 0, 0};
  c.counters = counters;
  const size_t max_nodes = (size_t)n / 2 + 8;
  std::vector<F4> node8(RB_NODE_F4 * max_nodes);
  tri_out.assign(3 * (size_t)n, F4{0, 0, 0, 0});
  c.node8 = node8.data();
  c.tri_isect = tri_out.data();
  for (uint32_t i = 0; i < n; ++i) bounds_body(c, i);
  for (uint32_t i = 0; i < n; ++i) morton_body(c, i);
  // radix sort stand-in: stable sort by key (cub::DeviceRadixSort is stable too)
  std::vector<uint32_t> perm(n);
  std::iota(perm.begin(), perm.end(), 0u);
  std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return morton[a] < morton[b]; });
  for (uint32_t i = 0; i < n; ++i) morton_sorted[i] = morton[perm[i]], order_sorted[i] = order[perm[i]];
  c.morton = morton_sorted.data();
  c.order = order_sorted.data();
  uint32_t depth = 0, n_nodes = 0;
  if (n <= RB_LEAF_MAX) {
    tiny_root_body(c);
    n_nodes = 1, depth = 1;
  } else {
    if (getenv("EMU_SAH")) {
      // EXPERIMENT (host only): binned-SAH binary tree instead of the Morton radix tree, to measure how much
      // traversal work a better topology would save. Not part of the product build.
      int next = 0;
      struct Job {
        int lo, hi, parent, is_right;
      };
      std::function<int(int, int, int)> build = [&](int lo, int hi, int par) -> int {
        const int id = next++;
        parent[id] = par;
        rlo[id] = lo, rhi[id] = hi;
        float cmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        auto cen = [&](int k, int a) {
          const uint32_t t = order_sorted[k];
          const float* L = &tlo[t].x;
          const float* H = &thi[t].x;
          return 0.5f * (L[a] + H[a]);
        };
        for (int k = lo; k <= hi; ++k)
          for (int a = 0; a < 3; ++a) cmin[a] = std::min(cmin[a], cen(k, a)), cmax[a] = std::max(cmax[a], cen(k, a));
        int mid = (lo + hi + 1) / 2;
        int bestAxis = -1;
        float bestCost = FLT_MAX;
        int bestBin = 0;
        const int NB = 16;
        for (int a = 0; a < 3; ++a) {
          if (!(cmax[a] > cmin[a])) continue;
          F4 blo[NB], bhi[NB];
          int cnt[NB];
          for (int b = 0; b < NB; ++b) blo[b] = F4{FLT_MAX, FLT_MAX, FLT_MAX, 0}, bhi[b] = F4{-FLT_MAX, -FLT_MAX, -FLT_MAX, 0}, cnt[b] = 0;
          const float sc = NB / (cmax[a] - cmin[a]);
          for (int k = lo; k <= hi; ++k) {
            int b = std::min(NB - 1, (int)((cen(k, a) - cmin[a]) * sc));
            const uint32_t t = order_sorted[k];
            blo[b] = F4{std::min(blo[b].x, tlo[t].x), std::min(blo[b].y, tlo[t].y), std::min(blo[b].z, tlo[t].z), 0};
            bhi[b] = F4{std::max(bhi[b].x, thi[t].x), std::max(bhi[b].y, thi[t].y), std::max(bhi[b].z, thi[t].z), 0};
            cnt[b]++;
          }
          float la[NB], ra[NB];
          int lc[NB], rc[NB];
          F4 l = blo[0], h = bhi[0];
          int cc = 0;
          for (int b = 0; b < NB; ++b) {
            if (cnt[b]) l = F4{std::min(l.x, blo[b].x), std::min(l.y, blo[b].y), std::min(l.z, blo[b].z), 0}, h = F4{std::max(h.x, bhi[b].x), std::max(h.y, bhi[b].y), std::max(h.z, bhi[b].z), 0};
            cc += cnt[b];
            la[b] = cc ? box_area(l, h) : 0, lc[b] = cc;
          }
          l = F4{FLT_MAX, FLT_MAX, FLT_MAX, 0}, h = F4{-FLT_MAX, -FLT_MAX, -FLT_MAX, 0};
          cc = 0;
          for (int b = NB - 1; b >= 0; --b) {
            if (cnt[b]) l = F4{std::min(l.x, blo[b].x), std::min(l.y, blo[b].y), std::min(l.z, blo[b].z), 0}, h = F4{std::max(h.x, bhi[b].x), std::max(h.y, bhi[b].y), std::max(h.z, bhi[b].z), 0};
            cc += cnt[b];
            ra[b] = cc ? box_area(l, h) : 0, rc[b] = cc;
          }
          for (int b = 0; b + 1 < NB; ++b) {
            if (lc[b] == 0 || rc[b + 1] == 0) continue;
            const float cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
            if (cost < bestCost) bestCost = cost, bestAxis = a, bestBin = b;
          }
        }
        if (bestAxis >= 0) {
          const int a = bestAxis;
          const float sc = NB / (cmax[a] - cmin[a]);
          auto it = std::partition(order_sorted.begin() + lo, order_sorted.begin() + hi + 1, [&](uint32_t t) {
            const float* L = &tlo[t].x;
            const float* H = &thi[t].x;
            int b = std::min(NB - 1, (int)((0.5f * (L[a] + H[a]) - cmin[a]) * sc));
            return b <= bestBin;
          });
          mid = (int)(it - order_sorted.begin());
          if (mid <= lo || mid > hi) mid = (lo + hi + 1) / 2;
        }
        // children: [lo, mid-1], [mid, hi]
        if (mid - 1 == lo) {
          left[id] = ~lo;
          leaf_parent[lo] = id;
        } else
          left[id] = build(lo, mid - 1, id);
        if (mid == hi) {
          right[id] = ~hi;
          leaf_parent[hi] = id;
        } else
          right[id] = build(mid, hi, id);
        return id;
      };
      build(0, (int)n - 1, -1);
    } else
    // run the per-thread bodies in a scrambled order to mimic unordered GPU scheduling
    for (uint32_t i = 0; i + 1 < n; ++i) karras_body(c, n - 2 - i);
    for (uint32_t i = 0; i < n; ++i) fit_body(c, n - 1 - i);
    std::vector<int> q[2];
    q[0].assign(2 * (size_t)n + 2, 0);
    q[1].assign(2 * (size_t)n + 2, 0);
    int len = 1, cur = 0;
    while (len > 0) {
      c.q_in = q[cur].data();
      c.q_out = q[cur ^ 1].data();
      c.q_in_len = len;
      counters[2] = 0;
      for (int w = len - 1; w >= 0; --w) collapse_body(c, (uint32_t)w);
      if ((size_t)counters[0] > max_nodes) return -2;
      len = counters[2];
      n_nodes = (uint32_t)counters[0];
      cur ^= 1;
      depth++;
      if (depth > 4096) return -2;
    }
    if ((uint32_t)counters[1] != n) return -2;
  }
  if (2 * depth + 2 > RB_STACK_MAX) return -4;
  node8.resize(RB_NODE_F4 * (size_t)n_nodes);
  node8_out = node8;
  *n_nodes_out = n_nodes;
  *depth_out = depth;
  return 0;
}
static int emu_build_bvh(Emu* E) {
  int rc = emu_build_bvh_impl(E->hs.pos.data(), nullptr, (uint32_t)E->hs.n, E->hs.maxabs, E->node8, E->tri_isect, &E->n_nodes, &E->depth);
  if (rc != 0) return rc;
  // emissive-only BVH (two-step BRDF-candidate rays), as rb_upload_scene builds it
  E->em_n_nodes = 0;
  const size_t NL = E->hs.emissive.size();
  if (NL > 0) {
    std::vector<float> epos(9 * NL);
    for (size_t i = 0; i < NL; ++i) memcpy(&epos[9 * i], &E->hs.pos[9 * (size_t)E->hs.emissive[i]], 36);
    uint32_t d = 0;
    rc = emu_build_bvh_impl(epos.data(), E->hs.emissive.data(), (uint32_t)NL, E->hs.maxabs, E->em_node8, E->em_tri_isect, &E->em_n_nodes, &d);
  }
  return rc;
}

extern "C" {

void* emu_create(int width, int height, uint32_t seed, int y0, int y1) {
  Emu* E = new Emu();
  E->width = width, E->height = height, E->seed = seed, E->y0 = y0, E->y1 = y1;
  const size_t n = (size_t)width * height;
  for (auto& g : E->gs) {
    g.a.assign(n, F4{0, 0, 0, 0}), g.b = g.a, g.c = g.a, g.d = g.a, g.e = g.a;
    g.ids.assign(n, U2{0xFFFFFFFFu, 0xFFFFFFFFu});
  }
  for (auto& r : E->rs) {
    r.a.assign(n, F4{0, 0, 0, 0}), r.b = r.a, r.c = r.a;
    r.li.assign(n, -1);
  }
  E->frame.assign(n * 3, 0.0f);
  return E;
}
void emu_destroy(void* h) { delete (Emu*)h; }
const char* emu_last_error(void* h) { return ((Emu*)h)->err.c_str(); }

int emu_upload_scene(void* h, const RbSceneDesc* sd) {
  Emu* E = (Emu*)h;
  int rc = flatten_scene(sd, E->hs, E->err);
  if (rc != RB_OK) return rc;
  E->n_nodes = 0;
  if (E->hs.n > 0) {
    rc = emu_build_bvh(E);
    if (rc != 0) {
      E->err = "emu_build_bvh failed";
      return rc;
    }
  }
  SceneDev& sc = E->sc;
  sc.node8 = E->node8.data();
  sc.tri_isect = E->tri_isect.data();
  sc.tri_normals = E->hs.nrm.data();
  sc.tri_info = E->hs.info.data();
  sc.mat = E->hs.mat.data();
  E->mat_const.resize(E->hs.mat.size() / 3);
  for (size_t i = 0; i < E->mat_const.size(); ++i) E->mat_const[i] = make_mat_const(E->hs.mat[3 * i].w);
  sc.mat_const = E->mat_const.data();
  sc.light = E->hs.light.data();
  sc.cdf = E->hs.cdf.data();
  sc.alias_prob = E->hs.alias_prob.data();
  sc.alias_idx = E->hs.alias_idx.data();
  sc.alias_pair = reinterpret_cast<const U2*>(E->hs.alias_pair.data());
  sc.light_cull = E->hs.light_cull.data();
  sc.maxabs = E->hs.maxabs;
  sc.n_lights = (uint32_t)E->hs.emissive.size();
  sc.n_tris = (uint32_t)E->hs.n;
  sc.n_nodes = E->n_nodes;
  sc.total_area = E->hs.totalSurface;
  sc.q7_base = 0x43000000u;
  sc.em_node8 = E->em_node8.data();
  sc.em_tri_isect = E->em_tri_isect.data();
  sc.em_n_nodes = E->em_n_nodes;
  sc.tri_uv = E->hs.uv.empty() ? nullptr : E->hs.uv.data();
  sc.tex = nullptr;
  sc.mat_tex = nullptr;
  sc.tri_tan = nullptr;
  E->tex_tab.clear(), E->tex_data.clear(), E->tex_slots.clear();
  E->haveScene = true;
  E->havePrev = false;
  return 0;
}
int emu_set_params(void* h, const RbParams* p) {
  Emu* E = (Emu*)h;
  if (p->useSkybox && E->sc.sky.data == nullptr) return RB_ERR_INVALID_ARGUMENT;
  if (p->spatialReuseNeighborCount > RB_MAX_NEIGHBORS) return RB_ERR_UNSUPPORTED;
  E->P = *p;
  return 0;
}
void emu_scene_stats(void* h, uint32_t* out4) {
  Emu* E = (Emu*)h;
  out4[0] = (uint32_t)E->hs.n, out4[1] = (uint32_t)E->hs.emissive.size(), out4[2] = E->n_nodes, out4[3] = E->depth;
}

}  // extern "C"

template <class F>
static void for_pixels(Emu* E, FrameCtx& fc, F&& f) {
  unsigned long long c0 = 0, c1 = 0, c2 = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : c0, c1, c2)
  for (int y = fc.y0; y < fc.y1; ++y)
    for (int x = 0; x < fc.width; ++x) {
      Cnt cnt = {0, 0, 0};
      f(x, y, cnt);
      c0 += cnt.closest, c1 += cnt.anyW, c2 += cnt.anyT;
    }
  E->counters[0] += c0, E->counters[1] += c1, E->counters[2] += c2;
}

extern "C" {

}  // extern "C"

// mirrors frame_begin / frame_spatial / frame_end of restir_b200.cu (the host-side pass schedule)
static int spatial_reach(const RbParams& P) { return (int)sqrtf(std::max(P.spatialReuseRadius, 0.0f)) + 1; }
static uint32_t PX(const FrameCtx& fc, int x, int y) { return (uint32_t)(y * fc.width + x); }

enum { EMU_CLOSEST = 0, EMU_ANY = 1, EMU_CLOSEST_EMISSIVE = 2, EMU_ANY_PRECEDES = 3 };
static void emu_trace_queue(Emu* E, int mode, float tnear = -1.0f) {
  FrameCtx& fc = E->fc;
  if (tnear < 0.0f) tnear = FLT_MIN + E->Pf.tnearOffset;
  const SceneDev em = emissive_view(fc.sc);
  // the "precedes" step consumes the queue the closest-emitter step filled (brdf_chain_push), as on the device
  const RayQ* q = mode == EMU_ANY_PRECEDES ? E->chain.data() : E->rays.data();
  const int64_t qn = mode == EMU_ANY_PRECEDES ? (int64_t)std::min<uint32_t>(E->chain_count[0], fc.wave.chain_capacity) : (int64_t)E->qcount;
  // EMU_QUEUE_STATS=1: size of every traced queue and how many of its any-hit rays were occluded (tools/queue_stats.py)
  static const bool stats = getenv("EMU_QUEUE_STATS") != nullptr;
  int64_t n_occ = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : n_occ)
  for (int64_t i = 0; i < qn; ++i) {
    const RayQ& r = q[i];
    const uint32_t dest = f2u(r.d_dest.w);
    HitRec hr;
    if (mode == EMU_ANY) {
      E->occ[dest] = trace8<true>(fc.sc, xyz(r.o_tfar), xyz(r.d_dest), tnear, r.o_tfar.w, &hr) ? 1 : 0;
      n_occ += E->occ[dest];
    } else if (mode == EMU_ANY_PRECEDES) {
      E->occ[dest] = trace8_precedes(fc.sc, xyz(r.o_tfar), xyz(r.d_dest), tnear, r.o_tfar.w, E->hits[dest].tri) ? 1 : 0;
      n_occ += E->occ[dest];
    } else {
      trace8<false>(mode == EMU_CLOSEST_EMISSIVE ? em : fc.sc, xyz(r.o_tfar), xyz(r.d_dest), tnear, r.o_tfar.w, &hr);
      E->hits[dest] = hr;
      if (mode == EMU_CLOSEST_EMISSIVE && hr.tri != 0xFFFFFFFFu) brdf_chain_push(fc.wave, xyz(r.o_tfar), xyz(r.d_dest), hr.t, dest);
    }
  }
  if (stats) {
    static const char* names[] = {"closest", "any", "closest_emissive", "any_precedes"};
    fprintf(stderr, "[emu queue] %s rays %lld occluded %lld\n", names[mode], (long long)qn, (long long)n_occ);
  }
}
template <class F>
static void emu_stream(Emu* E, F&& body) {  // stream half: emits rays, counts nothing
  E->qcount = 0;
  unsigned long long save[3] = {E->counters[0], E->counters[1], E->counters[2]};
  for_pixels(E, E->fc, body);
  E->counters[0] = save[0], E->counters[1] = save[1], E->counters[2] = save[2];
}
template <class F>
static void emu_rows(Emu* E, int ry0, int ry1, F&& body) {
  FrameCtx f = E->fc;
  f.y0 = ry0, f.y1 = ry1;
  for_pixels(E, f, body);
}
static void emu_bind(Emu* E) { E->fc.Rread = E->rp(E->rRead), E->fc.Rwrite = E->rp(E->rWrite), E->fc.Rlast = E->rp(E->rLast); }

extern "C" {

int emu_frame_begin(void* h, const RbCamera* cam, uint32_t frame_idx) {
  Emu* E = (Emu*)h;
  if (!E->haveScene) return RB_ERR_NO_SCENE;
  E->Pf = E->P;
  const RbParams& P = E->Pf;
  const bool banded = !(E->y0 == 0 && E->y1 == E->height);
  if (banded && P.doSpatialReuse && (E->y1 - E->y0) < spatial_reach(P)) return RB_ERR_UNSUPPORTED;
  FrameCtx& fc = E->fc;
  fc = FrameCtx{};
  fc.width = E->width, fc.height = E->height, fc.y0 = E->y0, fc.y1 = E->y1;
  fc.sc = E->sc;
  fc.P = P;
  fc.cam.pos = v3(cam->pos[0], cam->pos[1], cam->pos[2]);
  fc.cam.focal = cam->focal_px;
  memcpy(fc.cam.viewMat, cam->viewMat, 64);
  memcpy(fc.cam.invViewMat, cam->invViewMat, 64);
  fc.prevCam = E->havePrev ? E->prevCam : fc.cam;
  fc.G = E->gp(E->gCur);
  fc.Gprev = E->gp(E->gCur ^ 1);
  fc.frame = E->frame.data();
  fc.counters = E->counters;
  const int margin = banded ? std::max(16, spatial_reach(P)) : 0;
  fc.gy0 = std::max(0, E->y0 - margin);
  fc.gy1 = std::min(E->height, E->y1 + margin);
  fc.gpy0 = E->havePrev ? E->prevGy0 : 0;
  fc.gpy1 = E->havePrev ? E->prevGy1 : 0;
  memset(E->counters, 0, sizeof(E->counters));
  E->frame_idx = frame_idx;
  E->wave = P.wavefront != 0;
  E->wave_spatial = E->wave && P.spatialWeightCalc == RB_SW_CONSTANT && (size_t)E->width * E->height < (1u << RB_CAND_INDEX_BITS);
  const uint32_t npix = (uint32_t)(E->width * E->height);
  uint32_t slots = std::max<uint32_t>(std::max<uint32_t>(4u, (uint32_t)P.spatialReuseNeighborCount + 1u), (uint32_t)P.M_Brdf);
  const uint32_t staged_slots = spatial_staged_slots(P.spatialWeightCalc, (uint32_t)P.spatialReuseNeighborCount + 1u);
  E->wave_spatial_staged = E->wave && P.doSpatialReuse && P.spatialWeightCalc != RB_SW_CONSTANT && staged_slots <= 64u;
  if (E->wave_spatial_staged) slots = std::max(slots, staged_slots);
  if (E->wave) {
    E->rays.resize((size_t)npix * std::max<uint32_t>(slots, (uint32_t)std::max(P.M_Brdf, 1)));
    E->occ.assign((size_t)npix * slots, 0xCD);
    E->hits.resize((size_t)npix * std::max(P.M_Brdf, 1));
    fc.wave.rays = E->rays.data();
    fc.wave.count = &E->qcount;
    fc.wave.capacity = (uint32_t)E->rays.size();
    fc.wave.occ = E->occ.data();
    fc.wave.hits = E->hits.data();
    E->brdf_dir.resize((size_t)npix * std::max(P.M_Brdf, 1));
    fc.wave.brdf_dir = E->brdf_dir.data();
    fc.wave.npix = npix;
    fc.wave.brdf_two_step = (E->em_n_nodes > 0 && !(getenv("RB_TWO_STEP_BRDF") && atoi(getenv("RB_TWO_STEP_BRDF")) == 0)) ? 1u : 0u;
    size_t cand_slots = (E->wave_spatial && P.doSpatialReuse) ? (size_t)P.spatialReuseNeighborCount + 1 : 0;
    if (P.doTemporalReuse) cand_slots = std::max<size_t>(cand_slots, 2);
    E->cand.assign((size_t)npix * cand_slots, U4{0xCDCDCDCDu, 0xCDCDCDCDu, 0xCDCDCDCDu, 0xCDCDCDCDu});
    fc.wave.cand = E->cand.data();
    E->deferred.assign(npix, 0xFFFFFFFFu);
    E->deferred_count = 0;
    E->chain.assign((size_t)npix * std::max(P.M_Brdf, 1), RayQ{});
    E->chain_count[0] = E->chain_count[1] = 0;
    fc.wave.chain_rays = E->chain.data();
    fc.wave.chain_count = E->chain_count;
    fc.wave.chain_capacity = (uint32_t)E->chain.size();
    fc.wave.fuse_vis = 0u;
    fc.wave.fuse_shade = 0u;
    fc.wave.deferred = E->deferred.data();
    fc.wave.deferred_count = &E->deferred_count;
  }
  const bool wave = E->wave;
  emu_bind(E);
  fc.frame_key = rng_frame_key(E->seed, frame_idx, PASS_GBUF, 0);
  emu_rows(E, fc.gy0, fc.gy1, [&](int x, int y, Cnt& c) { gbuffer_pixel(fc, x, y, c); });
  fc.frame_key = rng_frame_key(E->seed, frame_idx, PASS_INITIAL, 0);
  if (wave) {
    if (P.M_Brdf > 0 && fc.sc.n_lights > 0) {
      emu_stream(E, [&](int x, int y, Cnt&) { initial_brdf_gen_pixel(fc, x, y, GenVis{&fc, PX(fc, x, y)}); });
      if (fc.wave.brdf_two_step) {
        emu_trace_queue(E, EMU_CLOSEST_EMISSIVE);
        emu_trace_queue(E, EMU_ANY_PRECEDES);
      } else {
        emu_trace_queue(E, EMU_CLOSEST);
      }
    }
    if (P.doVisibilityPass) {  // the resolve kernel also queues the visibility pass's rays (fresh queue)
      E->qcount = 0;
      fc.wave.fuse_vis = 1u;
      for_pixels(E, fc, [&](int x, int y, Cnt& c) { { uint32_t pk[kPickChunk]; initial_pixel(fc, x, y, ResolveVis{&fc, PX(fc, x, y)}, c, PickStore{pk, 1}); } });
      fc.wave.fuse_vis = 0u;
    } else
      for_pixels(E, fc, [&](int x, int y, Cnt& c) { { uint32_t pk[kPickChunk]; initial_pixel(fc, x, y, ResolveInlineShadowVis{&fc, PX(fc, x, y)}, c, PickStore{pk, 1}); } });
  } else {
    for_pixels(E, fc, [&](int x, int y, Cnt& c) { { uint32_t pk[kPickChunk]; initial_pixel(fc, x, y, InlineVis{&fc, PX(fc, x, y)}, c, PickStore{pk, 1}); } });
  }
  const bool temporal_runs = P.doTemporalReuse && frame_idx > 0 && E->havePrev;
  bool vis_in_temporal = false;
  if (P.doVisibilityPass) {
    if (wave) {
      emu_trace_queue(E, EMU_ANY);  // rays queued by the initial-pass resolve
      vis_in_temporal = temporal_runs;  // applied by the temporal stream pass when there is one
      if (!vis_in_temporal)
        for_pixels(E, fc, [&](int x, int y, Cnt& c) { visibility_pixel(fc, x, y, ResolveVis{&fc, PX(fc, x, y)}, c); });
    } else {
      for_pixels(E, fc, [&](int x, int y, Cnt& c) { visibility_pixel(fc, x, y, InlineVis{&fc, PX(fc, x, y)}, c); });
    }
  }
  if (temporal_runs) {
    std::swap(E->rRead, E->rWrite);
    emu_bind(E);
    fc.frame_key = rng_frame_key(E->seed, frame_idx, PASS_TEMPORAL, 0);
    if (wave) {
      E->qcount = 0;
      fc.wave.fuse_vis = vis_in_temporal ? 2u : 0u;
      if (banded) {  // as the device schedule: bulk launch with deferral, then the deferred pixels with the re-derivation code
        for_pixels(E, fc, [&](int x, int y, Cnt& c) { temporal_gen_pixel<2>(fc, x, y, GenVis{&fc, PX(fc, x, y)}, c); });
        unsigned long long c0 = 0, c1 = 0, c2 = 0;
        for (uint32_t i = 0; i < E->deferred_count; ++i) {
          const uint32_t pi = E->deferred[i];
          Cnt c = {0, 0, 0};
          temporal_gen_pixel<1>(fc, (int)(pi % (uint32_t)fc.width), (int)(pi / (uint32_t)fc.width), GenVis{&fc, pi}, c);
          c0 += c.closest, c1 += c.anyW, c2 += c.anyT;
        }
        E->counters[0] += c0, E->counters[1] += c1, E->counters[2] += c2;
        E->n_deferred_total += E->deferred_count;
      } else {
        for_pixels(E, fc, [&](int x, int y, Cnt& c) { temporal_gen_pixel<0>(fc, x, y, GenVis{&fc, PX(fc, x, y)}, c); });
      }
      fc.wave.fuse_vis = 0u;
      emu_trace_queue(E, EMU_ANY);
      for_pixels(E, fc, [&](int x, int y, Cnt& c) { temporal_merge_pixel(fc, x, y, c); });
    } else if (banded) {
      for_pixels(E, fc, [&](int x, int y, Cnt& c) { temporal_pixel<InlineVis, true>(fc, x, y, InlineVis{&fc, PX(fc, x, y)}, c); });
    } else {
      for_pixels(E, fc, [&](int x, int y, Cnt& c) { temporal_pixel<InlineVis, false>(fc, x, y, InlineVis{&fc, PX(fc, x, y)}, c); });
    }
  }
  E->open = true;
  E->shaded = false;
  return 0;
}

int emu_frame_spatial(void* h, int i) {
  Emu* E = (Emu*)h;
  if (!E->open) return -1;
  FrameCtx& fc = E->fc;
  std::swap(E->rRead, E->rWrite);
  emu_bind(E);
  fc.spatial_iter = i;
  fc.frame_key = rng_frame_key(E->seed, E->frame_idx, PASS_SPATIAL, (uint32_t)i);
  E->shaded = false;
  if (E->wave_spatial) {
    E->qcount = 0;
    E->shaded = (i == E->Pf.spatialPassCount - 1);  // the last pass's resolve also shades, as on the device
    fc.wave.fuse_shade = E->shaded ? 1u : 0u;
    for_pixels(E, fc, [&](int x, int y, Cnt& c) { spatial_gen_pixel(fc, x, y, GenVis{&fc, PX(fc, x, y)}, c); });
    emu_trace_queue(E, EMU_ANY);
    for_pixels(E, fc, [&](int x, int y, Cnt& c) { spatial_merge_pixel(fc, x, y, c); });
    fc.wave.fuse_shade = 0u;
  } else if (E->wave_spatial_staged) {  // as frame_spatial of restir_b200.cu: spatial_pixel staged (StagedVis)
    const int mode = E->Pf.spatialWeightCalc;
    emu_stream(E, [&](int x, int y, Cnt& c) { spatial_pixel(fc, x, y, StagedVis<1>{&fc, PX(fc, x, y), 0u, false}, c); });
    emu_trace_queue(E, EMU_ANY);
    if (mode == RB_SW_CONSTANT_DEBIAS_Z_TERM || mode == RB_SW_CONSTANT_DEBIAS_CONTRIB) {
      emu_stream(E, [&](int x, int y, Cnt& c) { spatial_pixel(fc, x, y, StagedVis<2>{&fc, PX(fc, x, y), 0u, false}, c); });
      emu_trace_queue(E, EMU_ANY);
    }
    for_pixels(E, fc, [&](int x, int y, Cnt& c) { spatial_pixel(fc, x, y, StagedVis<3>{&fc, PX(fc, x, y), 0u, false}, c); });
  } else {
    for_pixels(E, fc, [&](int x, int y, Cnt& c) { spatial_pixel(fc, x, y, InlineVis{&fc, PX(fc, x, y)}, c); });
  }
  return 0;
}

int emu_frame_end(void* h, float* rgb_out) {
  Emu* E = (Emu*)h;
  if (!E->open) return -1;
  FrameCtx& fc = E->fc;
  std::swap(E->rRead, E->rWrite);
  emu_bind(E);
  if (!E->shaded) for_pixels(E, fc, [&](int x, int y, Cnt& c) { shade_pixel(fc, x, y, InlineVis{&fc, PX(fc, x, y)}, c); });
  E->shaded = false;
  std::swap(E->rLast, E->rRead);
  E->gCur ^= 1;
  E->prevCam = fc.cam;
  E->prevGy0 = fc.gy0, E->prevGy1 = fc.gy1;
  E->havePrev = true;
  E->open = false;
  if (rgb_out) memcpy(rgb_out, E->frame.data(), E->frame.size() * sizeof(float));
  return 0;
}

// textured materials: the tables rb_set_textures uploads, on the host
int emu_set_textures(void* h, const RbTexture* textures, uint32_t n_textures, const RbMaterialTextures* per_material, uint32_t n_materials) {
  Emu* E = (Emu*)h;
  if (!E->haveScene || n_materials * 3 != E->hs.mat.size()) return -1;
  E->tex_tab.resize(n_textures);
  E->tex_data.resize(n_textures);
  for (uint32_t t = 0; t < n_textures; ++t) {
    const RbTexture& T = textures[t];
    const unsigned char* d = (const unsigned char*)T.data;
    E->tex_data[t].assign(d, d + (size_t)T.scan_width * T.height);
    E->tex_tab[t] = TexDev{E->tex_data[t].data(), T.width, T.height, T.scan_width, T.pixel_size};
  }
  E->tex_slots.resize(n_materials);
  memcpy(E->tex_slots.data(), per_material, n_materials * sizeof(I4));
  E->sc.tex = E->tex_tab.data();
  E->sc.mat_tex = E->tex_slots.data();
  bool any_normal = false;
  for (uint32_t m = 0; m < n_materials; ++m) any_normal = any_normal || per_material[m].normal >= 0;
  if (any_normal && E->hs.tan.empty()) return RB_ERR_INVALID_ARGUMENT;
  E->sc.tri_tan = any_normal ? E->hs.tan.data() : nullptr;
  E->havePrev = false;
  return 0;
}

// sky: the record rb_set_sky uploads, on the host
int emu_set_sky(void* h, const RbTexture* sky) {
  Emu* E = (Emu*)h;
  E->sky_data.clear();
  E->sc.sky = TexDev{nullptr, 0, 0, 0, 0, 0};
  E->havePrev = false;
  if (!sky) return 0;
  const unsigned char* d = (const unsigned char*)sky->data;
  E->sky_data.assign(d, d + (size_t)sky->scan_width * sky->height);
  E->sc.sky = TexDev{E->sky_data.data(), sky->width, sky->height, sky->scan_width, sky->pixel_size, 1};
  return 0;
}

// N2: the kernel bodies of rb_render_mis_frame (k_gbuffer + k_mis_direct) on the G-buffer the next frame overwrites
int emu_render_mis_frame(void* h, const RbCamera* cam, uint32_t frame_idx, uint32_t techniques, float* rgb_out) {
  Emu* E = (Emu*)h;
  if (!E->haveScene) return RB_ERR_NO_SCENE;
  FrameCtx fc{};
  fc.width = E->width, fc.height = E->height, fc.y0 = E->y0, fc.y1 = E->y1;
  fc.sc = E->sc;
  fc.P = E->P;
  fc.cam.pos = v3(cam->pos[0], cam->pos[1], cam->pos[2]);
  fc.cam.focal = cam->focal_px;
  memcpy(fc.cam.viewMat, cam->viewMat, 64);
  memcpy(fc.cam.invViewMat, cam->invViewMat, 64);
  fc.prevCam = fc.cam;
  fc.G = fc.Gprev = E->gp(E->gCur);  // (emu: two G-buffers, gCur is the one the next frame writes)
  fc.frame = E->frame.data();
  fc.counters = E->counters;
  fc.gy0 = E->y0, fc.gy1 = E->y1;
  fc.mis_flags = techniques;
  memset(E->counters, 0, sizeof(E->counters));
  fc.frame_key = rng_frame_key(E->seed, frame_idx, PASS_GBUF, 0);
  for_pixels(E, fc, [&](int x, int y, Cnt& c) { gbuffer_pixel(fc, x, y, c); });
  fc.frame_key = rng_frame_key(E->seed, frame_idx, PASS_MIS, 0);
  for_pixels(E, fc, [&](int x, int y, Cnt& c) { mis_direct_pixel(fc, x, y, InlineVis{&fc, PX(fc, x, y)}, c); });
  if (rgb_out) memcpy(rgb_out, E->frame.data(), E->frame.size() * sizeof(float));
  return 0;
}

int emu_render_frame(void* h, const RbCamera* cam, uint32_t frame_idx, float* rgb_out) {
  Emu* E = (Emu*)h;
  int rc = emu_frame_begin(h, cam, frame_idx);
  if (rc) return rc;
  if (E->Pf.doSpatialReuse)
    for (int i = 0; i < E->Pf.spatialPassCount; ++i) emu_frame_spatial(h, i);
  return emu_frame_end(h, rgb_out);
}

int emu_set_band(void* h, int y0, int y1) {
  Emu* E = (Emu*)h;
  if (y0 < 0 || y1 > E->height || y0 >= y1 || E->open) return RB_ERR_INVALID_ARGUMENT;
  if (E->havePrev && (y0 < E->prevGy0 || y1 > E->prevGy1)) return RB_ERR_UNSUPPORTED;
  E->y0 = y0, E->y1 = y1;
  return 0;
}
int emu_halo_rows(void* h) { return spatial_reach(((Emu*)h)->open ? ((Emu*)h)->Pf : ((Emu*)h)->P); }
// rows [y, y+rows) of the reservoirs the next spatial pass reads: 4 planes packed back to back (52 B / px)
static int emu_halo_copy(Emu* E, int y, int rows, char* host, bool to_host) {
  if (y < 0 || rows < 0 || y + rows > E->height) return -1;
  Emu::RStore& R = E->rs[E->open ? E->rWrite : E->rLast];
  const size_t px = (size_t)rows * E->width, off = (size_t)y * E->width;
  char* planes[4] = {(char*)(R.a.data() + off), (char*)(R.b.data() + off), (char*)(R.c.data() + off), (char*)(R.li.data() + off)};
  const size_t sz[4] = {px * 16, px * 16, px * 16, px * 4};
  for (int i = 0; i < 4; ++i) {
    if (to_host)
      memcpy(host, planes[i], sz[i]);
    else
      memcpy(planes[i], host, sz[i]);
    host += sz[i];
  }
  return 0;
}
int emu_halo_export(void* h, int y, int rows, void* dst) { return emu_halo_copy((Emu*)h, y, rows, (char*)dst, true); }
int emu_halo_import(void* h, int y, int rows, const void* src) { return emu_halo_copy((Emu*)h, y, rows, (char*)src, false); }

uint64_t emu_deferred_total(void* h) { return ((Emu*)h)->n_deferred_total; }
// calc_I_M with (use_const = 1) and without the per-material constants of make_mat_const: must be the same bits
float emu_calc_I_M(float nDotV, float n, int use_const) {
  if (!use_const) return calc_I_M(nDotV, n, nullptr);
  const MatConst mc = make_mat_const(n);
  return calc_I_M(nDotV, n, &mc);
}
// horizon pre-test of initial_pixel (rb_passes.cuh): switch the check mode, return and reset
// {pre-culled, confirmed by the exact test, violations, candidates, culled by the exact test}
void emu_horizon_cull_check(int mode, uint64_t* out4) {
  for (int i = 0; i < 5; ++i) {
    if (out4) out4[i] = g_horizon_cull_check.counts[i];
    g_horizon_cull_check.counts[i] = 0;
  }
  g_horizon_cull_check.check_mode = mode;
}
void emu_counters(void* h, uint64_t* out3) {
  Emu* E = (Emu*)h;
  out3[0] = E->counters[0], out3[1] = E->counters[1], out3[2] = E->counters[2];
}

// N1: the kernel body of k_accumulate_display, serial sums
int emu_accumulate_display(void* h, uint32_t acc_frame_ctr, int tonemap, int gamma_correct, float* display_rgba_out, double* stats4) {
  Emu* E = (Emu*)h;
  const size_t n = (size_t)E->width * E->height;
  if (E->accumulator.size() != n * 3) E->accumulator.assign(n * 3, 0.0f);
  if (E->display.size() != n) E->display.assign(n, F4{0, 0, 0, 0});
  const float mix_a = 1.0f / (float)(acc_frame_ctr + 1u);
  double s = 0, s2 = 0;
  for (size_t pi = (size_t)E->y0 * E->width; pi < (size_t)E->y1 * E->width; ++pi) {
    const float pm = accumulate_display_pixel(E->frame.data(), E->accumulator.data(), E->display.data(), pi, mix_a, tonemap != 0, gamma_correct != 0);
    s += (double)pm;
    s2 += (double)(pm * pm);
  }
  const double count = (double)(E->y1 - E->y0) * E->width;
  if (stats4) stats4[0] = s, stats4[1] = s2, stats4[2] = s / count, stats4[3] = s2 / count - (s / count) * (s / count);
  if (display_rgba_out) memcpy(display_rgba_out, E->display.data(), n * 16);
  return 0;
}

int emu_readback(void* h, int id, void* dst, size_t bytes) {
  Emu* E = (Emu*)h;
  const Emu::GStore& G = E->gs[E->gCur ^ 1];
  const Emu::RStore& R = E->rs[E->rLast];
  const void* src = nullptr;
  size_t need = 0;
  switch (id) {
    case RB_BUF_GBUF_POS_DEPTH: src = G.a.data(), need = G.a.size() * 16; break;
    case RB_BUF_GBUF_NORMAL_SHIN: src = G.b.data(), need = G.b.size() * 16; break;
    case RB_BUF_GBUF_DIFFUSE_IIM: src = G.c.data(), need = G.c.size() * 16; break;
    case RB_BUF_GBUF_SPEC_TYPE: src = G.d.data(), need = G.d.size() * 16; break;
    case RB_BUF_GBUF_EMISSION: src = G.e.data(), need = G.e.size() * 16; break;
    case RB_BUF_HIT_IDS: src = G.ids.data(), need = G.ids.size() * 8; break;
    case RB_BUF_RES_POINT_WSUM: src = R.a.data(), need = R.a.size() * 16; break;
    case RB_BUF_RES_NORMAL_W: src = R.b.data(), need = R.b.size() * 16; break;
    case RB_BUF_RES_LI_CONF: src = R.c.data(), need = R.c.size() * 16; break;
    case RB_BUF_RES_LIGHT_IDX: src = R.li.data(), need = R.li.size() * 4; break;
    case RB_BUF_FRAME_RGB: src = E->frame.data(), need = E->frame.size() * 4; break;
    case RB_BUF_ACCUMULATOR: src = E->accumulator.data(), need = E->accumulator.size() * 4; break;
    case RB_BUF_DISPLAY: src = E->display.data(), need = E->display.size() * 16; break;
    case RB_BUF_ALIAS_PROB: src = E->hs.alias_prob.data(), need = E->hs.alias_prob.size() * 4; break;
    case RB_BUF_ALIAS_IDX: src = E->hs.alias_idx.data(), need = E->hs.alias_idx.size() * 4; break;
    case RB_BUF_LIGHT_CDF: src = E->hs.cdf.data(), need = E->hs.cdf.size() * 4; break;
    default: return -1;
  }
  if (bytes < need) return -1;
  if (need) memcpy(dst, src, need);
  return 0;
}

int emu_trace_closest(void* h, const RbRay* rays, RbHit* hits, uint32_t n) {
  Emu* E = (Emu*)h;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const RbRay& r = rays[i];
    HitRec hr;
    const bool hit = trace8<false>(E->sc, v3(r.org_x, r.org_y, r.org_z), v3(r.dir_x, r.dir_y, r.dir_z), r.tnear, r.tfar, &hr);
    RbHit o;
    if (hit) {
      const U4 info = E->sc.tri_info[hr.tri];
      o.t = hr.t, o.u = hr.u, o.v = hr.v, o.primID = info.y, o.geomID = info.x;
    } else {
      o.t = r.tfar, o.u = 0, o.v = 0, o.primID = 0xFFFFFFFFu, o.geomID = 0xFFFFFFFFu;
    }
    hits[i] = o;
  }
  return 0;
}
int emu_trace_occluded(void* h, const RbRay* rays, uint8_t* occ, uint32_t n) {
  Emu* E = (Emu*)h;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const RbRay& r = rays[i];
    occ[i] = trace8<true>(E->sc, v3(r.org_x, r.org_y, r.org_z), v3(r.dir_x, r.dir_y, r.dir_z), r.tnear, r.tfar, nullptr) ? 1 : 0;
  }
  return 0;
}

// traversal statistics for BVH-quality work: average node visits (trav_step calls) per ray
double g_tri_tests = 0;
double emu_last_tri_tests() { return g_tri_tests; }
double emu_trace_steps(void* h, const RbRay* rays, uint32_t n, int any) {
  Emu* E = (Emu*)h;
  double steps = 0, tris = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : steps, tris)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const RbRay& r = rays[i];
    Trav T;
    U2 stack[RB_STACK_MAX];
    if (!trav_init(T, E->sc, v3(r.org_x, r.org_y, r.org_z), v3(r.dir_x, r.dir_y, r.dir_z), r.tnear, r.tfar)) continue;
    bool more = true;
    while (more) {
      more = any ? trav_step<true>(T, stack, E->sc) : trav_step<false>(T, stack, E->sc);
      steps += 1;
    }
#ifdef RB_TRAV_STATS
    tris += T.n_tri_tests;
#endif
  }
  g_tri_tests = tris / (n ? n : 1);
  return steps / (n ? n : 1);
}

// structural check of the wide BVH: every triangle reachable exactly once, every child box
// (decoded from the 8-bit grid) contains its subtree's triangles. Returns 0 if sound.
int emu_validate_bvh(void* h) {
  Emu* E = (Emu*)h;
  if (E->n_nodes == 0) return 0;
  std::vector<int> seen(E->hs.n, 0);
  struct Item {
    uint32_t node;
    float lo[3], hi[3];
  };
  std::vector<Item> st;
  st.push_back({0, {-FLT_MAX, -FLT_MAX, -FLT_MAX}, {FLT_MAX, FLT_MAX, FLT_MAX}});
  while (!st.empty()) {
    Item it = st.back();
    st.pop_back();
    const F4* np = E->node8.data() + RB_NODE_F4 * (size_t)it.node;
    const uint32_t eb = f2u(np[0].w), imask = eb >> 24;
    const float org[3] = {np[0].x, np[0].y, np[0].z};
    const float sc3[3] = {u2f(byte_of(eb, 0) << 23), u2f(byte_of(eb, 1) << 23), u2f(byte_of(eb, 2) << 23)};
    const uint32_t child_base = f2u(np[1].x), tri_base = f2u(np[1].y);
    const uint32_t tri_shift = f2u(np[1].z);
    uint32_t hitw[8];  // per-slot hit word in the traversal's hit-mask layout: internal 1 << (24 + s), leaf tri bits
    {
      const uint32_t pw[4] = {f2u(np[5].x), f2u(np[5].y), f2u(np[5].z), f2u(np[5].w)};
      uint32_t low_tris = 0;
      for (int s = 0; s < 8; ++s) {
        const uint32_t half = (pw[s & 3] >> (s < 4 ? 0 : 16)) & 0xFFFFu;
        if (half >> 12) {
          if (half != (1u << (12 + (s & 3)))) return 20;
          hitw[s] = 1u << (24 + s);
        } else {
          hitw[s] = s < 4 ? half : half << tri_shift;
          if (s < 4) low_tris += (uint32_t)__builtin_popcount(half);
        }
      }
      if (low_tris != tri_shift && (hitw[4] | hitw[5] | hitw[6] | hitw[7]) & 0xFFFFFFu) return 21;
    }
    const uint32_t q[6][2] = {{f2u(np[2].x), f2u(np[2].y)}, {f2u(np[2].z), f2u(np[2].w)}, {f2u(np[3].x), f2u(np[3].y)},
                              {f2u(np[3].z), f2u(np[3].w)}, {f2u(np[4].x), f2u(np[4].y)}, {f2u(np[4].z), f2u(np[4].w)}};
    int rel = 0;
    uint32_t tri_bits_seen = 0;
    for (int s = 0; s < 8; ++s) {
      const uint32_t hw = hitw[s];
      if (!hw) {
        if ((imask >> s) & 1) return 10;
        continue;
      }
      float lo[3], hi[3];
      for (int a = 0; a < 3; ++a) {
        lo[a] = org[a] + (128.0f + (float)byte_of(q[a][s >> 2], s & 3)) * sc3[a];
        hi[a] = org[a] + (128.0f + (float)byte_of(q[3 + a][s >> 2], s & 3)) * sc3[a];
        if (byte_of(q[a][s >> 2], s & 3) > 127u || byte_of(q[3 + a][s >> 2], s & 3) > 127u) return 18;
        // carry the intersection of all ancestor boxes: triangles must lie inside every one of them
        lo[a] = fmaxf(lo[a], it.lo[a]);
        hi[a] = fminf(hi[a], it.hi[a]);
        if (lo[a] > hi[a]) return 11;
      }
      if ((imask >> s) & 1) {
        if (hw != (1u << (24 + s))) return 12;
        Item ch;
        ch.node = child_base + rel++;
        if (ch.node >= E->n_nodes) return 13;
        memcpy(ch.lo, lo, 12), memcpy(ch.hi, hi, 12);
        st.push_back(ch);
      } else {
        // a leaf's hit word: a run of 1..3 one-bits inside the low 24 bits, disjoint from the other leaves' runs
        if (hw >> 24) return 14;
        if (hw & tri_bits_seen) return 19;
        tri_bits_seen |= hw;
        const int off = __builtin_ctz(hw), cnt = __builtin_popcount(hw);
        if (cnt > RB_LEAF_MAX || hw != (((1u << cnt) - 1u) << off)) return 14;
        for (int t = 0; t < cnt; ++t) {
          const F4* tp = E->tri_isect.data() + 3 * (size_t)(tri_base + off + t);
          const uint32_t id = f2u(tp[2].y);
          if (id >= E->hs.n) return 15;
          seen[id]++;
          const float* p = E->hs.pos.data() + 9 * (size_t)id;
          for (int v = 0; v < 3; ++v)
            for (int a = 0; a < 3; ++a)
              if (p[3 * v + a] < lo[a] || p[3 * v + a] > hi[a]) return 16;
        }
      }
    }
  }
  for (size_t i = 0; i < seen.size(); ++i)
    if (seen[i] != 1) return 17;
  return 0;
}
}
