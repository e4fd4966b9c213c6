// restir_b200.cu — kernels and the extern "C" shim of include/restir_b200.h.
//
// Hand-written CUDA for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a
// -fmad=false). No OptiX, no Triton, no CPU fallback: every entry point either
// runs on the GPU or returns an error code.
#include <cuda_runtime.h>

#include <algorithm>
#include <type_traits>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "rb_build.cuh"
#include "rb_host_scene.h"
#include "rb_obj_loader.h"
#include "rb_passes.cuh"

using namespace rb;

// =====================================================================================
// kernels
// =====================================================================================
namespace {

constexpr int kTileW = 8, kTileH = 32;  // one warp = 8x4 pixels, one CTA = 8x32

__device__ __forceinline__ void flush_counts(unsigned long long* ctr, const Cnt& c) {
  const unsigned a = __reduce_add_sync(0xFFFFFFFFu, c.closest);
  const unsigned b = __reduce_add_sync(0xFFFFFFFFu, c.anyW);
  const unsigned d = __reduce_add_sync(0xFFFFFFFFu, c.anyT);
  if ((threadIdx.x + threadIdx.y * blockDim.x) % 32 == 0) {
    if (a) atomicAdd(ctr + 0, (unsigned long long)a);
    if (b) atomicAdd(ctr + 1, (unsigned long long)b);
    if (d) atomicAdd(ctr + 2, (unsigned long long)d);
  }
}

// One thread per pixel, 8x4 pixels per warp. VIS is the visibility policy: InlineVis (trace on the spot),
// GenVis (stream half: emit rays), ResolveVis (resolve half: read traced results). COUNT: whether this launch
// contributes to the ray counters (the stream half of a pass does not: the resolve half counts the same rays).
#define RB_PIXEL_KERNEL(NAME, VIS, COUNT, MINB, CALL) RB_PIXEL_KERNEL_T(NAME, VIS, COUNT, kTileW* kTileH, MINB, CALL)
// THREADS = 8 x block height (launch_rows is told the same number): 128-thread CTAs give the register allocator a finer
// occupancy ladder (e.g. 94 registers: two 256-thread CTAs = 16 warps, but five 128-thread CTAs = 20 warps)
#define RB_PIXEL_KERNEL_T(NAME, VIS, COUNT, THREADS, MINB, CALL)               \
  __global__ void __launch_bounds__(THREADS, MINB) NAME(FrameCtx fc) {         \
    const int x = blockIdx.x * kTileW + threadIdx.x;                           \
    const int y = fc.y0 + blockIdx.y * blockDim.y + threadIdx.y;               \
    Cnt cnt = {0, 0, 0};                                                       \
    if (fc.wave.reset_pair != nullptr && (blockIdx.x | blockIdx.y | threadIdx.x | threadIdx.y) == 0)   \
      fc.wave.reset_pair[0] = 0u, fc.wave.reset_pair[1] = 0u;                  \
    if (x < fc.width && y < fc.y1) {                                           \
      const VIS vis = {&fc, (uint32_t)(y * fc.width + x)};                     \
      (void)vis;                                                               \
      CALL;                                                                    \
    }                                                                          \
    if (COUNT) flush_counts(fc.counters, cnt);                                 \
  }

// initial_pixel's per-thread list of picked lights: [candidate][thread] words of shared memory (conflict-free)
template <int THREADS>
__device__ __forceinline__ PickStore smem_picks() {
  __shared__ uint32_t s_pick[kPickChunk * THREADS];
  return PickStore{s_pick + threadIdx.y * blockDim.x + threadIdx.x, THREADS};
}

#ifndef RB_GB_MINB
#define RB_GB_MINB 7
#endif
// 128-thread CTAs, seven per SM (72 registers, no spills): measured 0.85 (96 regs, 5 CTAs) -> 0.72 ms with the material constants
RB_PIXEL_KERNEL_T(k_gbuffer, InlineVis, true, 128, RB_GB_MINB, gbuffer_pixel(fc, x, y, cnt))
RB_PIXEL_KERNEL(k_initial, InlineVis, true, 1, initial_pixel(fc, x, y, vis, cnt, smem_picks<kTileW * kTileH>()))
RB_PIXEL_KERNEL(k_visibility, InlineVis, true, 1, visibility_pixel(fc, x, y, vis, cnt))
RB_PIXEL_KERNEL(k_temporal, InlineVis, true, 1, (temporal_pixel<InlineVis, false>(fc, x, y, vis, cnt)))
RB_PIXEL_KERNEL(k_temporal_banded, InlineVis, true, 1, (temporal_pixel<InlineVis, true>(fc, x, y, vis, cnt)))
RB_PIXEL_KERNEL(k_spatial, InlineVis, true, 1, spatial_pixel(fc, x, y, vis, cnt))
RB_PIXEL_KERNEL(k_shade, InlineVis, true, 4, shade_pixel(fc, x, y, vis, cnt))
// N2: one-sample MIS direct lighting (ground truth next to the path); its two rays per pixel are traced inline
RB_PIXEL_KERNEL(k_mis_direct, InlineVis, true, 1, mis_direct_pixel(fc, x, y, vis, cnt))
// wavefront halves
RB_PIXEL_KERNEL(k_initial_brdf_stream, GenVis, false, 1, initial_brdf_gen_pixel(fc, x, y, vis))
// 92 regs, no spills; the normal-map branch of the candidates' emitter hits would not fit (96 + spills), so the frame
// launches the second instantiation only while a material of the scene has a normal map
#ifndef RB_INIT_MINB
#define RB_INIT_MINB 5
#endif
RB_PIXEL_KERNEL_T(k_initial_resolve, ResolveVisFlat, true, 128, RB_INIT_MINB, initial_pixel(fc, x, y, vis, cnt, smem_picks<128>()))
RB_PIXEL_KERNEL_T(k_initial_resolve_nmap, ResolveVis, true, 128, 5, initial_pixel(fc, x, y, vis, cnt, smem_picks<128>()))

RB_PIXEL_KERNEL(k_initial_resolve_inline_shadow, ResolveInlineShadowVis, true, 1, initial_pixel(fc, x, y, vis, cnt, smem_picks<kTileW * kTileH>()))
RB_PIXEL_KERNEL(k_visibility_resolve, ResolveVis, true, 1, visibility_pixel(fc, x, y, vis, cnt))
// 128-thread CTAs give the register allocator a finer occupancy ladder: seven per SM at 72 registers (0.265 -> 0.256 ms)
#ifndef RB_TS_THREADS
#define RB_TS_THREADS 128
#define RB_TS_MINB 7
#endif
RB_PIXEL_KERNEL_T(k_temporal_stream, GenVis, true, RB_TS_THREADS, RB_TS_MINB, temporal_gen_pixel<0>(fc, x, y, vis, cnt))
// bands: the bulk launch defers the few pixels whose reprojection leaves the rows held here ...
RB_PIXEL_KERNEL_T(k_temporal_stream_banded, GenVis, true, RB_TS_THREADS, RB_TS_MINB, temporal_gen_pixel<2>(fc, x, y, vis, cnt))
// ... to this one, which carries the G-buffer re-derivation (traversal) code; grid-stride over the deferred list
__global__ void __launch_bounds__(128) k_temporal_stream_deferred(FrameCtx fc) {
  const uint32_t n = *fc.wave.deferred_count;
  Cnt cnt = {0, 0, 0};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t pi = fc.wave.deferred[i];
    const GenVis vis = {&fc, pi};
    temporal_gen_pixel<1>(fc, (int)(pi % (uint32_t)fc.width), (int)(pi / (uint32_t)fc.width), vis, cnt);
  }
  flush_counts(fc.counters, cnt);
}
#ifndef RB_TR_THREADS
#define RB_TR_THREADS 256
#define RB_TR_MINB 4
#endif
RB_PIXEL_KERNEL_T(k_temporal_resolve, ResolveVis, true, RB_TR_THREADS, RB_TR_MINB, temporal_merge_pixel(fc, x, y, cnt))
// eight 128-thread CTAs per SM at 62 registers, no spills (0.42 -> 0.375 ms against three 256-thread CTAs)
#ifndef RB_SS_THREADS
#define RB_SS_THREADS 128
#define RB_SS_MINB 8
#endif
RB_PIXEL_KERNEL_T(k_spatial_stream, GenVis, true, RB_SS_THREADS, RB_SS_MINB, spatial_gen_pixel(fc, x, y, vis, cnt))
#ifndef RB_SR_THREADS
#define RB_SR_THREADS 256
#define RB_SR_MINB 4
#endif
RB_PIXEL_KERNEL_T(k_spatial_resolve, ResolveVis, true, RB_SR_THREADS, RB_SR_MINB, spatial_merge_pixel(fc, x, y, cnt))
// the other spatial MIS modes in the wavefront schedule: spatial_pixel staged (StagedVis, rb_passes.cuh)
#define RB_STAGED_KERNEL(NAME, STAGE, COUNT)                                                           \
  __global__ void __launch_bounds__(kTileW* kTileH, 2) NAME(FrameCtx fc) {                            \
    const int x = blockIdx.x * kTileW + threadIdx.x;                                                   \
    const int y = fc.y0 + blockIdx.y * blockDim.y + threadIdx.y;                                       \
    Cnt cnt = {0, 0, 0};                                                                               \
    if (fc.wave.reset_pair != nullptr && (blockIdx.x | blockIdx.y | threadIdx.x | threadIdx.y) == 0)   \
      fc.wave.reset_pair[0] = 0u, fc.wave.reset_pair[1] = 0u;                                          \
    if (x < fc.width && y < fc.y1) {                                                                   \
      const StagedVis<STAGE> vis = {&fc, (uint32_t)(y * fc.width + x), 0u, false};                     \
      spatial_pixel(fc, x, y, vis, cnt);                                                               \
    }                                                                                                  \
    if (COUNT) flush_counts(fc.counters, cnt);                                                         \
  }
RB_STAGED_KERNEL(k_spatial_mis_stage1, 1, false)
RB_STAGED_KERNEL(k_spatial_mis_stage2, 2, false)
RB_STAGED_KERNEL(k_spatial_mis_stage3, 3, true)

// accumulate + tonemap + gamma + statistics (N1). One thread per pixel; the two double sums are reduced per warp
// and added with one atomic pair per warp (summation order is not fixed: results agree with a serial sum to
// ~1e-15 relative, which is the stated tolerance of the statistics).
__global__ void __launch_bounds__(256) k_accumulate_display(const float* __restrict__ frame, float* __restrict__ accumulator,
                                                            F4* __restrict__ display, int width, int y0, int y1, float mix_a,
                                                            int tonemap, int gamma, double* __restrict__ sums) {
  const size_t first = (size_t)y0 * width, count = (size_t)(y1 - y0) * width;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0, m2 = 0.0;
  if (i < count) {
    const float pm = accumulate_display_pixel(frame, accumulator, display, first + i, mix_a, tonemap != 0, gamma != 0);
    m = (double)pm;
    m2 = (double)(pm * pm);  // float product, as the reference squares the float mean before widening
  }
  for (int o = 16; o > 0; o >>= 1) {
    m += __shfl_down_sync(0xFFFFFFFFu, m, o);
    m2 += __shfl_down_sync(0xFFFFFFFFu, m2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sums + 0, m);
    atomicAdd(sums + 1, m2);
  }
}

// calc_I_M's per-material constants, computed on the device by the functions the pixels would call (same bits)
__global__ void k_material_constants(const F4* __restrict__ mat, MatConst* __restrict__ out, uint32_t n_mat) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_mat) out[i] = make_mat_const(mat[3 * (size_t)i].w);
}

// ---- persistent traversal kernels over the ray queue ------------------------------------------------
// One ray per lane. Three things keep the warps full in this divergent workload:
//   * dynamic fetch: warps pull rays from the queue with one aggregated atomic and refill the lanes whose rays
//     have terminated as soon as fewer than refill_lanes lanes are busy;
//   * phase scheduling: every iteration runs a NODE phase (all lanes with a pending node group test eight child
//     boxes in lockstep) and, only when at least tri_lanes lanes hold pending leaf triangles (or node work is
//     running out), a TRIANGLE phase in which each of those lanes tests one triangle. Leaf hits wait in a
//     per-lane triangle stack (the top end of the traversal stack) until then. Traversal results do not depend on
//     the processing order, so this is invisible to parity;
//   * a short per-lane stack in local memory (L1-resident), ray state in registers.
// Grid = SM count x resident blocks, sized by the host.
constexpr int kTraceThreads = 128;

#ifndef RB_TRACE_MINB
#define RB_TRACE_MINB 8
#endif
// where the rays come from and where the results go: the frame's ray queue, or the RbRay / RbHit arrays of the ray seam
struct QueueIO {
  const RayQ* rays;
  const uint32_t* count_ptr;
  uint32_t capacity;
  uint8_t* occ;
  HitRec* hits;
  float tnear;
  WaveBufs chain;  // chain.chain_rays != nullptr: a closest-hit ray that found something queues its follow-up ray there
  __device__ uint32_t count() const { return min(*count_ptr, capacity); }
  __device__ void fetch(uint32_t i, V3* o, V3* d, float* tn, float* tf, uint32_t* dest) const {
    const float4 a = __ldg(reinterpret_cast<const float4*>(&rays[i].o_tfar));
    const float4 b = __ldg(reinterpret_cast<const float4*>(&rays[i].d_dest));
    *o = v3(a.x, a.y, a.z), *d = v3(b.x, b.y, b.z), *tn = tnear, *tf = a.w, *dest = __float_as_uint(b.w);
  }
  __device__ uint32_t tie_id(uint32_t dest) const { return hits[dest].tri; }  // written by the previous trace
  __device__ void put_any(uint32_t dest, bool hit) const { occ[dest] = hit ? 1 : 0; }
  __device__ void put_closest(const SceneDev&, uint32_t dest, const Trav& T) const {
    hits[dest] = T.best;
    if (chain.chain_rays != nullptr && T.best.tri != 0xFFFFFFFFu) brdf_chain_push(chain, T.o, T.d, T.best.t, dest);
  }
};
struct SeamIO {
  const RbRay* rays;
  uint32_t n;
  uint8_t* occ;
  RbHit* hits;
  __device__ uint32_t count() const { return n; }
  __device__ void fetch(uint32_t i, V3* o, V3* d, float* tn, float* tf, uint32_t* dest) const {
    const float4* p = reinterpret_cast<const float4*>(rays + i);  // RTCRay layout, 48 bytes
    const float4 a = __ldg(p), b = __ldg(p + 1);
    const float c = __ldg(reinterpret_cast<const float*>(p + 2));
    *o = v3(a.x, a.y, a.z), *d = v3(b.x, b.y, b.z), *tn = a.w, *tf = c, *dest = i;
  }
  __device__ uint32_t tie_id(uint32_t) const { return 0u; }
  __device__ void put_any(uint32_t dest, bool hit) const { occ[dest] = hit ? 1 : 0; }
  __device__ void put_closest(const SceneDev& sc, uint32_t dest, const Trav& T) const {
    RbHit o;
    if (T.best.tri != 0xFFFFFFFFu) {
      const U4 info = sc.tri_info[T.best.tri];
      o.t = T.best.t, o.u = T.best.u, o.v = T.best.v, o.primID = info.y, o.geomID = info.x;
    } else {
      o.t = T.tfar, o.u = 0, o.v = 0, o.primID = 0xFFFFFFFFu, o.geomID = 0xFFFFFFFFu;
    }
    hits[dest] = o;
  }
};

// SMEM: the per-lane stack lives in dynamic shared memory, stack_n entries of kTraceThreads x 8 bytes (the host passes
// 2 * tree depth + 2, what a traversal can need); otherwise in local memory (deep trees: RB_STACK_MAX entries).
template <bool ANY, bool TIE, class IO, bool SMEM>
__global__ void __launch_bounds__(kTraceThreads, RB_TRACE_MINB) k_trace_queue(SceneDev sc, IO io, uint32_t* __restrict__ next,
                                                                              int refill_lanes, int tri_lanes, int stack_n) {
  const uint32_t count = io.count();
  const unsigned lane = threadIdx.x & 31u;
  const unsigned FULL = 0xFFFFFFFFu;
  Trav T;
  // node groups grow up from 0 (T.sp), pending triangle groups grow down from the top (tsp)
  extern __shared__ U2 smem_stack[];
  U2 local_stack[SMEM ? 1 : RB_STACK_MAX];
  typename std::conditional<SMEM, SmemStack, U2*>::type stack;
  if constexpr (SMEM)
    stack = SmemStack{smem_stack + threadIdx.x, kTraceThreads};
  else
    stack = local_stack;
  const int STACK_TOP = SMEM ? stack_n : RB_STACK_MAX;
  int tsp = STACK_TOP;
  U2 tg = U2{0u, 0u};  // triangle group currently being consumed
  bool active = false;
  bool exhausted = false;  // warp-uniform
  uint32_t dest = 0;
  while (true) {
    // ---- refill idle lanes ---------------------------------------------------------------------
    const unsigned idle = __ballot_sync(FULL, !active);
    if (!exhausted && idle != 0) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(next, (uint32_t)__popc(idle));
      base = __shfl_sync(FULL, base, 0);
      if (base + __popc(idle) >= count) exhausted = true;
      if (!active) {
        const uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
        if (i < count) {
          V3 o, d;
          float tn, tf;
          io.fetch(i, &o, &d, &tn, &tf, &dest);
          active = trav_init(T, sc, o, d, tn, tf);
          if (TIE) T.tie_id = io.tie_id(dest);
          tsp = STACK_TOP;
          tg = U2{0u, 0u};
          if (!active) {  // cannot hit anything
            if (ANY)
              io.put_any(dest, false);
            else
              io.put_closest(sc, dest, T);
          }
        }
      }
    }
    if (__ballot_sync(FULL, active) == 0) {
      if (exhausted) break;
      continue;
    }
    while (true) {
      bool done = false;
      // ---- NODE phase ------------------------------------------------------------------------------
      if (active && has_node_work(T)) {
        U2 g = trav_node_step<ANY>(T, stack, sc);
        if (g.y != 0) {
          if (tg.y == 0)
            tg = g;
          else if (tsp - T.sp > 2)
            stack[--tsp] = g;
          else  // stack almost full: drain this group right away
            while (g.y != 0 && !done) done = trav_tri_one<ANY, TIE>(T, g, sc);
        }
        if (!has_node_work(T) && T.sp > 0) T.ngroup = stack[--T.sp];
      }
      // ---- TRIANGLE phase, when it is worth a warp-wide pass ---------------------------------------------
      const bool has_tri = active && !done && tg.y != 0;
      const int n_tri = __popc(__ballot_sync(FULL, has_tri));
      const int n_node = __popc(__ballot_sync(FULL, active && !done && has_node_work(T)));
      if (n_tri != 0 && (n_tri >= tri_lanes || n_node < tri_lanes)) {
        if (has_tri) {
          done = trav_tri_one<ANY, TIE>(T, tg, sc);
          if (!done && tg.y == 0 && tsp < STACK_TOP) tg = stack[tsp++];
        }
      }
      // ---- finished rays ------------------------------------------------------------------------------------
      if (active && (done || (!has_node_work(T) && tg.y == 0))) {
        if (ANY)
          io.put_any(dest, T.hit_any);
        else
          io.put_closest(sc, dest, T);
        active = false;
      }
      const int busy = __popc(__ballot_sync(FULL, active));
      if (busy == 0 || (busy < refill_lanes && !exhausted)) break;
    }
  }
}
__global__ void k_reset_queue(uint32_t* count, uint32_t* next) {
  *count = 0;
  *next = 0;
}

// ---- band halos over peer memory --------------------------------------------------------------
// One process per GPU; the neighbours' reservoir planes are mapped into this process with CUDA IPC (NVLink /
// NVSwitch peer memory). k_halo_push stores this band's boundary rows of the four reservoir planes straight into
// the halo rows of the rank above and the rank below (same row offsets on every rank) and then publishes a
// stamp in the peer's flag word: every thread fences its peer stores system-wide, the last block to finish
// (device-scope counter) releases the flags. The consumer's k_halo_wait polls its own flag words (they live in
// its HBM, the peer's release store lands in its L2) before the kernels that read the halo rows are allowed to
// start. No rendezvous and no host involvement: a rank waits only when the neighbour's rows are really late.
struct HaloPush {
  const char* src[4];                // this rank's planes of R[rWrite]
  char* dst[2][4];                   // the same planes of the rank above [0] / below [1] (peer mappings), null = none
  unsigned long long off[2][4];      // byte offset of the rows that go up / down
  unsigned long long bytes[2][4];
  uint32_t* flag[2];                 // peer flag words: above's "from below", below's "from above"
  uint32_t* done;                    // local block counter (returns to zero)
  uint32_t stamp;
};
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(256) k_halo_push(HaloPush p) {
  const int plane = blockIdx.y & 3, dir = blockIdx.y >> 2;
  char* dst = p.dst[dir][plane];
  if (dst != nullptr) {
    const size_t off = p.off[dir][plane], bytes = p.bytes[dir][plane];
    const size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    if (plane < 3) {  // 16-byte records
      const uint4* s = reinterpret_cast<const uint4*>(p.src[plane] + off);
      uint4* d = reinterpret_cast<uint4*>(dst + off);
      for (size_t k = first; k < bytes / 16; k += step) d[k] = s[k];
    } else {  // light indices: 4-byte records (a row need not be a multiple of 16 bytes)
      const uint32_t* s = reinterpret_cast<const uint32_t*>(p.src[plane] + off);
      uint32_t* d = reinterpret_cast<uint32_t*>(dst + off);
      for (size_t k = first; k < bytes / 4; k += step) d[k] = s[k];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    if (atomicAdd(p.done, 1u) == total - 1u) {
      *p.done = 0u;
      __threadfence_system();
      if (p.flag[0]) st_release_sys(p.flag[0], p.stamp);
      if (p.flag[1]) st_release_sys(p.flag[1], p.stamp);
    }
  }
}
// flags[0]: stamp of the last halo pushed by the rank above, flags[1]: by the rank below. Gives up after
// timeout_ns (a dead neighbour must not hang the GPU) and reports through *err (mapped host memory).
__global__ void k_halo_wait(const uint32_t* flags, uint32_t stamp, int need_above, int need_below, uint32_t* err,
                            unsigned long long timeout_ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int s = 0; s < 2; ++s) {
    if (!(s == 0 ? need_above : need_below)) continue;
    while ((int32_t)(ld_acquire_sys(flags + s) - stamp) < 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) {
        *err = stamp ? stamp : 1u;
        return;
      }
    }
  }
}

// ---- BVH build -----------------------------------------------------------------------
__global__ void k_bounds(BuildCtx c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c.n) bounds_body(c, i);
}
__global__ void k_morton(BuildCtx c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c.n) morton_body(c, i);
}
__global__ void k_karras(BuildCtx c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < c.n) karras_body(c, i);
}
__global__ void k_fit(BuildCtx c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c.n) fit_body(c, i);
}
__global__ void k_collapse(BuildCtx c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (uint32_t)c.q_in_len) collapse_body(c, i);
}
__global__ void k_tiny_root(BuildCtx c) {
  if (blockIdx.x == 0 && threadIdx.x == 0) tiny_root_body(c);
}

// implicit FMA contraction must be off for CPU/GPU parity: a*b+c with a single rounding differs here
__global__ void k_selfcheck(float a, float b, float c, float* out) { out[0] = a * b + c; }

}  // namespace

// =====================================================================================
// context
// =====================================================================================
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct FrameMark {
  int pass, kind;  // kind 0 = streaming kernel, 1 = traversal kernel
};
struct FrameState {
  FrameCtx fc{};   // back half (and everything host code reads back)
  FrameCtx ffc{};  // front half
  RbParams P{};
  bool open = false, timed = false, wave = false, wave_spatial = false;
  bool wave_spatial_staged = false;  // non-constant spatial MIS mode in the wavefront schedule (StagedVis)
  bool shaded = false;  // the last spatial pass's resolve kernel has written frame_data already
  uint32_t launches = 0, frame_idx = 0;
  std::vector<FrameMark> marks;
};

struct RbContext {
  RbCreateInfo info{};
  RbParams params{};
  std::string err;
  cudaStream_t stream = nullptr;
  std::vector<void*> allocs;

  // frame state
  // Frame pipeline: a frame is a FRONT half (G-buffer, initial candidates incl. the BRDF-candidate rays; depends on the
  // camera only) and a BACK half (visibility trace, temporal and spatial reuse, shading; depends on the previous
  // frame's reservoirs). The front half of frame n+1 is issued on its own stream and overlaps the back half of frame
  // n, which fills the tails and ramps at the back half's many kernel boundaries and the waits for neighbour bands.
  // That takes one more G-buffer (3: being written / current / previous) and one more reservoir buffer (4: the front
  // half's output + the three the back half rotates), and separate ray queues for the two halves.
  GBufPlanes G[3]{};
  int gCur = 2;  // G-buffer of the frame rendered last (the next frame takes (gCur + 1) % 3)
  ResPlanes R[4]{};
  int rRead = 0, rWrite = 1, rLast = 2, rFree = 3;
  cudaStream_t fstream = nullptr;  // front halves (lower priority than `stream`)
  cudaEvent_t evFrontDone = nullptr, evBackDone[2]{};
  bool backRecorded[2]{};
  bool frontRecorded = false;
  uint32_t frameSeq = 0;
  bool overlap = true;  // RB_OVERLAP=0: both halves on `stream`
  bool maxFramesInFlight = true;  // RB_FRAMES_IN_FLIGHT=0: the host may run ahead without bound
  // RB_OVERLAP_DEBUG=1: timestamps of the halves of the first frames, printed by rb_destroy
  bool ovDebug = false;
  static constexpr int kOvFrames = 24;
  cudaEvent_t evOv[kOvFrames][4]{};  // front begin / end, back begin / end
  WaveBufs fwave{};     // front half: BRDF-candidate rays (+ chained "precedes" rays), their hits / occlusion bytes
  RayQ* visRays[2]{};   // visibility rays of frame parity p: queued by the front half, traced by the back half
  size_t fwaveRayCap = 0, fwaveHitCap = 0, visRayCap = 0;
  unsigned long long* counters2 = nullptr;  // [2][8]: per frame parity {closest, any as written, any traced, -, deferred
                                            // count, chain queue pair, BRDF queue pair, visibility queue pair}
  float* frame = nullptr;
  // rb_render_frame_async: the frame's rows go to the host on a copy stream while the next frame renders. The kernel
  // that writes frame_data of the NEXT frame (the last resolve / k_shade) waits for the copy; nothing else does.
  cudaStream_t copyStream = nullptr;
  static constexpr int kCopyRing = 4;
  cudaEvent_t evCopyDone[kCopyRing]{}, evFrameReady = nullptr;
  uint32_t asyncSeq = 0;       // async frames issued so far
  bool copyPending = false;    // the latest async copy may still be reading h->frame
  float* accumulator = nullptr;  // N1: running mean of frame_data
  F4* display = nullptr;         // N1: tonemapped / gamma-compressed display_data
  double* statSums = nullptr;
  unsigned long long* counters = nullptr;
  CamState prevCam{};
  bool havePrev = false;
  int prevGy0 = 0, prevGy1 = 0;  // G-buffer rows held for the previous frame
  FrameState fs;

  // multi-GPU bands: NCCL communicator (loaded with dlopen, see rb_comm_init), comm stream and events
  void* comm = nullptr;
  int commRank = 0, commSize = 1;
  cudaStream_t commStream = nullptr;
  cudaEvent_t evHaloReady = nullptr, evHaloDone = nullptr, evHaloT0 = nullptr;  // T0..Done time the exchange itself
  bool haloTimed = false;
  // halo rows over peer memory (k_halo_push / k_halo_wait): the neighbours' reservoir planes and flag words mapped with
  // CUDA IPC in rb_comm_init; RB_HALO=nccl (or a failed mapping on any rank) keeps the grouped ncclSend/ncclRecv path
  bool p2p = false;
  // member of an in-process multi-device group (rb_multi_create): the neighbours' planes are plain peer pointers
  // (cudaDeviceEnablePeerAccess), no communicator, no IPC mappings; static bands (no balancer)
  bool localGroup = false;
  void* peerBase[2][17]{};      // cudaIpcOpenMemHandle results, [above, below][16 planes + flags]
  ResPlanes peerR[2][4]{};
  uint32_t* peerFlags[2]{};
  uint32_t* haloFlags = nullptr;  // [0] stamp pushed by the rank above, [1] by the rank below, [2] push block counter
  uint32_t* haloErr = nullptr;    // mapped pinned host word: a wait that timed out leaves its stamp here
  uint32_t haloSeq = 0;
  // load balancing of the bands (rb_comm_init): every frame end the last-frame reservoirs of kShipRows rows on both
  // sides of each boundary and the ranks' frame costs go to the neighbours; boundaries follow the cost difference
  static constexpr int kShipRows = 16, kBalRing = 8, kMaxStalls = 8;  // a boundary moves <= 15 rows per period (G-buffer margin: 16)
  bool balance = true;
  int balPeriod = 8;                             // ship + decide every balPeriod frames (RB_BAL_PERIOD)
  bool balDebug = false;                         // RB_BAL_DEBUG: print costs and bands to stderr
  uint32_t balFrame = 0;                         // frames rendered since rb_comm_init
  cudaEvent_t evShipReady = nullptr, evShipDone = nullptr, evAccMoved = nullptr;
  cudaEvent_t evBalCopied[kBalRing]{};           // costs of ring slot arrived in balHost
  cudaEvent_t evFrameB[kBalRing]{}, evFrameE[kBalRing]{};
  cudaEvent_t evStallA[kBalRing][kMaxStalls]{}, evStallB[kBalRing][kMaxStalls]{};
  int nStalls[kBalRing]{};
  static constexpr int kMaxRanks = 64;
  float* balDev = nullptr;                       // [2 + 2 * kMaxRanks]: this rank's {cost, rows}, then the all-gathered pairs
  float* balHost = nullptr;                      // pinned, [kBalRing][2 + 2 * kMaxRanks]: the same, per ring slot
  bool shipPending = false;

  // scene
  bool haveScene = false;
  bool hasDielectric = false;  // a material the N2 estimator does not cover
  F4* sceneUv = nullptr;       // per-triangle texture coordinates (owned by sceneAllocs), null = none uploaded
  F4* sceneTan = nullptr;      // per-triangle tangents (owned by sceneAllocs), null = none uploaded
  uint32_t nMaterials = 0;
  std::vector<void*> texAllocs;  // rb_set_textures
  std::vector<void*> skyAllocs;  // rb_set_sky
  SceneDev sc{};
  std::vector<void*> sceneAllocs;
  RbSceneStats stats{};

  cudaEvent_t ev[16]{};
  static constexpr int kMaxEvents = 96;
  cudaEvent_t fev[kMaxEvents]{};  // per-kernel frame events
  bool evCreated = false;
  int numSMs = 148;
  // traversal tuning (overridable for experiments: RB_REFILL, RB_POSTPONE, RB_TRACE_BLOCKS)
  int refillLanes = 26, postponeLanes = 8, traceBlocksPerSM = 8;
  // RB_SMEM_STACK=1: per-lane traversal stack in shared memory. Measured (profiles/r2_experiments.md): the traversal
  // launches take the same time with either stack (the L1-resident local stack's latency is hidden), but 24 KB per CTA
  // crowd the other stream's kernels out of the SMs (frame pipeline: 126.0 instead of 129.3 fps) — off by default
  bool smemStack = false;
  int smemStackEntries = 0;  // per-lane stack entries in shared memory for the uploaded scene (0 = local-memory stack)
  bool twoStepBrdf = true;  // RB_TWO_STEP_BRDF=0 traces the BRDF-candidate rays against the full BVH instead

  // wavefront buffers
  WaveBufs wave{};
  size_t waveRayCap = 0, waveOccCap = 0, waveCandCap = 0;
  uint32_t* waveCounters = nullptr;
  int wavePair = 0;
};

static thread_local std::string g_create_error;

#define RB_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
      return e__ == cudaErrorMemoryAllocation ? RB_ERR_OUT_OF_MEMORY : RB_ERR_CUDA;        \
    }                                                                                      \
  } while (0)

template <class T>
static int dev_alloc(RbContext* h, T** out, size_t count, std::vector<void*>& owner) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  RB_CUDA(cudaMalloc(&p, bytes));
  owner.push_back(p);
  *out = (T*)p;
  return RB_OK;
}
#define RB_TRY(expr)            \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != RB_OK) return rc__; \
  } while (0)

static void free_list(std::vector<void*>& l) {
  for (void* p : l) cudaFree(p);
  l.clear();
}

static int spatial_reach(const RbParams& P) { return (int)sqrtf(std::max(P.spatialReuseRadius, 0.0f)) + 1; }

// ---- NCCL, loaded at run time so that the library has no link-time dependency on it ------------------------
#include <dlfcn.h>
namespace {
struct NcclId128 {
  char b[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId128 /*ncclUniqueId by value*/, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
bool load_nccl(std::string& err) {
  if (g_nccl.lib) return true;
  // torch's bundled libnccl.so.2 is reused when it is already in the process; otherwise the system one
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    err = std::string("rb_comm: cannot load libnccl: ") + dlerror();
    return false;
  }
  auto sym = [&](const char* n) { return dlsym(lib, n); };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
  g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart ||
      !g_nccl.GroupEnd || !g_nccl.AllReduce || !g_nccl.AllGather) {  // the balancer and the transport agreement need the collectives
    err = "rb_comm: libnccl lacks a required symbol";
    return false;
  }
  g_nccl.lib = lib;
  return true;
}
}  // namespace

#define RB_NCCL(call)                                                                                   \
  do {                                                                                                  \
    int r__ = (call);                                                                                   \
    if (r__ != 0) {                                                                                     \
      h->err = std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?");  \
      return RB_ERR_COMM;                                                                               \
    }                                                                                                   \
  } while (0)

// Exchange the halo rows of R[rWrite] with rank-1 / rank+1: one grouped send/recv pair per neighbour and plane
// (rows of a plane are contiguous and live at the same offsets on every rank). Runs on the comm stream after the
// producing kernels; frame_spatial streams the interior rows meanwhile.
static int halo_push_p2p(RbContext* h);
static int halo_wait_p2p(RbContext* h);
static int halo_exchange_begin(RbContext* h) {
  if (!h->comm && !h->localGroup) return RB_OK;
  if (h->p2p) return halo_push_p2p(h);
  const int y0 = h->info.band_y0, y1 = h->info.band_y1, W = h->info.width, H = h->info.height;
  const int R = spatial_reach(h->fs.P);
  const ResPlanes& P = h->R[h->rWrite];
  RB_CUDA(cudaEventRecord(h->evHaloReady, h->stream));
  RB_CUDA(cudaStreamWaitEvent(h->commStream, h->evHaloReady, 0));
  RB_CUDA(cudaEventRecord(h->evHaloT0, h->commStream));
  h->haloTimed = true;
  char* planes[4] = {(char*)P.point_wsum, (char*)P.normal_W, (char*)P.Li_conf, (char*)P.light_idx};
  const size_t esz[4] = {16, 16, 16, 4};
  RB_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < 4; ++i) {
    if (h->commRank > 0) {  // neighbour above owns rows < y0
      const int r = std::min(R, y1 - y0), ru = std::min(R, y0);
      RB_NCCL(g_nccl.Send(planes[i] + (size_t)y0 * W * esz[i], (size_t)r * W * esz[i], /*ncclChar*/ 0, h->commRank - 1, h->comm,
                          h->commStream));
      RB_NCCL(g_nccl.Recv(planes[i] + (size_t)(y0 - ru) * W * esz[i], (size_t)ru * W * esz[i], 0, h->commRank - 1, h->comm,
                          h->commStream));
    }
    if (h->commRank + 1 < h->commSize) {  // neighbour below owns rows >= y1
      const int r = std::min(R, y1 - y0), rd = std::min(R, H - y1);
      RB_NCCL(g_nccl.Send(planes[i] + (size_t)(y1 - r) * W * esz[i], (size_t)r * W * esz[i], 0, h->commRank + 1, h->comm,
                          h->commStream));
      RB_NCCL(g_nccl.Recv(planes[i] + (size_t)y1 * W * esz[i], (size_t)rd * W * esz[i], 0, h->commRank + 1, h->comm, h->commStream));
    }
  }
  RB_NCCL(g_nccl.GroupEnd());
  RB_CUDA(cudaEventRecord(h->evHaloDone, h->commStream));
  return RB_OK;
}
// main stream waits for an event of the comm stream; the wait is bracketed by two events so that the frame's cost
// (what the balancer equalises) can leave the time spent waiting for a slower neighbour out
static int wait_recording_stall(RbContext* h, cudaEvent_t ev) {
  const int slot = (int)(h->balFrame % RbContext::kBalRing);
  const bool rec = h->balance && h->nStalls[slot] < RbContext::kMaxStalls;
  if (rec) RB_CUDA(cudaEventRecord(h->evStallA[slot][h->nStalls[slot]], h->stream));
  RB_CUDA(cudaStreamWaitEvent(h->stream, ev, 0));
  if (rec) RB_CUDA(cudaEventRecord(h->evStallB[slot][h->nStalls[slot]++], h->stream));
  return RB_OK;
}
static int halo_exchange_wait(RbContext* h) {
  if (!h->comm && !h->localGroup) return RB_OK;
  if (h->p2p) return halo_wait_p2p(h);
  return wait_recording_stall(h, h->evHaloDone);
}

// Peer-memory halo exchange: push this band's boundary rows of R[rWrite] into the neighbours' halo rows (main stream,
// right behind the kernel that produced them) ...
static int halo_push_p2p(RbContext* h) {
  const int y0 = h->info.band_y0, y1 = h->info.band_y1, W = h->info.width;
  const int R = spatial_reach(h->fs.P);
  const int r = std::min(R, y1 - y0);
  const ResPlanes& P = h->R[h->rWrite];
  HaloPush a{};
  const char* planes[4] = {(const char*)P.point_wsum, (const char*)P.normal_W, (const char*)P.Li_conf, (const char*)P.light_idx};
  const size_t esz[4] = {16, 16, 16, 4};
  const bool has[2] = {h->commRank > 0, h->commRank + 1 < h->commSize};
  const int first_row[2] = {y0, y1 - r};  // rows that go up / down
  for (int d = 0; d < 2; ++d) {
    const ResPlanes& Q = h->peerR[d][h->rWrite];
    char* dst[4] = {(char*)Q.point_wsum, (char*)Q.normal_W, (char*)Q.Li_conf, (char*)Q.light_idx};
    for (int i = 0; i < 4; ++i) {
      a.dst[d][i] = has[d] ? dst[i] : nullptr;
      a.off[d][i] = (unsigned long long)first_row[d] * W * esz[i];
      a.bytes[d][i] = (unsigned long long)r * W * esz[i];
    }
  }
  for (int i = 0; i < 4; ++i) a.src[i] = planes[i];
  a.flag[0] = has[0] ? h->peerFlags[0] + 1 : nullptr;  // the rank above sees me as "below"
  a.flag[1] = has[1] ? h->peerFlags[1] + 0 : nullptr;
  a.done = h->haloFlags + 2;
  a.stamp = ++h->haloSeq;
  const unsigned bx = (unsigned)std::max<size_t>(1, std::min<size_t>(32, ((size_t)r * W + 1023) / 1024));
  k_halo_push<<<dim3(bx, 8), 256, 0, h->stream>>>(a);
  h->fs.launches++;
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}
// ... and, after the interior rows have been streamed, hold the main stream until both neighbours' rows are in.
// A k_halo_wait that gave up leaves its stamp in mapped host memory. Checked wherever the host has just synchronised
// with the frame's stream (and before the next exchange): the frame whose boundary rows read stale halo rows is reported
// as failed, and the word is cleared so that a transient stall does not poison the handle for good.
static int halo_check(RbContext* h) {
  if (!h->haloErr) return RB_OK;
  const uint32_t stamp = *(volatile uint32_t*)h->haloErr;
  if (stamp == 0u) return RB_OK;
  *(volatile uint32_t*)h->haloErr = 0u;
  h->err = "band halo exchange: a neighbour's rows did not arrive in time (exchange " + std::to_string(stamp) +
           "); the frame's boundary rows are invalid";
  return RB_ERR_COMM;
}
static int halo_wait_p2p(RbContext* h) {
  RB_TRY(halo_check(h));
  const int slot = (int)(h->balFrame % RbContext::kBalRing);
  const bool rec = h->balance && h->nStalls[slot] < RbContext::kMaxStalls;
  const bool timed = h->fs.timed;
  if (rec) RB_CUDA(cudaEventRecord(h->evStallA[slot][h->nStalls[slot]], h->stream));
  if (timed) RB_CUDA(cudaEventRecord(h->evHaloT0, h->stream));
  // RB_HALO_TIMEOUT_MS (default 5 s); the first exchanges wait 12 x longer: ranks allocate their wavefront buffers and
  // NCCL sets its channels up during the first frames
  static const double timeout_ms = getenv("RB_HALO_TIMEOUT_MS") ? atof(getenv("RB_HALO_TIMEOUT_MS")) : 5000.0;
  const unsigned long long timeout_ns = (unsigned long long)(timeout_ms * (h->haloSeq <= 8u ? 12.0 : 1.0) * 1.0e6);
  k_halo_wait<<<1, 1, 0, h->stream>>>(h->haloFlags, h->haloSeq, h->commRank > 0 ? 1 : 0, h->commRank + 1 < h->commSize ? 1 : 0,
                                      h->haloErr, timeout_ns);
  h->fs.launches++;
  if (timed) {
    RB_CUDA(cudaEventRecord(h->evHaloDone, h->stream));
    h->haloTimed = true;
  }
  if (rec) RB_CUDA(cudaEventRecord(h->evStallB[slot][h->nStalls[slot]++], h->stream));
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

// End of frame: ship the final reservoirs (R[rLast]) of kShipRows rows on both sides of each boundary plus this
// rank's cost of an earlier frame to the neighbours (one NCCL group on the comm stream, overlapped with the next
// frame's G-buffer / initial / visibility passes).
static int balance_ship_begin(RbContext* h) {
  if (!h->comm || !h->balance) return RB_OK;
  // once per balPeriod frames: the boundary rows' reservoirs must be those of the frame right before the move, so the
  // exchange happens at the end of the period's last frame and is applied at the start of the next one
  if (h->balFrame % (uint32_t)h->balPeriod != (uint32_t)h->balPeriod - 1u) return RB_OK;
  const int y0 = h->info.band_y0, y1 = h->info.band_y1, W = h->info.width, H = h->info.height;
  const int S = RbContext::kShipRows;
  const int slot = (int)(h->balFrame % RbContext::kBalRing);
  // my cost: GPU time of the frame before last, minus its waits (its events are complete or nearly so; this also keeps
  // the host at most two frames ahead of the device)
  float cost = 0.0f;
  if (h->balFrame >= 2) {
    RB_CUDA(cudaEventSynchronize(h->evFrameE[(h->balFrame - 2) % RbContext::kBalRing]));
    // mean over up to four completed frames of this period (same band): single frames are noisy. A frame's cost is
    // the time between the ends of two consecutive back halves (front halves overlap the previous frame, so start-to-end
    // times would count shared time twice) minus what its main stream spent waiting for neighbour bands.
    int used = 0;
    for (uint32_t j = 2; j <= 5 && j + 1 <= h->balFrame; ++j) {
      if (j > 2 && (int)j > h->balPeriod - 1) break;
      const int ps = (int)((h->balFrame - j) % RbContext::kBalRing), pp = (int)((h->balFrame - j - 1) % RbContext::kBalRing);
      float t = 0;
      RB_CUDA(cudaEventElapsedTime(&t, h->evFrameE[pp], h->evFrameE[ps]));
      for (int i = 0; i < h->nStalls[ps]; ++i) {
        float st = 0;
        RB_CUDA(cudaEventElapsedTime(&st, h->evStallA[ps][i], h->evStallB[ps][i]));
        t -= st;
      }
      cost += t;
      ++used;
    }
    cost = used ? cost / (float)used : 0.0f;
    if (!(cost > 0.0f)) cost = 0.0f;
  }
  constexpr int BS = 2 + 2 * RbContext::kMaxRanks;
  float* hostSlot = h->balHost + BS * slot;
  hostSlot[0] = cost, hostSlot[1] = (float)(y1 - y0);
  for (int i = 2; i < BS; ++i) hostSlot[i] = -1.0f;
  const ResPlanes& P = h->R[h->rLast];
  RB_CUDA(cudaEventRecord(h->evShipReady, h->stream));
  RB_CUDA(cudaStreamWaitEvent(h->commStream, h->evShipReady, 0));
  RB_CUDA(cudaMemcpyAsync(h->balDev, hostSlot, 8, cudaMemcpyHostToDevice, h->commStream));
  char* planes[4] = {(char*)P.point_wsum, (char*)P.normal_W, (char*)P.Li_conf, (char*)P.light_idx};
  const size_t esz[4] = {16, 16, 16, 4};
  RB_NCCL(g_nccl.GroupStart());
  if (h->commRank > 0) {
    const int r = std::min(S, y1 - y0), ru = std::min(S, y0);
    for (int i = 0; i < 4; ++i) {
      RB_NCCL(g_nccl.Send(planes[i] + (size_t)y0 * W * esz[i], (size_t)r * W * esz[i], 0, h->commRank - 1, h->comm, h->commStream));
      RB_NCCL(g_nccl.Recv(planes[i] + (size_t)(y0 - ru) * W * esz[i], (size_t)ru * W * esz[i], 0, h->commRank - 1, h->comm, h->commStream));
    }
  }
  if (h->commRank + 1 < h->commSize) {
    const int r = std::min(S, y1 - y0), rd = std::min(S, H - y1);
    for (int i = 0; i < 4; ++i) {
      RB_NCCL(g_nccl.Send(planes[i] + (size_t)(y1 - r) * W * esz[i], (size_t)r * W * esz[i], 0, h->commRank + 1, h->comm, h->commStream));
      RB_NCCL(g_nccl.Recv(planes[i] + (size_t)y1 * W * esz[i], (size_t)rd * W * esz[i], 0, h->commRank + 1, h->comm, h->commStream));
    }
  }
  RB_NCCL(g_nccl.GroupEnd());
  // every rank's {cost, rows}: all ranks then derive the same new boundaries from the same numbers
  RB_NCCL(g_nccl.AllGather(h->balDev, h->balDev + 2, 2, /*ncclFloat32*/ 7, h->comm, h->commStream));
  RB_CUDA(cudaMemcpyAsync(hostSlot + 2, h->balDev + 2, 2 * h->commSize * sizeof(float), cudaMemcpyDeviceToHost, h->commStream));
  RB_CUDA(cudaEventRecord(h->evBalCopied[slot], h->commStream));
  RB_CUDA(cudaEventRecord(h->evShipDone, h->commStream));
  h->shipPending = true;
  return RB_OK;
}

// Start of frame: move the band boundaries. Every rank holds the same all-gathered {cost, rows} pairs (exchanged at the
// previous frame end) and runs the same arithmetic on them, so all ranks agree on every boundary without another
// message. The target partition gives every band the same share of the summed cost, taking the cost per row as uniform
// inside each current band; a boundary moves half of the way towards its target, at most kShipRows - 1 rows per period:
// the rows a rank gains had their last-frame reservoirs shipped at the previous frame end and their previous G-buffer
// rendered as margin. (A global rule instead of pairwise diffusion: a chain of N bands would need O(N^2) periods.)
static void balance_targets(const float* pairs, int n, int H, int* bounds /* n + 1, in: current, out: new */) {
  const int lim = RbContext::kShipRows - 1, min_rows = 2 * RbContext::kShipRows;
  double total = 0.0;
  for (int r = 0; r < n; ++r) {
    const float cost = pairs[2 * r], rows = pairs[2 * r + 1];
    if (!(cost > 0.0f) || !(rows > 0.0f) || (int)rows != bounds[r + 1] - bounds[r]) return;  // not a consistent snapshot
    if (bounds[r + 1] - bounds[r] < min_rows + 2 * lim) return;                              // bands too thin to move safely
    total += (double)cost;
  }
  int nb[RbContext::kMaxRanks + 1];
  nb[0] = 0, nb[n] = H;
  int band = 0;
  double before = 0.0;  // cost of the bands above `band`
  for (int k = 1; k < n; ++k) {
    const double want = total * (double)k / (double)n;
    while (band < n - 1 && before + (double)pairs[2 * band] < want) before += (double)pairs[2 * band++];
    const double frac = (want - before) / (double)pairs[2 * band];
    const double ideal = (double)bounds[band] + frac * (double)(bounds[band + 1] - bounds[band]);
    int step = (int)lrint(0.5 * (ideal - (double)bounds[k]));
    step = std::max(-lim, std::min(lim, step));
    nb[k] = bounds[k] + step;
  }
  for (int k = 1; k <= n; ++k)  // (cannot trigger with the thickness test above; kept as a guard)
    if (nb[k] - nb[k - 1] < min_rows) return;
  for (int k = 1; k < n; ++k) bounds[k] = nb[k];
}
static int balance_update_band(RbContext* h) {
  if (!h->comm || !h->balance || h->balFrame < 4) return RB_OK;
  // the exchange made at the end of the previous frame (the last of its period) is applied now; waiting for it drains
  // the host/device pipeline once per period
  if (h->balFrame % (uint32_t)h->balPeriod != 0) return RB_OK;
  const int slot = (int)((h->balFrame - 1) % RbContext::kBalRing);
  RB_CUDA(cudaEventSynchronize(h->evBalCopied[slot]));
  constexpr int BS = 2 + 2 * RbContext::kMaxRanks;
  const float* c = h->balHost + BS * slot;  // own {cost, rows} as sent, then all ranks' pairs
  const int n = h->commSize;
  // current boundaries, reconstructed from the gathered heights (the ranks' bands tile the image top to bottom)
  int bounds[RbContext::kMaxRanks + 1];
  bounds[0] = 0;
  bool ok = true;
  for (int r = 0; r < n; ++r) {
    const float rows = c[2 + 2 * r + 1];
    if (!(rows > 0.0f)) ok = false;
    bounds[r + 1] = bounds[r] + (ok ? (int)rows : 0);
  }
  if (ok && bounds[n] == h->info.height && bounds[h->commRank] == h->info.band_y0 && bounds[h->commRank + 1] == h->info.band_y1) {
    balance_targets(c + 2, n, h->info.height, bounds);
    // Rows that change owner take their ACCUMULATOR along (rb_accumulate_display keeps the running mean per band):
    // every rank derives the same new boundaries, so both sides of a boundary know which rows move and in which
    // direction. One grouped send / recv on the comm stream, ordered after whatever the main stream did to the
    // accumulator so far; the main stream continues once the rows are in (they are <= 15 rows, once per period).
    const int oy0 = h->info.band_y0, oy1 = h->info.band_y1, ny0 = bounds[h->commRank], ny1 = bounds[h->commRank + 1];
    if (ny0 != oy0 || ny1 != oy1) {
      const size_t row = (size_t)h->info.width * 3;
      RB_CUDA(cudaEventRecord(h->evShipReady, h->stream));
      RB_CUDA(cudaStreamWaitEvent(h->commStream, h->evShipReady, 0));
      RB_NCCL(g_nccl.GroupStart());
      if (ny0 > oy0)  // rows [oy0, ny0) go to the rank above
        RB_NCCL(g_nccl.Send(h->accumulator + row * oy0, row * (size_t)(ny0 - oy0), /*ncclFloat32*/ 7, h->commRank - 1, h->comm, h->commStream));
      if (ny0 < oy0)  // rows [ny0, oy0) come from the rank above
        RB_NCCL(g_nccl.Recv(h->accumulator + row * ny0, row * (size_t)(oy0 - ny0), 7, h->commRank - 1, h->comm, h->commStream));
      if (ny1 < oy1)  // rows [ny1, oy1) go to the rank below
        RB_NCCL(g_nccl.Send(h->accumulator + row * ny1, row * (size_t)(oy1 - ny1), 7, h->commRank + 1, h->comm, h->commStream));
      if (ny1 > oy1)  // rows [oy1, ny1) come from the rank below
        RB_NCCL(g_nccl.Recv(h->accumulator + row * oy1, row * (size_t)(ny1 - oy1), 7, h->commRank + 1, h->comm, h->commStream));
      RB_NCCL(g_nccl.GroupEnd());
      RB_CUDA(cudaEventRecord(h->evAccMoved, h->commStream));
      RB_CUDA(cudaStreamWaitEvent(h->stream, h->evAccMoved, 0));
    }
    h->info.band_y0 = ny0;
    h->info.band_y1 = ny1;
  }
  if (h->balDebug && h->balFrame % 8 == 0) {
    std::string all;
    for (int r = 0; r < n; ++r) all += " " + std::to_string(c[2 + 2 * r]).substr(0, 5) + "/" + std::to_string((int)c[2 + 2 * r + 1]);
    fprintf(stderr, "[rb balance] rank %d frame %u: cost/rows%s -> band [%d, %d)\n", h->commRank, h->balFrame, all.c_str(),
            h->info.band_y0, h->info.band_y1);
  }
  return RB_OK;
}

// Map the neighbours' reservoir planes and flag words into this process (CUDA IPC handles travel through the NCCL
// communicator that is already up). Every rank takes part in the exchange and in the agreement, whatever its own
// outcome: the peer-memory path is used only when it works on ALL ranks.
static int halo_p2p_setup(RbContext* h) {
  const char* mode = getenv("RB_HALO");
  const bool want = !(mode && strcmp(mode, "nccl") == 0);
  RB_CUDA(cudaMalloc((void**)&h->haloFlags, 4 * sizeof(uint32_t)));
  RB_CUDA(cudaMemset(h->haloFlags, 0, 4 * sizeof(uint32_t)));
  RB_CUDA(cudaHostAlloc((void**)&h->haloErr, sizeof(uint32_t), cudaHostAllocMapped));
  *h->haloErr = 0u;
  h->haloSeq = 0;
  constexpr int NH = 17;
  cudaIpcMemHandle_t mine[NH], theirs[2][NH];
  memset(mine, 0, sizeof(mine));
  memset(theirs, 0, sizeof(theirs));
  void* ptrs[NH];
  for (int b = 0; b < 4; ++b) {
    ptrs[4 * b + 0] = h->R[b].point_wsum, ptrs[4 * b + 1] = h->R[b].normal_W, ptrs[4 * b + 2] = h->R[b].Li_conf,
                 ptrs[4 * b + 3] = h->R[b].light_idx;
  }
  ptrs[16] = h->haloFlags;
  int ok = want && g_nccl.AllReduce ? 1 : 0;
  for (int i = 0; i < NH && ok; ++i)
    if (cudaIpcGetMemHandle(&mine[i], ptrs[i]) != cudaSuccess) ok = 0;
  (void)cudaGetLastError();
  const bool has[2] = {h->commRank > 0, h->commRank + 1 < h->commSize};
  const int peer[2] = {h->commRank - 1, h->commRank + 1};
  char* stage = nullptr;
  const size_t HB = sizeof(mine);
  RB_CUDA(cudaMalloc((void**)&stage, 3 * HB + 16));
  struct StageGuard {  // freed on every path out of this function
    char*& p;
    ~StageGuard() {
      if (p) cudaFree(p);
      p = nullptr;
    }
  } stage_guard{stage};
  RB_CUDA(cudaMemcpyAsync(stage, mine, HB, cudaMemcpyHostToDevice, h->commStream));
  RB_NCCL(g_nccl.GroupStart());
  for (int d = 0; d < 2; ++d)
    if (has[d]) {
      RB_NCCL(g_nccl.Send(stage, HB, 0, peer[d], h->comm, h->commStream));
      RB_NCCL(g_nccl.Recv(stage + (1 + d) * HB, HB, 0, peer[d], h->comm, h->commStream));
    }
  RB_NCCL(g_nccl.GroupEnd());
  RB_CUDA(cudaMemcpyAsync(theirs, stage + HB, 2 * HB, cudaMemcpyDeviceToHost, h->commStream));
  RB_CUDA(cudaStreamSynchronize(h->commStream));
  for (int d = 0; d < 2 && ok; ++d) {
    if (!has[d]) continue;
    for (int i = 0; i < NH && ok; ++i)
      if (cudaIpcOpenMemHandle(&h->peerBase[d][i], theirs[d][i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        h->peerBase[d][i] = nullptr;
        ok = 0;
      }
  }
  (void)cudaGetLastError();
  int all_ok = 0;
  if (g_nccl.AllReduce) {
    int* agree = reinterpret_cast<int*>(stage + 3 * HB);
    RB_CUDA(cudaMemcpyAsync(agree, &ok, sizeof(int), cudaMemcpyHostToDevice, h->commStream));
    RB_NCCL(g_nccl.AllReduce(agree, agree, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, h->comm, h->commStream));
    RB_CUDA(cudaMemcpyAsync(&all_ok, agree, sizeof(int), cudaMemcpyDeviceToHost, h->commStream));
    RB_CUDA(cudaStreamSynchronize(h->commStream));
  }
  if (all_ok) {
    for (int d = 0; d < 2; ++d) {
      if (!has[d]) continue;
      for (int b = 0; b < 4; ++b) {
        h->peerR[d][b].point_wsum = (F4*)h->peerBase[d][4 * b + 0];
        h->peerR[d][b].normal_W = (F4*)h->peerBase[d][4 * b + 1];
        h->peerR[d][b].Li_conf = (F4*)h->peerBase[d][4 * b + 2];
        h->peerR[d][b].light_idx = (int*)h->peerBase[d][4 * b + 3];
      }
      h->peerFlags[d] = (uint32_t*)h->peerBase[d][16];
    }
    h->p2p = true;
  } else {
    for (int d = 0; d < 2; ++d)
      for (int i = 0; i < NH; ++i)
        if (h->peerBase[d][i]) {
          cudaIpcCloseMemHandle(h->peerBase[d][i]);
          h->peerBase[d][i] = nullptr;
        }
    (void)cudaGetLastError();
    h->p2p = false;
    if (want && h->commRank == 0)
      fprintf(stderr, "[rb comm] peer-memory halo exchange unavailable on at least one rank: using ncclSend/ncclRecv\n");
  }
  if (getenv("RB_BAL_DEBUG") && h->commRank == 0) fprintf(stderr, "[rb comm] halo exchange: %s\n", h->p2p ? "peer memory (CUDA IPC)" : "NCCL send/recv");
  return RB_OK;
}

// Everything rb_comm_init made, in reverse. `barrier`: a normal shutdown — all ranks meet (one small all-reduce) after
// their own streams have drained and BEFORE any peer mapping is closed or any plane freed, so that no neighbour can
// still be storing halo rows into this rank's memory. (Skipped when tearing down a failed init: peers may be gone.)
static void comm_teardown(RbContext* h, bool barrier) {
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->fstream) cudaStreamSynchronize(h->fstream);
  if (h->commStream) cudaStreamSynchronize(h->commStream);
  if (barrier && h->comm && h->commStream && h->balDev && g_nccl.AllReduce) {
    if (g_nccl.AllReduce(h->balDev, h->balDev, 1, /*ncclFloat32*/ 7, /*ncclSum*/ 0, h->comm, h->commStream) == 0)
      cudaStreamSynchronize(h->commStream);
  }
  for (int d = 0; d < 2; ++d)
    for (int i = 0; i < 17; ++i)
      if (h->peerBase[d][i]) {
        cudaIpcCloseMemHandle(h->peerBase[d][i]);
        h->peerBase[d][i] = nullptr;
      }
  h->p2p = false;
  if (h->haloFlags) cudaFree(h->haloFlags);
  if (h->haloErr) cudaFreeHost(h->haloErr);
  h->haloFlags = nullptr, h->haloErr = nullptr;
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  h->comm = nullptr;
  h->commRank = 0, h->commSize = 1;
  if (h->commStream) cudaStreamDestroy(h->commStream);
  h->commStream = nullptr;
  cudaEvent_t* single[] = {&h->evHaloReady, &h->evHaloDone, &h->evHaloT0, &h->evShipReady, &h->evShipDone, &h->evAccMoved};
  for (cudaEvent_t* e : single)
    if (*e) {
      cudaEventDestroy(*e);
      *e = nullptr;
    }
  for (int i = 0; i < RbContext::kBalRing; ++i) {
    cudaEvent_t* ring[] = {&h->evBalCopied[i], &h->evFrameB[i], &h->evFrameE[i]};
    for (cudaEvent_t* e : ring)
      if (*e) {
        cudaEventDestroy(*e);
        *e = nullptr;
      }
    for (int j = 0; j < RbContext::kMaxStalls; ++j) {
      if (h->evStallA[i][j]) cudaEventDestroy(h->evStallA[i][j]);
      if (h->evStallB[i][j]) cudaEventDestroy(h->evStallB[i][j]);
      h->evStallA[i][j] = h->evStallB[i][j] = nullptr;
    }
  }
  if (h->balDev) cudaFree(h->balDev);
  if (h->balHost) cudaFreeHost(h->balHost);
  h->balDev = nullptr, h->balHost = nullptr;
  h->shipPending = false;
  (void)cudaGetLastError();
}

extern "C" {

uint32_t rb_abi_version(void) { return RB_ABI_VERSION; }

const char* rb_last_error(RbHandle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void rb_default_params(RbParams* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->M_Area = 1;
  p->M_Brdf = 1;
  p->spatialReuseNeighborCount = 5;
  p->spatialPassCount = 1;
  p->confidenceCap = 20;
  p->spatialReuseRadius = 30.0f;
  p->minNormalSimilarity = 0.85f;
  p->maxDepthDifference = 0.2f;
  p->spatialWeightCalc = RB_SW_CONSTANT;
  p->tnearOffset = 0.01f;
  p->tfarOffset = 0.001f;
  p->normalOffset = 0.001f;
  p->bgColor[0] = p->bgColor[1] = p->bgColor[2] = 0.5f;
  p->useSkybox = 0;
  p->lightSampler = RB_LS_CDF;
  p->wavefront = 0;
}

int rb_create(const RbCreateInfo* info, RbHandle* out) {
  if (!info || !out || info->width <= 0 || info->height <= 0 || info->band_y0 < 0 || info->band_y1 > info->height ||
      info->band_y0 >= info->band_y1) {
    g_create_error = "rb_create: invalid RbCreateInfo";
    return RB_ERR_INVALID_ARGUMENT;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("rb_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
    return RB_ERR_CUDA;
  }
  if (info->device < 0 || info->device >= ndev) {
    g_create_error = "rb_create: device ordinal out of range";
    return RB_ERR_INVALID_ARGUMENT;
  }
  RbContext* h = new RbContext();
  h->info = *info;
  rb_default_params(&h->params);
  auto fail = [&](int rc) {
    g_create_error = h->err;
    free_list(h->allocs);
    delete h;
    return rc;
  };
  auto body = [&]() -> int {
    RB_CUDA(cudaSetDevice(info->device));
    {
      int lo = 0, hi = 0;  // the back half (critical path) outranks the front half of the next frame
      RB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      RB_CUDA(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi));
    }
    const size_t n = (size_t)info->width * info->height;
    for (int i = 0; i < 3; ++i) {
      RB_TRY(dev_alloc(h, &h->G[i].pos_depth, n, h->allocs));
      RB_TRY(dev_alloc(h, &h->G[i].normal_shin, n, h->allocs));
      RB_TRY(dev_alloc(h, &h->G[i].diffuse_iim, n, h->allocs));
      RB_TRY(dev_alloc(h, &h->G[i].spec_type, n, h->allocs));
      RB_TRY(dev_alloc(h, &h->G[i].emission, n, h->allocs));
      RB_TRY(dev_alloc(h, &h->G[i].hit_ids, n, h->allocs));
      RB_CUDA(cudaMemsetAsync(h->G[i].pos_depth, 0, n * 16, h->stream));
      RB_CUDA(cudaMemsetAsync(h->G[i].normal_shin, 0, n * 16, h->stream));
      RB_CUDA(cudaMemsetAsync(h->G[i].diffuse_iim, 0, n * 16, h->stream));
      RB_CUDA(cudaMemsetAsync(h->G[i].spec_type, 0, n * 16, h->stream));
      RB_CUDA(cudaMemsetAsync(h->G[i].emission, 0, n * 16, h->stream));
      RB_CUDA(cudaMemsetAsync(h->G[i].hit_ids, 0xFF, n * 8, h->stream));
    }
    for (int i = 0; i < 4; ++i) {
      RB_TRY(dev_alloc(h, &h->R[i].point_wsum, n, h->allocs));
      RB_TRY(dev_alloc(h, &h->R[i].normal_W, n, h->allocs));
      RB_TRY(dev_alloc(h, &h->R[i].Li_conf, n, h->allocs));
      RB_TRY(dev_alloc(h, &h->R[i].light_idx, n, h->allocs));
      RB_CUDA(cudaMemsetAsync(h->R[i].point_wsum, 0, n * 16, h->stream));
      RB_CUDA(cudaMemsetAsync(h->R[i].normal_W, 0, n * 16, h->stream));
      RB_CUDA(cudaMemsetAsync(h->R[i].Li_conf, 0, n * 16, h->stream));
      RB_CUDA(cudaMemsetAsync(h->R[i].light_idx, 0xFF, n * 4, h->stream));
    }
    RB_TRY(dev_alloc(h, &h->frame, n * 3, h->allocs));
    RB_CUDA(cudaMemsetAsync(h->frame, 0, n * 12, h->stream));
    RB_TRY(dev_alloc(h, &h->accumulator, n * 3, h->allocs));
    RB_CUDA(cudaMemsetAsync(h->accumulator, 0, n * 12, h->stream));
    RB_TRY(dev_alloc(h, &h->display, n, h->allocs));
    RB_CUDA(cudaMemsetAsync(h->display, 0, n * 16, h->stream));
    RB_TRY(dev_alloc(h, &h->statSums, 2, h->allocs));
    RB_TRY(dev_alloc(h, &h->counters2, 16, h->allocs));
    RB_CUDA(cudaMemsetAsync(h->counters2, 0, 128, h->stream));
    h->counters = h->counters2;
    {
      int lo = 0, hi = 0;
      RB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      // the front half yields to the back half (RB_FRONT_PRIO=1: same priority — measured, no better)
      const bool same = getenv("RB_FRONT_PRIO") && atoi(getenv("RB_FRONT_PRIO")) != 0;
      RB_CUDA(cudaStreamCreateWithPriority(&h->fstream, cudaStreamNonBlocking, same ? hi : lo));
    }
    RB_CUDA(cudaEventCreateWithFlags(&h->evFrontDone, cudaEventDisableTiming));
    for (auto& ev : h->evBackDone) RB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : h->ev) RB_CUDA(cudaEventCreate(&ev));
    for (auto& ev : h->fev) RB_CUDA(cudaEventCreate(&ev));
    h->evCreated = true;
    RB_CUDA(cudaDeviceGetAttribute(&h->numSMs, cudaDevAttrMultiProcessorCount, info->device));
    if (const char* e = getenv("RB_REFILL")) h->refillLanes = atoi(e);
    if (const char* e = getenv("RB_POSTPONE")) h->postponeLanes = atoi(e);
    if (const char* e = getenv("RB_TRACE_BLOCKS")) h->traceBlocksPerSM = std::max(1, atoi(e));
    if (const char* e = getenv("RB_SMEM_STACK")) h->smemStack = atoi(e) != 0;
    if (const char* e = getenv("RB_OVERLAP")) h->overlap = atoi(e) != 0;
    if (const char* e = getenv("RB_FRAMES_IN_FLIGHT")) h->maxFramesInFlight = atoi(e) != 0;
    if (const char* e = getenv("RB_OVERLAP_DEBUG")) h->ovDebug = atoi(e) != 0;
    if (h->ovDebug)
      for (auto& f : h->evOv)
        for (auto& ev : f) RB_CUDA(cudaEventCreate(&ev));
    if (const char* e = getenv("RB_TWO_STEP_BRDF")) h->twoStepBrdf = atoi(e) != 0;
    // arithmetic self-check: implicit contraction must be off
    float* d = nullptr;
    RB_TRY(dev_alloc(h, &d, 1, h->allocs));
    const float a = 1.0f + 1.0f / 4096.0f;
    k_selfcheck<<<1, 1, 0, h->stream>>>(a, a, -1.0f, d);
    float r = 0;
    RB_CUDA(cudaMemcpyAsync(&r, d, 4, cudaMemcpyDeviceToHost, h->stream));
    RB_CUDA(cudaStreamSynchronize(h->stream));
    if (r != 1.0f / 2048.0f) {
      h->err = "rb_create: kernels were built with FMA contraction enabled (need nvcc -fmad=false)";
      return RB_ERR_UNSUPPORTED;
    }
    return RB_OK;
  };
  int rc = body();
  if (rc != RB_OK) return fail(rc);
  *out = h;
  return RB_OK;
}

void rb_destroy(RbHandle h) {
  if (!h) return;
  cudaSetDevice(h->info.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  comm_teardown(h, /*barrier=*/true);  // first: the ranks meet before any plane a neighbour may still write is freed
  if (h->ovDebug) {
    const int n = (int)std::min<uint32_t>(h->frameSeq, RbContext::kOvFrames);
    for (int f = 1; f < n; ++f) {
      float t[4];
      for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], h->evOv[1][0], h->evOv[f][k]);
      fprintf(stderr, "[rb overlap] frame %2d: front %8.3f .. %8.3f   back issued %8.3f, ends %8.3f ms\n", f, t[0], t[1], t[2], t[3]);
    }
    for (auto& f : h->evOv)
      for (auto& ev : f) cudaEventDestroy(ev);
  }
  free_list(h->allocs);
  free_list(h->sceneAllocs);
  free_list(h->texAllocs);
  free_list(h->skyAllocs);
  if (h->evCreated) {
    for (auto& ev : h->ev) cudaEventDestroy(ev);
    for (auto& ev : h->fev) cudaEventDestroy(ev);
  }
  if (h->wave.rays) cudaFree(h->wave.rays);
  if (h->wave.occ) cudaFree(h->wave.occ);
  if (h->fwave.rays) cudaFree(h->fwave.rays);
  if (h->fwave.hits) cudaFree(h->fwave.hits);
  if (h->fwave.occ) cudaFree(h->fwave.occ);
  if (h->fwave.brdf_dir) cudaFree(h->fwave.brdf_dir);
  for (auto q : h->visRays)
    if (q) cudaFree(q);
  if (h->fstream) {
    cudaStreamSynchronize(h->fstream);
    cudaStreamDestroy(h->fstream);
  }
  if (h->copyStream) {
    cudaStreamSynchronize(h->copyStream);
    cudaStreamDestroy(h->copyStream);
  }
  for (auto ev : h->evCopyDone)
    if (ev) cudaEventDestroy(ev);
  if (h->evFrameReady) cudaEventDestroy(h->evFrameReady);
  if (h->evFrontDone) cudaEventDestroy(h->evFrontDone);
  for (auto ev : h->evBackDone)
    if (ev) cudaEventDestroy(ev);
  if (h->wave.cand) cudaFree(h->wave.cand);
  if (h->wave.deferred) cudaFree(h->wave.deferred);
  if (h->waveCounters) cudaFree(h->waveCounters);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int rb_set_params(RbHandle h, const RbParams* p) {
  if (!h || !p) return RB_ERR_INVALID_ARGUMENT;
  if (p->useSkybox && h->sc.sky.data == nullptr) {  // the reference would dereference a null Scene::skybox
    h->err = "rb_set_params: useSkybox=1 needs the sky texture (rb_set_sky)";
    return RB_ERR_INVALID_ARGUMENT;
  }
  if (p->M_Area < 0 || p->M_Brdf < 0 || p->spatialReuseNeighborCount < 0 || p->spatialPassCount < 0 ||
      p->spatialWeightCalc < 0 || p->spatialWeightCalc > 4 || p->lightSampler < 0 || p->lightSampler > 1) {
    h->err = "rb_set_params: parameter out of range";
    return RB_ERR_INVALID_ARGUMENT;
  }
  if (p->spatialReuseNeighborCount > RB_MAX_NEIGHBORS) {
    h->err = "rb_set_params: spatialReuseNeighborCount > 32 is not supported";
    return RB_ERR_UNSUPPORTED;
  }
  h->params = *p;
  return RB_OK;
}

// -------------------------------------------------------------------------------------
// scene upload + BVH build
// -------------------------------------------------------------------------------------
static int build_bvh(RbContext* h, const float* d_tri_pos, const uint32_t* d_id_map, uint32_t n, float maxabs, F4** node8_out,
                     F4** tri_isect_out, uint32_t* n_nodes_out, uint32_t* depth_out) {
  std::vector<void*> tmp;
  auto cleanup = [&]() { free_list(tmp); };
  BuildCtx c{};
  c.n = n;
  c.tri_pos = d_tri_pos;
  c.id_map = d_id_map;
  c.pad = maxabs * (1.0f / 262144.0f) + 1e-30f;  // 2^-18 of the scene scale (DESIGN.md "conservative culling")
  auto body = [&]() -> int {
    RB_TRY(dev_alloc(h, &c.scene_bounds, 6, tmp));
    RB_TRY(dev_alloc(h, &c.tbox_lo, n, tmp));
    RB_TRY(dev_alloc(h, &c.tbox_hi, n, tmp));
    uint64_t* morton_in = nullptr;
    uint32_t* order_in = nullptr;
    RB_TRY(dev_alloc(h, &morton_in, n, tmp));
    RB_TRY(dev_alloc(h, &order_in, n, tmp));
    RB_TRY(dev_alloc(h, &c.morton, n, tmp));
    RB_TRY(dev_alloc(h, &c.order, n, tmp));
    RB_TRY(dev_alloc(h, &c.left, n, tmp));
    RB_TRY(dev_alloc(h, &c.right, n, tmp));
    RB_TRY(dev_alloc(h, &c.parent, n, tmp));
    RB_TRY(dev_alloc(h, &c.leaf_parent, n, tmp));
    RB_TRY(dev_alloc(h, &c.range_lo, n, tmp));
    RB_TRY(dev_alloc(h, &c.range_hi, n, tmp));
    RB_TRY(dev_alloc(h, &c.ibox_lo, n, tmp));
    RB_TRY(dev_alloc(h, &c.ibox_hi, n, tmp));
    RB_TRY(dev_alloc(h, &c.visit, n, tmp));
    RB_TRY(dev_alloc(h, &c.dp_cost, 7 * (size_t)n, tmp));
    RB_TRY(dev_alloc(h, &c.dp_dec, 2 * (size_t)n, tmp));
    c.c_prim = RB_COLLAPSE_C_PRIM;
    RB_TRY(dev_alloc(h, &c.counters, 4, tmp));
    int* q[2] = {nullptr, nullptr};
    RB_TRY(dev_alloc(h, &q[0], 2 * (size_t)n + 2, tmp));
    RB_TRY(dev_alloc(h, &q[1], 2 * (size_t)n + 2, tmp));
    const size_t max_nodes = (size_t)n / 2 + 8;
    F4* node8_big = nullptr;
    RB_TRY(dev_alloc(h, &node8_big, RB_NODE_F4 * max_nodes, tmp));
    F4* tri_isect = nullptr;
    RB_TRY(dev_alloc(h, &tri_isect, 3 * (size_t)n, h->sceneAllocs));
    c.node8 = node8_big;
    c.tri_isect = tri_isect;

    const int B = 256;
    const unsigned grid = (n + B - 1) / B;
    const int init_bounds[6] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    RB_CUDA(cudaMemcpyAsync(c.scene_bounds, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, h->stream));
    RB_CUDA(cudaMemsetAsync(c.visit, 0, sizeof(int) * (size_t)n, h->stream));
    k_bounds<<<grid, B, 0, h->stream>>>(c);
    // Morton codes go to the sort inputs
    BuildCtx cm = c;
    cm.morton = morton_in;
    cm.order = order_in;
    k_morton<<<grid, B, 0, h->stream>>>(cm);
    size_t sort_bytes = 0;
    RB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, morton_in, c.morton, order_in, c.order, (int)n, 0, 63,
                                            h->stream));
    void* sort_tmp = nullptr;
    RB_CUDA(cudaMalloc(&sort_tmp, std::max<size_t>(sort_bytes, 16)));
    tmp.push_back(sort_tmp);
    RB_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, morton_in, c.morton, order_in, c.order, (int)n, 0, 63,
                                            h->stream));
    uint32_t n_nodes = 0, depth = 0;
    if (n <= RB_LEAF_MAX) {
      k_tiny_root<<<1, 1, 0, h->stream>>>(c);
      n_nodes = 1;
      depth = 1;
    } else {
      k_karras<<<grid, B, 0, h->stream>>>(c);
      k_fit<<<grid, B, 0, h->stream>>>(c);
      // level-synchronous collapse, root binary node 0 -> node8 0
      const int first[2] = {0, 0};
      RB_CUDA(cudaMemcpyAsync(q[0], first, sizeof(first), cudaMemcpyHostToDevice, h->stream));
      const int init_counters[4] = {1, 0, 0, 0};
      RB_CUDA(cudaMemcpyAsync(c.counters, init_counters, sizeof(init_counters), cudaMemcpyHostToDevice, h->stream));
      int len = 1, cur = 0;
      while (len > 0) {
        c.q_in = q[cur];
        c.q_out = q[cur ^ 1];
        c.q_in_len = len;
        RB_CUDA(cudaMemsetAsync(c.counters + 2, 0, sizeof(int), h->stream));
        k_collapse<<<(len + 127) / 128, 128, 0, h->stream>>>(c);
        int hc[4];
        RB_CUDA(cudaMemcpyAsync(hc, c.counters, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
        RB_CUDA(cudaStreamSynchronize(h->stream));
        if ((size_t)hc[0] > max_nodes) {
          h->err = "build_bvh: node pool overflow";
          return RB_ERR_CUDA;
        }
        len = hc[2];
        n_nodes = (uint32_t)hc[0];
        cur ^= 1;
        depth++;
        if (depth > 4096) {
          h->err = "build_bvh: collapse did not terminate";
          return RB_ERR_CUDA;
        }
      }
      int hc[4];
      RB_CUDA(cudaMemcpyAsync(hc, c.counters, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
      RB_CUDA(cudaStreamSynchronize(h->stream));
      if ((uint32_t)hc[1] != n) {
        h->err = "build_bvh: emitted " + std::to_string(hc[1]) + " leaf triangles for " + std::to_string(n);
        return RB_ERR_CUDA;
      }
    }
    RB_CUDA(cudaGetLastError());
    if (2 * depth + 2 > RB_STACK_MAX) {
      h->err = "build_bvh: tree depth " + std::to_string(depth) + " exceeds the traversal stack";
      return RB_ERR_UNSUPPORTED;
    }
    F4* node8 = nullptr;
    RB_TRY(dev_alloc(h, &node8, RB_NODE_F4 * (size_t)n_nodes, h->sceneAllocs));
    RB_CUDA(cudaMemcpyAsync(node8, node8_big, 16 * RB_NODE_F4 * (size_t)n_nodes, cudaMemcpyDeviceToDevice, h->stream));
    RB_CUDA(cudaStreamSynchronize(h->stream));
    *node8_out = node8;
    *tri_isect_out = tri_isect;
    *n_nodes_out = n_nodes;
    *depth_out = depth;
    return RB_OK;
  };
  int rc = body();
  cudaStreamSynchronize(h->stream);
  cleanup();
  return rc;
}

int rb_upload_scene(RbHandle h, const RbSceneDesc* sd) {
  if (!h || !sd || (sd->n_surfaces && !sd->surfaces) || (sd->n_materials && !sd->materials)) return RB_ERR_INVALID_ARGUMENT;
  RB_CUDA(cudaSetDevice(h->info.device));
  RB_CUDA(cudaStreamSynchronize(h->stream));
  free_list(h->sceneAllocs);
  free_list(h->texAllocs);
  h->haveScene = false;
  h->havePrev = false;

  h->hasDielectric = false;
  for (uint32_t m = 0; m < sd->n_materials; ++m)
    if (sd->materials[m].type == RB_MAT_DIELECTRIC) h->hasDielectric = true;
  HostScene hs;
  {
    int frc = flatten_scene(sd, hs, h->err);
    if (frc != RB_OK) return frc;
  }
  const size_t n = hs.n;
  const size_t NL = hs.emissive.size();
  const float maxabs = hs.maxabs, totalSurface = hs.totalSurface;
  std::vector<float>& pos = hs.pos;
  std::vector<F4>&nrm = hs.nrm, &mat = hs.mat, &light = hs.light;
  std::vector<U4>& info = hs.info;
  std::vector<float>&cdf = hs.cdf, &alias_prob = hs.alias_prob;
  std::vector<uint32_t>& alias_idx = hs.alias_idx;

  // ---- upload -------------------------------------------------------------------------------
  cudaEvent_t t0, t1;
  RB_CUDA(cudaEventCreate(&t0));
  RB_CUDA(cudaEventCreate(&t1));
  float* d_pos = nullptr;
  F4 *d_nrm = nullptr, *d_mat = nullptr, *d_light = nullptr;
  U4* d_info = nullptr;
  float *d_cdf = nullptr, *d_ap = nullptr;
  uint32_t* d_ai = nullptr;
  std::vector<void*> tmp;
  auto body = [&]() -> int {
    RB_TRY(dev_alloc(h, &d_pos, 9 * n, tmp));
    RB_TRY(dev_alloc(h, &d_nrm, 3 * n, h->sceneAllocs));
    RB_TRY(dev_alloc(h, &d_info, n, h->sceneAllocs));
    RB_TRY(dev_alloc(h, &d_mat, mat.size(), h->sceneAllocs));
    RB_TRY(dev_alloc(h, &d_light, light.size(), h->sceneAllocs));
    RB_TRY(dev_alloc(h, &d_cdf, NL, h->sceneAllocs));
    RB_TRY(dev_alloc(h, &d_ap, NL, h->sceneAllocs));
    RB_TRY(dev_alloc(h, &d_ai, NL, h->sceneAllocs));
    U2* d_pair = nullptr;
    RB_TRY(dev_alloc(h, &d_pair, NL, h->sceneAllocs));

    auto up = [&](void* d, const void* s, size_t bytes) -> cudaError_t {
      return bytes ? cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream) : cudaSuccess;
    };
    RB_CUDA(up(d_pos, pos.data(), pos.size() * 4));
    RB_CUDA(up(d_nrm, nrm.data(), nrm.size() * 16));
    RB_CUDA(up(d_info, info.data(), info.size() * 16));
    RB_CUDA(up(d_mat, mat.data(), mat.size() * 16));
    RB_CUDA(up(d_light, light.data(), light.size() * 16));
    RB_CUDA(up(d_cdf, cdf.data(), NL * 4));
    RB_CUDA(up(d_ap, alias_prob.data(), NL * 4));
    RB_CUDA(up(d_ai, alias_idx.data(), NL * 4));
    RB_CUDA(up(d_pair, hs.alias_pair.data(), NL * 8));
    F4* d_cull = nullptr;
    RB_TRY(dev_alloc(h, &d_cull, NL, h->sceneAllocs));
    RB_CUDA(up(d_cull, hs.light_cull.data(), NL * 16));
    F4* d_uv = nullptr;
    if (!hs.uv.empty()) {
      RB_TRY(dev_alloc(h, &d_uv, hs.uv.size(), h->sceneAllocs));
      RB_CUDA(up(d_uv, hs.uv.data(), hs.uv.size() * 16));
    }
    h->sceneUv = d_uv;
    F4* d_tan = nullptr;
    if (!hs.tan.empty()) {
      RB_TRY(dev_alloc(h, &d_tan, hs.tan.size(), h->sceneAllocs));
      RB_CUDA(up(d_tan, hs.tan.data(), hs.tan.size() * 16));
    }
    h->sceneTan = d_tan;
    h->nMaterials = sd->n_materials;
    MatConst* d_mc = nullptr;
    if (sd->n_materials > 0) {
      RB_TRY(dev_alloc(h, &d_mc, sd->n_materials, h->sceneAllocs));
      k_material_constants<<<(sd->n_materials + 127) / 128, 128, 0, h->stream>>>(d_mat, d_mc, sd->n_materials);
    }
    RB_CUDA(cudaStreamSynchronize(h->stream));
    F4 *node8 = nullptr, *tri_isect = nullptr;
    uint32_t n_nodes = 0, depth = 0;
    RB_CUDA(cudaEventRecord(t0, h->stream));
    if (n > 0) RB_TRY(build_bvh(h, d_pos, nullptr, (uint32_t)n, maxabs, &node8, &tri_isect, &n_nodes, &depth));
    // emissive-only BVH (two-step BRDF-candidate rays): same builder over the emitter subset, leaves keep scene ids
    F4 *em_node8 = nullptr, *em_tri_isect = nullptr;
    uint32_t em_nodes = 0, em_depth = 0;
    if (NL > 0) {
      std::vector<float> epos(9 * NL);
      for (size_t i = 0; i < NL; ++i) memcpy(&epos[9 * i], &pos[9 * (size_t)hs.emissive[i]], 36);
      float* d_epos = nullptr;
      uint32_t* d_emap = nullptr;
      RB_TRY(dev_alloc(h, &d_epos, 9 * NL, tmp));
      RB_TRY(dev_alloc(h, &d_emap, NL, tmp));
      RB_CUDA(up(d_epos, epos.data(), epos.size() * 4));
      RB_CUDA(up(d_emap, hs.emissive.data(), NL * 4));
      RB_CUDA(cudaStreamSynchronize(h->stream));
      RB_TRY(build_bvh(h, d_epos, d_emap, (uint32_t)NL, maxabs, &em_node8, &em_tri_isect, &em_nodes, &em_depth));
    }
    RB_CUDA(cudaEventRecord(t1, h->stream));
    RB_CUDA(cudaEventSynchronize(t1));
    float ms = 0;
    RB_CUDA(cudaEventElapsedTime(&ms, t0, t1));
    SceneDev& sc = h->sc;
    sc.node8 = node8;
    sc.tri_isect = tri_isect;
    sc.tri_normals = d_nrm;
    sc.tri_info = d_info;
    sc.mat = d_mat;
    sc.mat_const = d_mc;
    sc.light = d_light;
    sc.cdf = d_cdf;
    sc.alias_prob = d_ap;
    sc.alias_idx = d_ai;
    sc.alias_pair = d_pair;
    sc.light_cull = d_cull;
    sc.maxabs = maxabs;
    sc.n_lights = (uint32_t)NL;
    sc.n_tris = (uint32_t)n;
    sc.n_nodes = n_nodes;
    sc.total_area = totalSurface;
    sc.q7_base = 0x43000000u;
    sc.em_node8 = em_node8;
    sc.em_tri_isect = em_tri_isect;
    sc.em_n_nodes = em_nodes;
    sc.tri_uv = h->sceneUv;
    sc.tex = nullptr;  // a new scene drops the textures (rb_set_textures)
    sc.mat_tex = nullptr;
    sc.tri_tan = nullptr;
    RbSceneStats& st = h->stats;
    memset(&st, 0, sizeof(st));
    st.n_triangles = (uint32_t)n;
    st.n_emissive = (uint32_t)NL;
    st.n_bvh_nodes = n_nodes;
    st.bvh_depth = depth;
    {  // shared-memory traversal stack: 2 * depth + 2 entries per lane (what build_bvh checks against RB_STACK_MAX), if
       // that fits next to the other resident CTAs of an SM (227 KB, 1 KB reserved per CTA)
      const int need = 2 * (int)std::max(depth, em_depth) + 2;
      const size_t per_cta = (size_t)need * kTraceThreads * sizeof(U2) + 1024;
      h->smemStackEntries = (h->smemStack && per_cta * (size_t)h->traceBlocksPerSM <= 227u * 1024u) ? need : 0;
    }
    st.build_ms = ms;
    st.total_emissive_area = totalSurface;
    for (int a = 0; a < 3; ++a) st.bounds_lo[a] = FLT_MAX, st.bounds_hi[a] = -FLT_MAX;
    for (size_t i = 0; i < pos.size(); ++i) {
      st.bounds_lo[i % 3] = std::min(st.bounds_lo[i % 3], pos[i]);
      st.bounds_hi[i % 3] = std::max(st.bounds_hi[i % 3], pos[i]);
    }
    return RB_OK;
  };
  int rc = body();
  cudaStreamSynchronize(h->stream);
  free_list(tmp);
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  if (rc != RB_OK) {
    free_list(h->sceneAllocs);
    return rc;
  }
  h->haveScene = true;
  return RB_OK;
}

// Textured materials: texel arrays and per-material slots; see include/restir_b200.h
int rb_set_textures(RbHandle h, const RbTexture* textures, uint32_t n_textures, const RbMaterialTextures* per_material,
                    uint32_t n_materials) {
  if (!h || (n_textures && !textures) || !per_material) return RB_ERR_INVALID_ARGUMENT;
  if (!h->haveScene) {
    h->err = "rb_set_textures: no scene uploaded";
    return RB_ERR_NO_SCENE;
  }
  if (n_materials != h->nMaterials) {
    h->err = "rb_set_textures: one RbMaterialTextures per material of the uploaded scene is required";
    return RB_ERR_INVALID_ARGUMENT;
  }
  bool any = false, any_normal = false;
  for (uint32_t m = 0; m < n_materials; ++m) {
    const int32_t s[4] = {per_material[m].diffuse, per_material[m].specular, per_material[m].shininess, per_material[m].normal};
    for (int k = 0; k < 4; ++k) {
      if (s[k] >= (int32_t)n_textures || s[k] < -1) {
        h->err = "rb_set_textures: texture index out of range";
        return RB_ERR_INVALID_ARGUMENT;
      }
      any = any || s[k] >= 0;
    }
    any_normal = any_normal || s[3] >= 0;
  }
  for (uint32_t t = 0; t < n_textures; ++t) {
    const RbTexture& T = textures[t];
    const bool fmt = T.pixel_size == 3 || T.pixel_size == 4 || T.pixel_size == 12 || T.pixel_size == 16;
    if (T.width <= 0 || T.height <= 0 || !fmt || T.scan_width < T.width * T.pixel_size || !T.data) {
      h->err = "rb_set_textures: bad texture " + std::to_string(t);
      return RB_ERR_INVALID_ARGUMENT;
    }
  }
  if (any && !h->sceneUv) {
    h->err = "rb_set_textures: the uploaded scene carries no texture coordinates (RbSurface.uv)";
    return RB_ERR_INVALID_ARGUMENT;
  }
  if (any_normal && !h->sceneTan) {
    h->err = "rb_set_textures: normal maps need the per-vertex tangents of the uploaded scene (RbSurface.tangent)";
    return RB_ERR_INVALID_ARGUMENT;
  }
  RB_CUDA(cudaSetDevice(h->info.device));
  RB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->fstream) RB_CUDA(cudaStreamSynchronize(h->fstream));
  free_list(h->texAllocs);
  h->sc.tex = nullptr;
  h->sc.mat_tex = nullptr;
  h->sc.tri_tan = nullptr;
  if (!any) return RB_OK;
  std::vector<TexDev> tab(n_textures);
  for (uint32_t t = 0; t < n_textures; ++t) {
    const RbTexture& T = textures[t];
    unsigned char* d = nullptr;
    const size_t bytes = (size_t)T.scan_width * T.height;
    RB_TRY(dev_alloc(h, &d, bytes, h->texAllocs));
    RB_CUDA(cudaMemcpyAsync(d, T.data, bytes, cudaMemcpyHostToDevice, h->stream));
    tab[t] = TexDev{d, T.width, T.height, T.scan_width, T.pixel_size};
  }
  TexDev* d_tab = nullptr;
  I4* d_slots = nullptr;
  RB_TRY(dev_alloc(h, &d_tab, std::max<uint32_t>(n_textures, 1), h->texAllocs));
  RB_TRY(dev_alloc(h, &d_slots, n_materials, h->texAllocs));
  RB_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(TexDev), cudaMemcpyHostToDevice, h->stream));
  static_assert(sizeof(RbMaterialTextures) == sizeof(I4), "slot record layout");
  RB_CUDA(cudaMemcpyAsync(d_slots, per_material, n_materials * sizeof(I4), cudaMemcpyHostToDevice, h->stream));
  RB_CUDA(cudaStreamSynchronize(h->stream));
  h->sc.tex = d_tab;
  h->sc.mat_tex = d_slots;
  h->sc.tri_tan = any_normal ? h->sceneTan : nullptr;
  h->havePrev = false;  // the previous frame's G-buffer was made with other materials
  return RB_OK;
}

// Sky texture (Scene::setSkybox, P/Scene.cpp:47-50; SphericalMap::getTexel, P/SphericalMap.cpp:10-14); see include/restir_b200.h
int rb_set_sky(RbHandle h, const RbTexture* sky) {
  if (!h) return RB_ERR_INVALID_ARGUMENT;
  if (!sky && h->params.useSkybox) {
    h->err = "rb_set_sky: the sky cannot be removed while useSkybox=1 (rb_set_params first)";
    return RB_ERR_INVALID_ARGUMENT;
  }
  if (sky) {
    const bool fmt = sky->pixel_size == 3 || sky->pixel_size == 4 || sky->pixel_size == 12 || sky->pixel_size == 16;
    if (sky->width <= 0 || sky->height <= 0 || !fmt || sky->scan_width < sky->width * sky->pixel_size || !sky->data) {
      h->err = "rb_set_sky: bad texture";
      return RB_ERR_INVALID_ARGUMENT;
    }
  }
  RB_CUDA(cudaSetDevice(h->info.device));
  RB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->fstream) RB_CUDA(cudaStreamSynchronize(h->fstream));
  free_list(h->skyAllocs);
  h->sc.sky = TexDev{nullptr, 0, 0, 0, 0, 0};
  h->havePrev = false;  // the previous frame's G-buffer was made with another background
  if (!sky) return RB_OK;
  unsigned char* d = nullptr;
  const size_t bytes = (size_t)sky->scan_width * sky->height;
  RB_TRY(dev_alloc(h, &d, bytes, h->skyAllocs));
  RB_CUDA(cudaMemcpyAsync(d, sky->data, bytes, cudaMemcpyHostToDevice, h->stream));
  RB_CUDA(cudaStreamSynchronize(h->stream));
  h->sc.sky = TexDev{d, sky->width, sky->height, sky->scan_width, sky->pixel_size, /*CLAMP_TO_EDGE*/ 1};
  return RB_OK;
}

int rb_scene_stats(RbHandle h, RbSceneStats* out) {
  if (!h || !out) return RB_ERR_INVALID_ARGUMENT;
  if (!h->haveScene) {
    h->err = "rb_scene_stats: no scene";
    return RB_ERR_NO_SCENE;
  }
  *out = h->stats;
  return RB_OK;
}

// -------------------------------------------------------------------------------------
// frame
// -------------------------------------------------------------------------------------
static CamState cam_state(const RbCamera* c) {
  CamState s;
  s.pos = v3(c->pos[0], c->pos[1], c->pos[2]);
  s.focal = c->focal_px;
  memcpy(s.viewMat, c->viewMat, sizeof(s.viewMat));
  memcpy(s.invViewMat, c->invViewMat, sizeof(s.invViewMat));
  return s;
}

// three {count, next} pairs: two alternate between the passes of a frame (a stream kernel zeroes the pair of the
// previous pass, so no reset launches), the third belongs to the ray seam
static int ensure_counters(RbContext* h) {
  if (h->waveCounters) return RB_OK;
  RB_CUDA(cudaMalloc((void**)&h->waveCounters, 6 * sizeof(uint32_t)));
  RB_CUDA(cudaMemsetAsync(h->waveCounters, 0, 6 * sizeof(uint32_t), h->stream));
  h->wavePair = 0;
  h->wave.count = h->waveCounters;
  h->wave.reset_pair = nullptr;
  return RB_OK;
}

// wavefront buffers sized for `slots` rays per band pixel
static int ensure_wave(RbContext* h, uint32_t slots, uint32_t brdf_slots, uint32_t cand_slots) {
  const size_t npix = (size_t)h->info.width * h->info.height;
  if ((size_t)cand_slots * npix > h->waveCandCap) {
    if (h->wave.cand) cudaFree(h->wave.cand);
    h->wave.cand = nullptr;
    h->waveCandCap = 0;
    RB_CUDA(cudaMalloc((void**)&h->wave.cand, (size_t)cand_slots * npix * sizeof(U4)));
    h->waveCandCap = (size_t)cand_slots * npix;
  }
  // sized for the band plus head-room: the balancer moves boundaries by a few rows every period, and growing a queue
  // means cudaFree + cudaMalloc (a device-wide synchronisation of several ms) in the middle of the frame loop
  const bool whole = h->info.band_y0 == 0 && h->info.band_y1 == h->info.height;
  const int band_rows_alloc = whole ? h->info.height : std::min(h->info.height, (h->info.band_y1 - h->info.band_y0) + 4 * RbContext::kShipRows);
  const size_t band_px_now = (size_t)h->info.width * (h->info.band_y1 - h->info.band_y0);
  const size_t band_px = (band_px_now * std::max<uint32_t>(slots, 1u) <= h->waveRayCap && band_px_now <= h->visRayCap &&
                          band_px_now * 2u * brdf_slots <= h->fwaveRayCap)
                             ? band_px_now  // fits what is there: keep it
                             : (size_t)h->info.width * band_rows_alloc;
  // back half: the queues of the temporal and spatial passes
  const size_t need_rays = band_px * slots;
  if (need_rays > h->waveRayCap) {
    if (h->wave.rays) cudaFree(h->wave.rays);
    h->wave.rays = nullptr;
    h->waveRayCap = 0;
    if (need_rays > 0xFFFFFFF0ull) {
      h->err = "wavefront ray queue would exceed 2^32 rays";
      return RB_ERR_UNSUPPORTED;
    }
    RB_CUDA(cudaMalloc((void**)&h->wave.rays, need_rays * sizeof(RayQ)));
    h->waveRayCap = need_rays;
  }
  if ((size_t)slots * npix > h->waveOccCap) {
    if (h->wave.occ) cudaFree(h->wave.occ);
    h->wave.occ = nullptr;
    RB_CUDA(cudaMalloc((void**)&h->wave.occ, (size_t)slots * npix));
    h->waveOccCap = (size_t)slots * npix;
  }
  // front half: BRDF-candidate rays in the first half of its queue, the chained "precedes" rays in the second
  const size_t need_front = band_px * 2u * brdf_slots;
  if (need_front > h->fwaveRayCap) {
    if (h->fwave.rays) cudaFree(h->fwave.rays);
    h->fwave.rays = nullptr;
    h->fwaveRayCap = 0;
    RB_CUDA(cudaMalloc((void**)&h->fwave.rays, need_front * sizeof(RayQ)));
    h->fwaveRayCap = need_front;
  }
  if ((size_t)brdf_slots * npix > h->fwaveHitCap) {
    if (h->fwave.hits) cudaFree(h->fwave.hits);
    if (h->fwave.occ) cudaFree(h->fwave.occ);
    if (h->fwave.brdf_dir) cudaFree(h->fwave.brdf_dir);
    h->fwave.hits = nullptr, h->fwave.occ = nullptr, h->fwave.brdf_dir = nullptr;
    h->fwaveHitCap = 0;
    RB_CUDA(cudaMalloc((void**)&h->fwave.brdf_dir, (size_t)brdf_slots * npix * sizeof(F4)));
    RB_CUDA(cudaMalloc((void**)&h->fwave.hits, (size_t)brdf_slots * npix * sizeof(HitRec)));
    RB_CUDA(cudaMalloc((void**)&h->fwave.occ, (size_t)brdf_slots * npix));
    h->fwaveHitCap = (size_t)brdf_slots * npix;
  }
  if (band_px > h->visRayCap) {  // visibility rays: one per pixel, one queue per frame parity
    for (auto& q : h->visRays) {
      if (q) cudaFree(q);
      q = nullptr;
    }
    h->visRayCap = 0;
    for (auto& q : h->visRays) RB_CUDA(cudaMalloc((void**)&q, band_px * sizeof(RayQ)));
    h->visRayCap = band_px;
  }
  const bool banded = !(h->info.band_y0 == 0 && h->info.band_y1 == h->info.height);
  if (banded && !h->wave.deferred) RB_CUDA(cudaMalloc((void**)&h->wave.deferred, npix * sizeof(uint32_t)));
  RB_TRY(ensure_counters(h));
  h->wave.capacity = (uint32_t)h->waveRayCap;
  h->wave.npix = (uint32_t)npix;
  h->wave.fuse_vis = h->wave.fuse_shade = 0u;
  h->wave.chain_rays = nullptr, h->wave.chain_count = nullptr, h->wave.chain_capacity = 0;
  h->wave.hits = nullptr;
  h->fwave.npix = (uint32_t)npix;
  h->fwave.capacity = (uint32_t)(band_px * brdf_slots);
  h->fwave.chain_rays = h->fwave.rays + band_px * brdf_slots;
  h->fwave.chain_capacity = (uint32_t)(band_px * brdf_slots);
  h->fwave.reset_pair = nullptr;
  h->fwave.cand = nullptr, h->fwave.deferred = nullptr, h->fwave.deferred_count = nullptr;
  h->fwave.fuse_vis = h->fwave.fuse_shade = 0u;
  return RB_OK;
}

}  // extern "C"

// ---- frame phases -----------------------------------------------------------------------------------
// A frame is: begin (G-buffer, initial candidates, visibility, temporal) -> spatial pass x N -> end (shade,
// buffer rotation). Between phases the reservoirs every later pass reads sit in R[rWrite]; that is where
// band halos are exchanged (NCCL inside rb_render_frame, or rb_halo_export/import by a host that drives the
// phases itself).

static void fs_mark(RbContext* h, int pass, int kind) {
  FrameState& F = h->fs;
  if (!F.timed || (int)F.marks.size() + 1 >= RbContext::kMaxEvents) return;
  cudaEventRecord(h->fev[F.marks.size() + 1], h->stream);
  F.marks.push_back({pass, kind});
}
static void fs_bind(RbContext* h) {
  FrameCtx& fc = h->fs.fc;
  fc.Rread = h->R[h->rRead];
  fc.Rwrite = h->R[h->rWrite];
  fc.Rlast = h->R[h->rLast];
}
static dim3 rows_grid(const FrameCtx& fc, int ry0, int ry1, int tileH) {
  return dim3((fc.width + kTileW - 1) / kTileW, (std::max(ry1 - ry0, 0) + tileH - 1) / tileH);
}
// launch a pixel kernel over image rows [ry0, ry1) — back half (h->fs.fc on h->stream) unless told otherwise
template <class K>
static void launch_rows(RbContext* h, K kernel, int ry0, int ry1, int threads = kTileW * kTileH, const FrameCtx* ctx = nullptr,
                        cudaStream_t st = nullptr) {
  if (ry1 <= ry0) return;
  FrameCtx f = ctx ? *ctx : h->fs.fc;
  f.y0 = ry0;
  f.y1 = ry1;
  const int tileH = threads / kTileW;
  kernel<<<rows_grid(f, ry0, ry1, tileH), dim3(kTileW, tileH), 0, st ? st : h->stream>>>(f);
  h->fs.launches++;
}
enum TraceMode { TRACE_CLOSEST = 0, TRACE_ANY = 1, TRACE_CLOSEST_EMISSIVE = 2, TRACE_ANY_PRECEDES = 3 };
// trace the queue of wave buffers `w` (the back half's by default)
static void fs_trace(RbContext* h, int mode, int pass, float tnear = -1.0f, const WaveBufs* wsel = nullptr, cudaStream_t st = nullptr) {
  const RbParams& P = h->fs.P;
  if (tnear < 0.0f) tnear = FLT_MIN + P.tnearOffset;
  const int trace_grid = h->numSMs * h->traceBlocksPerSM;
  const WaveBufs& w = wsel ? *wsel : h->wave;
  if (!st) st = h->stream;
  QueueIO io{w.rays, w.count, w.capacity, w.occ, w.hits, tnear, WaveBufs{}};
  uint32_t* next = w.count + 1;
  if (mode == TRACE_CLOSEST_EMISSIVE) io.chain = w;  // queues the "precedes" rays of its hits (brdf_chain_push) ...
  if (mode == TRACE_ANY_PRECEDES) {                  // ... which this launch traces
    io.rays = w.chain_rays, io.count_ptr = w.chain_count, io.capacity = w.chain_capacity;
    next = w.chain_count + 1;
  }
  // the per-lane stack goes to shared memory when 2 * depth + 2 entries of every resident CTA fit (h->smemStackEntries)
  const int sn = h->smemStackEntries;
  const size_t sb = (size_t)sn * kTraceThreads * sizeof(U2);
#define RB_TRACE_LAUNCH(ANY, TIE, SC)                                                                                     \
  do {                                                                                                                    \
    if (sn > 0)                                                                                                           \
      k_trace_queue<ANY, TIE, QueueIO, true><<<trace_grid, kTraceThreads, sb, st>>>(SC, io, next, h->refillLanes, h->postponeLanes, sn); \
    else                                                                                                                  \
      k_trace_queue<ANY, TIE, QueueIO, false><<<trace_grid, kTraceThreads, 0, st>>>(SC, io, next, h->refillLanes, h->postponeLanes, 0); \
  } while (0)
  if (mode == TRACE_ANY)
    RB_TRACE_LAUNCH(true, false, h->sc);
  else if (mode == TRACE_CLOSEST)
    RB_TRACE_LAUNCH(false, false, h->sc);
  else if (mode == TRACE_CLOSEST_EMISSIVE)
    RB_TRACE_LAUNCH(false, false, emissive_view(h->sc));
  else
    RB_TRACE_LAUNCH(true, true, h->sc);
#undef RB_TRACE_LAUNCH
  if (!wsel) {
    h->wave.reset_pair = nullptr;
    h->fs.fc.wave.reset_pair = nullptr;
  }
  h->fs.launches++;
  fs_mark(h, pass, 1);
}
// start a new ray queue: switch to the other counter pair (already zero) and let the coming stream kernel zero this one
static void fs_reset_queue(RbContext* h) {
  uint32_t* old_pair = h->waveCounters + 2 * h->wavePair;
  h->wavePair ^= 1;
  h->wave.count = h->waveCounters + 2 * h->wavePair;
  h->wave.reset_pair = old_pair;
  h->fs.fc.wave.count = h->wave.count;
  h->fs.fc.wave.reset_pair = old_pair;
}

// frame_data is about to be overwritten: an asynchronous copy of the previous frame (rb_render_frame_async) must be through
static int wait_frame_copy(RbContext* h) {
  if (!h->copyPending) return RB_OK;
  RB_CUDA(cudaStreamWaitEvent(h->stream, h->evCopyDone[(h->asyncSeq - 1u) % RbContext::kCopyRing], 0));
  h->copyPending = false;
  return RB_OK;
}

// everything a frame may have to ALLOCATE for the current parameters (wavefront queues); frame_begin repeats the call,
// which is then a no-op. rb_multi_render_frame_async runs it on every device before it queues the first kernel.
static uint32_t frame_wave_slots(const RbContext* h, const RbParams& P, uint32_t* brdf_slots, uint32_t* cand_slots, bool* staged) {
  const int H = h->info.height;
  const bool wave = P.wavefront != 0;
  const bool wave_spatial = wave && P.spatialWeightCalc == RB_SW_CONSTANT && (size_t)h->info.width * H < (1u << RB_CAND_INDEX_BITS);
  const uint32_t staged_slots = spatial_staged_slots(P.spatialWeightCalc, (uint32_t)P.spatialReuseNeighborCount + 1u);
  *staged = wave && P.doSpatialReuse && P.spatialWeightCalc != RB_SW_CONSTANT && staged_slots <= 64u &&
            (size_t)staged_slots * h->info.width * H < 0xFFFFFFF0ull && getenv("RB_STAGED_SPATIAL_OFF") == nullptr;
  uint32_t slots = std::max<uint32_t>(std::max<uint32_t>(4u, (uint32_t)P.spatialReuseNeighborCount + 1u), (uint32_t)P.M_Brdf);
  if (*staged) slots = std::max(slots, staged_slots);
  *cand_slots = (wave_spatial && P.doSpatialReuse) ? (uint32_t)P.spatialReuseNeighborCount + 1u : 0u;
  if (P.doTemporalReuse) *cand_slots = std::max(*cand_slots, 2u);
  *brdf_slots = (uint32_t)std::max(P.M_Brdf, 1);
  return slots;
}
static int frame_prepare(RbContext* h) {
  if (!h->haveScene || !h->params.wavefront) return RB_OK;
  RB_CUDA(cudaSetDevice(h->info.device));
  uint32_t brdf_slots = 1, cand_slots = 0;
  bool staged = false;
  const uint32_t slots = frame_wave_slots(h, h->params, &brdf_slots, &cand_slots, &staged);
  return ensure_wave(h, slots, brdf_slots, cand_slots);
}

static int frame_begin(RbHandle h, const RbCamera* cam, uint32_t frame_idx, bool timed) {
  if (!h || !cam) return RB_ERR_INVALID_ARGUMENT;
  if (!h->haveScene) {
    h->err = "rb_render_frame: no scene uploaded";
    return RB_ERR_NO_SCENE;
  }
  RB_CUDA(cudaSetDevice(h->info.device));
  FrameState& F = h->fs;
  F.P = h->params;  // snapshot (the GUI thread may edit the host copy)
  const RbParams& P = F.P;
  const int y0 = h->info.band_y0, y1 = h->info.band_y1, H = h->info.height;
  const bool banded = !(y0 == 0 && y1 == H);
  if (banded && P.doSpatialReuse && (y1 - y0) < spatial_reach(P)) {
    h->err = "rb_render_frame: band is thinner than the spatial reuse reach";
    return RB_ERR_UNSUPPORTED;
  }
  if (banded && P.doTemporalReuse && P.temporalFetchReprojected) {
    h->err = "rb_render_frame: temporalFetchReprojected needs the whole image on one handle (last-frame reservoirs of other bands)";
    return RB_ERR_UNSUPPORTED;
  }
  F.timed = timed;
  F.marks.clear();
  F.launches = 0;
  F.frame_idx = frame_idx;
  F.open = true;
  F.shaded = false;
  // The stream -> trace -> resolve split covers the passes whose rays do not depend on visibility results:
  // BRDF-candidate rays, the visibility pass, temporal reuse, spatial reuse with constant weights.
  F.wave = P.wavefront != 0;
  // (the candidate records of the split spatial pass keep a pixel index in 27 bits)
  F.wave_spatial = F.wave && P.spatialWeightCalc == RB_SW_CONSTANT && (size_t)h->info.width * H < (1u << RB_CAND_INDEX_BITS);
  // ... and, staged, the other spatial MIS modes while their slots per pixel stay moderate (BALANCE_HEURISTIC needs
  // (k + 1)^2 + k + 1: up to k = 6; beyond that the pass traces inline as before)
  {
    uint32_t brdf_slots = 1, cand_slots = 0;
    const uint32_t slots = frame_wave_slots(h, P, &brdf_slots, &cand_slots, &F.wave_spatial_staged);
    if (F.wave) RB_TRY(ensure_wave(h, slots, brdf_slots, cand_slots));
  }
  cudaStream_t st = h->stream;
  // ---- FRONT half: G-buffer + initial candidates of this frame (depends on the camera only) ---------------------
  const int par = (int)(h->frameSeq & 1u);
  cudaStream_t sf = (h->overlap && !timed) ? h->fstream : st;
  // its outputs (G[(gCur+1)%3], R[rFree], the parity's visibility queue and counters) were last used by the back half
  // of the frame before last; its ray queues by the previous front half
  if (h->backRecorded[par]) {
    // at most two frames in flight: the host waits for the back half of the frame before last (whose buffers this
    // front half reuses anyway). Without it a host that never reads a frame back runs hundreds of launches ahead,
    // which measurably slows the banded frame loop (N=8: 606 fps against 791 with one sync per frame, r2 bench).
    if (h->maxFramesInFlight) RB_CUDA(cudaEventSynchronize(h->evBackDone[par]));
    RB_CUDA(cudaStreamWaitEvent(sf, h->evBackDone[par], 0));
  }
  if (h->frontRecorded) RB_CUDA(cudaStreamWaitEvent(sf, h->evFrontDone, 0));
  const bool ovd = h->ovDebug && h->frameSeq < (uint32_t)RbContext::kOvFrames;
  if (ovd) cudaEventRecord(h->evOv[h->frameSeq][0], sf);
  unsigned long long* ctr = h->counters2 + 8 * par;
  h->counters = ctr;
  RB_CUDA(cudaMemsetAsync(ctr, 0, 64, sf));
  uint32_t* brdf_pair = reinterpret_cast<uint32_t*>(ctr + 6);
  uint32_t* vis_pair = reinterpret_cast<uint32_t*>(ctr + 7);
  const int gNew = (h->gCur + 1) % 3;
  // G-buffer rows kept by this handle: the band plus a margin that covers the spatial reach and most reprojections
  static const int margin_env = getenv("RB_GBUF_MARGIN") ? atoi(getenv("RB_GBUF_MARGIN")) : 16;
  const int margin = banded ? std::max(margin_env, spatial_reach(P)) : 0;
  FrameCtx& ff = F.ffc;
  ff = FrameCtx{};
  ff.width = h->info.width;
  ff.height = H;
  ff.y0 = y0;
  ff.y1 = y1;
  ff.sc = h->sc;
  ff.P = P;
  ff.cam = cam_state(cam);
  ff.prevCam = ff.cam;
  ff.G = h->G[gNew];
  ff.Gprev = h->G[gNew];
  ff.Rread = ff.Rlast = ff.Rwrite = h->R[h->rFree];
  ff.frame = h->frame;
  ff.counters = ctr;
  ff.gy0 = std::max(0, y0 - margin);
  ff.gy1 = std::min(H, y1 + margin);
  ff.wave = h->fwave;
  ff.wave.count = brdf_pair;
  ff.wave.chain_count = reinterpret_cast<uint32_t*>(ctr + 5);
  ff.wave.brdf_two_step = (F.wave && h->twoStepBrdf && h->sc.em_n_nodes > 0) ? 1u : 0u;
  if (timed) cudaEventRecord(h->fev[0], st);
  // G-buffer: primary rays are coherent, the inline kernel beats queue + persistent traversal + resolve (measured)
  ff.frame_key = rng_frame_key(h->info.seed, frame_idx, PASS_GBUF, 0);
  launch_rows(h, k_gbuffer, ff.gy0, ff.gy1, 128, &ff, sf);
  fs_mark(h, 0, 0);
  // initial candidates
  ff.frame_key = rng_frame_key(h->info.seed, frame_idx, PASS_INITIAL, 0);
  if (F.wave) {
    if (P.M_Brdf > 0 && h->sc.n_lights > 0) {
      launch_rows(h, k_initial_brdf_stream, y0, y1, kTileW * kTileH, &ff, sf);
      fs_mark(h, 1, 0);
      if (ff.wave.brdf_two_step) {
        // closest EMITTER along each ray (small BVH); the traversal queues "does anything precede it?" for the few rays
        // that found one, and the second launch traces those against the full BVH
        fs_trace(h, TRACE_CLOSEST_EMISSIVE, 1, -1.0f, &ff.wave, sf);
        fs_trace(h, TRACE_ANY_PRECEDES, 1, -1.0f, &ff.wave, sf);
      } else {
        fs_trace(h, TRACE_CLOSEST, 1, -1.0f, &ff.wave, sf);
      }
    }
    // shadow rays of the candidates (visibility pass off) are traced inline by the resolve kernel; with the visibility
    // pass on, the resolve kernel also queues that pass's ray for the reservoir it has just produced
    if (P.doVisibilityPass) {
      ff.wave.rays = h->visRays[par];
      ff.wave.count = vis_pair;
      ff.wave.capacity = (uint32_t)h->visRayCap;
      ff.wave.fuse_vis = 1u;
      launch_rows(h, ff.sc.tri_tan != nullptr ? k_initial_resolve_nmap : k_initial_resolve, y0, y1, 128, &ff, sf);
    } else {
      launch_rows(h, k_initial_resolve_inline_shadow, y0, y1, kTileW * kTileH, &ff, sf);
    }
  } else {
    launch_rows(h, k_initial, y0, y1, kTileW * kTileH, &ff, sf);
  }
  fs_mark(h, 1, 0);
  if (ovd) cudaEventRecord(h->evOv[h->frameSeq][1], sf);
  RB_CUDA(cudaEventRecord(h->evFrontDone, sf));
  h->frontRecorded = true;
  RB_CUDA(cudaGetLastError());

  // ---- BACK half: needs the front half and the previous frame -------------------------------------------------------
  if (ovd) cudaEventRecord(h->evOv[h->frameSeq][2], st);  // (before the wait for the front half: end of the previous back half)
  RB_CUDA(cudaStreamWaitEvent(st, h->evFrontDone, 0));
  h->gCur = gNew;
  {  // the front half's output joins the rotation as "written last"; the buffer it replaces is the next front's target
    const int w = h->rWrite;
    h->rWrite = h->rFree;
    h->rFree = w;
  }
  FrameCtx& fc = F.fc;
  fc = FrameCtx{};
  fc.width = h->info.width;
  fc.height = H;
  fc.y0 = y0;
  fc.y1 = y1;
  fc.sc = h->sc;
  fc.P = P;
  fc.cam = ff.cam;
  fc.prevCam = h->havePrev ? h->prevCam : fc.cam;
  fc.G = h->G[h->gCur];
  fc.Gprev = h->G[(h->gCur + 2) % 3];
  fc.frame = h->frame;
  fc.counters = ctr;
  fc.wave = h->wave;
  fc.wave.deferred_count = reinterpret_cast<uint32_t*>(ctr + 4);
  fc.gy0 = ff.gy0;
  fc.gy1 = ff.gy1;
  fc.gpy0 = h->havePrev ? h->prevGy0 : 0;
  fc.gpy1 = h->havePrev ? h->prevGy1 : 0;

  const bool balancing = h->comm && h->balance;
  const int bslot = (int)(h->balFrame % RbContext::kBalRing);
  if (balancing) {
    h->nStalls[bslot] = 0;
    RB_CUDA(cudaEventRecord(h->evFrameB[bslot], st));
  }
  fs_bind(h);
  // ---- visibility ---------------------------------------------------------------------------------
  const bool temporal_runs = P.doTemporalReuse && frame_idx > 0 && h->havePrev;
  bool vis_in_temporal = false;
  if (P.doVisibilityPass) {
    if (F.wave) {
      WaveBufs vw = h->wave;  // the rays were queued by the front half's k_initial_resolve; results go to the back half's bytes
      vw.rays = h->visRays[par];
      vw.count = vis_pair;
      vw.capacity = (uint32_t)h->visRayCap;
      fs_trace(h, TRACE_ANY, 2, -1.0f, &vw, st);
      // the result is applied by the temporal stream kernel when there is one, else by the pass's own resolve kernel
      vis_in_temporal = temporal_runs;
      if (!vis_in_temporal) launch_rows(h, k_visibility_resolve, y0, y1);
    } else {
      launch_rows(h, k_visibility, y0, y1);
    }
    fs_mark(h, 2, 0);
  }
  // the rows this band may have gained take their last-frame reservoirs from the neighbour: shipped at the previous
  // frame end on the comm stream, needed from here on
  if (balancing && h->shipPending) {
    RB_TRY(wait_recording_stall(h, h->evShipDone));
    h->shipPending = false;
  }
  // ---- temporal reuse ----------------------------------------------------------------------------------
  if (temporal_runs) {
    std::swap(h->rRead, h->rWrite);  // swapReservoirBuffers, P/simpleguidx11.h:116
    fs_bind(h);
    fc.frame_key = rng_frame_key(h->info.seed, frame_idx, PASS_TEMPORAL, 0);
    if (F.wave) {
      fs_reset_queue(h);
      fc.wave.fuse_vis = vis_in_temporal ? 2u : 0u;
      if (banded) {
        launch_rows(h, k_temporal_stream_banded, y0, y1, RB_TS_THREADS);
        k_temporal_stream_deferred<<<h->numSMs, 128, 0, st>>>(fc);
        F.launches++;
      } else {
        launch_rows(h, k_temporal_stream, y0, y1, RB_TS_THREADS);
      }
      fc.wave.fuse_vis = 0u;
      fs_mark(h, 3, 0);
      fs_trace(h, TRACE_ANY, 3);
      launch_rows(h, k_temporal_resolve, y0, y1, RB_TR_THREADS);
    } else {
      if (banded)
        launch_rows(h, k_temporal_banded, y0, y1);
      else
        launch_rows(h, k_temporal, y0, y1);
    }
    fs_mark(h, 3, 0);
  }
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}


// One spatial pass. With `overlap_halo` the interior rows (which need no halo) are streamed while the halo rows
// of the neighbouring bands are in flight; the boundary rows follow once they have arrived.
static int frame_spatial(RbHandle h, int i, bool overlap_halo) {
  if (!h || !h->fs.open) return RB_ERR_INVALID_ARGUMENT;
  FrameState& F = h->fs;
  const RbParams& P = F.P;
  FrameCtx& fc = F.fc;
  const int y0 = h->info.band_y0, y1 = h->info.band_y1;
  const int R = spatial_reach(P);
  const int iy0 = (overlap_halo && y0 > 0) ? std::min(y0 + R, y1) : y0;
  const int iy1 = (overlap_halo && y1 < h->info.height) ? std::max(y1 - R, iy0) : y1;
  if (overlap_halo) RB_TRY(halo_exchange_begin(h));
  std::swap(h->rRead, h->rWrite);
  fs_bind(h);
  fc.spatial_iter = i;
  fc.frame_key = rng_frame_key(h->info.seed, F.frame_idx, PASS_SPATIAL, (uint32_t)i);
  F.shaded = false;
  if (F.wave_spatial) {
    fs_reset_queue(h);
    // the last pass's resolve kernel shades the pixel from the reservoir it holds in registers
    F.shaded = (i == P.spatialPassCount - 1);
    fc.wave.fuse_shade = F.shaded ? 1u : 0u;
    auto kss = k_spatial_stream;
    launch_rows(h, kss, iy0, iy1, RB_SS_THREADS);
    if (overlap_halo) {
      RB_TRY(halo_exchange_wait(h));
      launch_rows(h, kss, y0, iy0, RB_SS_THREADS);
      launch_rows(h, kss, iy1, y1, RB_SS_THREADS);
    }
    fs_mark(h, 4, 0);
    fs_trace(h, TRACE_ANY, 4);
    if (F.shaded) RB_TRY(wait_frame_copy(h));
    launch_rows(h, k_spatial_resolve, y0, y1, RB_SR_THREADS);
    fc.wave.fuse_shade = 0u;
  } else if (F.wave_spatial_staged) {
    // stage 1 (rays asked for before the selection) -> trace -> [debias modes: stage 2 (rays that depend on the
    // selection) -> trace] -> stage 3 (everything answered, reservoir stored)
    fs_reset_queue(h);
    launch_rows(h, k_spatial_mis_stage1, iy0, iy1);
    if (overlap_halo) {
      RB_TRY(halo_exchange_wait(h));
      launch_rows(h, k_spatial_mis_stage1, y0, iy0);
      launch_rows(h, k_spatial_mis_stage1, iy1, y1);
    }
    fs_mark(h, 4, 0);
    fs_trace(h, TRACE_ANY, 4);
    if (P.spatialWeightCalc == RB_SW_CONSTANT_DEBIAS_Z_TERM || P.spatialWeightCalc == RB_SW_CONSTANT_DEBIAS_CONTRIB) {
      fs_reset_queue(h);
      launch_rows(h, k_spatial_mis_stage2, y0, y1);
      fs_mark(h, 4, 0);
      fs_trace(h, TRACE_ANY, 4);
    }
    launch_rows(h, k_spatial_mis_stage3, y0, y1);
  } else {
    launch_rows(h, k_spatial, iy0, iy1);
    if (overlap_halo) {
      RB_TRY(halo_exchange_wait(h));
      launch_rows(h, k_spatial, y0, iy0);
      launch_rows(h, k_spatial, iy1, y1);
    }
  }
  fs_mark(h, 4, 0);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

static int frame_end(RbHandle h, RbTimings* timings) {
  if (!h || !h->fs.open) return RB_ERR_INVALID_ARGUMENT;
  FrameState& F = h->fs;
  FrameCtx& fc = F.fc;
  cudaStream_t st = h->stream;
  std::swap(h->rRead, h->rWrite);
  fs_bind(h);
  if (!F.shaded) {
    RB_TRY(wait_frame_copy(h));
    launch_rows(h, k_shade, h->info.band_y0, h->info.band_y1);
  }
  F.shaded = false;
  fs_mark(h, 5, 0);
  RB_CUDA(cudaGetLastError());
  // memcpy(reservoirsLastFrame, ...) + gBufferLastFrame.setDataFrom(gBuffer) (P/simpleguidx11.cpp:478-481) by rotation
  std::swap(h->rLast, h->rRead);
  h->prevCam = fc.cam;
  h->prevGy0 = fc.gy0;
  h->prevGy1 = fc.gy1;
  h->havePrev = true;
  F.open = false;
  {  // the front half of the frame after next may reuse this frame's buffers once this point is reached
    const int par = (int)(h->frameSeq & 1u);
    if (h->ovDebug && h->frameSeq < (uint32_t)RbContext::kOvFrames) cudaEventRecord(h->evOv[h->frameSeq][3], st);
    RB_CUDA(cudaEventRecord(h->evBackDone[par], st));
    h->backRecorded[par] = true;
    h->frameSeq++;
  }
  if (h->comm && h->balance) {
    RB_CUDA(cudaEventRecord(h->evFrameE[h->balFrame % RbContext::kBalRing], st));
    RB_TRY(balance_ship_begin(h));
    h->balFrame++;
  }

  if (timings) {
    memset(timings, 0, sizeof(*timings));
    unsigned long long hc[8];
    RB_CUDA(cudaMemcpyAsync(hc, h->counters, 64, cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    RB_TRY(halo_check(h));
    timings->rays_closest = hc[0];
    timings->rays_any_as_written = hc[1];
    timings->rays_any_traced = hc[2];
    timings->kernel_launches = F.launches;
    if (F.timed) {
      float* per_pass[6] = {&timings->ms_gbuffer, &timings->ms_initial, &timings->ms_visibility,
                            &timings->ms_temporal, &timings->ms_spatial, &timings->ms_shade};
      for (size_t i = 0; i < F.marks.size(); ++i) {
        float ms = 0;
        RB_CUDA(cudaEventElapsedTime(&ms, h->fev[i], h->fev[i + 1]));
        *per_pass[F.marks[i].pass] += ms;
        if (F.marks[i].kind == 1) {
          timings->ms_trace[F.marks[i].pass] += ms;
          timings->ms_trace_any += ms;
        } else {
          timings->ms_stream[F.marks[i].pass] += ms;
        }
      }
      RB_CUDA(cudaEventElapsedTime(&timings->ms_total, h->fev[0], h->fev[F.marks.size()]));
      if (h->comm && h->haloTimed) RB_CUDA(cudaEventElapsedTime(&timings->ms_halo, h->evHaloT0, h->evHaloDone));
    }
  }
  return RB_OK;
}

static int render_frame_impl(RbHandle h, const RbCamera* cam, uint32_t frame_idx, RbTimings* timings) {
  if (!h) return RB_ERR_INVALID_ARGUMENT;
  const bool timed = h->info.collect_timings != 0 && timings != nullptr;
  RB_TRY(balance_update_band(h));
  RB_TRY(frame_begin(h, cam, frame_idx, timed));
  if (h->fs.P.doSpatialReuse)
    for (int i = 0; i < h->fs.P.spatialPassCount; ++i) RB_TRY(frame_spatial(h, i, h->comm != nullptr || h->localGroup));
  return frame_end(h, timings);
}

extern "C" {

int rb_frame_begin(RbHandle h, const RbCamera* cam, uint32_t frame_idx) {
  return frame_begin(h, cam, frame_idx, h && h->info.collect_timings != 0);
}
int rb_frame_spatial(RbHandle h, int32_t pass_index) { return frame_spatial(h, pass_index, false); }
int rb_get_band(RbHandle h, int32_t* y0, int32_t* y1) {
  if (!h || !y0 || !y1) return RB_ERR_INVALID_ARGUMENT;
  *y0 = h->info.band_y0;
  *y1 = h->info.band_y1;
  return RB_OK;
}
int rb_set_band(RbHandle h, int32_t y0, int32_t y1) {
  if (!h || y0 < 0 || y1 > h->info.height || y0 >= y1) return RB_ERR_INVALID_ARGUMENT;
  if (h->fs.open) {
    h->err = "rb_set_band: a frame is open";
    return RB_ERR_INVALID_ARGUMENT;
  }
  // the temporal pass reads last frame's G-buffer at its own pixels without a fallback: stay inside what was rendered
  if (h->havePrev && (y0 < h->prevGy0 || y1 > h->prevGy1)) {
    h->err = "rb_set_band: the band may only move within the rows whose G-buffer this handle rendered last frame";
    return RB_ERR_UNSUPPORTED;
  }
  h->info.band_y0 = y0;
  h->info.band_y1 = y1;
  return RB_OK;
}
int rb_frame_end(RbHandle h, float* frame_rgb_out, RbTimings* timings) {
  int rc = frame_end(h, timings);
  if (rc != RB_OK) return rc;
  if (frame_rgb_out) {
    const size_t row = (size_t)h->info.width * 3;
    const size_t off = row * h->info.band_y0, cnt = row * (h->info.band_y1 - h->info.band_y0);
    RB_CUDA(cudaMemcpyAsync(frame_rgb_out + off, h->frame + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RB_CUDA(cudaStreamSynchronize(h->stream));
    RB_TRY(halo_check(h));
  }
  return RB_OK;
}

int rb_render_frame(RbHandle h, const RbCamera* cam, uint32_t frame_idx, float* frame_rgb_out, RbTimings* timings) {
  int rc = render_frame_impl(h, cam, frame_idx, timings);
  if (rc != RB_OK) return rc;
  if (frame_rgb_out) {
    const size_t row = (size_t)h->info.width * 3;
    const size_t off = row * h->info.band_y0, cnt = row * (h->info.band_y1 - h->info.band_y0);
    RB_CUDA(cudaMemcpyAsync(frame_rgb_out + off, h->frame + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RB_CUDA(cudaStreamSynchronize(h->stream));
    RB_TRY(halo_check(h));
  }
  return RB_OK;
}

// Pipelined form of rb_render_frame (include/restir_b200.h): the frame is issued, its rows are copied to the host on a
// copy stream behind the back half, and the call returns. The next frame's kernels overlap the copy.
int rb_render_frame_async(RbHandle h, const RbCamera* cam, uint32_t frame_idx, float* frame_rgb_out) {
  if (!h || !frame_rgb_out) return RB_ERR_INVALID_ARGUMENT;
  if (!h->copyStream) {
    RB_CUDA(cudaSetDevice(h->info.device));
    RB_CUDA(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
    for (auto& ev : h->evCopyDone) RB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    RB_CUDA(cudaEventCreateWithFlags(&h->evFrameReady, cudaEventDisableTiming));
  }
  int rc = render_frame_impl(h, cam, frame_idx, nullptr);
  if (rc != RB_OK) return rc;
  const size_t row = (size_t)h->info.width * 3;
  const size_t off = row * h->info.band_y0, cnt = row * (h->info.band_y1 - h->info.band_y0);
  RB_CUDA(cudaEventRecord(h->evFrameReady, h->stream));
  RB_CUDA(cudaStreamWaitEvent(h->copyStream, h->evFrameReady, 0));
  RB_CUDA(cudaMemcpyAsync(frame_rgb_out + off, h->frame + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, h->copyStream));
  RB_CUDA(cudaEventRecord(h->evCopyDone[h->asyncSeq % RbContext::kCopyRing], h->copyStream));
  h->asyncSeq++;
  h->copyPending = true;
  return RB_OK;
}
int rb_frame_wait(RbHandle h, uint32_t frames_in_flight) {
  if (!h) return RB_ERR_INVALID_ARGUMENT;
  if (!h->copyStream || h->asyncSeq <= frames_in_flight) return RB_OK;
  if (frames_in_flight >= (uint32_t)RbContext::kCopyRing) {
    h->err = "rb_frame_wait: at most 3 frames can be left in flight";
    return RB_ERR_INVALID_ARGUMENT;
  }
  RB_CUDA(cudaSetDevice(h->info.device));
  RB_CUDA(cudaEventSynchronize(h->evCopyDone[(h->asyncSeq - 1u - frames_in_flight) % RbContext::kCopyRing]));
  return halo_check(h);
}

int rb_render_frame_device(RbHandle h, const RbCamera* cam, uint32_t frame_idx, float* frame_rgb_dev, RbTimings* timings) {
  int rc = render_frame_impl(h, cam, frame_idx, timings);
  if (rc != RB_OK) return rc;
  if (frame_rgb_dev) {
    const size_t row = (size_t)h->info.width * 3;
    const size_t off = row * h->info.band_y0, cnt = row * (h->info.band_y1 - h->info.band_y0);
    RB_CUDA(cudaMemcpyAsync(frame_rgb_dev + off, h->frame + off, cnt * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  }
  return RB_OK;
}

// N2: the reference's ground-truth estimator on the same boundary. Independent of the ReSTIR state: it
// renders its own G-buffer into the buffer the next frame's front half will overwrite anyway, and writes frame_data
// (so rb_accumulate_display converges it exactly like the reference's Producer loop does).
int rb_render_mis_frame(RbHandle h, const RbCamera* cam, uint32_t frame_idx, uint32_t techniques, float* frame_rgb_out) {
  if (!h || !cam || (techniques & ~3u)) return RB_ERR_INVALID_ARGUMENT;
  if (!h->haveScene) {
    h->err = "rb_render_mis_frame: no scene uploaded";
    return RB_ERR_NO_SCENE;
  }
  if (h->fs.open) {
    h->err = "rb_render_mis_frame: a ReSTIR frame is open (rb_frame_begin without rb_frame_end)";
    return RB_ERR_INVALID_ARGUMENT;
  }
  if (h->hasDielectric) {
    h->err = "rb_render_mis_frame: dielectric materials are not covered by the MIS ground-truth estimator";
    return RB_ERR_UNSUPPORTED;
  }
  RB_CUDA(cudaSetDevice(h->info.device));
  if (h->frontRecorded) RB_CUDA(cudaStreamWaitEvent(h->stream, h->evFrontDone, 0));
  const int y0 = h->info.band_y0, y1 = h->info.band_y1;
  FrameCtx fc{};
  fc.width = h->info.width;
  fc.height = h->info.height;
  fc.y0 = y0;
  fc.y1 = y1;
  fc.sc = h->sc;
  fc.P = h->params;
  fc.cam = cam_state(cam);
  fc.prevCam = fc.cam;
  fc.G = fc.Gprev = h->G[(h->gCur + 1) % 3];
  fc.Rread = fc.Rwrite = fc.Rlast = h->R[h->rFree];
  fc.frame = h->frame;
  unsigned long long* ctr = h->counters2 + 8 * (h->frameSeq & 1u);
  // (the parity block of the NEXT ReSTIR frame: its front half zeroes it again, and this stream is idle by then)
  fc.counters = ctr;
  fc.gy0 = y0, fc.gy1 = y1;
  fc.mis_flags = techniques;
  RB_CUDA(cudaMemsetAsync(ctr, 0, 64, h->stream));
  fc.frame_key = rng_frame_key(h->info.seed, frame_idx, PASS_GBUF, 0);
  launch_rows(h, k_gbuffer, y0, y1, 128, &fc);
  fc.frame_key = rng_frame_key(h->info.seed, frame_idx, PASS_MIS, 0);
  RB_TRY(wait_frame_copy(h));
  launch_rows(h, k_mis_direct, y0, y1, kTileW * kTileH, &fc);
  RB_CUDA(cudaGetLastError());
  // the next ReSTIR front half (other stream) must not start on this G-buffer before the MIS kernels are done
  RB_CUDA(cudaEventRecord(h->evFrontDone, h->stream));
  h->frontRecorded = true;
  if (frame_rgb_out) {
    const size_t off = (size_t)y0 * h->info.width * 3, cnt = (size_t)(y1 - y0) * h->info.width * 3;
    RB_CUDA(cudaMemcpyAsync(frame_rgb_out + off, h->frame + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RB_CUDA(cudaStreamSynchronize(h->stream));
  }
  return RB_OK;
}
int rb_synchronize(RbHandle h) {
  if (!h) return RB_ERR_INVALID_ARGUMENT;
  RB_CUDA(cudaSetDevice(h->info.device));
  RB_CUDA(cudaStreamSynchronize(h->stream));
  return halo_check(h);
}

int rb_timer_begin(RbHandle h) {
  if (!h) return RB_ERR_INVALID_ARGUMENT;
  RB_CUDA(cudaSetDevice(h->info.device));
  RB_CUDA(cudaEventRecord(h->ev[12], h->stream));
  return RB_OK;
}
int rb_timer_end(RbHandle h, float* ms_out) {
  if (!h || !ms_out) return RB_ERR_INVALID_ARGUMENT;
  RB_CUDA(cudaSetDevice(h->info.device));
  RB_CUDA(cudaEventRecord(h->ev[13], h->stream));
  RB_CUDA(cudaEventSynchronize(h->ev[13]));
  RB_CUDA(cudaEventElapsedTime(ms_out, h->ev[12], h->ev[13]));
  return RB_OK;
}

int rb_readback(RbHandle h, int id, void* dst, size_t bytes) {
  if (!h || !dst) return RB_ERR_INVALID_ARGUMENT;
  RB_CUDA(cudaSetDevice(h->info.device));
  const size_t n = (size_t)h->info.width * h->info.height;
  const GBufPlanes& G = h->G[h->gCur];  // the frame just rendered
  const ResPlanes& R = h->R[h->rLast];
  const void* src = nullptr;
  size_t need = 0;
  switch (id) {
    case RB_BUF_GBUF_POS_DEPTH: src = G.pos_depth, need = n * 16; break;
    case RB_BUF_GBUF_NORMAL_SHIN: src = G.normal_shin, need = n * 16; break;
    case RB_BUF_GBUF_DIFFUSE_IIM: src = G.diffuse_iim, need = n * 16; break;
    case RB_BUF_GBUF_SPEC_TYPE: src = G.spec_type, need = n * 16; break;
    case RB_BUF_GBUF_EMISSION: src = G.emission, need = n * 16; break;
    case RB_BUF_HIT_IDS: src = G.hit_ids, need = n * 8; break;
    case RB_BUF_RES_POINT_WSUM: src = R.point_wsum, need = n * 16; break;
    case RB_BUF_RES_NORMAL_W: src = R.normal_W, need = n * 16; break;
    case RB_BUF_RES_LI_CONF: src = R.Li_conf, need = n * 16; break;
    case RB_BUF_RES_LIGHT_IDX: src = R.light_idx, need = n * 4; break;
    case RB_BUF_FRAME_RGB: src = h->frame, need = n * 12; break;
    case RB_BUF_ALIAS_PROB: src = h->sc.alias_prob, need = (size_t)h->sc.n_lights * 4; break;
    case RB_BUF_ALIAS_IDX: src = h->sc.alias_idx, need = (size_t)h->sc.n_lights * 4; break;
    case RB_BUF_LIGHT_CDF: src = h->sc.cdf, need = (size_t)h->sc.n_lights * 4; break;
    case RB_BUF_ACCUMULATOR: src = h->accumulator, need = n * 12; break;
    case RB_BUF_DISPLAY: src = h->display, need = n * 16; break;
    default: h->err = "rb_readback: unknown buffer id"; return RB_ERR_INVALID_ARGUMENT;
  }
  if (id >= RB_BUF_ALIAS_PROB && id <= RB_BUF_LIGHT_CDF && !h->haveScene) {
    h->err = "rb_readback: no scene";
    return RB_ERR_NO_SCENE;
  }
  if (bytes < need) {
    h->err = "rb_readback: destination too small";
    return RB_ERR_INVALID_ARGUMENT;
  }
  if (need) RB_CUDA(cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, h->stream));
  RB_CUDA(cudaStreamSynchronize(h->stream));
  return halo_check(h);
}

// -------------------------------------------------------------------------------------
// after the path: accumulate / tonemap / statistics
// -------------------------------------------------------------------------------------
int rb_accumulate_display(RbHandle h, uint32_t acc_frame_ctr, int32_t tonemap, int32_t gamma_correct, float* display_rgba_out,
                          RbImageStats* stats) {
  if (!h) return RB_ERR_INVALID_ARGUMENT;
  RB_CUDA(cudaSetDevice(h->info.device));
  const int W = h->info.width, y0 = h->info.band_y0, y1 = h->info.band_y1;
  const size_t count = (size_t)(y1 - y0) * W;
  const float mix_a = 1.0f / (float)(acc_frame_ctr + 1u);  // 1.0f / static_cast<float>(accFrameCtr + 1), :251
  RB_CUDA(cudaMemsetAsync(h->statSums, 0, 16, h->stream));
  k_accumulate_display<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(h->frame, h->accumulator, h->display, W, y0, y1, mix_a,
                                                                              tonemap, gamma_correct, h->statSums);
  RB_CUDA(cudaGetLastError());
  if (display_rgba_out) {
    const size_t off = (size_t)y0 * W * 4;
    RB_CUDA(cudaMemcpyAsync(display_rgba_out + off, (float*)h->display + off, count * 16, cudaMemcpyDeviceToHost, h->stream));
  }
  if (stats) {
    double s[2];
    RB_CUDA(cudaMemcpyAsync(s, h->statSums, 16, cudaMemcpyDeviceToHost, h->stream));
    RB_CUDA(cudaStreamSynchronize(h->stream));
    stats->sum = s[0];
    stats->sum_sq = s[1];
    stats->pixels = count;
    stats->mean = s[0] / (double)count;
    stats->variance = s[1] / (double)count - stats->mean * stats->mean;  // D(X) = E(X^2) - E(X)^2, :324
  } else if (display_rgba_out) {
    RB_CUDA(cudaStreamSynchronize(h->stream));
  }
  return RB_OK;
}

// -------------------------------------------------------------------------------------
// ray seam
// -------------------------------------------------------------------------------------
static int trace_device(RbHandle h, const RbRay* rays, void* out, uint32_t n, bool any, float* ms_out) {
  if (!h || (n && (!rays || !out))) return RB_ERR_INVALID_ARGUMENT;
  if (!h->haveScene) {
    h->err = "rb_trace: no scene uploaded";
    return RB_ERR_NO_SCENE;
  }
  RB_CUDA(cudaSetDevice(h->info.device));
  if (n == 0) {
    if (ms_out) *ms_out = 0;
    return RB_OK;
  }
  // the same persistent, phase-scheduled traversal kernel as the frame's ray queue, reading RTCRay records
  RB_TRY(ensure_counters(h));
  uint32_t* seam_pair = h->waveCounters + 4;
  if (ms_out) RB_CUDA(cudaEventRecord(h->ev[14], h->stream));
  k_reset_queue<<<1, 1, 0, h->stream>>>(seam_pair, seam_pair + 1);
  const int trace_grid = h->numSMs * h->traceBlocksPerSM;
  const SeamIO io{rays, n, (uint8_t*)out, (RbHit*)out};
  const int sn = h->smemStackEntries;
  const size_t sb = (size_t)sn * kTraceThreads * sizeof(U2);
  if (any && sn > 0)
    k_trace_queue<true, false, SeamIO, true><<<trace_grid, kTraceThreads, sb, h->stream>>>(h->sc, io, seam_pair + 1, h->refillLanes, h->postponeLanes, sn);
  else if (any)
    k_trace_queue<true, false, SeamIO, false><<<trace_grid, kTraceThreads, 0, h->stream>>>(h->sc, io, seam_pair + 1, h->refillLanes, h->postponeLanes, 0);
  else if (sn > 0)
    k_trace_queue<false, false, SeamIO, true><<<trace_grid, kTraceThreads, sb, h->stream>>>(h->sc, io, seam_pair + 1, h->refillLanes, h->postponeLanes, sn);
  else
    k_trace_queue<false, false, SeamIO, false><<<trace_grid, kTraceThreads, 0, h->stream>>>(h->sc, io, seam_pair + 1, h->refillLanes, h->postponeLanes, 0);
  RB_CUDA(cudaGetLastError());
  if (ms_out) {
    RB_CUDA(cudaEventRecord(h->ev[15], h->stream));
    RB_CUDA(cudaEventSynchronize(h->ev[15]));
    RB_CUDA(cudaEventElapsedTime(ms_out, h->ev[14], h->ev[15]));
  }
  return RB_OK;
}
static int trace_host(RbHandle h, const RbRay* rays, void* out, size_t out_elem, uint32_t n, bool any) {
  if (!h || (n && (!rays || !out))) return RB_ERR_INVALID_ARGUMENT;
  if (n == 0) return RB_OK;
  RB_CUDA(cudaSetDevice(h->info.device));
  RbRay* d_rays = nullptr;
  void* d_out = nullptr;
  RB_CUDA(cudaMalloc(&d_rays, (size_t)n * sizeof(RbRay)));
  cudaError_t e = cudaMalloc(&d_out, (size_t)n * out_elem);
  if (e != cudaSuccess) {
    cudaFree(d_rays);
    h->err = "rb_trace: out of device memory";
    return RB_ERR_OUT_OF_MEMORY;
  }
  int rc = RB_OK;
  auto body = [&]() -> int {
    RB_CUDA(cudaMemcpyAsync(d_rays, rays, (size_t)n * sizeof(RbRay), cudaMemcpyHostToDevice, h->stream));
    RB_TRY(trace_device(h, d_rays, d_out, n, any, nullptr));
    RB_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * out_elem, cudaMemcpyDeviceToHost, h->stream));
    RB_CUDA(cudaStreamSynchronize(h->stream));
    return RB_OK;
  };
  rc = body();
  cudaFree(d_rays);
  cudaFree(d_out);
  return rc;
}
int rb_trace_closest(RbHandle h, const RbRay* rays, RbHit* hits, uint32_t n) {
  return trace_host(h, rays, hits, sizeof(RbHit), n, false);
}
int rb_trace_occluded(RbHandle h, const RbRay* rays, uint8_t* occluded, uint32_t n) {
  return trace_host(h, rays, occluded, 1, n, true);
}
int rb_trace_closest_device(RbHandle h, const RbRay* rays_dev, RbHit* hits_dev, uint32_t n, float* ms_out) {
  return trace_device(h, rays_dev, hits_dev, n, false, ms_out);
}
int rb_trace_occluded_device(RbHandle h, const RbRay* rays_dev, uint8_t* occ_dev, uint32_t n, float* ms_out) {
  return trace_device(h, rays_dev, occ_dev, n, true, ms_out);
}

// -------------------------------------------------------------------------------------
// multi-GPU bands: reservoir halo rows
// -------------------------------------------------------------------------------------
size_t rb_halo_bytes(RbHandle h, int32_t rows) {
  if (!h || rows < 0) return 0;
  return (size_t)rows * h->info.width * 52;
}
// rows [y, y+rows) of the reservoirs the next spatial pass reads (R[rWrite] between frame phases)
static int halo_copy(RbHandle h, int32_t y, int32_t rows, void* host, bool to_host) {
  if (!h || !host || y < 0 || rows < 0 || y + rows > h->info.height) return RB_ERR_INVALID_ARGUMENT;
  RB_CUDA(cudaSetDevice(h->info.device));
  const ResPlanes& R = h->R[h->fs.open ? h->rWrite : h->rLast];
  const size_t px = (size_t)rows * h->info.width, off = (size_t)y * h->info.width;
  char* p = (char*)host;
  void* planes[4] = {R.point_wsum + off, R.normal_W + off, R.Li_conf + off, R.light_idx + off};
  const size_t sz[4] = {px * 16, px * 16, px * 16, px * 4};
  for (int i = 0; i < 4; ++i) {
    if (to_host)
      RB_CUDA(cudaMemcpyAsync(p, planes[i], sz[i], cudaMemcpyDeviceToHost, h->stream));
    else
      RB_CUDA(cudaMemcpyAsync(planes[i], p, sz[i], cudaMemcpyHostToDevice, h->stream));
    p += sz[i];
  }
  RB_CUDA(cudaStreamSynchronize(h->stream));
  return RB_OK;
}
int rb_halo_export(RbHandle h, int32_t y, int32_t rows, void* dst_host) { return halo_copy(h, y, rows, dst_host, true); }
int rb_halo_import(RbHandle h, int32_t y, int32_t rows, const void* src_host) {
  return halo_copy(h, y, rows, const_cast<void*>(src_host), false);
}
int32_t rb_halo_rows(RbHandle h) { return h ? spatial_reach(h->fs.open ? h->fs.P : h->params) : 0; }

int rb_comm_unique_id(void* out_id, size_t id_bytes) {
  if (!out_id || id_bytes < 128) return RB_ERR_INVALID_ARGUMENT;
  std::string err;
  if (!load_nccl(err)) {
    g_create_error = err;
    return RB_ERR_COMM;
  }
  return g_nccl.GetUniqueId(out_id) == 0 ? RB_OK : RB_ERR_COMM;
}
// host-only arithmetic of the band balancer, exported so that it can be tested without GPUs: `pairs` = {cost, rows} per
// rank (rows must tile [0, height)); bounds_out gets the n + 1 new boundaries every rank would compute
// The frame's ray queues, for measurements on the rays the path REALLY traces (BASELINE configs[3]): which = 0 the back
// half's current queue (after rb_frame_begin: the temporal pass's rays; after rb_frame_spatial(i): that pass's), 1 = the
// visibility rays of the open frame. Rays come out in RTCRay layout, queue order. Only between the phases of an open
// frame (rb_frame_begin .. rb_frame_end), wavefront mode.
int rb_debug_ray_queue(RbHandle h, int32_t which, RbRay* dst, uint32_t capacity, uint32_t* count_out) {
  if (!h || !count_out || which < 0 || which > 1) return RB_ERR_INVALID_ARGUMENT;
  if (!h->fs.open || !h->fs.wave) {
    h->err = "rb_debug_ray_queue: needs an open frame in wavefront mode";
    return RB_ERR_INVALID_ARGUMENT;
  }
  RB_CUDA(cudaSetDevice(h->info.device));
  RB_CUDA(cudaStreamSynchronize(h->stream));
  const int par = (int)(h->frameSeq & 1u);
  const RayQ* src = which == 0 ? h->wave.rays : h->visRays[par];
  const uint32_t* cnt = which == 0 ? h->wave.count : reinterpret_cast<const uint32_t*>(h->counters2 + 8 * par + 7);
  const uint32_t cap = which == 0 ? (uint32_t)h->waveRayCap : (uint32_t)h->visRayCap;
  uint32_t n = 0;
  RB_CUDA(cudaMemcpy(&n, cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  n = std::min(n, cap);
  *count_out = n;
  if (!dst || n == 0) return RB_OK;
  const uint32_t m = std::min(n, capacity);
  std::vector<RayQ> tmp(m);
  RB_CUDA(cudaMemcpy(tmp.data(), src, (size_t)m * sizeof(RayQ), cudaMemcpyDeviceToHost));
  const float tnear = FLT_MIN + h->fs.P.tnearOffset;
  for (uint32_t i = 0; i < m; ++i) {
    RbRay r;
    memset(&r, 0, sizeof(r));
    r.org_x = tmp[i].o_tfar.x, r.org_y = tmp[i].o_tfar.y, r.org_z = tmp[i].o_tfar.z;
    r.dir_x = tmp[i].d_dest.x, r.dir_y = tmp[i].d_dest.y, r.dir_z = tmp[i].d_dest.z;
    r.tnear = tnear;
    r.tfar = tmp[i].o_tfar.w;
    r.mask = 0xFFFFFFFFu;
    dst[i] = r;
  }
  return RB_OK;
}
int rb_debug_balance_step(const float* pairs, int32_t n_ranks, int32_t height, int32_t* bounds_out) {
  if (!pairs || !bounds_out || n_ranks < 1 || n_ranks > RbContext::kMaxRanks) return RB_ERR_INVALID_ARGUMENT;
  bounds_out[0] = 0;
  for (int r = 0; r < n_ranks; ++r) bounds_out[r + 1] = bounds_out[r] + (int)pairs[2 * r + 1];
  if (bounds_out[n_ranks] != height) return RB_ERR_INVALID_ARGUMENT;
  balance_targets(pairs, n_ranks, height, bounds_out);
  return RB_OK;
}
// ---- scene ingestion (first half): Wavefront OBJ / MTL with the reference's conventions ------------
struct RbObjScene {
  rbobj::Scene sc;
};
int rb_obj_load(const char* obj_path, int32_t gamma_correct, RbObjScene** out, char* err, size_t err_bytes) {
  if (!obj_path || !out) return RB_ERR_INVALID_ARGUMENT;
  RbObjScene* s = new RbObjScene();
  std::string e;
  if (!rbobj::load_obj(obj_path, gamma_correct != 0, s->sc, e)) {
    if (err && err_bytes) {
      strncpy(err, e.c_str(), err_bytes - 1);
      err[err_bytes - 1] = 0;
    }
    delete s;
    return RB_ERR_INVALID_ARGUMENT;
  }
  *out = s;
  return RB_OK;
}
const RbSceneDesc* rb_obj_scene_desc(const RbObjScene* s) { return s ? &s->sc.desc : nullptr; }
const char* rb_obj_material_name(const RbObjScene* s, uint32_t i) {
  return (s && i < s->sc.material_names.size()) ? s->sc.material_names[i].c_str() : nullptr;
}
const char* rb_obj_texture_name(const RbObjScene* s, uint32_t i, int32_t slot) {
  if (!s || i >= s->sc.materials.size()) return nullptr;
  const std::vector<std::string>* v[4] = {&s->sc.map_kd, &s->sc.map_ks, &s->sc.map_ns, &s->sc.map_kn};
  return (slot >= 0 && slot < 4) ? (*v[slot])[i].c_str() : nullptr;
}
void rb_obj_free(RbObjScene* s) { delete s; }

int32_t rb_comm_transport(RbHandle h) { return !h ? 0 : (h->localGroup ? 1 : (!h->comm ? 0 : (h->p2p ? 1 : 2))); }
static int comm_init_body(RbHandle h, int32_t rank, int32_t nranks, const void* nccl_unique_id) {
  RB_CUDA(cudaSetDevice(h->info.device));
  NcclId128 id;
  memcpy(id.b, nccl_unique_id, 128);
  RB_NCCL(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
  h->commRank = rank;
  h->commSize = nranks;
  {
    int lo = 0, hi = 0;  // highest priority: the exchange kernels should not queue behind a frame kernel's CTAs
    RB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const int prio = getenv("RB_COMM_PRIO") ? atoi(getenv("RB_COMM_PRIO")) : hi;
    RB_CUDA(cudaStreamCreateWithPriority(&h->commStream, cudaStreamNonBlocking, prio));
  }
  RB_CUDA(cudaEventCreateWithFlags(&h->evHaloReady, cudaEventDisableTiming));
  RB_CUDA(cudaEventCreate(&h->evHaloDone));
  RB_CUDA(cudaEventCreate(&h->evHaloT0));
  h->balance = true;
  if (const char* e = getenv("RB_BALANCE")) h->balance = atoi(e) != 0;
  if (const char* e = getenv("RB_BAL_PERIOD")) h->balPeriod = std::max(1, atoi(e));
  if (const char* e = getenv("RB_BAL_DEBUG")) h->balDebug = atoi(e) != 0;
  RB_CUDA(cudaEventCreateWithFlags(&h->evShipReady, cudaEventDisableTiming));
  RB_CUDA(cudaEventCreateWithFlags(&h->evShipDone, cudaEventDisableTiming));
  RB_CUDA(cudaEventCreateWithFlags(&h->evAccMoved, cudaEventDisableTiming));
  for (int i = 0; i < RbContext::kBalRing; ++i) {
    RB_CUDA(cudaEventCreateWithFlags(&h->evBalCopied[i], cudaEventDisableTiming));
    RB_CUDA(cudaEventCreate(&h->evFrameB[i]));
    RB_CUDA(cudaEventCreate(&h->evFrameE[i]));
    for (int j = 0; j < RbContext::kMaxStalls; ++j) {
      RB_CUDA(cudaEventCreate(&h->evStallA[i][j]));
      RB_CUDA(cudaEventCreate(&h->evStallB[i][j]));
    }
  }
  RB_CUDA(cudaMalloc((void**)&h->balDev, (2 + 2 * RbContext::kMaxRanks) * sizeof(float)));
  RB_CUDA(cudaMallocHost((void**)&h->balHost, RbContext::kBalRing * (2 + 2 * RbContext::kMaxRanks) * sizeof(float)));
  h->balFrame = 0;
  h->shipPending = false;
  {  // The balancer ships kShipRows rows across every boundary with matching send / recv counts, which needs every band
     // to be at least that thick (and it refuses to move bands thinner than 2 * kShipRows + 2 * 15 anyway): decided
     // once, for all ranks together — either every rank balances or none does.
    int rows = h->info.band_y1 - h->info.band_y0, min_rows = 0;
    int* d = reinterpret_cast<int*>(h->balDev);
    RB_CUDA(cudaMemcpyAsync(d, &rows, sizeof(int), cudaMemcpyHostToDevice, h->commStream));
    RB_NCCL(g_nccl.AllReduce(d, d, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, h->comm, h->commStream));
    RB_CUDA(cudaMemcpyAsync(&min_rows, d, sizeof(int), cudaMemcpyDeviceToHost, h->commStream));
    RB_CUDA(cudaStreamSynchronize(h->commStream));
    if (min_rows < 2 * RbContext::kShipRows) {
      if (h->balance && rank == 0)
        fprintf(stderr, "[rb comm] thinnest band has %d rows (< %d): band balancing is off\n", min_rows, 2 * RbContext::kShipRows);
      h->balance = false;
    }
  }
  RB_TRY(halo_p2p_setup(h));
  return RB_OK;
}
int rb_comm_init(RbHandle h, int32_t rank, int32_t nranks, const void* nccl_unique_id, size_t id_bytes) {
  if (!h || !nccl_unique_id || id_bytes < 128 || rank < 0 || rank >= nranks) return RB_ERR_INVALID_ARGUMENT;
  if (nranks > RbContext::kMaxRanks) {
    h->err = "rb_comm_init: more than 64 ranks";
    return RB_ERR_UNSUPPORTED;
  }
  if (h->comm || h->localGroup) {
    h->err = "rb_comm_init: this handle already has a communicator (or belongs to an rb_multi group)";
    return RB_ERR_INVALID_ARGUMENT;
  }
  if (h->fs.open) {
    h->err = "rb_comm_init: a frame is open";
    return RB_ERR_INVALID_ARGUMENT;
  }
  if (!load_nccl(h->err)) return RB_ERR_COMM;
  const int rc = comm_init_body(h, rank, nranks, nccl_unique_id);
  if (rc != RB_OK) {  // leave the handle as it was: single-band, no half-made communicator
    const std::string keep = h->err;
    comm_teardown(h, /*barrier=*/false);
    h->err = keep;
  }
  return rc;
}

}  // extern "C"

// -------------------------------------------------------------------------------------
// one handle, several GPUs, one host thread (the reference is ONE process with ONE Producer thread,
// P/simpleguidx11.cpp:499). rb_multi_* drives N band handles — one per listed device — from the calling thread: the
// scene is uploaded to every device, a frame is issued band by band (everything asynchronous), the halo rows travel
// through the same k_halo_push / k_halo_wait kernels as between processes, over plain peer pointers
// (cudaDeviceEnablePeerAccess; no NCCL, no IPC), and every band's rows land in ONE host frame_data buffer.
// Static bands of equal height (the in-process balancer is not built: rows that change owner would have to take
// their last-frame reservoirs and accumulator rows along, which rb_comm_init's balancer does through NCCL).
// -------------------------------------------------------------------------------------
struct RbMulti {
  std::vector<RbContext*> m;
  std::string err;
  int width = 0, height = 0;
};
static thread_local std::string g_multi_create_error;

static int multi_fail(RbMulti* M, RbContext* h, int rc) {
  M->err = h ? h->err : std::string("rb_multi: invalid argument");
  return rc;
}
// neighbours see each other's reservoir planes and flag words directly
static int local_group_link(RbMulti* M) {
  const int n = (int)M->m.size();
  for (int i = 0; i < n; ++i) {
    RbContext* h = M->m[i];
    RB_CUDA(cudaSetDevice(h->info.device));
    RB_CUDA(cudaMalloc((void**)&h->haloFlags, 4 * sizeof(uint32_t)));
    RB_CUDA(cudaMemset(h->haloFlags, 0, 4 * sizeof(uint32_t)));
    RB_CUDA(cudaHostAlloc((void**)&h->haloErr, sizeof(uint32_t), cudaHostAllocMapped));
    *h->haloErr = 0u;
    h->haloSeq = 0;
    h->commRank = i;
    h->commSize = n;
    h->localGroup = true;
    h->balance = false;
    h->p2p = true;
  }
  for (int i = 0; i < n; ++i) {
    RbContext* h = M->m[i];
    RB_CUDA(cudaSetDevice(h->info.device));
    for (int d = 0; d < 2; ++d) {
      const int j = d == 0 ? i - 1 : i + 1;
      if (j < 0 || j >= n) continue;
      RbContext* q = M->m[j];
      if (q->info.device != h->info.device) {
        int can = 0;
        RB_CUDA(cudaDeviceCanAccessPeer(&can, h->info.device, q->info.device));
        if (!can) {
          h->err = "rb_multi_create: devices " + std::to_string(h->info.device) + " and " + std::to_string(q->info.device) + " have no peer access";
          return RB_ERR_UNSUPPORTED;
        }
        const cudaError_t e = cudaDeviceEnablePeerAccess(q->info.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          h->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
          return RB_ERR_CUDA;
        }
        (void)cudaGetLastError();
      }
      for (int b = 0; b < 4; ++b) h->peerR[d][b] = q->R[b];
      h->peerFlags[d] = q->haloFlags;
    }
  }
  return RB_OK;
}

extern "C" {

const char* rb_multi_last_error(RbMulti* M) { return M ? M->err.c_str() : g_multi_create_error.c_str(); }
int32_t rb_multi_device_count(RbMulti* M) { return M ? (int32_t)M->m.size() : 0; }
RbHandle rb_multi_member(RbMulti* M, int32_t i) { return (M && i >= 0 && i < (int32_t)M->m.size()) ? M->m[i] : nullptr; }

void rb_multi_destroy(RbMulti* M) {
  if (!M) return;
  for (RbContext* h : M->m) {  // every device idle before any plane a neighbour may still be writing is freed
    cudaSetDevice(h->info.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->fstream) cudaStreamSynchronize(h->fstream);
  }
  for (RbContext* h : M->m) rb_destroy(h);
  delete M;
}

int rb_multi_create(const RbCreateInfo* info, const int32_t* device_ordinals, int32_t n_devices, RbMulti** out) {
  if (!info || !device_ordinals || !out || n_devices < 1 || n_devices > RbContext::kMaxRanks || info->height < n_devices) {
    g_multi_create_error = "rb_multi_create: invalid argument";
    return RB_ERR_INVALID_ARGUMENT;
  }
  RbMulti* M = new RbMulti();
  M->width = info->width, M->height = info->height;
  const int rows = (info->height + n_devices - 1) / n_devices;
  for (int i = 0; i < n_devices; ++i) {
    RbCreateInfo ci = *info;
    ci.device = device_ordinals[i];
    ci.band_y0 = std::min(info->height, i * rows);
    ci.band_y1 = std::min(info->height, (i + 1) * rows);
    RbContext* h = nullptr;
    const int rc = ci.band_y0 < ci.band_y1 ? rb_create(&ci, &h) : RB_ERR_INVALID_ARGUMENT;
    if (rc != RB_OK) {
      g_multi_create_error = rc == RB_ERR_INVALID_ARGUMENT && ci.band_y0 >= ci.band_y1 ? "rb_multi_create: more devices than band rows"
                                                                                       : g_create_error;
      rb_multi_destroy(M);
      return rc;
    }
    M->m.push_back(h);
  }
  if (n_devices > 1) {
    auto link = [&]() -> int {
      RbContext* h = M->m[0];  // (RB_CUDA reports through a context)
      int rc = local_group_link(M);
      if (rc != RB_OK)
        for (RbContext* q : M->m)
          if (!q->err.empty()) h->err = q->err;
      return rc;
    };
    const int rc = link();
    if (rc != RB_OK) {
      g_multi_create_error = M->m[0]->err;
      rb_multi_destroy(M);
      return rc;
    }
  }
  *out = M;
  return RB_OK;
}

#define RB_MULTI_EACH(CALL)                          \
  do {                                               \
    if (!M) return RB_ERR_INVALID_ARGUMENT;          \
    for (RbContext* h : M->m) {                      \
      const int rc__ = (CALL);                       \
      if (rc__ != RB_OK) return multi_fail(M, h, rc__); \
    }                                                \
    return RB_OK;                                    \
  } while (0)

int rb_multi_upload_scene(RbMulti* M, const RbSceneDesc* sd) { RB_MULTI_EACH(rb_upload_scene(h, sd)); }
int rb_multi_set_params(RbMulti* M, const RbParams* p) { RB_MULTI_EACH(rb_set_params(h, p)); }
int rb_multi_set_textures(RbMulti* M, const RbTexture* textures, uint32_t n_textures, const RbMaterialTextures* per_material,
                          uint32_t n_materials) {
  RB_MULTI_EACH(rb_set_textures(h, textures, n_textures, per_material, n_materials));
}
int rb_multi_set_sky(RbMulti* M, const RbTexture* sky) { RB_MULTI_EACH(rb_set_sky(h, sky)); }
int rb_multi_synchronize(RbMulti* M) { RB_MULTI_EACH(rb_synchronize(h)); }
int rb_multi_frame_wait(RbMulti* M, uint32_t frames_in_flight) { RB_MULTI_EACH(rb_frame_wait(h, frames_in_flight)); }
int rb_multi_accumulate_display(RbMulti* M, uint32_t acc_frame_ctr, int32_t tonemap, int32_t gamma_correct, float* display_rgba_out,
                                RbImageStats* stats) {
  if (!M) return RB_ERR_INVALID_ARGUMENT;
  RbImageStats total{};
  for (RbContext* h : M->m) {  // every band writes its own rows of display_rgba_out; the sums combine
    RbImageStats st{};
    const int rc = rb_accumulate_display(h, acc_frame_ctr, tonemap, gamma_correct, display_rgba_out, stats ? &st : nullptr);
    if (rc != RB_OK) return multi_fail(M, h, rc);
    total.sum += st.sum, total.sum_sq += st.sum_sq, total.pixels += st.pixels;
  }
  if (stats) {
    total.mean = total.pixels ? total.sum / (double)total.pixels : 0.0;
    total.variance = total.pixels ? total.sum_sq / (double)total.pixels - total.mean * total.mean : 0.0;
    *stats = total;
  }
  return RB_OK;
}

// Issue one frame on every device: first everything that may allocate (a device-wide synchronisation must not meet a
// neighbour's halo wait that is still spinning for rows this thread has not queued yet), then the frames themselves,
// band by band, each followed by the asynchronous copy of its rows into the ONE host buffer.
int rb_multi_render_frame_async(RbMulti* M, const RbCamera* cam, uint32_t frame_idx, float* frame_rgb_out) {
  if (!M || !cam || !frame_rgb_out) return RB_ERR_INVALID_ARGUMENT;
  for (RbContext* h : M->m) {
    const int rc = frame_prepare(h);
    if (rc != RB_OK) return multi_fail(M, h, rc);
  }
  for (RbContext* h : M->m) {
    const int rc = rb_render_frame_async(h, cam, frame_idx, frame_rgb_out);
    if (rc != RB_OK) return multi_fail(M, h, rc);
  }
  return RB_OK;
}
int rb_multi_render_frame(RbMulti* M, const RbCamera* cam, uint32_t frame_idx, float* frame_rgb_out) {
  const int rc = rb_multi_render_frame_async(M, cam, frame_idx, frame_rgb_out);
  if (rc != RB_OK) return rc;
  return rb_multi_frame_wait(M, 0);
}
// per-pixel buffers: every device contributes the rows of its band (the ids of rb_readback; light tables from device 0)
int rb_multi_readback(RbMulti* M, int id, void* dst, size_t bytes) {
  if (!M || !dst) return RB_ERR_INVALID_ARGUMENT;
  if (id >= RB_BUF_ALIAS_PROB && id <= RB_BUF_LIGHT_CDF) {
    const int rc = rb_readback(M->m[0], id, dst, bytes);
    return rc == RB_OK ? rc : multi_fail(M, M->m[0], rc);
  }
  const size_t n = (size_t)M->width * M->height;
  size_t px = 16;
  if (id == RB_BUF_HIT_IDS) px = 8;
  if (id == RB_BUF_RES_LIGHT_IDX) px = 4;
  if (id == RB_BUF_FRAME_RGB || id == RB_BUF_ACCUMULATOR) px = 12;
  if (bytes < n * px) {
    M->err = "rb_multi_readback: destination too small";
    return RB_ERR_INVALID_ARGUMENT;
  }
  std::vector<char> tmp(n * px);
  for (RbContext* h : M->m) {
    const int rc = rb_readback(h, id, tmp.data(), tmp.size());
    if (rc != RB_OK) return multi_fail(M, h, rc);
    const size_t off = (size_t)h->info.band_y0 * M->width * px, cnt = (size_t)(h->info.band_y1 - h->info.band_y0) * M->width * px;
    memcpy((char*)dst + off, tmp.data() + off, cnt);
  }
  return RB_OK;
}

}  // extern "C"

