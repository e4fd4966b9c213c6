// rb_passes.cuh — per-pixel logic of the six ReSTIR passes (the body of
// SimpleGuiDX11::produceRestir's loops, P/simpleguidx11.cpp:359-487, and the
// ReSTIRIntegrator statics they call, P/ReSTIRIntegrator.cpp:89-732).
//
// GPU-first restructuring relative to the reference, all result-preserving:
//   * 1/I_M (Phong normalisation: incomplete beta + 2 lgamma + pow,
//     P/MaterialPhong.cpp:228-248) is evaluated once per pixel in the G-buffer
//     pass and stored, instead of on every BRDF evaluation (it depends only on
//     the element and its own frame's camera at every call site).
//   * identical p-hat evaluations inside a pass are computed once (temporal 7 -> 4,
//     spatial k+2 -> k+1, initial M+1 -> M); shadow rays whose unshadowed
//     contribution is exactly zero are not traced (x*V is x either way).
//   * visibility is obtained through a policy object so the same code runs with
//     inline traversal, or as the generate / resolve halves of the wavefront split.
#ifndef RB_PASSES_CUH_
#define RB_PASSES_CUH_

#include "rb_scene.cuh"

namespace rb {

#define RB_MAX_NEIGHBORS 32

// Wavefront buffers (stream -> trace -> resolve split): a compact ray queue filled by the "stream" half of a
// pass, traced by the persistent traversal kernels, and per-(slot, pixel) results read by the "resolve" half.
struct RayQ {
  F4 o_tfar;  // origin.xyz, tfar
  F4 d_dest;  // dir.xyz, bits(destination index = slot * npix + pixel)
};
struct WaveBufs {
  RayQ* rays;
  uint32_t* count;  // {rays queued, next ray to fetch} (device counters); two such pairs alternate between passes
  uint32_t* reset_pair;  // stream kernels zero the pair of the PREVIOUS pass here (nullptr otherwise): no reset launches
  uint32_t capacity;
  uint8_t* occ;  // any-hit results   [slot * npix + pixel]
  HitRec* hits;  // closest results   [slot * npix + pixel]
  F4* brdf_dir;  // {omega_i, pdf} of the BRDF-sampled candidates [slot * npix + pixel]: written by the stream half of the
                 // initial pass, read by its resolve half (which then carries no sampleBRDF code)
  uint32_t npix;
  uint32_t brdf_two_step;  // BRDF-candidate hits come from the emissive-only BVH; occ[] says whether something precedes them
  U4* cand;      // spatial pass, constant weights: candidate records [slot * npix + pixel] (spatial_gen_pixel)
  // banded temporal stream: pixels whose reprojection leaves the G-buffer rows this handle holds are listed here and
  // handled by a second, small launch that carries the re-derivation (traversal) code; the bulk kernel stays lean
  uint32_t* deferred;        // pixel indices
  uint32_t* deferred_count;  // device counter, zeroed with the frame's ray counters
  // two-step BRDF-candidate rays: the closest-EMITTER traversal itself queues the "does anything precede this hit?"
  // ray of every ray that found an emitter (brdf_chain_push) — no stream kernel between the two traversals
  RayQ* chain_rays;
  uint32_t* chain_count;  // {queued, next to fetch}, zeroed with the frame's ray counters
  uint32_t chain_capacity;
  // visibility pass folded into its neighbours (wavefront schedule): bit 0 = the initial-pass resolve kernel queues the
  // visibility ray of the reservoir it has just produced; bit 1 = the temporal stream kernel applies the traced result
  // (W = 0 where occluded) before it uses the reservoir. Either bit clear = k_visibility_stream / _resolve do it.
  uint32_t fuse_vis;
  uint32_t fuse_shade;  // the last spatial pass's resolve kernel also shades the pixel (no k_shade launch)
};

struct FrameCtx {
  int width, height;
  int y0, y1;  // rows rendered by this handle (band)
  SceneDev sc;
  RbParams P;
  CamState cam, prevCam;
  GBufPlanes G, Gprev;
  ResPlanes Rread, Rwrite, Rlast;
  float* frame;  // w*h*3
  uint32_t frame_key;
  int spatial_iter;
  // image rows for which this handle holds the G-buffer of the current / previous frame (its band plus a margin);
  // elements outside are re-derived on demand (they are a pure function of camera, pixel and scene)
  int gy0, gy1, gpy0, gpy1;
  unsigned long long* counters;  // [0] closest, [1] any-hit as written, [2] any-hit traced
  WaveBufs wave;
  uint32_t mis_flags;  // mis_direct_pixel: bit 0 = sample the BRDF, bit 1 = sample the light sources
};

struct Cnt {
  uint32_t closest, anyW, anyT;
};

RB_HD SurfaceHit no_hit() {
  SurfaceHit h;
  h.didHit = false;
  h.normal = v3(0);
  h.hitPoint = v3(0);
  h.t = FLT_MAX;
  h.tri = h.geomID = h.primID = 0xFFFFFFFFu;
  h.material = 0;
  h.emissiveId = -1;
  h.tex_u = h.tex_v = 0.0f;
  return h;
}
// reserve one queue slot; on the device the atomic is aggregated over the currently converged lanes
RB_HD uint32_t queue_reserve(uint32_t* counter) {
#if defined(__CUDA_ARCH__)
  const unsigned m = __activemask();
  const int lane = threadIdx.x + threadIdx.y * blockDim.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + __popc(m & ((1u << lane) - 1u));
#else
  return __atomic_fetch_add(counter, 1u, __ATOMIC_RELAXED);
#endif
}

// reserve n queue slots (n may differ per lane, n < 64); on the device ONE atomic for the converged lanes: the per-lane
// offsets come from ballots of the bits of n (no shuffle chain, no dependence on the atomic's result until the end)
RB_HD uint32_t queue_reserve_n(uint32_t* counter, uint32_t n) {
#if defined(__CUDA_ARCH__)
  const unsigned m = __activemask();
  const int lane = threadIdx.x + threadIdx.y * blockDim.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    const unsigned bal = __ballot_sync(m, (n >> b) & 1u);
    before += (uint32_t)__popc(bal & lt) << b;
    total += (uint32_t)__popc(bal) << b;
  }
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader && total != 0) base = atomicAdd(counter, total);
  base = __shfl_sync(m, base, leader);
  return base + before;
#else
  return __atomic_fetch_add(counter, n, __ATOMIC_RELAXED);
#endif
}

// step 2 of the two-step BRDF-candidate rays, queued by the traversal that found the emitter hit (t_hit, hits[dest].tri)
RB_HD void brdf_chain_push(const WaveBufs& w, const V3& o, const V3& d, float t_hit, uint32_t dest) {
  const uint32_t i = queue_reserve(w.chain_count);
  if (i < w.chain_capacity) {
    st4(&w.chain_rays[i].o_tfar, f4(o, t_hit));
    st4(&w.chain_rays[i].d_dest, f4(d, u2f(dest)));
  }
}

// ---- visibility / closest-hit policies ----------------------------------------------
// InlineVis: every query is traced on the spot by the calling thread.
struct InlineVis {
  static constexpr bool kStore = true;
  const FrameCtx* fc;
  uint32_t pixel;
  RB_HD void begin_post_selection() const {}
  RB_HD bool visible(int /*slot*/, const V3& from, const V3& to) const {
    return !test_occlusion(fc->sc, from, to, fc->P.tnearOffset, fc->P.tfarOffset);
  }
  RB_HD SurfaceHit closest(int /*slot*/, const V3& org, const V3& dir, float tnear, float tfar) const {
    return intersect_surface(fc->sc, org, dir, tnear, tfar);
  }
  RB_HD void put_brdf_sample(int, const V3&, float) const {}
  RB_HD bool get_brdf_sample(int, V3*, float*) const { return false; }
};
// GenVis: the stream half. Queries are appended to the ray queue and answered "visible" / "miss"; the pass body
// runs only to enumerate its rays, its stores are suppressed (kStore).
struct GenVis {
  static constexpr bool kStore = false;
  const FrameCtx* fc;
  uint32_t pixel;
  RB_HD void begin_post_selection() const {}
  RB_HD void push(int slot, const V3& o, const V3& d, float tfar) const {
    const WaveBufs& w = fc->wave;
    const uint32_t i = queue_reserve(w.count);
    if (i < w.capacity) {
      st4(&w.rays[i].o_tfar, f4(o, tfar));
      st4(&w.rays[i].d_dest, f4(d, u2f((uint32_t)slot * w.npix + pixel)));
    }
  }
  RB_HD bool visible(int slot, const V3& from, const V3& to) const {
    V3 dir;
    float tfar;
    shadow_ray(from, to, fc->P.tfarOffset, &dir, &tfar);
    push(slot, from, dir, tfar);
    return true;
  }
  RB_HD SurfaceHit closest(int slot, const V3& org, const V3& dir, float /*tnear*/, float tfar) const {
    push(slot, org, dir, tfar);
    return no_hit();
  }
  RB_HD void put_brdf_sample(int slot, const V3& wi, float pdf) const {
    st4(fc->wave.brdf_dir + (size_t)slot * fc->wave.npix + pixel, f4(wi, pdf));
  }
  RB_HD bool get_brdf_sample(int, V3*, float*) const { return false; }
  // the same shadow ray as visible(), into a slot reserved beforehand (reserve: all of a pixel's rays with one atomic
  // per warp instead of one per ray — the reservation round trip was the largest stall of the reuse stream kernels)
  RB_HD uint32_t reserve(uint32_t n) const { return queue_reserve_n(fc->wave.count, n); }
  RB_HD void visible_at(uint32_t i, int slot, const V3& from, const V3& to) const {
    V3 dir;
    float tfar;
    shadow_ray(from, to, fc->P.tfarOffset, &dir, &tfar);
    const WaveBufs& w = fc->wave;
    if (i < w.capacity) {
      st4(&w.rays[i].o_tfar, f4(from, tfar));
      st4(&w.rays[i].d_dest, f4(dir, u2f((uint32_t)slot * w.npix + pixel)));
    }
  }
};
// ResolveVis: the resolve half. The same queries, in the same order, read the traced results.
// INLINE_SHADOW: shadow rays are traced on the spot instead (initial pass with the visibility pass off, whose
// 33 rays per pixel are not queued); closest-hit queries are always looked up.
template <bool INLINE_SHADOW, bool NMAP = true>
struct ResolveVisT {
  static constexpr bool kStore = true;
  const FrameCtx* fc;
  uint32_t pixel;
  RB_HD void begin_post_selection() const {}
  RB_HD bool visible(int slot, const V3& from, const V3& to) const {
    const WaveBufs& w = fc->wave;
    if (INLINE_SHADOW) return !test_occlusion(fc->sc, from, to, fc->P.tnearOffset, fc->P.tfarOffset);
    return w.occ[(size_t)slot * w.npix + pixel] == 0;
  }
  RB_HD SurfaceHit closest(int slot, const V3& org, const V3& dir, float /*tnear*/, float /*tfar*/) const {
    const WaveBufs& w = fc->wave;
    HitRec hr = w.hits[(size_t)slot * w.npix + pixel];
    // two-step BRDF rays: the emitter hit only stands if nothing precedes it (a miss and a non-emissive closest hit
    // are the same thing to brdfSampleLight, P/ReSTIRIntegrator.cpp:143)
    if (w.brdf_two_step && hr.tri != 0xFFFFFFFFu && w.occ[(size_t)slot * w.npix + pixel] != 0) hr.tri = 0xFFFFFFFFu;
    return surface_from_hit<NMAP>(fc->sc, org, dir, hr);
  }
  RB_HD void put_brdf_sample(int, const V3&, float) const {}
  RB_HD bool get_brdf_sample(int slot, V3* wi, float* pdf) const {  // what the stream half computed (same bits, not recomputed)
    const F4 v = ld4(fc->wave.brdf_dir + (size_t)slot * fc->wave.npix + pixel);
    *wi = xyz(v), *pdf = v.w;
    return true;
  }
};
// StagedVis: the wavefront schedule of the spatial MIS modes whose ray count is not k + 1 (BALANCE_HEURISTIC: every
// sample at every source pixel, O(k^2); PAIRWISE_MIS: O(k); the two CONSTANT_DEBIAS modes: k + 1 more rays that depend on
// which sample was selected, P/ReSTIRIntegrator.cpp:407-467,494-538). spatial_pixel itself is run two or three times
// with this policy; every visibility query takes the next slot of the pixel (the sequence of queries is the same in every
// stage: before the selection it depends on the inputs only, after it on a selection that stages 2 and 3 both make from
// traced answers):
//   stage 1  queues the rays asked for before the selection, answers "visible"; rays after it are not queued;
//   stage 2  (debias modes only) answers the former from the traced bytes — so the selection is the real one — and
//            queues the rays asked for after it;
//   stage 3  answers everything from the traced bytes and stores the reservoir.
// The arithmetic is spatial_pixel's own in every stage, so the result is the inline kernel's bit for bit.
template <int STAGE>
struct StagedVis {
  static constexpr bool kStore = STAGE == 3;
  const FrameCtx* fc;
  uint32_t pixel;
  mutable uint32_t seq;
  mutable bool post;
  RB_HD void begin_post_selection() const { post = true; }
  RB_HD bool visible(int /*slot*/, const V3& from, const V3& to) const {
    const WaveBufs& w = fc->wave;
    const uint32_t dest_slot = seq++;
    if (STAGE == 3 || (STAGE == 2 && !post)) return w.occ[(size_t)dest_slot * w.npix + pixel] == 0;
    if ((STAGE == 1 && !post) || (STAGE == 2 && post)) {
      V3 dir;
      float tfar;
      shadow_ray(from, to, fc->P.tfarOffset, &dir, &tfar);
      const uint32_t i = queue_reserve(w.count);
      if (i < w.capacity) {
        st4(&w.rays[i].o_tfar, f4(from, tfar));
        st4(&w.rays[i].d_dest, f4(dir, u2f(dest_slot * w.npix + pixel)));
      }
    }
    return true;
  }
};
// visibility queries spatial_pixel can make per pixel with n = k + 1 resampling sources (slots of the staged schedule)
RB_HD uint32_t spatial_staged_slots(int mode, uint32_t n) {
  if (mode == RB_SW_BALANCE_HEURISTIC) return n * n + n;
  if (mode == RB_SW_PAIRWISE_MIS) return 4u * n;
  return 2u * n;  // CONSTANT_DEBIAS_Z_TERM / _CONTRIB: n resampling rays + n after the selection
}
typedef ResolveVisT<false> ResolveVis;
typedef ResolveVisT<false, false> ResolveVisFlat;  // scenes without normal maps (see surface_from_hit)
typedef ResolveVisT<true> ResolveInlineShadowVis;

// ---- Phong / Lambert statics (P/MaterialPhong.cpp:122-248, P/MaterialLambert.cpp:33-53,
//      P/Distribution.h) ------------------------------------------------------------
RB_HD float max_component(const V3& v) { return gmax(gmax(v.x, v.y), v.z); }  // P/utils.h:61-63

// The parts of calc_I_M that depend on the shininess alone — log B(n/2, 1/2) inside the incomplete beta function and the
// gamma quotient: five lgamma evaluations and an exp, half of the function's cost — are material constants. They are
// computed once per material at upload BY THESE SAME FUNCTIONS (on the device for the library, on the host for the
// host build), so a pixel that uses them gets the bits it would have computed itself; a pixel whose shininess does not
// come from the material constant (roughness map) computes them as before.
RB_HD double im_lbeta(float n) { return dm::ibeta_lbeta((double)(0.5f * n), (double)0.5f); }
RB_HD float im_gamma_quot(float n) {  // gamma_quot(halfn + 0.5, halfn + 1), P/MaterialPhong.cpp:224-226
  const float halfn = 0.5f * n;
  return dm::expf_(dm::lgammaf_(halfn + 0.5f) - dm::lgammaf_(halfn + 1.0f));
}
RB_HD MatConst make_mat_const(float n) {
  MatConst c;
  c.lbeta = n >= 1e-18f ? im_lbeta(n) : 0.0;
  c.gq = im_gamma_quot(n);
  c.shininess = n;
  return c;
}
RB_HD_NOINLINE float calc_I_M(float nDotV, float n, const MatConst* mc = nullptr) {  // MaterialPhong::calc_I_M, :228-244
  float costerm = nDotV;
  float sinterm_sq = 1.0f - costerm * costerm;
  float halfn = 0.5f * n;
  float negterm = costerm;
  sinterm_sq = gclamp(sinterm_sq, 0.0f, 1.0f);
  const bool known = mc != nullptr && f2u(mc->shininess) == f2u(n);
  if (n >= 1e-18f) {
    const double lbeta = known ? mc->lbeta : im_lbeta(n);
    negterm *= halfn * (float)dm::ibeta_full_lb((double)halfn, (double)0.5f, (double)sinterm_sq, lbeta);
  }
  const float gq = known ? mc->gq : im_gamma_quot(n);
  return (RB_TWO_PI * costerm + RB_ROOT_PI * gq * (dm::powf_(sinterm_sq, halfn) - negterm)) / (n + 2.0f);
}
RB_HD float inv_I_M(const V3& pos, const V3& normal, float shininess, const V3& camPos, const MatConst* mc = nullptr) {
  const V3 V = normalize(camPos - pos);
  float nDotV = dot(V, normal);
  return frcp_(calc_I_M(nDotV, shininess, mc));
}
RB_HD bool uses_phong_brdf(uint32_t t) { return t == RB_MAT_PHONG || t == RB_MAT_DIELECTRIC; }

// Per-pixel shading context: the sub-expressions of evalBRDF / evalPdf / sampleBRDF that depend only on the
// G-buffer element and its own frame's camera, hoisted out of the per-candidate code. omega_r is
// normalize(reflect(omega_o, n)) with omega_o = normalize(pos - cam); evalBRDF spells it reflect(-V, n) with
// V = normalize(cam - pos), which is the same bits (x - y == -(y - x) exactly).
struct Shading {
  V3 omega_r;
  float pdfFactor;  // maxDiffuse / (maxDiffuse + maxSpecular), P/MaterialPhong.cpp:159
  bool phong;       // BRDF has the Phong lobe (PHONG, DIELECTRIC); otherwise Lambert (P/ReSTIRIntegrator.h:32-41)
  bool finite;      // 1/I_M, pdfFactor and shininess are finite numbers
};
RB_HD bool finitef_(float x) { return (f2u(x) & 0x7F800000u) != 0x7F800000u; }
RB_HD Shading make_shading(const GElem& g, const V3& camPos) {
  Shading sh;
  const V3 omega_o = normalize(g.pos - camPos);
  sh.omega_r = normalize(reflect(omega_o, g.normal));
  const float maxDiffuse = max_component(g.diffuse);
  const float maxSpecular = max_component(g.specular);
  sh.pdfFactor = maxDiffuse / (maxDiffuse + maxSpecular);
  sh.phong = uses_phong_brdf(g.matType);
  sh.finite = finitef_(g.invIM) && finitef_(sh.pdfFactor) && finitef_(g.shininess);
  return sh;
}
// The Phong lobe term pow(max(dot(omega_i, omega_r), 0), n) is shared by evalBRDF (:144) and evalPdf's
// CosineLobeDistribution::getPdf (P/Distribution.h:65-67). The two spell the clamp as max(d,0) and max(0,d):
// identical except for NaN d, which is kept apart.
struct Lobe {
  float brdf, pdf;
};
RB_HD Lobe phong_lobe(const GElem& g, const Shading& sh, const V3& omega_i) {
  const float d = dot(omega_i, sh.omega_r);
  Lobe l;
  l.brdf = dm::powf_(gmax(d, 0.0f), g.shininess);
  l.pdf = (d != d) ? dm::powf_(gmax(0.0f, d), g.shininess) : l.brdf;
  return l;
}
// getMaterialBRDFEvalFunc dispatch, P/ReSTIRIntegrator.h:32-41
RB_HD V3 brdf_from_lobe(const GElem& g, const Shading& sh, float lobe) {
  V3 f_r = g.diffuse * RB_ONE_OVER_PI;
  if (sh.phong) f_r = f_r + g.specular * g.invIM * lobe;
  return f_r;
}
RB_HD float cosw_pdf(const V3& n, const V3& wi) { return gmax(dot(n, wi), 0.0f) * RB_ONE_OVER_PI; }
// MaterialPhong::evalPdf, :150-172 (getMaterialPDFEvalFunc always returns it, P/ReSTIRIntegrator.h:54-59)
RB_HD float pdf_from_lobe(const GElem& g, const Shading& sh, const V3& wi, float lobe) {
  float pdf = cosw_pdf(g.normal, wi) * sh.pdfFactor;
  pdf += ((g.shininess + 1.0f) * RB_ONE_OVER_TWO_PI * lobe) * (1.0f - sh.pdfFactor);
  return pdf;
}
RB_HD V3 orthogonal(const V3& v) {  // Utils::orthogonal, P/utils.cpp:204-207
  return fabsf_(v.x) > fabsf_(v.z) ? v3(v.y, -v.x, 0.0f) : v3(0.0f, v.z, -v.y);
}
RB_HD V3 to_world(const V3& n, const V3& s) {  // glm::mat3{o1,o2,n} * s with the basis of P/Distribution.h:21-28
  V3 o2 = normalize(orthogonal(n));
  V3 o1 = normalize(cross(n, o2));
  o2 = normalize(cross(o1, n));
  return v3(o1.x * s.x + o2.x * s.y + n.x * s.z, o1.y * s.x + o2.y * s.y + n.y * s.z, o1.z * s.x + o2.z * s.y + n.z * s.z);
}
RB_HD V3 cosw_sample(const V3& n, float r1, float r2) {  // CosineWeightedDistribution::sample, :10-31
  float sn, cs;
  dm::sincosf_(RB_PI * 2.0f * r1, &sn, &cs);
  float x = cs * sqrtf_(1.0f - r2);
  float y = sn * sqrtf_(1.0f - r2);
  float z = sqrtf_(r2);
  return to_world(n, normalize(v3(x, y, z)));
}
RB_HD V3 lobe_sample(const V3& wr, float gamma, float r1, float r2) {  // CosineLobeDistribution::sample, :43-63
  float sn, cs;
  dm::sincosf_(2.0f * RB_PI * r1, &sn, &cs);
  float pw = dm::powf_(r2, 2.0f / (gamma + 1.0f));
  float x = cs * sqrtf_(1.0f - pw);
  float y = sn * sqrtf_(1.0f - pw);
  float z = dm::powf_(r2, 1.0f / (gamma + 1.0f));
  return to_world(wr, normalize(v3(x, y, z)));
}
// getMaterialSampleFunc dispatch (:43-52): Lambert for LAMBERT, Phong for everything else.
// Returns omega_i and its pdf (the f_r the reference also computes is unused by its ReSTIR caller).
RB_HD V3 brdf_sample(const GElem& g, const Shading& sh, uint32_t key, uint32_t base, float* pdf) {
  const float r1 = rng_value(key, base + 1, 0, 1), r2 = rng_value(key, base + 2, 0, 1);
  if (g.matType == RB_MAT_LAMBERT) {  // MaterialLambert::sampleBRDF, P/MaterialLambert.cpp:43-53
    V3 wi = cosw_sample(g.normal, r1, r2);
    *pdf = cosw_pdf(g.normal, wi);
    return wi;
  }
  // MaterialPhong::sampleBRDF, P/MaterialPhong.cpp:174-222
  float maxDiffuse = max_component(g.diffuse);
  float maxSpecular = max_component(g.specular);
  float r0 = rng_value(key, base + 0, 0.0f, maxDiffuse + maxSpecular);
  V3 wi = (r0 < maxDiffuse) ? cosw_sample(g.normal, r1, r2) : lobe_sample(sh.omega_r, g.shininess, r1, r2);
  float pdfDiffuse = cosw_pdf(g.normal, wi) * sh.pdfFactor;
  // CosineLobeDistribution::getPdf(omega_i, omega_r, gamma), P/Distribution.h:65-67
  float pdfSpecular = ((g.shininess + 1.0f) * RB_ONE_OVER_TWO_PI * dm::powf_(gmax(0.0f, dot(wi, sh.omega_r)), g.shininess)) *
                      (1.0f - sh.pdfFactor);
  *pdf = pdfDiffuse + pdfSpecular;
  return wi;
}

// ---- evaluateF / evaluatePHat (P/ReSTIRIntegrator.cpp:180-211) ---------------------
// What is already known about the visibility term of an evaluation (all exact, DESIGN.md "ray elimination"):
//   VIS_TRACE      nothing: trace the shadow ray.
//   VIS_KNOWN      the sample is the reservoir this very pixel produced in the previous pass of this frame and
//                  its W is > 0. W = w_sum / p-hat(sample) with the shadowed p-hat, so W > 0 implies that the
//                  identical ray (same origin, same target) was found unoccluded: V = 1.
//   VIS_IRRELEVANT the value is only ever multiplied by a reservoir weight W that is exactly 0: any finite value
//                  gives the same product, so the ray is skipped when the unshadowed value is finite.
enum VisMode { VIS_TRACE = 0, VIS_KNOWN = 1, VIS_IRRELEVANT = 2 };
RB_HD VisMode vis_mode_from_W(float W, bool own_pixel_this_frame) {
  if (W == 0.0f) return VIS_IRRELEVANT;
  if (own_pixel_this_frame && W > 0.0f) return VIS_KNOWN;
  return VIS_TRACE;  // includes NaN W
}
RB_HD bool finite3(const V3& v) { return finitef_(v.x) && finitef_(v.y) && finitef_(v.z); }

// Unshadowed value L_i * f_r * G; *wants_ray tells whether the reference would trace here.
RB_HD V3 eval_F0(const LightSample& s, const GElem& g, const Shading& sh, bool* wants_ray) {
  *wants_ray = false;
  if (!sample_valid(s) || g.isEmissive) return v3(0);
  V3 lightDir = s.samplePoint - g.pos;
  float r_sqr = dot(lightDir, lightDir);
  lightDir = normalize(lightDir);
  float cosThetaI = gmax(dot(lightDir, g.normal), 0.0f);
  float cosThetaY = fabsf_(dot(-lightDir, s.sampleNormal));
  float G = cosThetaI * cosThetaY / r_sqr;
  float lobe = 0.0f;
  if (sh.phong) lobe = dm::powf_(gmax(dot(lightDir, sh.omega_r), 0.0f), g.shininess);
  V3 f_r = brdf_from_lobe(g, sh, lobe);
  *wants_ray = true;
  return s.L_i * f_r * G;
}
template <class Vis>
RB_HD V3 shadow_F0(const V3& F0, const V3& from, const V3& to, const Vis& vis, int slot, Cnt& cnt, int written_copies,
                   VisMode mode) {
  cnt.anyW += (uint32_t)written_copies;
  if (F0.x == 0.0f && F0.y == 0.0f && F0.z == 0.0f) return F0;  // F0 * V == F0 for V in {0,1}
  if (mode == VIS_KNOWN) return F0;
  if (mode == VIS_IRRELEVANT && finite3(F0)) return F0;
  cnt.anyT++;
  float V = vis.visible(slot, from, to) ? 1.0f : 0.0f;
  return F0 * V;
}
template <class Vis>
RB_HD V3 eval_F(const LightSample& s, const GElem& g, const Shading& sh, bool testVisibility, const Vis& vis, int slot,
                Cnt& cnt, int written_copies, VisMode mode = VIS_TRACE, bool* wants_out = nullptr) {
  bool wants;
  V3 F0 = eval_F0(s, g, sh, &wants);
  if (wants_out) *wants_out = wants;
  if (!wants || !testVisibility) return F0;
  return shadow_F0(F0, g.pos, s.samplePoint, vis, slot, cnt, written_copies, mode);
}
template <class Vis>
RB_HD float eval_phat(const LightSample& s, const GElem& g, const Shading& sh, bool testVisibility, const Vis& vis,
                      int slot, Cnt& cnt, int written_copies, VisMode mode = VIS_TRACE, bool* wants_out = nullptr) {
  return length(eval_F(s, g, sh, testVisibility, vis, slot, cnt, written_copies, mode, wants_out));
}

// ---- light sampling (P/TriangleCDF.cpp:36-54 / alias seam) --------------------------
struct LightPick {
  uint32_t idx;
  float pdf;  // probability of choosing this triangle
};
RB_HD LightPick pick_light(const SceneDev& sc, int sampler, uint32_t key, uint32_t slot) {
  LightPick p;
  const uint32_t N = sc.n_lights;
  if (sampler == RB_LS_ALIAS) {
    const uint32_t h = rng_bits(key, slot);
    const uint32_t i = (uint32_t)(((uint64_t)h * (uint64_t)N) >> 32);
    const float frac = bits_to_unit(rng_bits(key, slot | 0x40000000u));
    const U2 a = ldg2(sc.alias_pair + i);
    p.idx = (frac < u2f(a.x)) ? i : a.y;
    p.pdf = -1.0f;  // area / total: word 3 of the light's second record, which the caller loads anyway
    return p;
  }
  const float ksi = rng_value(key, slot, 0.0f, 1.0f);
  // std::lower_bound: first element >= ksi
  uint32_t lo = 0, hi = N;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (sc.cdf[mid] < ksi)
      lo = mid + 1;
    else
      hi = mid;
  }
  uint32_t index = lo;
  if (index >= N) index = N - 1;
  p.idx = index;
  p.pdf = index == 0 ? sc.cdf[0] : sc.cdf[index] - sc.cdf[index - 1];
  return p;
}
RB_HD float m_area(const RbParams& P, float pdfArea, float pdfBrdf) {  // P/ReSTIRIntegrator.h:62-67
  if (pdfArea == 0.0f && pdfBrdf == 0.0f) return 0.0f;
  return pdfArea / ((float)P.M_Area * pdfArea + (float)P.M_Brdf * pdfBrdf);
}
RB_HD float m_brdf(const RbParams& P, float pdfBrdf, float pdfArea) {  // :69-74
  if (pdfArea == 0.0f && pdfBrdf == 0.0f) return 0.0f;
  return pdfBrdf / ((float)P.M_Area * pdfArea + (float)P.M_Brdf * pdfBrdf);
}

// =====================================================================================
// Pass 0: G-buffer (ReSTIRIntegrator::gBufferFillPass, :213-234; Camera::GenerateRay,
// P/camera.cpp:20-42 — pixel corner, the two discarded aperture draws have no effect)
// =====================================================================================
RB_HD void primary_ray(const CamState& cam, int width, int height, int x, int y, V3* dir) {
  const V3 d_c = v3((float)x - (float)width / 2.0f, (float)height / 2.0f - (float)y, -cam.focal);
  const float* m = cam.invViewMat;
  V3 d_w = v3(m[0] * d_c.x + m[4] * d_c.y + m[8] * d_c.z, m[1] * d_c.x + m[5] * d_c.y + m[9] * d_c.z,
              m[2] * d_c.x + m[6] * d_c.y + m[10] * d_c.z);
  *dir = normalize(d_w);
}
// the element of a pixel from its primary ray's closest hit (material fetch, emission, cached 1/I_M)
RB_HD GElem gbuffer_from_hit(const FrameCtx& fc, const CamState& cam, const V3& dir, const SurfaceHit& h, uint32_t* geomID,
                             uint32_t* primID) {
  GElem e;
  e.pos = e.normal = e.diffuse = e.specular = e.emission = v3(0);
  e.shininess = e.depth = e.invIM = 0;
  e.matType = 0;
  *geomID = *primID = 0xFFFFFFFFu;
  if (h.didHit) {
    const F4 m0 = ldg4(fc.sc.mat + 3 * (size_t)h.material), m1 = ldg4(fc.sc.mat + 3 * (size_t)h.material + 1),
             m2 = ldg4(fc.sc.mat + 3 * (size_t)h.material + 2);
    e.pos = h.hitPoint;
    e.normal = h.normal;
    e.depth = length(h.hitPoint - cam.pos);
    e.matType = f2u(m1.w);
    e.diffuse = xyz(m0);
    e.specular = xyz(m1);
    e.emission = xyz(m2);
    e.shininess = m0.w;
    if (fc.sc.mat_tex != nullptr) {  // Material::getDiffuseColor / getSpecularColor / getShininess, P/material.cpp:105-134
      const I4 slots = fc.sc.mat_tex[h.material];
      if (slots.x >= 0) e.diffuse = tex_sample(fc.sc.tex[slots.x], h.tex_u, h.tex_v);
      if (slots.y >= 0) e.specular = tex_sample(fc.sc.tex[slots.y], h.tex_u, h.tex_v);
      if (slots.z >= 0) {  // roughness -> shininess
        const V3 texel = tex_sample(fc.sc.tex[slots.z], h.tex_u, h.tex_v);
        e.shininess = 2.0f / (texel.x * texel.x) - 2.0f;
      }
    }
    *geomID = h.geomID;
    *primID = h.primID;
  } else if (fc.P.useSkybox) {  // scene.getSkybox().getTexel(ray.getDir()), :231
    e.emission = sky_texel(fc.sc.sky, dir);
  } else {
    e.emission = v3(fc.P.bgColor[0], fc.P.bgColor[1], fc.P.bgColor[2]);
  }
  e.isEmissive = emissive3(e.emission);
  if (h.didHit && !e.isEmissive && uses_phong_brdf(e.matType))
    e.invIM = inv_I_M(e.pos, e.normal, e.shininess, cam.pos, fc.sc.mat_const != nullptr ? fc.sc.mat_const + h.material : nullptr);
  return e;
}
#define RB_PRIMARY_TNEAR (FLT_MIN + 0.01f)  // Ray ctor defaults, P/Ray.h:8
RB_HD GElem gbuffer_element(const FrameCtx& fc, const CamState& cam, int x, int y, uint32_t* geomID, uint32_t* primID,
                            Cnt& cnt) {
  V3 dir;
  primary_ray(cam, fc.width, fc.height, x, y, &dir);
  cnt.closest++;
  const SurfaceHit h = intersect_surface(fc.sc, cam.pos, dir, RB_PRIMARY_TNEAR, FLT_MAX);
  return gbuffer_from_hit(fc, cam, dir, h, geomID, primID);
}
// G-buffer access for pixels that may lie outside this handle's rows (multi-GPU bands, SURVEY §8e "local
// re-trace"): the element is recomputed from that frame's camera, bit-identical to what its owner stored.
// (the slow paths are kept out of line so that they do not inflate the register budget of the callers)
RB_HD_NOINLINE void retrace_gelem(const FrameCtx& fc, bool prev, int x, int y, GElem* out, uint32_t* n_closest) {
  uint32_t g, p;
  Cnt c = {0, 0, 0};
  *out = gbuffer_element(fc, prev ? fc.prevCam : fc.cam, x, y, &g, &p, c);
  *n_closest += c.closest;
}
RB_HD_NOINLINE void retrace_gpos(const FrameCtx& fc, bool prev, int x, int y, V3* out, uint32_t* n_closest) {
  const CamState& cam = prev ? fc.prevCam : fc.cam;
  V3 dir;
  primary_ray(cam, fc.width, fc.height, x, y, &dir);
  *n_closest += 1;
  HitRec r;
  *out = trace8<false>(fc.sc, cam.pos, dir, FLT_MIN + 0.01f, FLT_MAX, &r) ? cam.pos + dir * r.t : v3(0);
}
RB_HD GElem fetch_gelem(const FrameCtx& fc, bool prev, int x, int y, Cnt& cnt) {
  const int r0 = prev ? fc.gpy0 : fc.gy0, r1 = prev ? fc.gpy1 : fc.gy1;
  if (y >= r0 && y < r1) return load_gelem(prev ? fc.Gprev : fc.G, (size_t)y * fc.width + x);
  GElem e;
  retrace_gelem(fc, prev, x, y, &e, &cnt.closest);
  return e;
}
RB_HD V3 fetch_gpos(const FrameCtx& fc, bool prev, int x, int y, Cnt& cnt) {
  const int r0 = prev ? fc.gpy0 : fc.gy0, r1 = prev ? fc.gpy1 : fc.gy1;
  if (y >= r0 && y < r1) return xyz(ld4((prev ? fc.Gprev : fc.G).pos_depth + (size_t)y * fc.width + x));
  V3 p;
  retrace_gpos(fc, prev, x, y, &p, &cnt.closest);
  return p;
}

RB_HD void gbuffer_pixel(const FrameCtx& fc, int x, int y, Cnt& cnt) {
  uint32_t g, p;
  const GElem e = gbuffer_element(fc, fc.cam, x, y, &g, &p, cnt);
  store_gelem(fc.G, (size_t)y * fc.width + x, e, g, p);
}
// =====================================================================================
// Pass 1: initial candidates (ReSTIRIntegrator::initialRenderPass, :236-298;
// areaSampleLight :89-124; brdfSampleLight :126-177; Sampling::sampleTriangle P/Sampling.cpp:63-76)
// RNG slots: candidate c uses 4c..4c+3 = {pick | lobe, r1, r2, accept}.
// =====================================================================================
// Where initial_pixel parks the light indices of a chunk of candidates between its two loops: one word per candidate,
// `stride` words apart (kernels: shared memory, [candidate][thread]; host: a local array).
struct PickStore {
  uint32_t* p;
  int stride;
  RB_HD void put(int j, uint32_t v) const { p[j * stride] = v; }
  RB_HD uint32_t get(int j) const { return p[j * stride]; }
};
constexpr int kPickChunk = 32;
#ifndef RB_PICK_BATCH
#define RB_PICK_BATCH 2
#endif
constexpr int kPickBatch = RB_PICK_BATCH;  // divides kPickChunk
// P/TriangleCDF.cpp:36-54: the probability of the triangle a CDF pick returned (pick_light's own expression)
RB_HD float cdf_pick_pdf(const SceneDev& sc, uint32_t index) { return index == 0 ? sc.cdf[0] : sc.cdf[index] - sc.cdf[index - 1]; }

// Horizon pre-test of the initial pass. The loop below gives a candidate the weight +0 without evaluating it when
// its EXACT cull test holds (cosThetaI == 0 and everything finite). Half of the candidates of a typical pixel end there,
// but only after the sample point, both normalisations and the cosines were computed, and — one pixel per lane — the
// warp pays for the full evaluation as long as one lane needs it. This test decides from the light's bounding sphere
// {c, R} alone and is SUFFICIENT for the exact test (never the other way round: a candidate it lets through is simply
// handled as before), so the result cannot change. With M >= every |coordinate| of the light's vertices and of pos,
// n1 = |n.x| + |n.y| + |n.z| >= |n|, u = 2^-24, and H = (c - pos).n + R n1 >= (v - pos).n for every point v of the light:
//   * the computed sample point is sum(b_i v_i) + e with computed weights b_i >= 0, |sum(b_i) - 1| <= 4u, |e| <= 5uM
//     per component; lightDir = fl(point - pos) adds <= 2uM; so (computed lightDir).n <= (1 - 4u) H + 11 u M n1;
//   * normalising multiplies by a positive number, and the computed dot of the unit vector with n is off by <= 4 u n1;
//   so H < -2^-14 M n1 (a thousand times the error terms, including those of evaluating H itself in float) makes the
//   computed dot(lightDir, n) negative: cosThetaI = max(., 0) == 0. Also |lightDir| >= 2^-15 M, so r_sqr is a normal,
//   finite number for 1e-12 < M < 1e15, and lightDir's unit vector is finite;
//   * R is finite only for lights whose pick probability, 1/area and emission are finite and positive and whose vertex
//     normals interpolate to something normalisable (rb_host_scene.h), which makes d_ny finite and the sample valid.
struct HorizonCull {
  bool on;
  float n1, pad;
};
RB_HD HorizonCull make_horizon_cull(const SceneDev& sc, const GElem& g, const Shading& sh) {
  HorizonCull hc;
  const float M = gmax(gmax(sc.maxabs, fabsf_(g.pos.x)), gmax(fabsf_(g.pos.y), fabsf_(g.pos.z)));
  hc.n1 = fabsf_(g.normal.x) + fabsf_(g.normal.y) + fabsf_(g.normal.z);
  hc.pad = M * (1.0f / 16384.0f);
  hc.on = sh.finite && sc.light_cull != nullptr && M > 1e-12f && M < 1e15f && hc.n1 < 1e15f;
  return hc;
}
RB_HD bool surely_below_horizon(const HorizonCull& hc, const F4& sphere, const GElem& g) {
  const float H = dot(xyz(sphere) - g.pos, g.normal) + (sphere.w + hc.pad) * hc.n1;
  return H < 0.0f;  // false for R = +inf and for NaN
}

// One area-sampled candidate of initialRenderPass (areaSampleLight :89-124 + evaluatePHat :180-211 + the weight, :247-262):
// light `idx` was picked for candidate i. Returns false when the candidate's weight is exactly +0 by the horizon
// argument below (it then only counts); otherwise the sample, its weight w for addSample, p-hat and validity.
struct AreaEval {
  LightSample s;
  float w, p_hat;
  bool valid;
};
template <class Vis>
RB_HD bool eval_area_candidate(const FrameCtx& fc, const GElem& g, const Shading& sh, uint32_t key, int i, uint32_t idx,
                               float inv_MArea, bool testVis, const Vis& vis, Cnt& cnt, AreaEval* out) {
  const RbParams& P = fc.P;
  const uint32_t base = 4u * (uint32_t)i;
  const F4* L = fc.sc.light + 6 * (size_t)idx;
  const F4 l0 = ldg4(L), l1 = ldg4(L + 1), l2 = ldg4(L + 2), l3 = ldg4(L + 3), l4 = ldg4(L + 4), l5 = ldg4(L + 5);
  const float pick_pdf = P.lightSampler == RB_LS_ALIAS ? l1.w : cdf_pick_pdf(fc.sc, idx);
  const float r1 = rng_value(key, base + 1, 0, 1), r2 = rng_value(key, base + 2, 0, 1);
  const float sq = sqrtf_(r1);
  const float bx = 1.0f - sq;
  const float by = sq * (1.0f - r2);
  const float bz = sq * r2;
  LightSample& s = out->s;
  s.samplePoint = xyz(l0) * bx + xyz(l1) * by + xyz(l2) * bz;
  s.sampleNormal = normalize(xyz(l3) * bx + xyz(l4) * by + xyz(l5) * bz);
  s.L_i = v3(l3.w, l4.w, l5.w);
  s.lightIdx = (int)idx;
  // shared by areaSampleLight (:102-107) and evaluateF (:192-199): same expressions, same bits
  V3 lightDir = s.samplePoint - g.pos;
  const float r_sqr = dot(lightDir, lightDir);
  lightDir = normalize(lightDir);
  const float d_ny = dot(-lightDir, s.sampleNormal);
  const float cosThetaI = gmax(dot(lightDir, g.normal), 0.0f);
  const bool valid = sample_valid(s);  // false only for a non-positive emission
  out->valid = valid;
  if (valid && testVis) cnt.anyW++;
  // A light below the horizon has G == 0: with finite BRDF, pdf and weights the candidate's w is exactly +0,
  // so addSample only counts it (w_sum += 0; no draw consumed affects the result: rand < 0 is false).
  if (sh.finite && cosThetaI == 0.0f && r_sqr > 0.0f && finitef_(r_sqr) && pick_pdf > 0.0f && finitef_(l2.w) && l2.w > 0.0f &&
      finitef_(d_ny) && finite3(s.L_i))
    return false;
  const float triPointPdf = l2.w;  // 1.0f / area, computed with the same division at upload
  const float pdf_area = pick_pdf * triPointPdf;
  const float cosThetaY_pdf = gmax(d_ny, 0.0f);
  const float areaMeasureFactor = cosThetaY_pdf / r_sqr;
  const Lobe lobe = phong_lobe(g, sh, lightDir);
  const float pdfAsIfBrdfAreaMeasure = pdf_from_lobe(g, sh, lightDir, lobe.pdf) * areaMeasureFactor;
  const float W = frcp_(pdf_area);
  const float misWeight = m_area(P, pdf_area, pdfAsIfBrdfAreaMeasure);
  // evaluatePHat(sample, pixel)
  float p_hat = 0.0f;
  if (valid) {
    const float cosThetaY = fabsf_(d_ny);
    const float G = cosThetaI * cosThetaY / r_sqr;
    const V3 f_r = brdf_from_lobe(g, sh, lobe.brdf);
    V3 F = s.L_i * f_r * G;
    if (testVis) {
      cnt.anyW--;  // shadow_F0 counts it again
      F = shadow_F0(F, g.pos, s.samplePoint, vis, i, cnt, 1, VIS_TRACE);
    }
    p_hat = length(F);
  }
  out->p_hat = p_hat;
  out->w = (P.M_Brdf > 0) ? misWeight * p_hat * W : inv_MArea * p_hat * W;
  return true;
}
template <class Vis>
RB_HD void initial_finish(const FrameCtx& fc, size_t pi, const GElem& g, const Shading& sh, uint32_t key, bool testVis, Reservoir& r,
                          float p_sel, bool sel_wants, const Vis& vis, Cnt& cnt);

#if !defined(__CUDACC__)
// Host builds only (tests/emu): with check_mode set, a pre-culled candidate is evaluated all the same and the exact test
// must cull it too; counts = {pre-culled, confirmed by the exact test, violations, candidates,
// culled by the exact test (with or without the pre-test)}.
struct HorizonCullCheck {
  int check_mode = 0;
  unsigned long long counts[5] = {0, 0, 0, 0, 0};
};
inline HorizonCullCheck g_horizon_cull_check;
#define RB_CULL_CHECK(...) __VA_ARGS__
#else
#define RB_CULL_CHECK(...)
#endif

// First loop of the initial pass over candidates c0 .. c0 + cn - 1: picks every candidate's light (parked in `picks`) and
// runs the horizon pre-test. Returns the mask of the candidates that are left to evaluate; a pre-culled one only counts.
RB_HD uint32_t pick_and_precull(const FrameCtx& fc, const GElem& g, const HorizonCull& hc, uint32_t key, int c0, int cn,
                                const PickStore& picks, bool testVis, int& confidence, Cnt& cnt, uint32_t* pre_out) {
  const RbParams& P = fc.P;
  uint32_t todo = 0;
  RB_CULL_CHECK(uint32_t pre = 0;)
  // kPickBatch candidates at a time: their alias records, then their bounding spheres, are independent gathers
  for (int j0 = 0; j0 < cn; j0 += kPickBatch) {
    LightPick pk[kPickBatch];
    F4 sphere[kPickBatch];
#pragma unroll
    for (int u = 0; u < kPickBatch; ++u) pk[u] = pick_light(fc.sc, P.lightSampler, key, 4u * (uint32_t)(c0 + imin(j0 + u, cn - 1)));
#pragma unroll
    for (int u = 0; u < kPickBatch; ++u)
      if (hc.on) sphere[u] = ldg4(fc.sc.light_cull + pk[u].idx);
#pragma unroll
    for (int u = 0; u < kPickBatch; ++u) {
      const int j = j0 + u;
      if (j >= cn) break;
      picks.put(j, pk[u].idx);
      bool culled = false;
      if (hc.on) culled = surely_below_horizon(hc, sphere[u], g) && (P.lightSampler == RB_LS_ALIAS || pk[u].pdf > 0.0f);
      RB_CULL_CHECK(__atomic_fetch_add(&g_horizon_cull_check.counts[3], 1ull, __ATOMIC_RELAXED);
                    if (culled) __atomic_fetch_add(&g_horizon_cull_check.counts[0], 1ull, __ATOMIC_RELAXED);
                    if (culled && g_horizon_cull_check.check_mode) pre |= 1u << j, culled = false;)
      if (culled) {
        confidence += 1;
        if (testVis) cnt.anyW++;  // the sample is valid (eligible lights emit), the reference would have traced
      } else {
        todo |= 1u << j;
      }
    }
  }
  RB_CULL_CHECK(*pre_out = pre;)
  (void)pre_out;
  return todo;
}

template <class Vis>
RB_HD void initial_pixel(const FrameCtx& fc, int x, int y, const Vis& vis, Cnt& cnt, const PickStore& picks) {
  const size_t pi = (size_t)y * fc.width + x;
  const GElem g = load_gelem(fc.G, pi);
  Reservoir r = empty_reservoir();
  if (g.isEmissive || fc.sc.n_lights == 0) {
    if (Vis::kStore) store_reservoir(fc.Rwrite, pi, r);
    return;
  }
  const RbParams& P = fc.P;
  const uint32_t key = rng_pixel_key(fc.frame_key, (uint32_t)pi);
  const Shading sh = make_shading(g, fc.cam.pos);
  const bool testVis = !P.doVisibilityPass;
  float p_sel = 0.0f;  // p-hat of the currently selected sample (== the final evaluatePHat, :289)
  bool sel_wants = false;

  const float inv_MArea = P.M_Area > 0 ? 1.0f / (float)P.M_Area : 0.0f;
  const HorizonCull hc = make_horizon_cull(fc.sc, g, sh);
  // Two loops per chunk of candidates. The first picks every candidate's light (the counter RNG makes the picks
  // independent of each other and of the reservoir) and runs the horizon pre-test: cheap and the same for every lane.
  // The second evaluates only the candidates that are left, in candidate order, so addSample sees the same sequence of
  // non-zero weights; a pre-culled candidate only counts (w == +0: w_sum += 0, no draw can select it).
  for (int c0 = 0; c0 < P.M_Area; c0 += kPickChunk) {
    const int cn = imin(kPickChunk, P.M_Area - c0);
    uint32_t pre = 0;
    uint32_t todo = pick_and_precull(fc, g, hc, key, c0, cn, picks, testVis, r.confidence, cnt, &pre);
    (void)pre;
    while (todo != 0) {
      const int j = lowest_bit(todo);
      todo &= todo - 1u;
      const int i = c0 + j;
      AreaEval e;
      if (!eval_area_candidate(fc, g, sh, key, i, picks.get(j), inv_MArea, testVis, vis, cnt, &e)) {
        RB_CULL_CHECK(if (pre >> j & 1u) __atomic_fetch_add(&g_horizon_cull_check.counts[1], 1ull, __ATOMIC_RELAXED);
                      __atomic_fetch_add(&g_horizon_cull_check.counts[4], 1ull, __ATOMIC_RELAXED);)
        r.confidence += 1;
        continue;
      }
      RB_CULL_CHECK(if (pre >> j & 1u) __atomic_fetch_add(&g_horizon_cull_check.counts[2], 1ull, __ATOMIC_RELAXED);)
      if (add_sample(r, e.s, e.w, 1, key, 4u * (uint32_t)i + 3u)) {
        p_sel = e.p_hat;
        sel_wants = e.valid;
      }
    }
  }
  initial_finish(fc, pi, g, sh, key, testVis, r, p_sel, sel_wants, vis, cnt);
}

// The rest of initialRenderPass after the area candidates: the BRDF-sampled candidates (:266-286), the final W (:288-291),
// the store and — fused — the stream half of the visibility pass.
template <class Vis>
RB_HD void initial_finish(const FrameCtx& fc, size_t pi, const GElem& g, const Shading& sh, uint32_t key, bool testVis, Reservoir& r,
                          float p_sel, bool sel_wants, const Vis& vis, Cnt& cnt) {
  const RbParams& P = fc.P;
  const float inv_MBrdf = P.M_Brdf > 0 ? 1.0f / (float)P.M_Brdf : 0.0f;
  for (int i = 0; i < P.M_Brdf; ++i) {
    const uint32_t base = 4u * (uint32_t)(P.M_Area + i);
    float pdf;
    V3 wi;
    if (!vis.get_brdf_sample(i, &wi, &pdf)) wi = brdf_sample(g, sh, key, base, &pdf);
    const V3 org = g.pos + P.normalOffset * g.normal;
    cnt.closest++;
    const SurfaceHit h = vis.closest(i, org, wi, FLT_MIN + P.tnearOffset, FLT_MAX);
    LightSample s = invalid_sample();
    float W = 0, misWeight = 0;
    if (h.didHit && h.emissiveId >= 0) {
      V3 lightDir = h.hitPoint - g.pos;
      const float r_sqr = dot(lightDir, lightDir);
      lightDir = normalize(lightDir);
      const float cosThetaY = gmax(dot(-lightDir, h.normal), 0.0f);
      const float areaMeasureFactor = cosThetaY / r_sqr;
      const F4* L = fc.sc.light + 6 * (size_t)h.emissiveId;
      const F4 l0 = ldg4(L);
      float pdf_area = l0.w / fc.sc.total_area;  // TriangleCDF::getPDFForTriangle, P/TriangleCDF.h:25-31
      pdf_area *= frcp_(l0.w);
      const float brdfPdfAreaMeasure = pdf * areaMeasureFactor;
      s.samplePoint = h.hitPoint;
      s.sampleNormal = h.normal;
      const F4 m2 = ldg4(fc.sc.mat + 3 * (size_t)h.material + 2);
      s.L_i = xyz(m2);
      s.lightIdx = h.emissiveId;
      W = frcp_(brdfPdfAreaMeasure);
      misWeight = m_brdf(P, brdfPdfAreaMeasure, pdf_area);
    }
    bool wants;
    const float p_hat = eval_phat(s, g, sh, testVis, vis, P.M_Area + i, cnt, 1, VIS_TRACE, &wants);
    const float w = (P.M_Area > 0) ? misWeight * p_hat * W : inv_MBrdf * p_hat * W;
    if (add_sample(r, s, w, 1, key, base + 3)) {
      p_sel = p_hat;
      sel_wants = wants;
    }
  }
  // final evaluatePHat(r.bestSample) — the selected candidate's value again (:289)
  if (testVis && sel_wants) cnt.anyW++;
  const float p_hat = p_sel;
  r.W = p_hat > 0.0f ? frcp_(p_hat) * r.w_sum : 0.0f;
  r.confidence = imin(r.confidence, P.confidenceCap);
  if (Vis::kStore) store_reservoir(fc.Rwrite, pi, r);
  // stream half of the visibility pass (visibility_pixel<GenVis>) for the reservoir just produced: same origin,
  // target and W test as the pass would read back from memory
  if (Vis::kStore && (fc.wave.fuse_vis & 1u) && r.W != 0.0f) {
    const GenVis gv = {&fc, (uint32_t)pi};
    (void)gv.visible(0, g.pos, r.bestSample.samplePoint);
  }
}

// Stream half of the initial pass: only the BRDF-sampled closest-hit rays (brdfSampleLight, :136-141). The
// directions depend on the G-buffer element and the RNG alone, so they are emitted up front, traced by the
// closest-hit kernel, and consumed by initial_pixel through ResolveVis::closest.
RB_HD void initial_brdf_gen_pixel(const FrameCtx& fc, int x, int y, const GenVis& vis) {
  const size_t pi = (size_t)y * fc.width + x;
  const RbParams& P = fc.P;
  const GElem g = load_gelem(fc.G, pi);
  if (g.isEmissive || fc.sc.n_lights == 0) {
    // no ray: leave "miss" records so that the second step of the two-step scheme has nothing to do here
    if (fc.wave.brdf_two_step)
      for (int i = 0; i < P.M_Brdf; ++i) fc.wave.hits[(size_t)i * fc.wave.npix + pi].tri = 0xFFFFFFFFu;
    return;
  }
  const uint32_t key = rng_pixel_key(fc.frame_key, (uint32_t)pi);
  const Shading sh = make_shading(g, fc.cam.pos);
  for (int i = 0; i < P.M_Brdf; ++i) {
    const uint32_t base = 4u * (uint32_t)(P.M_Area + i);
    float pdf;
    const V3 wi = brdf_sample(g, sh, key, base, &pdf);
    const V3 org = g.pos + P.normalOffset * g.normal;
    vis.put_brdf_sample(i, wi, pdf);
    (void)vis.closest(i, org, wi, FLT_MIN + P.tnearOffset, FLT_MAX);
  }
}

// Second step of the two-step BRDF-candidate rays: where the emissive-only BVH reported a hit (t, id), the full BVH is
// asked whether anything precedes it (trace8_precedes / k_trace_queue<true, true>). Those rays are queued by the first
// traversal itself (brdf_chain_push); few rays reach this step: most BRDF-sampled directions do not point at an emitter.

// =====================================================================================
// Pass 2: visibility (ReSTIRIntegrator::visibilityPass, :302-312). The reference traces for every pixel, even
// with an empty reservoir (sample point = -FLT_MAX sentinel); the only effect of the ray is W = 0, so it is
// not traced where W is already exactly 0.
// =====================================================================================
template <class Vis>
RB_HD void visibility_pixel(const FrameCtx& fc, int x, int y, const Vis& vis, Cnt& cnt) {
  const size_t pi = (size_t)y * fc.width + x;
  cnt.anyW++;
  F4 nw = ld4(fc.Rwrite.normal_W + pi);
  if (nw.w == 0.0f) return;
  const F4 pw = ld4(fc.Rwrite.point_wsum + pi);
  const F4 pd = ld4(fc.G.pos_depth + pi);
  cnt.anyT++;
  const bool V = vis.visible(0, xyz(pd), xyz(pw));
  if (!V) {
    nw.w = 0.0f;
    if (Vis::kStore) st4(fc.Rwrite.normal_W + pi, nw);
  }
}

// =====================================================================================
// Pass 3: temporal reuse (ReSTIRIntegrator::temporalReusePass, :625-732;
// reprojectBackward / reprojectForward :544-587)
// =====================================================================================
RB_HD bool reproject(const CamState& cam, int width, int height, const V3& wsPos, int* sx, int* sy) {
  const float* m = cam.viewMat;  // glm mat4*vec4: (m0*x + m1*y) + (m2*z + m3*w)
  const float vx = (m[0] * wsPos.x + m[4] * wsPos.y) + (m[8] * wsPos.z + m[12] * 1.0f);
  const float vy = (m[1] * wsPos.x + m[5] * wsPos.y) + (m[9] * wsPos.z + m[13] * 1.0f);
  const float vz = (m[2] * wsPos.x + m[6] * wsPos.y) + (m[10] * wsPos.z + m[14] * 1.0f);
  if (vz >= 0) return false;
  const float fx = roundf((-vx / vz) * cam.focal + (float)width / 2.0f);  // glm::round: half away from zero
  const float fy = roundf((vy / vz) * cam.focal + (float)height / 2.0f);
  if (!(fx >= -1.0f)) return false;
  if (!(fy >= -1.0f)) return false;
  if (fx > (float)width || fy > (float)height) return false;
  const int screenX = (int)fx, screenY = (int)fy;
  if (screenX < 0 || screenX > width - 1 || screenY < 0 || screenY > height - 1) return false;
  *sx = screenX;
  *sy = screenY;
  return true;
}

// BANDED: the handle renders a band of the image, so reprojected pixels may lie outside its G-buffer rows and are
// re-derived (fetch_*); the single-band instantiation reads the planes directly and carries no traversal code.
template <class Vis, bool BANDED>
RB_HD void temporal_pixel(const FrameCtx& fc, int x, int y, const Vis& vis, Cnt& cnt) {
  const size_t pi = (size_t)y * fc.width + x;
  const RbParams& P = fc.P;
  const Reservoir cur = load_reservoir(fc.Rread, pi);
  const GElem curElem = load_gelem(fc.G, pi);
  int px, py;
  if (!reproject(fc.prevCam, fc.width, fc.height, curElem.pos, &px, &py)) {
    if (Vis::kStore) store_reservoir(fc.Rwrite, pi, cur);
    return;
  }
  const GElem prevElem = BANDED ? fetch_gelem(fc, true, px, py, cnt) : load_gelem(fc.Gprev, (size_t)py * fc.width + px);
  const V3 curCam = fc.cam.pos, prevCam = fc.prevCam.pos;
  const float currentDepth = length(curElem.pos - curCam);
  const float prevDepth = length(prevElem.pos - prevCam);
  const float depthRatio = currentDepth > prevDepth ? prevDepth / currentDepth : currentDepth / prevDepth;
  if (depthRatio < 0.9f) {
    if (Vis::kStore) store_reservoir(fc.Rwrite, pi, cur);
    return;
  }
  const V3 prevPosAtCurrent = xyz(ld4(fc.Gprev.pos_depth + pi));
  int fx, fy;
  if (!reproject(fc.cam, fc.width, fc.height, prevPosAtCurrent, &fx, &fy)) {
    if (Vis::kStore) store_reservoir(fc.Rwrite, pi, cur);
    return;
  }
  const V3 fwPos = BANDED ? fetch_gpos(fc, false, fx, fy, cnt) : xyz(ld4(fc.G.pos_depth + (size_t)fy * fc.width + fx));
  const float currentDepthP = length(prevPosAtCurrent - prevCam);
  const float prevDepthP = length(fwPos - curCam);
  const float depthRatioP = currentDepthP > prevDepthP ? prevDepthP / currentDepthP : currentDepthP / prevDepthP;
  if (depthRatioP < 0.9f) {
    if (Vis::kStore) store_reservoir(fc.Rwrite, pi, cur);
    return;
  }
  // same pixel, not the reprojected one (:641) — unless the repaired variant is asked for
  const Reservoir prev = load_reservoir(fc.Rlast, P.temporalFetchReprojected ? (size_t)py * fc.width + px : pi);
  const uint32_t key = rng_pixel_key(fc.frame_key, (uint32_t)pi);
  const Shading shCur = make_shading(curElem, curCam), shPrev = make_shading(prevElem, prevCam);

  // Four distinct p-hat values feed the seven evaluations of the reference. A and B only reach the output through
  // w_cur = m_cur * A * cur.W, C and D only through w_prev = m_prev * C * prev.W (and the final p-hat, which is A or C
  // and only when that sample was selected, i.e. its W is not 0).
  const VisMode modeA = vis_mode_from_W(cur.W, true);
  const VisMode modeB = cur.W == 0.0f ? VIS_IRRELEVANT : VIS_TRACE;
  const VisMode modeCD = prev.W == 0.0f ? VIS_IRRELEVANT : VIS_TRACE;
  bool wA, wC;
  const float A = eval_phat(cur.bestSample, curElem, shCur, true, vis, 0, cnt, 2, modeA, &wA);     // p_cur, p_hat_cur
  const float B = eval_phat(cur.bestSample, prevElem, shPrev, true, vis, 1, cnt, 1, modeB);        // p_prev
  const float C = eval_phat(prev.bestSample, curElem, shCur, true, vis, 2, cnt, 2, modeCD, &wC);   // p_cur', p_hat_prev
  const float D = eval_phat(prev.bestSample, prevElem, shPrev, true, vis, 3, cnt, 1, modeCD);      // p_prev'

  Reservoir out = empty_reservoir();
  float m_cur = A * (float)cur.confidence / (A * (float)cur.confidence + B * (float)prev.confidence);
  if (!(m_cur > 0)) m_cur = 0.0f;
  const float w_cur = m_cur * A * cur.W;
  int selected = -1;
  if (add_sample(out, cur.bestSample, w_cur, cur.confidence, key, 0)) selected = 0;
  float m_prev = D * (float)prev.confidence / (C * (float)cur.confidence + D * (float)prev.confidence);
  if (!(m_prev > 0)) m_prev = 0.0f;
  const float w_prev = m_prev * C * prev.W;
  if (add_sample(out, prev.bestSample, w_prev, prev.confidence, key, 1)) selected = 1;
  out.confidence = imin(out.confidence, P.confidenceCap);
  float final_p_hat = 0.0f;
  if (selected == 0) {
    final_p_hat = A;
    if (wA) cnt.anyW++;
  } else if (selected == 1) {
    final_p_hat = C;
    if (wC) cnt.anyW++;
  }
  out.W = final_p_hat > 0.0f ? out.w_sum / final_p_hat : 0.0f;
  if (Vis::kStore) store_reservoir(fc.Rwrite, pi, out);
}

// ---- temporal reuse split for the wavefront schedule (same idea as the spatial split below) -----------------------
// Stream half: validity tests, the four unshadowed p-hat values, the shadow rays; it leaves two 16-byte records per
// pixel: slot 0 = {A, B, C, D} (values if visible), slot 1 = {flags, 0, 0, 0} with 4 flag bits per evaluation
// (RB_CAND_WANTS / _RAY / _PH0_NAN) and bit 16 = "keep the current reservoir" (a validity test failed).
// Resolve half: the two-way merge from the records, the traced bits and the two reservoirs.
#define RB_CAND_WANTS 1u    // the reference traces a ray for this evaluation
#define RB_CAND_RAY 2u      // a ray was queued: the value depends on the traced bit of this slot
#define RB_CAND_PH0_NAN 4u  // length(F0 * 0) is NaN (F0 not finite)
#define RB_CAND_COPY 8u     // emissive pixel: the reservoir is copied through (spatial, slot 0 only)
#define RB_CAND_INDEX_BITS 27
#define RB_TEMPORAL_KEEP (1u << 16)
// BANDED: 0 = the handle holds the whole image; 1 = band, out-of-rows elements re-derived on the spot (fetch_*);
// 2 = band, bulk launch: a pixel that would need a re-derivation only puts itself on the deferred list (the small
// second launch runs those pixels with BANDED = 1). Same arithmetic in all three.
RB_HD void temporal_defer(const WaveBufs& wv, size_t pi) { wv.deferred[queue_reserve(wv.deferred_count)] = (uint32_t)pi; }
template <int BANDED>
RB_HD void temporal_gen_pixel(const FrameCtx& fc, int x, int y, const GenVis& vis, Cnt& cnt) {
  const size_t pi = (size_t)y * fc.width + x;
  const WaveBufs& wv = fc.wave;
  U4* rec_flags = wv.cand + (size_t)wv.npix + pi;
  if (BANDED != 1 && (wv.fuse_vis & 2u)) {
    // resolve half of the visibility pass (visibility_pixel<ResolveVis>) on this pixel's reservoir, in place, before
    // anything reads it: W = 0 where the ray queued by the initial pass was occluded. (Only the bulk launch does it;
    // a deferred pixel's second run finds the reservoir already final.)
    cnt.anyW++;
    F4 nw = ld4(fc.Rread.normal_W + pi);
    if (nw.w != 0.0f) {
      cnt.anyT++;
      if (wv.occ[pi] != 0) {
        nw.w = 0.0f;
        st4(fc.Rread.normal_W + pi, nw);
      }
    }
  }
  const GElem curElem = load_gelem(fc.G, pi);
  int px, py;
  if (!reproject(fc.prevCam, fc.width, fc.height, curElem.pos, &px, &py)) {
    *rec_flags = U4{RB_TEMPORAL_KEEP, 0u, 0u, 0u};
    return;
  }
  if (BANDED == 2 && !(py >= fc.gpy0 && py < fc.gpy1)) return temporal_defer(wv, pi);
  const GElem prevElem = BANDED == 1 ? fetch_gelem(fc, true, px, py, cnt) : load_gelem(fc.Gprev, (size_t)py * fc.width + px);
  const V3 curCam = fc.cam.pos, prevCam = fc.prevCam.pos;
  const float currentDepth = length(curElem.pos - curCam);
  const float prevDepth = length(prevElem.pos - prevCam);
  const float depthRatio = currentDepth > prevDepth ? prevDepth / currentDepth : currentDepth / prevDepth;
  if (depthRatio < 0.9f) {
    *rec_flags = U4{RB_TEMPORAL_KEEP, 0u, 0u, 0u};
    return;
  }
  const V3 prevPosAtCurrent = xyz(ld4(fc.Gprev.pos_depth + pi));
  int fx, fy;
  if (!reproject(fc.cam, fc.width, fc.height, prevPosAtCurrent, &fx, &fy)) {
    *rec_flags = U4{RB_TEMPORAL_KEEP, 0u, 0u, 0u};
    return;
  }
  if (BANDED == 2 && !(fy >= fc.gy0 && fy < fc.gy1)) return temporal_defer(wv, pi);
  const V3 fwPos = BANDED == 1 ? fetch_gpos(fc, false, fx, fy, cnt) : xyz(ld4(fc.G.pos_depth + (size_t)fy * fc.width + fx));
  const float currentDepthP = length(prevPosAtCurrent - prevCam);
  const float prevDepthP = length(fwPos - curCam);
  const float depthRatioP = currentDepthP > prevDepthP ? prevDepthP / currentDepthP : currentDepthP / prevDepthP;
  if (depthRatioP < 0.9f) {
    *rec_flags = U4{RB_TEMPORAL_KEEP, 0u, 0u, 0u};
    return;
  }
  const Reservoir cur = load_reservoir(fc.Rread, pi);
  // same pixel, not the reprojected one (:641) — unless the repaired variant is asked for (the resolve half finds the
  // pixel index in the flags record)
  const uint32_t prev_pi = fc.P.temporalFetchReprojected ? (uint32_t)(py * fc.width + px) : (uint32_t)pi;
  const Reservoir prev = load_reservoir(fc.Rlast, prev_pi);
  const Shading shCur = make_shading(curElem, curCam), shPrev = make_shading(prevElem, prevCam);
  const VisMode modeA = vis_mode_from_W(cur.W, true);
  const VisMode modeB = cur.W == 0.0f ? VIS_IRRELEVANT : VIS_TRACE;
  const VisMode modeCD = prev.W == 0.0f ? VIS_IRRELEVANT : VIS_TRACE;
  float ph[4];
  uint32_t flags = 0, raymask = 0;
  for (int e = 0; e < 4; ++e) {  // A: cur sample @ cur pixel, B: cur @ prev, C: prev @ cur, D: prev @ prev
    const LightSample& smp = e < 2 ? cur.bestSample : prev.bestSample;
    const GElem& g = (e & 1) ? prevElem : curElem;
    const Shading& sh = (e & 1) ? shPrev : shCur;
    const VisMode vm = e == 0 ? modeA : e == 1 ? modeB : modeCD;
    const int copies = (e & 1) ? 1 : 2;  // A and C stand for two evaluations of the reference each
    bool wants;
    const V3 F0 = eval_F0(smp, g, sh, &wants);
    uint32_t f = wants ? RB_CAND_WANTS : 0u;
    if (wants) {  // shadow_F0
      cnt.anyW += (uint32_t)copies;
      const bool zero = F0.x == 0.0f && F0.y == 0.0f && F0.z == 0.0f;
      if (!zero && vm != VIS_KNOWN && !(vm == VIS_IRRELEVANT && finite3(F0))) {
        cnt.anyT++;
        raymask |= 1u << e;
        f |= RB_CAND_RAY;
        if (!finite3(F0)) f |= RB_CAND_PH0_NAN;
      }
    }
    ph[e] = length(F0 * 1.0f);
    flags |= f << (4 * e);
  }
  {  // the rays of this pixel, one queue reservation
    uint32_t qi = vis.reserve((uint32_t)popc(raymask));
    for (int e = 0; e < 4; ++e)
      if (raymask & (1u << e))
        vis.visible_at(qi++, e, ((e & 1) ? prevElem : curElem).pos, (e < 2 ? cur.bestSample : prev.bestSample).samplePoint);
  }
  wv.cand[pi] = U4{f2u(ph[0]), f2u(ph[1]), f2u(ph[2]), f2u(ph[3])};
  *rec_flags = U4{flags, prev_pi, 0u, 0u};
}

RB_HD void temporal_merge_pixel(const FrameCtx& fc, int x, int y, Cnt& cnt) {
  const size_t pi = (size_t)y * fc.width + x;
  const RbParams& P = fc.P;
  const WaveBufs& wv = fc.wave;
  const U4 frec = wv.cand[(size_t)wv.npix + pi];
  const uint32_t flags = frec.x;
  const Reservoir cur = load_reservoir(fc.Rread, pi);
  if (flags & RB_TEMPORAL_KEEP) {
    store_reservoir(fc.Rwrite, pi, cur);
    return;
  }
  const Reservoir prev = load_reservoir(fc.Rlast, P.temporalFetchReprojected ? (size_t)frec.y : pi);
  const U4 rec = wv.cand[pi];
  const uint32_t key = rng_pixel_key(fc.frame_key, (uint32_t)pi);
  float v[4] = {u2f(rec.x), u2f(rec.y), u2f(rec.z), u2f(rec.w)};
  for (int e = 0; e < 4; ++e) {
    const uint32_t f = (flags >> (4 * e)) & 0xFu;
    if ((f & RB_CAND_RAY) && wv.occ[(size_t)e * wv.npix + pi] != 0) v[e] = (f & RB_CAND_PH0_NAN) ? u2f(0x7FC00000u) : 0.0f;
  }
  const float A = v[0], B = v[1], C = v[2], D = v[3];
  const bool wA = (flags & RB_CAND_WANTS) != 0, wC = ((flags >> 8) & RB_CAND_WANTS) != 0;

  Reservoir out = empty_reservoir();
  float m_cur = A * (float)cur.confidence / (A * (float)cur.confidence + B * (float)prev.confidence);
  if (!(m_cur > 0)) m_cur = 0.0f;
  const float w_cur = m_cur * A * cur.W;
  int selected = -1;
  if (add_sample(out, cur.bestSample, w_cur, cur.confidence, key, 0)) selected = 0;
  float m_prev = D * (float)prev.confidence / (C * (float)cur.confidence + D * (float)prev.confidence);
  if (!(m_prev > 0)) m_prev = 0.0f;
  const float w_prev = m_prev * C * prev.W;
  if (add_sample(out, prev.bestSample, w_prev, prev.confidence, key, 1)) selected = 1;
  out.confidence = imin(out.confidence, P.confidenceCap);
  float final_p_hat = 0.0f;
  if (selected == 0) {
    final_p_hat = A;
    if (wA) cnt.anyW++;
  } else if (selected == 1) {
    final_p_hat = C;
    if (wC) cnt.anyW++;
  }
  out.W = final_p_hat > 0.0f ? out.w_sum / final_p_hat : 0.0f;
  store_reservoir(fc.Rwrite, pi, out);
}

// =====================================================================================
// Pass 4: spatial reuse (ReSTIRIntegrator::spatialReusePass, :316-542;
// Sampling::sampleDiskUniform P/Sampling.cpp:78-87)
// RNG slots: neighbour i uses 2i (theta) and 2i+1 (radius); candidate j accepts on 2k+j.
// =====================================================================================
template <class Vis>
RB_HD void spatial_pixel(const FrameCtx& fc, int x, int y, const Vis& vis, Cnt& cnt) {
  const size_t pi = (size_t)y * fc.width + x;
  const RbParams& P = fc.P;
  const GElem thisElem = load_gelem(fc.G, pi);
  if (thisElem.isEmissive) {
    if (Vis::kStore) store_reservoir(fc.Rwrite, pi, load_reservoir(fc.Rread, pi));
    return;
  }
  const uint32_t key = rng_pixel_key(fc.frame_key, (uint32_t)pi);
  const V3 cam = fc.cam.pos;
  const Shading shThis = make_shading(thisElem, cam);
  const int k = P.spatialReuseNeighborCount;
  uint32_t nb[RB_MAX_NEIGHBORS + 1];  // pixel index of every resampling source; [0] = this pixel
  nb[0] = (uint32_t)pi;
  int M = 1;
  for (int i = 0; i < k; ++i) {
    const float theta = rng_value(key, 2u * i, 0, 2.0f) * RB_PI;
    const float r = sqrtf_(rng_value(key, 2u * i + 1, 0, P.spatialReuseRadius));
    float sn, cs;
    dm::sincosf_(theta, &sn, &cs);
    const float ox = r * cs, oy = r * sn;
    int nx = x + (int)ox, ny = y + (int)oy;
    nx = imin(imax(nx, 0), fc.width - 1);
    ny = imin(imax(ny, 0), fc.height - 1);
    const size_t ni = (size_t)ny * fc.width + nx;
    const F4 st = ld4(fc.G.spec_type + ni);
    if (f2u(st.w) & 0x100u) continue;  // emissive neighbours never hold a reservoir (:345)
    if (P.rejectDissimilarNeighbors) {
      const F4 nn = ld4(fc.G.normal_shin + ni);
      const float normalSimilarity = dot(xyz(nn), thisElem.normal);
      if (normalSimilarity < P.minNormalSimilarity) continue;
      const float ndepth = ld4(fc.G.pos_depth + ni).w;
      float depthRatio = 0;
      if (ndepth > 0) depthRatio = thisElem.depth / ndepth;
      const float halfDepthDiff = P.maxDepthDifference * 0.5f;
      if (depthRatio < 1.0f - halfDepthDiff || depthRatio > 1.0f + halfDepthDiff) continue;
    }
    nb[M++] = (uint32_t)ni;
  }
  const int n = M;
  const float rcpM = M > 0 ? 1.0f / (float)M : 0.0f;
  const int mode = P.spatialWeightCalc;

  int confidenceSum = 0, confidenceSumNonCanonical = 0;
  if (mode == RB_SW_PAIRWISE_MIS) {
    for (int i = 0; i < n; ++i) {
      const int c = (int)f2u(ld4(fc.Rread.Li_conf + nb[i]).w);
      confidenceSum += c;
      if (i != 0) confidenceSumNonCanonical += c;
    }
  }

  Reservoir out = empty_reservoir();
  int selectedSampleIndex = 0;
  bool any_selected = false;
  float p_sel = 0.0f;
  bool sel_wants = false;
  for (int i = 0; i < n; ++i) {
    const Reservoir ri = load_reservoir(fc.Rread, nb[i]);
    const LightSample& si = ri.bestSample;
    float misWeight = rcpM;
    if (mode == RB_SW_BALANCE_HEURISTIC) {
      float misNom = 0, misDenom = 0;
      misWeight = 0.0f;
      for (int j = 0; j < n; ++j) {
        const GElem gj = load_gelem(fc.G, nb[j]);
        const Shading shj = make_shading(gj, cam);
        const int cj = (int)f2u(ld4(fc.Rread.Li_conf + nb[j]).w);
        const float p_hat = eval_phat(si, gj, shj, true, vis, 0, cnt, 1);
        misDenom += p_hat * cj;
        if (i == j) misNom = p_hat * ri.confidence;
      }
      if (misDenom > 0) misWeight = misNom / misDenom;
    }
    if (mode == RB_SW_PAIRWISE_MIS) {
      misWeight = 0.0f;
      if (i == 0) {
        float sum = 0.0f;
        const float p_hat_c = eval_phat(si, thisElem, shThis, true, vis, 0, cnt, 1) * (float)ri.confidence;
        for (int j = 1; j < n; ++j) {
          const GElem gj = load_gelem(fc.G, nb[j]);
          const Shading shj = make_shading(gj, cam);
          const int cj = (int)f2u(ld4(fc.Rread.Li_conf + nb[j]).w);
          const float p_hat_j = eval_phat(si, gj, shj, true, vis, 0, cnt, 1);
          const float denom = p_hat_c + p_hat_j * (float)confidenceSumNonCanonical;
          if (denom > 0) {
            const float confFract = (float)cj / (float)confidenceSum;
            sum += confFract * (p_hat_c / denom);
          }
        }
        misWeight = ((float)ri.confidence / (float)confidenceSum) + sum;
      } else {
        const GElem gi = load_gelem(fc.G, nb[i]);
        const Shading shi = make_shading(gi, cam);
        float p_hat_i = eval_phat(si, gi, shi, true, vis, 0, cnt, 1);
        const float p_hat_c = eval_phat(si, thisElem, shThis, true, vis, 0, cnt, 1);
        p_hat_i *= (float)confidenceSumNonCanonical;
        const int c0 = (int)f2u(ld4(fc.Rread.Li_conf + nb[0]).w);
        const float denom = p_hat_i + p_hat_c * (float)c0;
        if (denom > 0 && confidenceSum > 0) misWeight = ((float)ri.confidence / (float)confidenceSum) * (p_hat_i / denom);
      }
    }
    // resampling weight misWeight * p-hat(sample_i @ this pixel) * W_i: the ray is irrelevant when W_i is exactly 0, and
    // already known unoccluded when the sample is this pixel's own reservoir with W > 0
    const VisMode vm = vis_mode_from_W(ri.W, nb[i] == (uint32_t)pi);
    bool wants;
    const float resamplingPhat = eval_phat(si, thisElem, shThis, true, vis, i, cnt, 1, vm, &wants);
    const float resamplingWeight = misWeight * resamplingPhat * ri.W;
    if (add_sample(out, si, resamplingWeight, ri.confidence, key, 2u * k + i)) {
      selectedSampleIndex = i;
      any_selected = true;
      p_sel = resamplingPhat;
      sel_wants = wants;
    }
  }
  // final evaluatePHat(resultReservoir.bestSample @ this pixel): the selected candidate's value again (:481)
  if (any_selected && sel_wants) cnt.anyW++;
  const float final_p_hat = any_selected ? p_sel : 0.0f;
  if (mode == RB_SW_CONSTANT || mode == RB_SW_BALANCE_HEURISTIC || mode == RB_SW_PAIRWISE_MIS) {
    out.W = final_p_hat > 0.0f ? out.w_sum / final_p_hat : 0.0f;
  } else if (mode == RB_SW_CONSTANT_DEBIAS_Z_TERM) {
    int Z = 0;
    float correctionFactor = 1.0f;
    vis.begin_post_selection();
    for (int i = 0; i < n; ++i) {
      const V3 pj = xyz(ld4(fc.G.pos_depth + nb[i]));
      cnt.anyW++;
      cnt.anyT++;
      if (vis.visible(0, pj, out.bestSample.samplePoint)) Z += 1;
    }
    if (Z > 0 && M > 0) correctionFactor = (1.0f / (float)Z) / rcpM;
    out.W = final_p_hat > 0.0f ? correctionFactor * out.w_sum / final_p_hat : 0.0f;
  } else if (mode == RB_SW_CONSTANT_DEBIAS_CONTRIB) {
    vis.begin_post_selection();
    const Reservoir rs = load_reservoir(fc.Rread, nb[selectedSampleIndex]);
    float misNom = 0, misDenom = 0, contribWeight = 0, correctionFactor = 0;
    for (int i = 0; i < n; ++i) {
      const GElem gi = load_gelem(fc.G, nb[i]);
      const Shading shi = make_shading(gi, cam);
      const int ci = (int)f2u(ld4(fc.Rread.Li_conf + nb[i]).w);
      const float p_hat = eval_phat(rs.bestSample, gi, shi, true, vis, 0, cnt, 1);
      misDenom += p_hat * (float)ci;
      if (i == selectedSampleIndex) misNom = p_hat * (float)ci;
    }
    if (misDenom > 0) contribWeight = misNom / misDenom;
    if (M > 0) correctionFactor = contribWeight / rcpM;
    out.W = final_p_hat > 0.0f ? correctionFactor * out.w_sum / final_p_hat : 0.0f;
  }
  out.confidence = imin(out.confidence, P.confidenceCap);
  if (Vis::kStore) store_reservoir(fc.Rwrite, pi, out);
}

// ---- spatial reuse with constant weights, split for the wavefront schedule -------------------------------------
// The stream half (spatial_gen_pixel) does everything of spatial_pixel that does not need a traced result: neighbour
// choice, the unshadowed p-hat of every candidate, the shadow rays. It leaves one 16-byte record per candidate:
//   { p-hat if visible, W_i, confidence_i, neighbour pixel index | flags << 27 }      (slot 0: M instead of the index)
// The resolve half (spatial_merge_pixel) replays the reservoir updates from the records and the traced bits and
// fetches only the selected neighbour's sample: no G-buffer access, no p-hat evaluation.
// The arithmetic is spatial_pixel's, expression by expression (F0 * V with V in {0,1}: length(F0 * 1) and
// length(F0 * 0) are both formed from F0 in the stream half).

RB_HD void spatial_gen_pixel(const FrameCtx& fc, int x, int y, const GenVis& vis, Cnt& cnt) {
  const size_t pi = (size_t)y * fc.width + x;
  const RbParams& P = fc.P;
  const WaveBufs& wv = fc.wave;
  const GElem thisElem = load_gelem(fc.G, pi);
  if (thisElem.isEmissive) {
    wv.cand[pi] = U4{0u, 0u, 0u, RB_CAND_COPY << RB_CAND_INDEX_BITS};
    return;
  }
  const uint32_t key = rng_pixel_key(fc.frame_key, (uint32_t)pi);
  const Shading shThis = make_shading(thisElem, fc.cam.pos);
  const int k = P.spatialReuseNeighborCount;
  uint32_t nb[RB_MAX_NEIGHBORS + 1];
  nb[0] = (uint32_t)pi;
  int M = 1;
  for (int i = 0; i < k; ++i) {  // neighbour choice: as in spatial_pixel
    const float theta = rng_value(key, 2u * i, 0, 2.0f) * RB_PI;
    const float r = sqrtf_(rng_value(key, 2u * i + 1, 0, P.spatialReuseRadius));
    float sn, cs;
    dm::sincosf_(theta, &sn, &cs);
    const float ox = r * cs, oy = r * sn;
    int nx = x + (int)ox, ny = y + (int)oy;
    nx = imin(imax(nx, 0), fc.width - 1);
    ny = imin(imax(ny, 0), fc.height - 1);
    const size_t ni = (size_t)ny * fc.width + nx;
    const F4 st = ld4(fc.G.spec_type + ni);
    if (f2u(st.w) & 0x100u) continue;
    if (P.rejectDissimilarNeighbors) {
      const F4 nn = ld4(fc.G.normal_shin + ni);
      const float normalSimilarity = dot(xyz(nn), thisElem.normal);
      if (normalSimilarity < P.minNormalSimilarity) continue;
      const float ndepth = ld4(fc.G.pos_depth + ni).w;
      float depthRatio = 0;
      if (ndepth > 0) depthRatio = thisElem.depth / ndepth;
      const float halfDepthDiff = P.maxDepthDifference * 0.5f;
      if (depthRatio < 1.0f - halfDepthDiff || depthRatio > 1.0f + halfDepthDiff) continue;
    }
    nb[M++] = (uint32_t)ni;
  }
  uint64_t raymask = 0;
  for (int i = 0; i < M; ++i) {
    // only what the weight needs of the neighbour's reservoir: sample point/normal/L_i, W, confidence
    const F4 a = ld4(fc.Rread.point_wsum + nb[i]), b = ld4(fc.Rread.normal_W + nb[i]), c = ld4(fc.Rread.Li_conf + nb[i]);
    LightSample si;
    si.samplePoint = xyz(a), si.sampleNormal = xyz(b), si.L_i = xyz(c), si.lightIdx = 0;
    const float Wi = b.w;
    const VisMode vm = vis_mode_from_W(Wi, nb[i] == (uint32_t)pi);
    bool wants;
    const V3 F0 = eval_F0(si, thisElem, shThis, &wants);
    uint32_t flags = wants ? RB_CAND_WANTS : 0u;
    if (wants) {  // shadow_F0
      cnt.anyW += 1;
      const bool zero = F0.x == 0.0f && F0.y == 0.0f && F0.z == 0.0f;
      if (!zero && vm != VIS_KNOWN && !(vm == VIS_IRRELEVANT && finite3(F0))) {
        cnt.anyT++;
        raymask |= 1ull << i;
        flags |= RB_CAND_RAY;
        if (!finite3(F0)) flags |= RB_CAND_PH0_NAN;
      }
    }
    const float ph1 = length(F0 * 1.0f);
    const uint32_t w3 = (i == 0 ? (uint32_t)M : nb[i]) | (flags << RB_CAND_INDEX_BITS);
    wv.cand[(size_t)i * wv.npix + pi] = U4{f2u(ph1), f2u(Wi), f2u(c.w), w3};
  }
  // the rays of this pixel, one queue reservation; the sample points are read again (they are in L1 / L2)
  uint32_t n_rays = 0;
  for (uint64_t m = raymask; m; m &= m - 1) ++n_rays;
  uint32_t qi = vis.reserve(n_rays);
  for (int i = 0; i < M; ++i)
    if ((raymask >> i) & 1ull) vis.visible_at(qi++, i, thisElem.pos, xyz(ld4(fc.Rread.point_wsum + nb[i])));
}

// Final shading of one pixel from its reservoir (body of shade_pixel below, see there). The shadow ray of the
// reference can never change the pixel: W > 0 means the sample was found unoccluded from this pixel by the pass that
// produced the reservoir; with W == 0 or NaN every component of f * V * W is 0 or NaN for V = 0 and V = 1 alike, and a
// NaN component zeroes the pixel in sanitize. KnownVis therefore answers "visible" without tracing (it is reached only
// in the NaN / non-finite cases, where shadow_F0 still counts the ray the reference traces).
struct KnownVis {
  static constexpr bool kStore = true;
  RB_HD bool visible(int, const V3&, const V3&) const { return true; }
};
RB_HD void shade_store(const FrameCtx& fc, size_t pi, const Reservoir& r, Cnt& cnt) {
  V3 pixel;
  if (r.w_sum > 0.0f) {  // Reservoir::hasSample
    const GElem g = load_gelem(fc.G, pi);
    const Shading sh = make_shading(g, fc.cam.pos);
    const V3 f = eval_F(r.bestSample, g, sh, true, KnownVis{}, 0, cnt, 1, vis_mode_from_W(r.W, true));
    pixel = f * r.W;
  } else {
    pixel = xyz(ld4(fc.G.emission + pi));
  }
  if (pixel.x != pixel.x || pixel.y != pixel.y || pixel.z != pixel.z) pixel = v3(0);
  if (pixel.x < 0 || pixel.y < 0 || pixel.z < 0) pixel = v3(0);
  float* o = fc.frame + 3 * pi;
  o[0] = pixel.x;
  o[1] = pixel.y;
  o[2] = pixel.z;
}

RB_HD void spatial_merge_pixel(const FrameCtx& fc, int x, int y, Cnt& cnt) {
  const size_t pi = (size_t)y * fc.width + x;
  const RbParams& P = fc.P;
  const WaveBufs& wv = fc.wave;
  const U4 r0 = wv.cand[pi];
  if ((r0.w >> RB_CAND_INDEX_BITS) & RB_CAND_COPY) {
    const Reservoir keep = load_reservoir(fc.Rread, pi);
    store_reservoir(fc.Rwrite, pi, keep);
    if (wv.fuse_shade) shade_store(fc, pi, keep, cnt);
    return;
  }
  const uint32_t key = rng_pixel_key(fc.frame_key, (uint32_t)pi);
  const int k = P.spatialReuseNeighborCount;
  const int M = (int)(r0.w & ((1u << RB_CAND_INDEX_BITS) - 1u));
  const float rcpM = M > 0 ? 1.0f / (float)M : 0.0f;
  float w_sum = 0.0f, p_sel = 0.0f;
  int confidence = 0;
  uint32_t sel_pixel = 0;
  bool any_selected = false, sel_wants = false;
  // Reservoir::addSample replayed from one candidate record and its traced bit
  auto replay = [&](int i, const U4& r, uint32_t occluded) {
    const uint32_t flags = r.w >> RB_CAND_INDEX_BITS;
    float ph = u2f(r.x);
    if ((flags & RB_CAND_RAY) && occluded != 0)  // occluded: length(F0 * 0)
      ph = (flags & RB_CAND_PH0_NAN) ? u2f(0x7FC00000u) : 0.0f;
    const float w = rcpM * ph * u2f(r.y);
    w_sum += w;
    confidence += (int)r.z;
    if (w == 0 && w_sum == 0) return;
    if (rng_value(key, 2u * k + i, 0, 1) < w / w_sum) {
      sel_pixel = i == 0 ? (uint32_t)pi : (r.w & ((1u << RB_CAND_INDEX_BITS) - 1u));
      any_selected = true;
      p_sel = ph;
      sel_wants = (flags & RB_CAND_WANTS) != 0;
    }
  };
  // The kernel is a chain of dependent loads (record -> traced byte -> next record ...) at DRAM latency; the first
  // kPre records and bytes are fetched up front, independently of each other (slots >= M hold stale but allocated
  // data that is never used: the stream half allocated k + 1 slots per pixel).
  constexpr int kPre = 6;
  U4 pre[kPre];
  uint32_t pocc[kPre];
  pre[0] = r0;
  pocc[0] = wv.occ[pi];
#pragma unroll
  for (int i = 1; i < kPre; ++i) {
    const bool have = i <= k;
    pre[i] = have ? wv.cand[(size_t)i * wv.npix + pi] : U4{0u, 0u, 0u, 0u};
    pocc[i] = have ? (uint32_t)wv.occ[(size_t)i * wv.npix + pi] : 0u;
  }
#pragma unroll
  for (int i = 0; i < kPre; ++i)
    if (i < M) replay(i, pre[i], pocc[i]);
  for (int i = kPre; i < M; ++i) replay(i, wv.cand[(size_t)i * wv.npix + pi], wv.occ[(size_t)i * wv.npix + pi]);
  if (any_selected && sel_wants) cnt.anyW++;
  Reservoir out = empty_reservoir();
  if (any_selected) {
    const F4 a = ld4(fc.Rread.point_wsum + sel_pixel), b = ld4(fc.Rread.normal_W + sel_pixel), c = ld4(fc.Rread.Li_conf + sel_pixel);
    out.bestSample.samplePoint = xyz(a), out.bestSample.sampleNormal = xyz(b), out.bestSample.L_i = xyz(c);
    out.bestSample.lightIdx = fc.Rread.light_idx[sel_pixel];
  }
  const float final_p_hat = any_selected ? p_sel : 0.0f;
  out.w_sum = w_sum;
  out.W = final_p_hat > 0.0f ? w_sum / final_p_hat : 0.0f;
  out.confidence = imin(confidence, P.confidenceCap);
  store_reservoir(fc.Rwrite, pi, out);
  if (wv.fuse_shade) shade_store(fc, pi, out, cnt);
}

// =====================================================================================
// Pass 5: final shading (P/simpleguidx11.cpp:452-472) + Integrator::sanitize (P/Integrator.cpp:6-23).
// The reservoir read here is this pixel's own output of the last pass: W > 0 means its sample was found
// unoccluded there, W == 0 zeroes the pixel — the reference's shadow ray never changes the result.
// =====================================================================================
template <class Vis>
RB_HD void shade_pixel(const FrameCtx& fc, int x, int y, const Vis&, Cnt& cnt) {
  const size_t pi = (size_t)y * fc.width + x;
  shade_store(fc, pi, load_reservoir(fc.Rread, pi), cnt);
}

// =====================================================================================
// Ground truth next to the path: one-sample MIS direct lighting, the estimator the reference's own
// reference images were made with — NEEPathIntegrator::integrateImpl2 with calcDI, without calcGI
// (P/NEEPathIntegrator.cpp:76-131) around DirectMISIntegrator::calculateDirectLighting
// (P/DirectMISIntegrator.cpp:18-144): a BRDF sample (Material::evaluateLightingGI, closest-hit ray, power
// heuristic against the area pdf) plus a light sample (TriangleCDF / alias pick, Sampling::sampleTriangle, shadow ray,
// power heuristic against Material::getPdfForSample). The pixel's primary hit is read from the G-buffer (same
// Intersection::intersectEmbree result); the material virtuals take the PRIMARY RAY direction (ray.getDir()), not the
// camera-to-hit direction the ReSTIR statics use, so omega_r and 1/I_M are formed from it here.
// Draw order = slot map: 0 lobe select (Phong only), 1-2 BRDF direction, 4 light pick, 5-6 point on the triangle.
// Dielectric materials refract in evaluateLightingGI and are not covered (rb_render_mis_frame refuses such scenes).
// =====================================================================================
RB_HD float power_heuristic(float pdf, float pdfOther) {  // DirectMISIntegrator::powerHeuristic, :10-15
  const float pdf_sqr = pdf * pdf;
  const float pdfOther_sqr = pdfOther * pdfOther;
  return pdf_sqr / (pdfOther_sqr + pdf_sqr);
}
template <class Vis>
RB_HD void mis_direct_pixel(const FrameCtx& fc, int x, int y, const Vis& vis, Cnt& cnt) {
  const size_t pi = (size_t)y * fc.width + x;
  const RbParams& P = fc.P;
  float* o = fc.frame + 3 * pi;
  const GElem g = load_gelem(fc.G, pi);
  const bool miss = fc.G.hit_ids[pi].x == 0xFFFFFFFFu;
  if (miss || emissive3(g.emission)) {  // background colour (:130) / Material::isEmitter at the camera vertex (:93-97)
    const V3 e = xyz(ld4(fc.G.emission + pi));
    o[0] = e.x, o[1] = e.y, o[2] = e.z;
    return;
  }
  V3 rdir;
  primary_ray(fc.cam, fc.width, fc.height, x, y, &rdir);
  const uint32_t key = rng_pixel_key(fc.frame_key, (uint32_t)pi);
  const bool lambert = g.matType == RB_MAT_LAMBERT;
  const V3 n = g.normal;
  const V3 omega_r = normalize(reflect(rdir, n));
  const float maxDiffuse = max_component(g.diffuse), maxSpecular = max_component(g.specular);
  const float pdfFactor = maxDiffuse / (maxDiffuse + maxSpecular);
  // 1 / calc_I_M(dot(-ray.getDir(), n), shininess): MaterialPhong::evaluateLightingGI :49-50 and ::evaluateBRDF :84-85
  // spell the same expression; it is evaluated at most once here (and only where one of them would)
  float i_m = 0.0f;
  bool have_im = false;
  V3 L_direct = v3(0);

  if (fc.mis_flags & RB_MIS_SAMPLE_BRDF) {  // DirectMISIntegrator::evaluateBRDFSample, :92-144
    const float r1 = rng_value(key, 1, 0, 1), r2 = rng_value(key, 2, 0, 1);
    V3 wi, f_r;
    float pdf;
    if (lambert) {  // MaterialLambert::evaluateLightingGI, P/MaterialLambert.cpp:10-18
      wi = cosw_sample(n, r1, r2);
      pdf = cosw_pdf(n, wi);
      f_r = v3(g.diffuse.x / RB_PI, g.diffuse.y / RB_PI, g.diffuse.z / RB_PI);
    } else {  // MaterialPhong::evaluateLightingGI, P/MaterialPhong.cpp:18-67
      const float r0 = rng_value(key, 0, 0.0f, maxDiffuse + maxSpecular);
      if (r0 < maxDiffuse) {
        wi = cosw_sample(n, r1, r2);
        f_r = g.diffuse * RB_ONE_OVER_PI;
      } else {
        wi = lobe_sample(omega_r, g.shininess, r1, r2);
        i_m = frcp_(calc_I_M(dot(-rdir, n), g.shininess));
        have_im = true;
        f_r = g.specular * i_m * dm::powf_(gmax(dot(wi, omega_r), 0.0f), g.shininess);
      }
      const float pdfDiffuse = cosw_pdf(n, wi) * pdfFactor;
      const float pdfSpecular = ((g.shininess + 1.0f) * RB_ONE_OVER_TWO_PI * dm::powf_(gmax(0.0f, dot(wi, omega_r)), g.shininess)) *
                                (1.0f - pdfFactor);
      pdf = pdfDiffuse + pdfSpecular;
      if (dot(n, wi) < 0) f_r = v3(0);
    }
    const V3 org = g.pos + P.normalOffset * n;
    cnt.closest++;
    const SurfaceHit h = vis.closest(0, org, wi, FLT_MIN + P.tnearOffset, FLT_MAX);
    if (h.didHit) {
      const V3 Le = xyz(ldg4(fc.sc.mat + 3 * (size_t)h.material + 2));
      if (Le.x + Le.y + Le.z > 0) {  // Material::isEmissive, P/material.h:135-137
        V3 lightDir = h.hitPoint - g.pos;
        const float r_sqr = dot(lightDir, lightDir);
        lightDir = normalize(lightDir);
        const float cosThetaI = gmax(dot(lightDir, n), 0.0f);
        const float cosThetaY = gmax(dot(-lightDir, h.normal), 0.0f);
        const float areaMeasureFactor = cosThetaY / r_sqr;
        // TriangleCDF::getPDFForTriangle(*tris[hitTriId]), P/TriangleCDF.h:25-31 (hitTriId is 0 for a triangle that is
        // not in the emissive list, P/Intersection.h:103-106)
        const F4 l0 = ldg4(fc.sc.light + 6 * (size_t)(h.emissiveId >= 0 ? h.emissiveId : 0));
        float pdfAsIfLight = l0.w / fc.sc.total_area;
        pdfAsIfLight *= frcp_(l0.w);
        const float brdfPdfAreaMeasure = pdf * areaMeasureFactor;
        const float misWeight = power_heuristic(brdfPdfAreaMeasure, pdfAsIfLight);
        const V3 t = (Le * misWeight) * f_r * cosThetaI;
        L_direct = L_direct + v3(t.x / pdf, t.y / pdf, t.z / pdf);
      }
    }
  }
  if ((fc.mis_flags & RB_MIS_SAMPLE_LIGHTS) && fc.sc.n_lights > 0) {  // ::evaluateLightSample, :38-90
    const LightPick pick = pick_light(fc.sc, P.lightSampler, key, 4u);
    const F4* Lp = fc.sc.light + 6 * (size_t)pick.idx;
    const F4 l0 = ldg4(Lp), l1 = ldg4(Lp + 1), l2 = ldg4(Lp + 2), l3 = ldg4(Lp + 3), l4 = ldg4(Lp + 4), l5 = ldg4(Lp + 5);
    const float pick_pdf = P.lightSampler == RB_LS_ALIAS ? l1.w : pick.pdf;
    const float r1 = rng_value(key, 5, 0, 1), r2 = rng_value(key, 6, 0, 1);
    const float sq = sqrtf_(r1);
    const float bx = 1.0f - sq, by = sq * (1.0f - r2), bz = sq * r2;
    const V3 samplePoint = xyz(l0) * bx + xyz(l1) * by + xyz(l2) * bz;
    const V3 triNormal = normalize(xyz(l3) * bx + xyz(l4) * by + xyz(l5) * bz);
    const float lightPdf = pick_pdf * l2.w;
    V3 lightDir = samplePoint - g.pos;
    const float r_sqr = dot(lightDir, lightDir);
    lightDir = normalize(lightDir);
    if (lightPdf != 0.0f && r_sqr != 0.0f) {
      const float cosThetaI = gmax(dot(lightDir, n), 0.0f);
      const float cosThetaY = gmax(dot(-lightDir, triNormal), 0.0f);
      const float areaMeasureFactor = cosThetaY / r_sqr;
      bool lit = cosThetaI > 0 && cosThetaY > 0;
      if (lit) {
        cnt.anyW++;
        cnt.anyT++;
        lit = vis.visible(1, g.pos, samplePoint);
      }
      if (lit) {
        float pdfAsIfBrdf = cosw_pdf(n, lightDir);  // Material::getPdfForSample
        float lobe = 0.0f;
        if (!lambert) {
          lobe = dm::powf_(gmax(0.0f, dot(lightDir, omega_r)), g.shininess);
          pdfAsIfBrdf = pdfAsIfBrdf * pdfFactor;
          pdfAsIfBrdf += ((g.shininess + 1.0f) * RB_ONE_OVER_TWO_PI * lobe) * (1.0f - pdfFactor);
        }
        const float pdfAsIfBrdfAreaMeasure = pdfAsIfBrdf * areaMeasureFactor;
        const V3 L_i = v3(l3.w, l4.w, l5.w);
        const float misWeight = power_heuristic(lightPdf, pdfAsIfBrdfAreaMeasure);
        if (misWeight > 0.0f) {
          const float G = cosThetaI * cosThetaY / r_sqr;
          V3 f_r = g.diffuse * RB_ONE_OVER_PI;  // Material::evaluateBRDF
          if (!lambert) {
            if (!have_im) i_m = frcp_(calc_I_M(dot(-rdir, n), g.shininess));
            // pow(max(d, 0), n) here vs pow(max(0, d), n) in the pdf: the same value unless d is NaN
            const float d = dot(lightDir, omega_r);
            const float lobe_brdf = (d != d) ? dm::powf_(gmax(d, 0.0f), g.shininess) : lobe;
            f_r = f_r + g.specular * i_m * lobe_brdf;
          }
          const V3 t = (L_i * misWeight) * f_r * G;
          L_direct = L_direct + v3(t.x / lightPdf, t.y / lightPdf, t.z / lightPdf);
        }
      }
    }
  }
  // Integrator::sanitize (P/Integrator.cpp:6-23), then L_i_indirect (= 0) + L_i_direct (:130)
  if (L_direct.x != L_direct.x || L_direct.y != L_direct.y || L_direct.z != L_direct.z) L_direct = v3(0);
  if (L_direct.x < 0 || L_direct.y < 0 || L_direct.z < 0) L_direct = v3(0);
  const V3 px = v3(0) + L_direct;
  o[0] = px.x, o[1] = px.y, o[2] = px.z;
}

// =====================================================================================
// After the path: accumulate, tonemap, gamma (P/simpleguidx11.cpp:246-253, 262-295;
// Utils::aces / Utils::compress, P/utils.cpp:190-197, 220-230)
// =====================================================================================
RB_HD float aces1(float x) {  // one channel of Utils::aces (float constants, glm::clamp)
  const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
  return gclamp((x * (a * x + b)) / (x * (c * x + d) + e), 0.0f, 1.0f);
}
RB_HD float compress1(float u) {  // Utils::compress: linear -> sRGB; the 0.0031308 comparison is in double
  if (u <= 0.0f) return 0.0f;
  if (u >= 1.0f) return 1.0f;
  if ((double)u <= 0.0031308) return u * 12.92f;
  return 1.055f * dm::powf_(u, 1.0f / 2.4f) - 0.055f;
}
// returns the per-pixel channel mean of the new accumulator value (for the statistics)
RB_HD float accumulate_display_pixel(const float* frame, float* accumulator, F4* display, size_t pi, float mix_a, bool tonemap,
                                     bool gamma) {
  const V3 f = v3(frame[3 * pi], frame[3 * pi + 1], frame[3 * pi + 2]);
  const V3 acc0 = v3(accumulator[3 * pi], accumulator[3 * pi + 1], accumulator[3 * pi + 2]);
  const V3 acc = acc0 * (1.0f - mix_a) + f * mix_a;  // glm::mix(x, y, a) = x * (1 - a) + y * a
  accumulator[3 * pi] = acc.x, accumulator[3 * pi + 1] = acc.y, accumulator[3 * pi + 2] = acc.z;
  V3 pix = acc;
  if (tonemap) pix = v3(aces1(pix.x), aces1(pix.y), aces1(pix.z));
  if (gamma) pix = v3(compress1(pix.x), compress1(pix.y), compress1(pix.z));
  st4(display + pi, f4(pix, 1.0f));
  return (acc.x + acc.y + acc.z) / 3.0f;
}

}  // namespace rb
#endif
