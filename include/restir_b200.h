/*
 * restir_b200.h — C ABI of the B200-native ReSTIR DI hot path.
 *
 * Drop-in boundary for the per-pixel frame loop of Tonz24/restir-embree.
 * Path prefix P/ = template/src/pg/pg1_embree/ in the reference checkout.
 *
 *   Frame seam : replaces the body of SimpleGuiDX11::produceRestir(float t)
 *                (P/simpleguidx11.cpp:359-487, declared P/simpleguidx11.h:87).
 *   Ray seam   : replaces Intersection::intersectEmbree / testOcclusion
 *                (P/Intersection.h:8-41, 43-60), i.e. Embree's rtcIntersect1 /
 *                rtcOccluded1 (embree3/rtcore_scene.h:99,123), in the batched
 *                stream form of rtcIntersect1M / rtcOccluded1M (:111,135).
 *
 * Plain C, POD only: no C++ types, no torch types.  All functions return
 * RB_OK (0) or a negative RbStatus; none throws across the boundary.
 * rb_last_error() gives the text the reference would have thrown from its
 * Embree error callback (P/tutorials.cpp:6-24).
 *
 * Threading: one caller thread per handle (the reference's Producer thread,
 * P/simpleguidx11.cpp:223,499); not re-entrant; one frame in flight.
 * Parameters and camera are snapshotted at rb_render_frame entry.
 */
#ifndef RESTIR_B200_H_
#define RESTIR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_ABI_VERSION 2

typedef enum RbStatus {
  RB_OK = 0,
  RB_ERR_INVALID_ARGUMENT = -1,
  RB_ERR_CUDA = -2,
  RB_ERR_NO_SCENE = -3,
  RB_ERR_UNSUPPORTED = -4,
  RB_ERR_OUT_OF_MEMORY = -5,
  RB_ERR_COMM = -6
} RbStatus;

/* MaterialType, P/enums.h:3-11 (same numeric values). */
typedef enum RbMaterialType {
  RB_MAT_NORMAL = 0,
  RB_MAT_LAMBERT = 1,
  RB_MAT_PHONG = 2,
  RB_MAT_MIRROR = 3,
  RB_MAT_DIELECTRIC = 4,
  RB_MAT_DIELECTRIC_TRANSPARENT = 5,
  RB_MAT_UNSUPPORTED = 6
} RbMaterialType;

/* ReSTIRIntegrator::SpatialWeightCalculation, P/ReSTIRIntegrator.h:19-25
 * (same order, same numeric values). */
typedef enum RbSpatialWeightCalc {
  RB_SW_CONSTANT = 0,
  RB_SW_CONSTANT_DEBIAS_CONTRIB = 1,
  RB_SW_CONSTANT_DEBIAS_Z_TERM = 2,
  RB_SW_BALANCE_HEURISTIC = 3,
  RB_SW_PAIRWISE_MIS = 4
} RbSpatialWeightCalc;

/* Light-sampler seam: the reference inverts a float CDF with
 * std::lower_bound (P/TriangleCDF.cpp:36-54); north_star asks for an alias
 * table.  Both are implemented; both pick emissive triangles ∝ area. */
typedef enum RbLightSampler {
  RB_LS_CDF = 0,  /* P/TriangleCDF.cpp semantics, pdf = cdf[i]-cdf[i-1]      */
  RB_LS_ALIAS = 1 /* Vose alias table, single-draw, pdf = area_i/total_area   */
} RbLightSampler;

/* Material constants read by the path: Material::{diffuse,specular,emission,
 * shininess_,ior}, P/material.h:104-115; type from getType(). */
typedef struct RbMaterial {
  uint32_t type; /* RbMaterialType */
  float diffuse[3];
  float specular[3];
  float emission[3];
  float shininess;
  float ior;
} RbMaterial;

/* One surface == one Embree geometry (geomID = index, attach order),
 * exactly what ModelLoader::loadScene hands Embree, P/ModelLoader.cpp:227-318:
 * non-indexed triangle soup, 3*n_tris vertices, identity index buffer (:297-299).
 * uv may be NULL (treated as zero; read by textured materials, rb_set_textures); tangent may be NULL (attribute slot 3 of
 * the reference's geometry, P/ModelLoader.cpp:286-287: kept at upload and read by materials with a normal map, which
 * rb_set_textures refuses for a scene uploaded without tangents). */
typedef struct RbSurface {
  uint32_t n_tris;
  uint32_t material;    /* index into RbSceneDesc.materials */
  const float* pos;     /* [3*n_tris][3] */
  const float* normal;  /* [3*n_tris][3] */
  const float* uv;      /* [3*n_tris][2] or NULL */
  const float* tangent; /* [3*n_tris][3] or NULL */
} RbSurface;

typedef struct RbSceneDesc {
  uint32_t n_surfaces;
  const RbSurface* surfaces;
  uint32_t n_materials;
  const RbMaterial* materials;
} RbSceneDesc;

/* 1:1 with the ReSTIRIntegrator statics (P/ReSTIRIntegrator.cpp:13-35,
 * P/ReSTIRIntegrator.h:91-113) and the RenderParams fields the path reads
 * (P/RenderParams.h:8-17).  rb_default_params() fills the reference defaults. */
typedef struct RbParams {
  int32_t M_Area;                    /* 1  */
  int32_t M_Brdf;                    /* 1  */
  int32_t spatialReuseNeighborCount; /* 5  */
  int32_t spatialPassCount;          /* 1  */
  int32_t confidenceCap;             /* 20 */
  float spatialReuseRadius;          /* 30 (offsets are < sqrt(radius) px, P/Sampling.cpp:78-87; halo reach = floor(sqrt(radius)) + 1 rows) */
  float minNormalSimilarity;         /* 0.85 */
  float maxDepthDifference;          /* 0.2  */
  int32_t doSpatialReuse;            /* 0 */
  int32_t doTemporalReuse;           /* 0 */
  int32_t doVisibilityPass;          /* 0 */
  int32_t rejectDissimilarNeighbors; /* 0 */
  int32_t spatialWeightCalc;         /* RbSpatialWeightCalc, CONSTANT */
  float tnearOffset;                 /* 0.01  */
  float tfarOffset;                  /* 0.001 */
  float normalOffset;                /* 0.001 */
  float bgColor[3];                  /* 0.5   */
  int32_t useSkybox;                 /* reference default 1; here 0 until a sky texture was given (rb_set_sky) */
  int32_t lightSampler;              /* RbLightSampler; reference = CDF */
  int32_t wavefront;                 /* 0: every pass traces its rays inline; 1: stream->trace->resolve split */
  int32_t temporalFetchReprojected;  /* 0 = the reference: temporal reuse merges last frame's reservoir of the SAME pixel
                                        (getReservoirLastFrame(pixelCoords), P/ReSTIRIntegrator.cpp:641), although it
                                        validates the reprojected one; 1 = the repaired variant: the reservoir of the
                                        backward-reprojected pixel. Single band only (bands would need last-frame
                                        reservoir halos): rb_render_frame refuses 1 on a banded handle. ABI version 2. */
} RbParams;

/* What produceRestir copies out of Camera into GBuffer each frame
 * (P/simpleguidx11.cpp:362-365): position, view matrix, inverse view matrix
 * (column-major like glm::mat4), focal length in pixels (P/camera.cpp:81-84). */
typedef struct RbCamera {
  float pos[3];
  float focal_px;
  float viewMat[16];
  float invViewMat[16];
} RbCamera;

/* The reference's per-pass millisecond timers (P/simpleguidx11.h:120-127),
 * here from CUDA events, plus ray counters for Mrays/s. */
typedef struct RbTimings {
  float ms_gbuffer;
  float ms_initial;
  float ms_visibility;
  float ms_temporal;
  float ms_spatial;
  float ms_shade;
  float ms_total;
  float ms_trace_any;           /* sum of the traversal kernels of the frame (wavefront mode) */
  /* wavefront mode, per pass (0 gbuffer, 1 initial, 2 visibility, 3 temporal, 4 spatial, 5 shade):
   * time in the streaming kernels (stream + resolve halves) and in the traversal kernel */
  float ms_stream[6];
  float ms_trace[6];
  uint64_t rays_closest;        /* closest-hit rays traced this frame  */
  uint64_t rays_any_as_written; /* shadow rays the reference would issue */
  uint64_t rays_any_traced;     /* after exact-duplicate / zero-contribution elimination */
  uint32_t kernel_launches;
  float ms_halo;                /* multi-GPU: duration of the frame's last reservoir-halo exchange on the comm stream */
} RbTimings;

typedef struct RbCreateInfo {
  int32_t width;   /* full image, SimpleGuiDX11::width_  */
  int32_t height;  /* full image, SimpleGuiDX11::height_ */
  int32_t device;  /* CUDA ordinal */
  uint32_t seed;   /* counter-RNG seed (reference: mt19937{123}, P/utils.cpp:175) */
  int32_t band_y0; /* first image row rendered by this handle (multi-GPU bands); 0 */
  int32_t band_y1; /* one past the last row; height for a single GPU */
  int32_t collect_timings; /* record CUDA events per pass (adds syncs at readout only) */
  int32_t reserved;
} RbCreateInfo;

/* Device buffers readable for parity checks (rb_readback). */
typedef enum RbBufferId {
  RB_BUF_GBUF_POS_DEPTH = 0,   /* float4 {pos.xyz, depth}            per px */
  RB_BUF_GBUF_NORMAL_SHIN = 1, /* float4 {normal.xyz, shininess}     per px */
  RB_BUF_GBUF_DIFFUSE_IIM = 2, /* float4 {diffuse.rgb, 1/I_M}        per px */
  RB_BUF_GBUF_SPEC_TYPE = 3,   /* float4 {specular.rgb, bits(matType|emissive<<8)} */
  RB_BUF_GBUF_EMISSION = 4,    /* float4 {emission.rgb, 0}           per px */
  RB_BUF_HIT_IDS = 5,          /* uint2  {geomID, primID} of the primary hit (~0u = miss) */
  RB_BUF_RES_POINT_WSUM = 6,   /* float4 {samplePoint.xyz, w_sum}   final reservoir of the frame */
  RB_BUF_RES_NORMAL_W = 7,     /* float4 {sampleNormal.xyz, W}      */
  RB_BUF_RES_LI_CONF = 8,      /* float4 {L_i.rgb, bits(confidence)} */
  RB_BUF_RES_LIGHT_IDX = 9,    /* int32  emissive-triangle id of bestSample (-1 none) */
  RB_BUF_FRAME_RGB = 10,       /* float3 frame_data (linear HDR)    */
  RB_BUF_ALIAS_PROB = 11,      /* float  [n_emissive]                */
  RB_BUF_ALIAS_IDX = 12,       /* uint32 [n_emissive]                */
  RB_BUF_LIGHT_CDF = 13,       /* float  [n_emissive]                */
  RB_BUF_ACCUMULATOR = 14,     /* float3 running mean of frame_data (rb_accumulate_display) */
  RB_BUF_DISPLAY = 15          /* float4 tonemapped, gamma-compressed display_data          */
} RbBufferId;

/* Ray seam records.  RbRay has the 48-byte layout of RTCRay
 * (embree3/rtcore_ray.h:11-27) so a host can pass its Embree rays unchanged. */
typedef struct RbRay {
  float org_x, org_y, org_z, tnear;
  float dir_x, dir_y, dir_z, time;
  float tfar;
  uint32_t mask, id, flags;
} RbRay;

/* Subset of RTCHit (rtcore_ray.h:30-42) the path consumes. */
typedef struct RbHit {
  float t, u, v;
  uint32_t primID; /* triangle index within its surface; 0xFFFFFFFF = miss */
  uint32_t geomID; /* surface index (attach order);      0xFFFFFFFF = miss */
} RbHit;

typedef struct RbContext* RbHandle;

uint32_t rb_abi_version(void);
const char* rb_last_error(RbHandle h); /* h may be NULL: error of the last failed rb_create */
void rb_default_params(RbParams* out);

/* replaces SimpleGuiDX11::Init buffer allocation, P/simpleguidx11.cpp:113-119 */
int rb_create(const RbCreateInfo* info, RbHandle* out);
void rb_destroy(RbHandle h);

/* replaces ModelLoader::loadScene + TriangleCDF ctor + rtcCommitScene
 * (P/ModelLoader.cpp:218-321, P/TriangleCDF.cpp:8-34, P/Scene.cpp:8-16):
 * copies the scene to the GPU, builds the light tables and the wide BVH there. */
int rb_upload_scene(RbHandle h, const RbSceneDesc* scene);

/* replaces the ImGui edits of the statics, P/ReSTIRIntegrator.cpp:37-87 */
int rb_set_params(RbHandle h, const RbParams* params);

/* replaces SimpleGuiDX11::produceRestir, P/simpleguidx11.cpp:359-487.
 * frame_idx plays the role of frameCtr (temporal reuse only when > 0, :408).
 * frame_rgb_out: w*h*3 floats (band rows only are written), host pointer or
 * NULL (keep the frame on the device; fetch later with rb_readback).
 * timings may be NULL. */
int rb_render_frame(RbHandle h, const RbCamera* cam, uint32_t frame_idx,
                    float* frame_rgb_out, RbTimings* timings);

/* Same, output left in / written to DEVICE memory (frame_rgb_dev may be NULL). */
int rb_render_frame_device(RbHandle h, const RbCamera* cam, uint32_t frame_idx,
                           float* frame_rgb_dev, RbTimings* timings);

/* Pipelined form for a Producer loop that double-buffers frame_data (P/simpleguidx11.cpp:240-253 consumes the frame
 * right after produceRestir; with two buffers it can consume frame n-1 while frame n renders): the frame is issued and
 * its rows are copied into frame_rgb_out (HOST memory; page-locked for a truly asynchronous copy) on a copy stream
 * behind the frame's last kernel; the call returns without waiting, so the next frame's kernels overlap the copy.
 * rb_frame_wait(h, k) blocks until at most k of the frames issued this way are still in flight (k = 0: every buffer
 * is complete; k = 1: the buffer of the frame before the latest is complete). The caller must not reuse a buffer
 * before its frame has been waited for. ABI version 2. */
int rb_render_frame_async(RbHandle h, const RbCamera* cam, uint32_t frame_idx, float* frame_rgb_out);
int rb_frame_wait(RbHandle h, uint32_t frames_in_flight);

/* ---- Ground truth next to the path ---------------------------------------------------------
 * One frame of the reference's one-sample MIS direct-lighting estimator — NEEPathIntegrator with "Calculate DI" on and
 * "Calculate GI" off around DirectMISIntegrator (P/NEEPathIntegrator.cpp:76-131, P/DirectMISIntegrator.cpp:18-144), the
 * integrator behind the author's reference images (S/mis_reference*.png.txt): primary ray (the G-buffer kernel), then a
 * BRDF sample (Material::evaluateLightingGI + closest-hit ray) and a light sample (area pick + shadow ray), combined
 * with the power heuristic; Integrator::sanitize on the result. Unbiased, so its running mean (rb_accumulate_display)
 * is what ReSTIR's bias is measured against. `techniques`: bit 0 = "Sample BRDF", bit 1 = "Sample Light Sources"
 * (P/DirectMISIntegrator.cpp:32-36; the reference default is both = 3). Uses RbParams' offsets, bgColor and
 * lightSampler; does not touch the ReSTIR state (reservoirs, previous G-buffer), writes frame_data (band rows) and,
 * when frame_rgb_out (host, w*h*3) is given, copies the band rows out and waits. Counter RNG pass id 4, slots:
 * 0 lobe select, 1-2 BRDF direction, 4 light pick, 5-6 point on the emitter. Scenes with RB_MAT_DIELECTRIC
 * materials are refused (RB_ERR_UNSUPPORTED): their evaluateLightingGI refracts. */
#define RB_MIS_SAMPLE_BRDF 1u
#define RB_MIS_SAMPLE_LIGHTS 2u
int rb_render_mis_frame(RbHandle h, const RbCamera* cam, uint32_t frame_idx, uint32_t techniques, float* frame_rgb_out);

int rb_readback(RbHandle h, int buffer_id /*RbBufferId*/, void* dst, size_t bytes);
int rb_synchronize(RbHandle h);

/* Device-side stopwatch on the handle's stream (CUDA events): the per-frame timers of the
 * reference (P/simpleguidx11.h:120-127) cover one frame; these bracket any number of frames
 * without a host synchronisation in between. rb_timer_end waits for the stream. */
int rb_timer_begin(RbHandle h);
int rb_timer_end(RbHandle h, float* ms_out);

/* replaces Intersection::intersectEmbree's rtcIntersect1 (P/Intersection.h:63-83)
 * and Intersection::testOcclusion's rtcOccluded1 (:43-60); host pointers. */
int rb_trace_closest(RbHandle h, const RbRay* rays, RbHit* hits, uint32_t n);
int rb_trace_occluded(RbHandle h, const RbRay* rays, uint8_t* occluded, uint32_t n);
/* device-pointer variants (no copies; timed with CUDA events → *ms_out) */
int rb_trace_closest_device(RbHandle h, const RbRay* rays_dev, RbHit* hits_dev, uint32_t n, float* ms_out);
int rb_trace_occluded_device(RbHandle h, const RbRay* rays_dev, uint8_t* occ_dev, uint32_t n, float* ms_out);

/* Scene statistics after upload. */
typedef struct RbSceneStats {
  uint32_t n_triangles;
  uint32_t n_emissive;
  uint32_t n_bvh_nodes;
  uint32_t bvh_depth;
  float build_ms;
  float total_emissive_area;
  float bounds_lo[3];
  float bounds_hi[3];
} RbSceneStats;
int rb_scene_stats(RbHandle h, RbSceneStats* out);

/* ---- Textured materials (second half) -------------------------------------------------------------
 * Material::getDiffuseColor / getSpecularColor / getShininess (P/material.cpp:105-134) read a texture where the material
 * has one (Material::set_texture, slots kDiffuseMapSlot / kSpecularMapSlot / kShininessMapSlot); the G-buffer pass calls
 * them with the hit's interpolated uv (P/ReSTIRIntegrator.cpp:225-228). rb_set_textures hands the library the texel
 * arrays the host's image loader produced — exactly Texture's members (P/Texture.h:44-48): width, height, bytes per row,
 * bytes per pixel (3 or 4: 8-bit B,G,R[,A] as FreeImage delivers them, value / 255; 12 or 16: float R,G,B[,A]) and the
 * data, row 0 first — and, per material, a texture index per slot (-1 = the material constant). Sampling is
 * Texture::get_texel(uv) as ModelLoader::TextureProxy configures it (BILINEAR, REPEAT; P/Texture.cpp:72-107,170-194):
 * pixel = (u * w, (1 - v) * h), four get_texel(x, y) with abs(x % w), glm::mix in x then y. The shininess map is a
 * roughness map: n = 2 / r^2 - 2 (P/material.cpp:124-131). Call after rb_upload_scene (whose surfaces must carry uv);
 * a new rb_upload_scene drops the textures. Data is copied.
 * Normal maps (Material::kNormalMapSlot): where the hit's material has one, EVERY closest-hit query of the path (G-buffer,
 * BRDF-sampled candidates, rb_render_mis_frame) replaces the interpolated normal — already flipped towards the ray — by
 * mat3(T, B, n) * (texel * 2 - 1) with T = normalize(tangent - dot(tangent, n) * n), B = normalize(cross(n, T)) and the
 * tangent interpolated from RbSurface.tangent, exactly Intersection::intersectEmbree (P/Intersection.h:25-39): no
 * re-normalisation, no second flip. A scene uploaded without tangents cannot take a normal map (invalid argument). */
typedef struct RbTexture {
  int32_t width, height;
  int32_t scan_width; /* bytes per row */
  int32_t pixel_size; /* bytes per texel: 3, 4 (8-bit BGR[A]) or 12, 16 (float RGB[A]) */
  const void* data;
} RbTexture;
typedef struct RbMaterialTextures {
  int32_t diffuse, specular, shininess, normal; /* index into the texture array or -1 */
} RbMaterialTextures;
int rb_set_textures(RbHandle h, const RbTexture* textures, uint32_t n_textures, const RbMaterialTextures* per_material,
                    uint32_t n_materials);

/* ---- Sky ---------------------------------------------------------------------------------
 * With RenderParams::useSkybox (the reference's default), a primary ray that misses the scene writes
 * scene.getSkybox().getTexel(ray.getDir()) into the G-buffer's emission (P/ReSTIRIntegrator.cpp:231) — which is what
 * the frame shows for that pixel (P/simpleguidx11.cpp:466-467) and what rb_render_mis_frame returns for it
 * (P/NEEPathIntegrator.cpp:131). The sky is a SphericalMap (P/SphericalMap.cpp:10-14): x = 0.5f + 0.5f * atan2f(d.y, d.x)
 * / pi, y = 1.0f - acos(d.z) / pi (float products, the constant 1 / pi and the sums in double), then
 * Texture::get_texel(uv) with the Texture constructor's defaults: BILINEAR, CLAMP_TO_EDGE. rb_set_sky replaces
 * Scene::setSkybox (P/Scene.cpp:47-50): it takes the texel array of the decoded image (normally float R,G,B from an .hdr /
 * .exr; the formats of RbTexture) and copies it; NULL removes the sky. The sky belongs to the handle and survives
 * rb_upload_scene. rb_set_params with useSkybox = 1 fails until a sky was set. atan2f / acos are evaluated by
 * det_math.h (correctly rounded in 4e6 random samples), so the texel coordinates can differ from a libm's in the last
 * bit in rare cases; everything else is the reference's arithmetic. */
int rb_set_sky(RbHandle h, const RbTexture* sky);

/* ---- Scene ingestion (first half) -------------------------------------------------------------
 * Where the reference constructs `Scene{file_name, device}` (P/raytracer.cpp:36, ModelLoader::loadScene /
 * loadOBJ / loadMaterials, P/ModelLoader.cpp:41-321) through ASSIMP, rb_obj_load parses a Wavefront OBJ + MTL itself
 * and ends in the RbSceneDesc rb_upload_scene takes (host-only, no GPU needed). Kept from the reference: material type
 * from the MTL key `Pc` (0..5 = P/enums.h MaterialType, else UNSUPPORTED), Kd/Ks expanded from sRGB when gamma_correct
 * (Raytracer::gammaCorrect, default on; Utils::expand), Ke/Ns/Ni as written, ASSIMP's OBJ defaults for absent keys,
 * materials in MTL order, one non-indexed surface per material in order of first use, faces without a material are an
 * error (the reference indexes materials with mMaterialIndex - 1). Not pinnable without an ASSIMP binary: polygon
 * triangulation (a fan here), mesh order of files that interleave materials, tangents (generated only when a material
 * names a normal map, by the per-face step of ASSIMP's CalcTangentSpace without its smoothing across faces). The MTL's
 * texture file names are available through rb_obj_texture_name, slots 0..3 = map_Kd, map_Ks, map_Ns, map_Kn/bump; the
 * host decodes them and calls rb_set_textures. Faces without normals get the flat face normal. err (optional) receives a message. */
typedef struct RbObjScene RbObjScene;
int rb_obj_load(const char* obj_path, int32_t gamma_correct, RbObjScene** out, char* err, size_t err_bytes);
const RbSceneDesc* rb_obj_scene_desc(const RbObjScene* s); /* valid until rb_obj_free */
const char* rb_obj_material_name(const RbObjScene* s, uint32_t material);
const char* rb_obj_texture_name(const RbObjScene* s, uint32_t material, int32_t slot);
void rb_obj_free(RbObjScene* s);

/* ---- Multi-GPU bands -----------------------------------------------------------------
 * One handle per GPU renders the image rows [band_y0, band_y1); scene and BVH are replicated. The only
 * cross-band data are the reservoir rows within the spatial-reuse reach of a band edge ("halo rows"), which are
 * exchanged before every spatial pass. G-buffer elements of other bands that temporal reprojection may touch
 * are re-derived locally (they are a pure function of camera, pixel and scene), so temporal reuse needs no
 * communication. The counter RNG is keyed on the global pixel index: the image is bit-identical for any
 * number of bands.
 *
 * rb_comm_init attaches an NCCL communicator over the handles of all ranks (rank r owns band r; bands ordered
 * top to bottom); afterwards rb_render_frame performs the halo exchange itself. nccl_unique_id is the 128-byte
 * ncclUniqueId made by rb_comm_unique_id on rank 0 and distributed by the launcher (torch.distributed, MPI,
 * a file). NCCL is loaded with dlopen at that point; single-GPU use needs no NCCL at all.
 *
 * Transport. One process per GPU on one node: rb_comm_init maps the neighbours' reservoir planes into this process
 * (CUDA IPC handles sent through the communicator), and a halo exchange is then ONE kernel that stores this band's
 * boundary rows straight into the neighbours' halo rows over NVLink/NVSwitch peer memory and releases a stamp in
 * their flag word; the interior rows of the spatial pass are streamed meanwhile, and the boundary rows start once a
 * one-thread kernel has seen both neighbours' stamps (rb_comm_transport() == 1). If the mapping fails on any rank
 * (or RB_HALO=nccl) all ranks use grouped ncclSend/ncclRecv on a side stream instead (== 2).
 *
 * Load balancing: image bands do not cost the same (ceiling vs. floor), and a halo exchange is a rendezvous, so the
 * slowest band sets the frame rate. With a communicator attached the library therefore moves the band boundaries:
 * at the end of every 8th frame (RB_BAL_PERIOD) each rank sends its neighbours the last-frame reservoirs of the 16 rows
 * next to each boundary and ALL ranks all-gather {cost, rows} (cost = GPU time between the ends of two consecutive back
 * halves minus the time spent waiting for neighbours, CUDA events, mean of up to four frames). At the start of the next
 * frame every rank computes the same equal-cost partition from the same gathered numbers and moves each boundary half of
 * the way towards its target, at most 15 rows. Rows that change owner find their last-frame reservoirs already there,
 * their previous G-buffer in the over-rendered margin, and their ACCUMULATOR rows (rb_accumulate_display) are sent along
 * when the move is applied; the image and the running mean stay bit-identical to the one-band ones
 * (tools/check_nccl_bands.py). Balancing is switched off for all ranks together when any band is thinner than 32 rows.
 * RB_BALANCE=0 turns it off; rb_get_band reports the rows currently owned (rb_render_frame writes exactly those rows of
 * frame_rgb_out). A halo wait that times out (RB_HALO_TIMEOUT_MS, default 5 s; 60 s for the first exchanges) fails
 * the frame at its next synchronisation point (rb_render_frame, rb_frame_wait, rb_synchronize, rb_readback) with
 * RB_ERR_COMM and is then forgotten: the next frame may succeed. A second rb_comm_init on a handle is refused. */
int rb_comm_unique_id(void* out_id, size_t id_bytes);
int rb_comm_init(RbHandle h, int32_t rank, int32_t nranks, const void* nccl_unique_id, size_t id_bytes);
int32_t rb_comm_transport(RbHandle h); /* 0 = no communicator, 1 = peer memory (CUDA IPC), 2 = NCCL send/recv */
/* The balancer's partition rule alone (host arithmetic, no GPU): pairs = {cost, rows} per rank, rows tiling [0, height);
 * bounds_out[n_ranks + 1] = the boundaries every rank derives from them for the next period. For tests and tuning. */
int rb_debug_balance_step(const float* pairs, int32_t n_ranks, int32_t height, int32_t* bounds_out);
/* The rays a frame REALLY traces, for traversal measurements on the path's own ray mix (BASELINE configs[3]; replaces
 * nothing in the reference — the counterpart is the stream of rtcOccluded1 calls under P/Intersection.h:43-60). Between
 * the phases of an open frame in wavefront mode: which = 0 the back half's current queue (after rb_frame_begin: the
 * temporal pass's shadow rays; after rb_frame_spatial(i): that pass's), which = 1 the visibility pass's rays. Copies up to
 * `capacity` rays in queue order into dst (RTCRay layout; dst may be NULL to ask for the count only). */
int rb_debug_ray_queue(RbHandle h, int32_t which, RbRay* dst, uint32_t capacity, uint32_t* count_out);

/* The same frame in phases, for hosts that move the halo rows themselves (another transport, or several
 * bands in one process):  rb_frame_begin  (G-buffer, initial candidates, visibility, temporal)
 *                         { exchange halos; rb_frame_spatial(i) }  for each spatial pass
 *                         rb_frame_end    (shade, buffer rotation; frame_rgb_out host pointer or NULL).
 * Between phases rb_halo_export / rb_halo_import copy `rows` image rows starting at global row y of the
 * reservoirs the next spatial pass will read (4 planes packed back to back: 16+16+16+4 bytes per pixel) to /
 * from host memory. rb_halo_rows gives the reach in rows for the current parameters. */
int rb_frame_begin(RbHandle h, const RbCamera* cam, uint32_t frame_idx);
/* Between frames a band may be moved (load balancing): the new rows must lie inside the rows whose G-buffer this
 * handle rendered last frame (its band +- 16 rows), and the caller must first have brought over the LAST frame's
 * reservoirs of the rows it gains (rb_halo_export / rb_halo_import outside a frame act on exactly those). With
 * rb_comm_init the library does all of this itself (see below). */
int rb_set_band(RbHandle h, int32_t band_y0, int32_t band_y1);
int rb_get_band(RbHandle h, int32_t* band_y0, int32_t* band_y1); /* the rows the NEXT frame will render */
int rb_frame_spatial(RbHandle h, int32_t pass_index);
int rb_frame_end(RbHandle h, float* frame_rgb_out, RbTimings* timings);
int32_t rb_halo_rows(RbHandle h);
size_t rb_halo_bytes(RbHandle h, int32_t rows);
int rb_halo_export(RbHandle h, int32_t y, int32_t rows, void* dst_host);
int rb_halo_import(RbHandle h, int32_t y, int32_t rows, const void* src_host);

/* ---- One handle, several GPUs, one host thread -----------------------------------------------------
 * The reference is ONE process whose single Producer thread calls produceRestir (P/simpleguidx11.cpp:240-241, thread
 * created :499). rb_multi_* is the same seam for a host that wants N GPUs without N processes and without handing an
 * ncclUniqueId around: rb_multi_create makes one band handle per listed device (bands of ceil(height / n) rows, top to
 * bottom; the same ordinal may be listed several times — several bands on one GPU, used by the tests), connects
 * neighbours with peer access (cudaDeviceEnablePeerAccess: the halo rows are stored straight into the neighbour's
 * planes by the same k_halo_push / k_halo_wait kernels as between processes; no NCCL, no IPC), and every call below
 * fans out from the calling thread. rb_multi_render_frame[_async] issues the frame on every device and copies every
 * band's rows into the ONE host buffer frame_rgb_out (w*h*3 floats): the caller holds the assembled frame, as the
 * Producer expects. The image is bit-identical to the single-handle image. Bands are static in this mode (the cost
 * balancer of rb_comm_init moves rows between processes through NCCL; an in-process equivalent is not built).
 * rb_multi_member gives the band handle of device i for queries the family does not wrap (rb_scene_stats,
 * rb_get_band, rb_trace_*). Errors: rb_multi_last_error (NULL: of the last failed rb_multi_create). ABI version 2. */
typedef struct RbMulti* RbMultiHandle;
int rb_multi_create(const RbCreateInfo* info /* device, band_y0/1 ignored */, const int32_t* device_ordinals, int32_t n_devices,
                    RbMultiHandle* out);
void rb_multi_destroy(RbMultiHandle m);
const char* rb_multi_last_error(RbMultiHandle m);
int32_t rb_multi_device_count(RbMultiHandle m);
RbHandle rb_multi_member(RbMultiHandle m, int32_t i);
int rb_multi_upload_scene(RbMultiHandle m, const RbSceneDesc* scene);
int rb_multi_set_params(RbMultiHandle m, const RbParams* params);
int rb_multi_set_textures(RbMultiHandle m, const RbTexture* textures, uint32_t n_textures, const RbMaterialTextures* per_material,
                          uint32_t n_materials);
int rb_multi_set_sky(RbMultiHandle m, const RbTexture* sky);
int rb_multi_render_frame(RbMultiHandle m, const RbCamera* cam, uint32_t frame_idx, float* frame_rgb_out);
int rb_multi_render_frame_async(RbMultiHandle m, const RbCamera* cam, uint32_t frame_idx, float* frame_rgb_out);
int rb_multi_frame_wait(RbMultiHandle m, uint32_t frames_in_flight);
int rb_multi_synchronize(RbMultiHandle m);
int rb_multi_readback(RbMultiHandle m, int buffer_id, void* dst, size_t bytes); /* bands assembled */

/* ---- The step after the path: SimpleGuiDX11::Producer's accumulate / tonemap / statistics loops,
 * P/simpleguidx11.cpp:246-253 (accumulator = glm::mix(accumulator, frame_data, 1 / (accFrameCtr + 1))), :262-295
 * (display_data = {compress(aces(accumulator)), 1}: Utils::aces P/utils.cpp:190-197 when RenderParams::tonemap,
 * Utils::compress :220-230 when Raytracer::gammaCorrect) and :308-326 (accumulatorMean / accumulatorVariance over
 * the per-pixel channel mean, in double). Runs on the frame the handle rendered last; everything stays on the
 * device unless display_rgba_out (host, w*h*4 floats, band rows written) is given. OIDN denoising is not part
 * of it (closed binary). With bands, the statistics are those of the band's rows: sums are returned so that a
 * host can combine ranks. */
typedef struct RbImageStats {
  double sum;        /* sum over pixels of (r+g+b)/3 of the accumulator */
  double sum_sq;     /* sum of its squares                              */
  double mean;       /* sum / pixels            (accumulatorMean)       */
  double variance;   /* sum_sq / pixels - mean^2 (accumulatorVariance)   */
  uint64_t pixels;
} RbImageStats;
int rb_accumulate_display(RbHandle h, uint32_t acc_frame_ctr, int32_t tonemap, int32_t gamma_correct,
                          float* display_rgba_out, RbImageStats* stats);
int rb_multi_accumulate_display(RbMultiHandle m, uint32_t acc_frame_ctr, int32_t tonemap, int32_t gamma_correct,
                                float* display_rgba_out, RbImageStats* stats); /* all bands; statistics of the whole image */

#ifdef __cplusplus
}
#endif
#endif /* RESTIR_B200_H_ */
