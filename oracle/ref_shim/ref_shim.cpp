// ref_shim.cpp — runs the reference's OWN ReSTIR code (P/ReSTIRIntegrator.cpp, MaterialPhong.cpp,
// MaterialLambert.cpp, Sampling.cpp, TriangleCDF.cpp, camera.cpp, Reservoir.h, ... compiled where they
// lie under /root/reference by ref_shim/Makefile) so that the oracle's restatement can be pinned against it.
// TEST INFRASTRUCTURE ONLY; built only where the read-only reference checkout exists.
//
// What is NOT the reference here, and why:
//   * Embree (closed Windows binary): the five entry points the path calls are provided below on top of
//     the oracle's tracer (liboracle.so: same ray/triangle arithmetic contract, smallest-t hit).
//   * SimpleGuiDX11 (Win32/D3D11/ImGui/OIDN): only its static ReSTIR buffers exist (stubs/simpleguidx11.h);
//     the pass schedule of produceRestir (P/simpleguidx11.cpp:359-487) is replayed below in the serial
//     order of the reference's _DEBUG build (no OpenMP), which is the only deterministic order it has.
//   * Scene construction (ASSIMP): Scene::Scene is defined here and fills the same members from an
//     RbSceneDesc in ModelLoader::loadScene's order (P/ModelLoader.cpp:218-321).
#include "stdafx.h"

#include "ReSTIRIntegrator.h"
#include "DirectMISIntegrator.h"
#include "Intersection.h"
#include "MaterialMirror.h"
#include "MaterialNormal.h"
#include "MaterialTransparent.h"
#include "Sampling.h"
#include "Scene.h"
#include "SphericalMap.h"
#include "camera.h"
#include "simpleguidx11.h"
#include "Utils.h"

#include "../../include/restir_b200.h"

// ---- oracle tracer (liboracle.so) ------------------------------------------------------------
extern "C" {
void* orc_create(int width, int height, uint32_t seed, int rng_mode, int math_mode, int tracer_mode, int cache_iim);
void orc_destroy(void* h);
int orc_upload_scene(void* h, const RbSceneDesc* sd);
int orc_trace_closest(void* h, const RbRay* rays, RbHit* hits, uint32_t n);
int orc_trace_occluded(void* h, const RbRay* rays, uint8_t* occ, uint32_t n);
}

// ---- SimpleGuiDX11 statics (P/simpleguidx11.cpp:20-31) ------------------------------------------
Reservoir* SimpleGuiDX11::reservoirsPingPong[2]{};
Reservoir* SimpleGuiDX11::reservoirsLastFrame{};
GBuffer SimpleGuiDX11::gBuffer{};
GBuffer SimpleGuiDX11::gBufferLastFrame{};
int SimpleGuiDX11::readReservoirBufferIndex{0};
int SimpleGuiDX11::writeReservoirBufferIndex{1};
int SimpleGuiDX11::width_{0};
int SimpleGuiDX11::height_{0};
glm::vec<2, int> SimpleGuiDX11::debugPixel{-7, -7};

// ---- Embree stand-in ------------------------------------------------------------------------------
struct FakeGeometry {
  unsigned geomID;
  void* userData;
  std::vector<glm::vec3> normals;  // attribute slot 0, 3 per triangle
  std::vector<glm::vec2> uvs;      // attribute slot 1 (empty: zero)
  std::vector<float> ids;          // attribute slot 2
  std::vector<glm::vec3> tangents; // attribute slot 3 (empty: zero)
};
struct FakeScene {
  void* tracer = nullptr;  // oracle handle
  std::vector<FakeGeometry*> geoms;
};
static FakeScene* g_scene = nullptr;

extern "C" {
void rtcIntersect1(RTCScene, struct RTCIntersectContext*, struct RTCRayHit* rh) {
  RbRay r{};
  r.org_x = rh->ray.org_x, r.org_y = rh->ray.org_y, r.org_z = rh->ray.org_z, r.tnear = rh->ray.tnear;
  r.dir_x = rh->ray.dir_x, r.dir_y = rh->ray.dir_y, r.dir_z = rh->ray.dir_z, r.tfar = rh->ray.tfar;
  RbHit h{};
  orc_trace_closest(g_scene->tracer, &r, &h, 1);
  if (h.geomID != 0xFFFFFFFFu) {
    rh->ray.tfar = h.t;
    rh->hit.u = h.u, rh->hit.v = h.v;
    rh->hit.geomID = h.geomID, rh->hit.primID = h.primID;
  }
}
void rtcOccluded1(RTCScene, struct RTCIntersectContext*, struct RTCRay* ray) {
  RbRay r{};
  r.org_x = ray->org_x, r.org_y = ray->org_y, r.org_z = ray->org_z, r.tnear = ray->tnear;
  r.dir_x = ray->dir_x, r.dir_y = ray->dir_y, r.dir_z = ray->dir_z, r.tfar = ray->tfar;
  uint8_t occ = 0;
  orc_trace_occluded(g_scene->tracer, &r, &occ, 1);
  if (occ) ray->tfar = -INFINITY;  // Embree's convention: tfar = -inf when occluded
}
RTCGeometry rtcGetGeometry(RTCScene, unsigned int geomID) { return (RTCGeometry)g_scene->geoms[geomID]; }
void* rtcGetGeometryUserData(RTCGeometry g) { return ((FakeGeometry*)g)->userData; }
void rtcInterpolate(const struct RTCInterpolateArguments* a) {
  const FakeGeometry* g = (const FakeGeometry*)a->geometry;
  const float u = a->u, v = a->v, w = 1.0f - u - v;
  if (a->bufferSlot == 0) {  // shading normal: w*n0 + u*n1 + v*n2 (parity contract; Embree's own order is not public API)
    const glm::vec3 n = g->normals[3 * a->primID] * w + g->normals[3 * a->primID + 1] * u + g->normals[3 * a->primID + 2] * v;
    a->P[0] = n.x, a->P[1] = n.y, a->P[2] = n.z;
  } else if (a->bufferSlot == 1 && !g->uvs.empty()) {  // texture coordinates, same interpolation contract
    const glm::vec2 t = g->uvs[3 * a->primID] * w + g->uvs[3 * a->primID + 1] * u + g->uvs[3 * a->primID + 2] * v;
    a->P[0] = t.x, a->P[1] = t.y;
  } else if (a->bufferSlot == 2) {
    a->P[0] = g->ids[3 * a->primID] * w + g->ids[3 * a->primID + 1] * u + g->ids[3 * a->primID + 2] * v;
  } else if (a->bufferSlot == 3 && !g->tangents.empty()) {  // tangent (read by normal-mapped materials), same contract
    const glm::vec3 t = g->tangents[3 * a->primID] * w + g->tangents[3 * a->primID + 1] * u + g->tangents[3 * a->primID + 2] * v;
    a->P[0] = t.x, a->P[1] = t.y, a->P[2] = t.z;
  } else {
    for (unsigned i = 0; i < a->valueCount; ++i) a->P[i] = 0.0f;  // uv, tangent: unused by untextured materials
  }
}
}

// ---- Scene (members filled like ModelLoader::loadScene would, from an RbSceneDesc) -----------------
static const RbSceneDesc* g_pending = nullptr;

Scene::Scene(const char*, const RTCDevice&) {
  const RbSceneDesc* sd = g_pending;
  FakeScene* fs = new FakeScene();
  std::vector<Triangle*> emissiveTris;
  uint32_t triIdCtr = 0;
  for (uint32_t m = 0; m < sd->n_materials; ++m) {
    const RbMaterial& M = sd->materials[m];
    Material* mat = nullptr;
    if (M.type == RB_MAT_LAMBERT)
      mat = new MaterialLambert();
    else if (M.type == RB_MAT_DIELECTRIC)
      mat = new MaterialDielectric();
    else if (M.type == RB_MAT_MIRROR)  // the ReSTIR statics dispatch on getType() only (P/ReSTIRIntegrator.h:32-59)
      mat = new MaterialMirror();
    else if (M.type == RB_MAT_NORMAL)
      mat = new MaterialNormal();
    else if (M.type == RB_MAT_DIELECTRIC_TRANSPARENT)
      mat = new MaterialTransparent();
    else
      mat = new MaterialPhong();
    mat->diffuse = {M.diffuse[0], M.diffuse[1], M.diffuse[2]};
    mat->specular = {M.specular[0], M.specular[1], M.specular[2]};
    mat->emission = {M.emission[0], M.emission[1], M.emission[2]};
    mat->shininess_ = M.shininess;
    mat->ior = M.ior;
    materials.push_back(mat);
  }
  for (uint32_t s = 0; s < sd->n_surfaces; ++s) {
    const RbSurface& sf = sd->surfaces[s];
    Surface* surface = new Surface(std::string("surface") + std::to_string(s), sf.n_tris);
    surface->set_material(materials[sf.material]);
    FakeGeometry* g = new FakeGeometry();
    g->geomID = s;
    g->userData = materials[sf.material];
    for (uint32_t i = 0; i < sf.n_tris; ++i) {
      Vertex v[3];
      for (int k = 0; k < 3; ++k) {
        const float* p = sf.pos + 9 * (size_t)i + 3 * k;
        const float* n = sf.normal + 9 * (size_t)i + 3 * k;
        v[k] = Vertex(glm::vec3{p[0], p[1], p[2]}, glm::vec3{n[0], n[1], n[2]}, glm::vec3{0}, glm::vec3{0});
        g->normals.push_back(glm::vec3{n[0], n[1], n[2]});
        if (sf.uv) g->uvs.push_back(glm::vec2{sf.uv[6 * (size_t)i + 2 * k], sf.uv[6 * (size_t)i + 2 * k + 1]});
        if (sf.tangent) {
          const float* tg = sf.tangent + 9 * (size_t)i + 3 * k;
          g->tangents.push_back(glm::vec3{tg[0], tg[1], tg[2]});
        }
        g->ids.push_back(materials[sf.material]->isEmissive() ? static_cast<float>(triIdCtr) : 0.0f);
      }
      surface->get_triangle(i) = Triangle(v[0], v[1], v[2], surface);
      if (materials[sf.material]->isEmissive()) {
        emissiveTris.push_back(&surface->get_triangle(i));
        triangles.push_back(&surface->get_triangle(i));
        triIdCtr += 1;
      }
    }
    surfaces.push_back(surface);
    fs->geoms.push_back(g);
  }
  cdf = TriangleCDF(emissiveTris);
  fs->tracer = orc_create(8, 8, 123, 0, 0, /*bvh2*/ 1, 0);
  orc_upload_scene(fs->tracer, sd);
  g_scene = fs;
  scene = (RTCScene)fs;
}
const RTCScene& Scene::getRTCScene() const { return scene; }
const Sky& Scene::getSkybox() const { return *skybox; }
const TriangleCDF& Scene::getEmissiveCDF() const { return cdf; }

// (Texture is the reference's own P/Texture.cpp, compiled in place against stubs/freeimage.h)

// ---- C interface -------------------------------------------------------------------------------------
struct RefCtx {
  Scene* scene = nullptr;
  Camera* camera = nullptr;
  int w = 0, h = 0;
  std::vector<Reservoir> res[3];
  std::vector<glm::vec3> frame;
  int frameCtr = 0;
};

extern "C" {

void* ref_create(int width, int height, const RbSceneDesc* sd) {
  RefCtx* c = new RefCtx();
  c->w = width, c->h = height;
  g_pending = sd;
  RTCDevice dev{};
  c->scene = new Scene("", dev);
  g_pending = nullptr;
  SimpleGuiDX11::width_ = width;
  SimpleGuiDX11::height_ = height;
  for (auto& r : c->res) r.assign((size_t)width * height, Reservoir{});
  SimpleGuiDX11::reservoirsPingPong[0] = c->res[0].data();
  SimpleGuiDX11::reservoirsPingPong[1] = c->res[1].data();
  SimpleGuiDX11::reservoirsLastFrame = c->res[2].data();
  SimpleGuiDX11::readReservoirBufferIndex = 0;
  SimpleGuiDX11::writeReservoirBufferIndex = 1;
  SimpleGuiDX11::gBuffer = GBuffer(glm::vec<2, int>{width, height});
  SimpleGuiDX11::gBufferLastFrame = GBuffer(glm::vec<2, int>{width, height});
  c->frame.assign((size_t)width * height, glm::vec3{0});
  Utils::generator.seed(123);  // P/utils.cpp:175
  return c;
}

// Material::set_texture with Texture objects whose members are filled from the RbTexture arrays (what the constructor
// stores after FreeImage_ConvertToRawBits), configured like ModelLoader::TextureProxy does: BILINEAR, REPEAT
void ref_set_textures(void* h, const RbTexture* textures, uint32_t n_textures, const RbMaterialTextures* per_material, uint32_t n_materials) {
  RefCtx* c = (RefCtx*)h;
  std::vector<Texture*> tex(n_textures);
  for (uint32_t t = 0; t < n_textures; ++t) {
    Texture* T = new Texture("", BILINEAR, REPEAT);
    T->width_ = textures[t].width, T->height_ = textures[t].height;
    T->scan_width_ = textures[t].scan_width, T->pixel_size_ = textures[t].pixel_size;
    T->data_ = new BYTE[(size_t)T->scan_width_ * T->height_];
    memcpy(T->data_, textures[t].data, (size_t)T->scan_width_ * T->height_);
    tex[t] = T;
  }
  const std::vector<Material*>& mats = c->scene->materials;
  for (uint32_t m = 0; m < n_materials && m < mats.size(); ++m) {
    const RbMaterialTextures& s = per_material[m];
    if (s.diffuse >= 0) mats[m]->set_texture(Material::kDiffuseMapSlot, tex[s.diffuse]);
    if (s.specular >= 0) mats[m]->set_texture(Material::kSpecularMapSlot, tex[s.specular]);
    if (s.shininess >= 0) mats[m]->set_texture(Material::kShininessMapSlot, tex[s.shininess]);
    if (s.normal >= 0) mats[m]->set_texture(Material::kNormalMapSlot, tex[s.normal]);
  }
}

// Scene::setSkybox (P/Scene.cpp:47-50): a SphericalMap whose Texture (constructor defaults: BILINEAR, CLAMP_TO_EDGE; the
// FreeImage stand-in leaves it empty) gets its members from the RbTexture
void ref_set_sky(void* h, const RbTexture* sky) {
  RefCtx* c = (RefCtx*)h;
  std::unique_ptr<SphericalMap> map = std::make_unique<SphericalMap>("");
  Texture* T = map->texture.get();
  T->width_ = sky->width, T->height_ = sky->height;
  T->scan_width_ = sky->scan_width, T->pixel_size_ = sky->pixel_size;
  T->data_ = new BYTE[(size_t)T->scan_width_ * T->height_];
  memcpy(T->data_, sky->data, (size_t)T->scan_width_ * T->height_);
  std::unique_ptr<Sky> newSky = std::move(map);
  std::swap(c->scene->skybox, newSky);
}

void ref_set_params(void*, const RbParams* p) {
  ReSTIRIntegrator::M_Area = p->M_Area;
  ReSTIRIntegrator::M_Brdf = p->M_Brdf;
  ReSTIRIntegrator::spatialReuseNeighborCount = p->spatialReuseNeighborCount;
  ReSTIRIntegrator::spatialPassCount = p->spatialPassCount;
  ReSTIRIntegrator::confidenceCap = p->confidenceCap;
  ReSTIRIntegrator::spatialReuseRadius = p->spatialReuseRadius;
  ReSTIRIntegrator::minNormalSimilarity = p->minNormalSimilarity;
  ReSTIRIntegrator::maxDepthDifference = p->maxDepthDifference;
  ReSTIRIntegrator::doSpatialReuse = p->doSpatialReuse != 0;
  ReSTIRIntegrator::doTemporalReuse = p->doTemporalReuse != 0;
  ReSTIRIntegrator::doVisibilityPass = p->doVisibilityPass != 0;
  ReSTIRIntegrator::rejectDissimilarNeighbors = p->rejectDissimilarNeighbors != 0;
  // the enum is private: its numeric values are the contract of RbSpatialWeightCalc
  *reinterpret_cast<int*>(&ReSTIRIntegrator::spatialWeightCalc) = p->spatialWeightCalc;
  ReSTIRIntegrator::renderParams.tnearOffset = p->tnearOffset;
  ReSTIRIntegrator::renderParams.tfarOffset = p->tfarOffset;
  ReSTIRIntegrator::renderParams.normalOffset = p->normalOffset;
  ReSTIRIntegrator::renderParams.bgColor = {p->bgColor[0], p->bgColor[1], p->bgColor[2]};
  ReSTIRIntegrator::renderParams.useSkybox = p->useSkybox != 0;
}

// the reference Camera's own matrices (glm::lookAt / glm::inverse), exported so the oracle gets identical inputs
void ref_camera(void* h, float fov_deg, const float* from, const float* at, RbCamera* out) {
  RefCtx* c = (RefCtx*)h;
  delete c->camera;
  c->camera = new Camera(c->w, c->h, fov_deg, glm::vec3{from[0], from[1], from[2]}, glm::vec3{at[0], at[1], at[2]});
  const glm::vec3 p = c->camera->getPosition();
  out->pos[0] = p.x, out->pos[1] = p.y, out->pos[2] = p.z;
  out->focal_px = c->camera->getFocalLength();
  memcpy(out->viewMat, &c->camera->getViewMat()[0][0], 64);
  memcpy(out->invViewMat, &c->camera->getInvViewMat()[0][0], 64);
}

static void swapReservoirBuffers() {  // P/simpleguidx11.h:116
  std::swap(SimpleGuiDX11::readReservoirBufferIndex, SimpleGuiDX11::writeReservoirBufferIndex);
}

// produceRestir's schedule, serial (P/simpleguidx11.cpp:359-487 with _DEBUG defined)
void ref_produce_restir(void* h, float* rgb_out) {
  RefCtx* c = (RefCtx*)h;
  const Scene& scene = *c->scene;
  Camera& camera_ = *c->camera;
  GBuffer& gBuffer = SimpleGuiDX11::gBuffer;
  const int width_ = c->w, height_ = c->h;
  gBuffer.setViewMat(camera_.getViewMat());
  gBuffer.setInvViewMat(camera_.getInvViewMat());
  gBuffer.setCameraPos(camera_.getPosition());
  gBuffer.setFocalLength(camera_.getFocalLength());
  for (int y = 0; y < height_; ++y)
    for (int x = 0; x < width_; ++x) ReSTIRIntegrator::gBufferFillPass(scene, {x, y}, camera_);
  for (int y = 0; y < height_; ++y)
    for (int x = 0; x < width_; ++x) ReSTIRIntegrator::initialRenderPass(scene, {x, y});
  if (ReSTIRIntegrator::doVisibilityPass)
    for (int y = 0; y < height_; ++y)
      for (int x = 0; x < width_; ++x) ReSTIRIntegrator::visibilityPass(scene, {x, y});
  if (ReSTIRIntegrator::doTemporalReuse && c->frameCtr > 0) {
    swapReservoirBuffers();
    for (int y = 0; y < height_; ++y)
      for (int x = 0; x < width_; ++x) ReSTIRIntegrator::temporalReusePass({x, y}, scene, camera_);
  }
  if (ReSTIRIntegrator::doSpatialReuse) {
    for (int i = 0; i < ReSTIRIntegrator::spatialPassCount; ++i) {
      swapReservoirBuffers();
      for (int y = 0; y < height_; ++y)
        for (int x = 0; x < width_; ++x) ReSTIRIntegrator::spatialReusePass({x, y}, scene);
    }
  }
  swapReservoirBuffers();
  for (int y = 0; y < height_; ++y) {
    for (int x = 0; x < width_; ++x) {
      const int offset = (y * width_ + x);
      glm::vec3 pixel{};
      Reservoir& r = SimpleGuiDX11::getReservoirRead({x, y});
      if (r.hasSample()) {
        glm::vec3 f_value = ReSTIRIntegrator::evaluateF(r.bestSample, scene, gBuffer.getCameraPos(), gBuffer.getAt({x, y}), true);
        pixel = f_value * r.W;
      } else
        pixel = gBuffer.getAt({x, y}).emission;
      Integrator::sanitize(pixel, false);
      c->frame[offset] = pixel;
    }
  }
  memcpy(SimpleGuiDX11::reservoirsLastFrame, SimpleGuiDX11::reservoirsPingPong[SimpleGuiDX11::readReservoirBufferIndex],
         (size_t)width_ * height_ * sizeof(Reservoir));
  SimpleGuiDX11::gBufferLastFrame.setDataFrom(gBuffer);
  c->frameCtr++;
  if (rgb_out) memcpy(rgb_out, c->frame.data(), c->frame.size() * sizeof(glm::vec3));
}

// N2: the reference's ground-truth estimator. Raytracer::get_pixel (P/raytracer.cpp:40-46) for every pixel
// in the serial _DEBUG order; NEEPathIntegrator::integrateImpl2 at bounce 0 with "Calculate DI" on and "Calculate GI"
// off (P/NEEPathIntegrator.cpp:76-131) is replayed here (its class drags in every other integrator), the estimator
// itself is the reference's own DirectMISIntegrator::calculateDirectLighting (P/DirectMISIntegrator.cpp, compiled in
// place) with its defaults (sample the BRDF and the light sources).
void ref_produce_mis(void* h, float* rgb_out) {
  RefCtx* c = (RefCtx*)h;
  const Scene& scene = *c->scene;
  Camera& camera_ = *c->camera;
  const RenderParams& renderParams = ReSTIRIntegrator::renderParams;  // the offsets / bgColor ref_set_params stored
  static DirectMISIntegrator misIntegrator;
  for (int y = 0; y < c->h; ++y)
    for (int x = 0; x < c->w; ++x) {
      Ray ray{camera_.GenerateRay(glm::vec2{x, y})};
      glm::vec3 radiance{0};
      auto intResult = Intersection::intersectEmbree(scene, ray);
      const HitInfo& hitInfo = intResult.hitInfo;
      if (hitInfo.didHit) {
        if (intResult.material->isEmitter()) {
          radiance = intResult.material->emission;  // lastVertexType == CAMERA_VERTEX
        } else {
          glm::vec3 L_i_direct = misIntegrator.calculateDirectLighting(scene, ray, intResult, renderParams, {x, y});
          Integrator::sanitize(L_i_direct, false);
          glm::vec3 L_i_indirect{0};
          Integrator::sanitize(L_i_indirect, false);
          radiance = L_i_indirect + L_i_direct;
        }
      } else {
        radiance = renderParams.useSkybox ? scene.getSkybox().getTexel(ray.getDir()) : renderParams.bgColor;  // P/NEEPathIntegrator.cpp:131
      }
      c->frame[(size_t)y * c->w + x] = radiance;
    }
  if (rgb_out) memcpy(rgb_out, c->frame.data(), c->frame.size() * sizeof(glm::vec3));
}

// final reservoirs of the frame: {point3, normal3, Li3, w_sum, W, confidence(as float)} = 12 floats / px
void ref_reservoirs(void* h, float* out12) {
  RefCtx* c = (RefCtx*)h;
  const Reservoir* R = SimpleGuiDX11::reservoirsLastFrame;
  for (size_t i = 0; i < (size_t)c->w * c->h; ++i) {
    float* o = out12 + 12 * i;
    o[0] = R[i].bestSample.samplePoint.x, o[1] = R[i].bestSample.samplePoint.y, o[2] = R[i].bestSample.samplePoint.z;
    o[3] = R[i].bestSample.sampleNormal.x, o[4] = R[i].bestSample.sampleNormal.y, o[5] = R[i].bestSample.sampleNormal.z;
    o[6] = R[i].bestSample.L_i.x, o[7] = R[i].bestSample.L_i.y, o[8] = R[i].bestSample.L_i.z;
    o[9] = R[i].w_sum, o[10] = R[i].W, o[11] = (float)R[i].confidence;
  }
}
// G-buffer of the frame: {pos3, normal3, diffuse3, specular3, emission3, shininess, depth, matType} = 18 floats / px
void ref_gbuffer(void* h, float* out18) {
  RefCtx* c = (RefCtx*)h;
  const GBuffer& G = SimpleGuiDX11::gBufferLastFrame;
  for (int y = 0; y < c->h; ++y)
    for (int x = 0; x < c->w; ++x) {
      GBufferElement e = G.getAt({x, y});
      float* o = out18 + 18 * ((size_t)y * c->w + x);
      const glm::vec3* v[5] = {&e.worldSpacePos, &e.worldSpaceNormal, &e.diffuseColor, &e.specularColor, &e.emission};
      for (int k = 0; k < 5; ++k) o[3 * k] = v[k]->x, o[3 * k + 1] = v[k]->y, o[3 * k + 2] = v[k]->z;
      o[15] = e.shininess, o[16] = e.depth, o[17] = (float)e.materialType;
    }
}

// ---- leaf functions of the reference, for the golden fixtures ------------------------------------------
void ref_seed(uint32_t s) { Utils::generator.seed(s); }
float ref_random() { return Utils::getRandomValue(0.0f, 1.0f); }
void ref_sampleDiskUniform(float radius, float* out2) {
  glm::vec2 v = Sampling::sampleDiskUniform(radius);
  out2[0] = v.x, out2[1] = v.y;
}
// tri = {p0,p1,p2,n0,n1,n2}; out = {point3, normal3, pdf}
void ref_sampleTriangle(const float* t, float* out7) {
  Vertex v[3];
  for (int k = 0; k < 3; ++k)
    v[k] = Vertex(glm::vec3{t[3 * k], t[3 * k + 1], t[3 * k + 2]}, glm::vec3{t[9 + 3 * k], t[10 + 3 * k], t[11 + 3 * k]}, glm::vec3{0}, glm::vec3{0});
  Triangle tri(v[0], v[1], v[2], nullptr);
  TrianglePointSample s = Sampling::sampleTriangle(tri);
  out7[0] = s.samplePoint.x, out7[1] = s.samplePoint.y, out7[2] = s.samplePoint.z;
  out7[3] = s.normal.x, out7[4] = s.normal.y, out7[5] = s.normal.z, out7[6] = s.pdf;
}
static GBufferElement elem_from(const float* e) {
  GBufferElement g;
  g.worldSpacePos = {e[0], e[1], e[2]};
  g.worldSpaceNormal = {e[3], e[4], e[5]};
  g.diffuseColor = {e[6], e[7], e[8]};
  g.specularColor = {e[9], e[10], e[11]};
  g.shininess = e[12];
  g.materialType = PHONG;
  return g;
}
void ref_phong_evalBRDF(const float* elem, const float* cam, const float* wi, float* out3) {
  glm::vec3 r = MaterialPhong::evalBRDF(elem_from(elem), {cam[0], cam[1], cam[2]}, {wi[0], wi[1], wi[2]});
  out3[0] = r.x, out3[1] = r.y, out3[2] = r.z;
}
float ref_phong_evalPdf(const float* elem, const float* cam, const float* wi) {
  return MaterialPhong::evalPdf(elem_from(elem), {cam[0], cam[1], cam[2]}, {wi[0], wi[1], wi[2]});
}
void ref_phong_sampleBRDF(const float* elem, const float* cam, float* out4) {
  PTInfoGI s = MaterialPhong::sampleBRDF(elem_from(elem), {cam[0], cam[1], cam[2]});
  out4[0] = s.omega_i.x, out4[1] = s.omega_i.y, out4[2] = s.omega_i.z, out4[3] = s.pdf;
}
// light pick through the reference's TriangleCDF of the scene: out = {index into TriangleCDF::tris, pdf}
void ref_cdf_pick(void* h, float* out2) {
  RefCtx* c = (RefCtx*)h;
  const TriangleCDF& cdf = c->scene->getEmissiveCDF();
  TriangleCDFSample s = cdf.getTriangle();
  int idx = -1;
  for (size_t i = 0; i < cdf.tris.size(); ++i)
    if (cdf.tris[i] == &s.triangle) idx = (int)i;
  out2[0] = (float)idx, out2[1] = s.pdf;
}
// the Producer loop's post-path arithmetic: Utils::aces, Utils::compress (P/utils.cpp:190-197,
// 220-230) and the accumulator update glm::mix(acc, frame, 1 / (accFrameCtr + 1)) (P/simpleguidx11.cpp:251)
void ref_aces(float* rgb) {
  glm::vec3 v{rgb[0], rgb[1], rgb[2]};
  Utils::aces(v);
  rgb[0] = v.x, rgb[1] = v.y, rgb[2] = v.z;
}
float ref_compress(float u) {
  Utils::compress(u);
  return u;
}
void ref_accumulate_mix(float* acc, const float* frame, int accFrameCtr) {
  glm::vec3 a{acc[0], acc[1], acc[2]}, f{frame[0], frame[1], frame[2]};
  a = glm::mix(a, f, 1.0f / static_cast<float>(accFrameCtr + 1));
  acc[0] = a.x, acc[1] = a.y, acc[2] = a.z;
}
int ref_sanitize(float* rgb) {
  glm::vec3 v{rgb[0], rgb[1], rgb[2]};
  bool ch = Integrator::sanitize(v, false);
  rgb[0] = v.x, rgb[1] = v.y, rgb[2] = v.z;
  return ch ? 1 : 0;
}
}
