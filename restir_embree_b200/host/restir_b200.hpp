// restir_b200.hpp — C++ host-side mirror of the reference's interface for the ReSTIR path, over the C ABI
// (include/restir_b200.h). Header only. Names, argument meaning and error behaviour follow the reference so that a
// maintainer can swap the body of SimpleGuiDX11::produceRestir for one call (INTEGRATION.md shows the diff):
//
//   rb200::ReSTIRIntegrator   the static parameters of P/ReSTIRIntegrator.{h,cpp} (same names, same defaults)
//   rb200::Camera             P/camera.{h,cpp}: ctor(width, height, fov_y, view_from, view_at), setPosition, setFOV,
//                             getViewMat/getInvViewMat/getPosition/getFocalLength (z-up, glm::lookAt conventions)
//   rb200::Scene              what ModelLoader::loadScene produces: surfaces (non-indexed triangle soups) + materials
//   rb200::Renderer           owns the device state that SimpleGuiDX11 keeps in its static buffers;
//                             produceRestir(camera, frameCtr, frame_data) == P/simpleguidx11.cpp:359-487
//
// Errors surface as std::runtime_error, like the reference's Embree error callback (P/tutorials.cpp:6-24).
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/restir_b200.h"

namespace rb200 {

struct vec3 {
  float x = 0, y = 0, z = 0;
};
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline vec3 cross(vec3 x, vec3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
inline vec3 normalize(vec3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }

// ---- P/ReSTIRIntegrator.h:91-113, P/ReSTIRIntegrator.cpp:13-35 --------------------------------------------------
struct ReSTIRIntegrator {
  enum SpatialWeightCalculation { CONSTANT, CONSTANT_DEBIAS_CONTRIB, CONSTANT_DEBIAS_Z_TERM, BALANCE_HEURISTIC, PAIRWISE_MIS };
  int M_Area{1};
  int M_Brdf{1};
  int spatialReuseNeighborCount{5};
  int spatialPassCount{1};
  int confidenceCap{20};
  float spatialReuseRadius{30};
  float minNormalSimilarity{0.85f};
  float maxDepthDifference{0.2f};
  bool doSpatialReuse{false};
  bool doTemporalReuse{false};
  bool doVisibilityPass{false};
  bool rejectDissimilarNeighbors{false};
  SpatialWeightCalculation spatialWeightCalc{CONSTANT};
  // RenderParams, P/RenderParams.h:8-17
  struct {
    float tnearOffset{0.01f}, tfarOffset{0.001f}, normalOffset{0.001f};
    vec3 bgColor{0.5f, 0.5f, 0.5f};
    bool useSkybox{false};  // the reference's default is true; here it needs Renderer::setSkybox first
  } renderParams;
  // seams added by this build
  bool aliasLightSampler{true};
  bool wavefront{true};
  bool temporalFetchReprojected{false};  // true: merge the reservoir of the reprojected pixel (the reference reads the same pixel, :641)

  RbParams toAbi() const {
    RbParams p;
    rb_default_params(&p);
    p.M_Area = M_Area, p.M_Brdf = M_Brdf;
    p.spatialReuseNeighborCount = spatialReuseNeighborCount, p.spatialPassCount = spatialPassCount;
    p.confidenceCap = confidenceCap, p.spatialReuseRadius = spatialReuseRadius;
    p.minNormalSimilarity = minNormalSimilarity, p.maxDepthDifference = maxDepthDifference;
    p.doSpatialReuse = doSpatialReuse, p.doTemporalReuse = doTemporalReuse, p.doVisibilityPass = doVisibilityPass;
    p.rejectDissimilarNeighbors = rejectDissimilarNeighbors, p.spatialWeightCalc = (int)spatialWeightCalc;
    p.tnearOffset = renderParams.tnearOffset, p.tfarOffset = renderParams.tfarOffset;
    p.normalOffset = renderParams.normalOffset;
    p.bgColor[0] = renderParams.bgColor.x, p.bgColor[1] = renderParams.bgColor.y, p.bgColor[2] = renderParams.bgColor.z;
    p.useSkybox = renderParams.useSkybox;
    p.lightSampler = aliasLightSampler ? RB_LS_ALIAS : RB_LS_CDF;
    p.wavefront = wavefront;
    p.temporalFetchReprojected = temporalFetchReprojected;
    return p;
  }
};

// ---- P/camera.{h,cpp} ------------------------------------------------------------------------------------------------
class Camera {
 public:
  Camera(int width, int height, float fov_y, vec3 view_from, vec3 view_at)
      : width_(width), height_(height), view_from_(view_from), viewAt(view_at) {
    setFOV(fov_y);
    recalculate_m_c_w();
  }
  void setFOV(float newFOV) {  // :81-84
    fov_y_ = newFOV * 3.14159265358979323846f / 180.0f;
    f_y_ = (float)height_ / (2.0f * tanf(fov_y_ / 2.0f));
  }
  void setPosition(vec3 newPos) {
    view_from_ = newPos;
    recalculate_m_c_w();
  }
  void setViewAt(vec3 newViewAt) {
    viewAt = newViewAt;
    recalculate_m_c_w();
  }
  const vec3& getPosition() const { return view_from_; }
  float getFocalLength() const { return f_y_; }
  const std::array<float, 16>& getViewMat() const { return viewMat; }
  const std::array<float, 16>& getInvViewMat() const { return invViewMat; }

  void recalculate_m_c_w() {  // :44-58, glm::lookAtRH
    const vec3 up{0.0f, 0.0f, 1.0f};
    vec3 z_c = normalize(view_from_ - viewAt);
    vec3 x_c = normalize(cross(up, z_c));
    vec3 y_c = normalize(cross(z_c, x_c));
    const vec3 f = normalize(viewAt - view_from_);
    const vec3 s = normalize(cross(f, y_c));
    const vec3 u = cross(s, f);
    viewMat = {s.x, u.x, -f.x, 0, s.y, u.y, -f.y, 0, s.z, u.z, -f.z, 0, -dot(s, view_from_), -dot(u, view_from_), dot(f, view_from_), 1};
    // rigid transform: inverse = [R^T | eye]
    invViewMat = {s.x, s.y, s.z, 0, u.x, u.y, u.z, 0, -f.x, -f.y, -f.z, 0, view_from_.x, view_from_.y, view_from_.z, 1};
  }
  RbCamera toAbi() const {
    RbCamera c;
    c.pos[0] = view_from_.x, c.pos[1] = view_from_.y, c.pos[2] = view_from_.z;
    c.focal_px = f_y_;
    for (int i = 0; i < 16; ++i) c.viewMat[i] = viewMat[i], c.invViewMat[i] = invViewMat[i];
    return c;
  }

 private:
  int width_, height_;
  float fov_y_{}, f_y_{};
  vec3 view_from_, viewAt;
  std::array<float, 16> viewMat{}, invViewMat{};
};

// ---- scene as ModelLoader::loadScene leaves it (P/ModelLoader.cpp:218-321) ---------------------------------------------------
struct Material {
  RbMaterialType type{RB_MAT_PHONG};
  vec3 diffuse{0.5f, 0.5f, 0.5f}, specular{}, emission{};
  float shininess_{10.0f}, ior{1.0f};
  bool isEmissive() const { return emission.x + emission.y + emission.z > 0; }  // P/material.h:135-137
};
struct Surface {
  std::vector<float> positions;  // 9 floats per triangle
  std::vector<float> normals;    // 9 floats per triangle
  std::vector<float> uvs;        // 6 floats per triangle or empty (Vertex::texture_coords, P/vertex.h)
  std::vector<float> tangents;   // 9 floats per triangle or empty (Vertex::tangent; read by normal-mapped materials)
  uint32_t material{0};
  size_t no_triangles() const { return positions.size() / 9; }
};
struct Scene {
  std::vector<Surface> surfaces;
  std::vector<Material> materials;
};

// ---- the renderer (role of SimpleGuiDX11 around produceRestir) -----------------------------------------------------------------
class Renderer {
 public:
  Renderer(int width, int height, int device = 0, uint32_t seed = 123, int band_y0 = 0, int band_y1 = -1) : width_(width), height_(height) {
    RbCreateInfo ci{};
    ci.width = width, ci.height = height, ci.device = device, ci.seed = seed;
    ci.band_y0 = band_y0, ci.band_y1 = band_y1 < 0 ? height : band_y1;
    ci.collect_timings = 1;
    if (rb_create(&ci, &h_) != RB_OK) throw std::runtime_error(std::string("rb_create: ") + rb_last_error(nullptr));
  }
  ~Renderer() { rb_destroy(h_); }
  Renderer(const Renderer&) = delete;
  Renderer& operator=(const Renderer&) = delete;

  void LoadScene(const Scene& scene) {
    std::vector<RbSurface> s(scene.surfaces.size());
    std::vector<RbMaterial> m(scene.materials.size());
    for (size_t i = 0; i < m.size(); ++i) {
      const Material& a = scene.materials[i];
      m[i].type = a.type;
      m[i].diffuse[0] = a.diffuse.x, m[i].diffuse[1] = a.diffuse.y, m[i].diffuse[2] = a.diffuse.z;
      m[i].specular[0] = a.specular.x, m[i].specular[1] = a.specular.y, m[i].specular[2] = a.specular.z;
      m[i].emission[0] = a.emission.x, m[i].emission[1] = a.emission.y, m[i].emission[2] = a.emission.z;
      m[i].shininess = a.shininess_, m[i].ior = a.ior;
    }
    for (size_t i = 0; i < s.size(); ++i) {
      const Surface& a = scene.surfaces[i];
      s[i].n_tris = (uint32_t)a.no_triangles();
      s[i].material = a.material;
      s[i].pos = a.positions.data(), s[i].normal = a.normals.data();
      s[i].uv = a.uvs.size() == 6 * a.no_triangles() && !a.uvs.empty() ? a.uvs.data() : nullptr;
      s[i].tangent = a.tangents.size() == 9 * a.no_triangles() && !a.tangents.empty() ? a.tangents.data() : nullptr;
    }
    RbSceneDesc d{(uint32_t)s.size(), s.data(), (uint32_t)m.size(), m.data()};
    check(rb_upload_scene(h_, &d), "rb_upload_scene");
  }

  // Raytracer::LoadScene(file_name) (P/raytracer.cpp:35-39): `scene = Scene{file_name, device_}` goes through
  // ModelLoader::loadScene / loadOBJ / loadMaterials; here the library's own OBJ / MTL parser with the same conventions
  // (rb_obj_load, see include/restir_b200.h). gammaCorrect = Raytracer::gammaCorrect.
  void LoadScene(const std::string& file_name, bool gammaCorrect = true) {
    RbObjScene* s = nullptr;
    char err[512] = {0};
    if (rb_obj_load(file_name.c_str(), gammaCorrect ? 1 : 0, &s, err, sizeof(err)) != RB_OK)
      throw std::runtime_error(std::string("rb_obj_load: ") + err);
    const int rc = rb_upload_scene(h_, rb_obj_scene_desc(s));
    rb_obj_free(s);
    check(rc, "rb_upload_scene");
  }

  // Material::set_texture for the uploaded scene (P/material.h:79-84): the texel arrays of the host's Texture objects
  // (width_, height_, scan_width_, pixel_size_, data_) and, per material, the texture index of each slot or -1
  void setTextures(const std::vector<RbTexture>& textures, const std::vector<RbMaterialTextures>& perMaterial) {
    check(rb_set_textures(h_, textures.data(), (uint32_t)textures.size(), perMaterial.data(), (uint32_t)perMaterial.size()),
          "rb_set_textures");
  }

  // Scene::setSkybox (P/Scene.cpp:47-50): the texel array of the decoded sky image; nullptr removes it. Needed before
  // params.renderParams.useSkybox = true.
  void setSkybox(const RbTexture* sky) { check(rb_set_sky(h_, sky), "rb_set_sky"); }

  // SimpleGuiDX11::produceRestir: frame_data is width*height float3 (linear HDR), owned by the caller
  void produceRestir(const Camera& camera, uint32_t frameCtr, float* frame_data) {
    RbParams p = params.toAbi();
    check(rb_set_params(h_, &p), "rb_set_params");
    RbCamera c = camera.toAbi();
    check(rb_render_frame(h_, &c, frameCtr, frame_data, &timings), "rb_render_frame");
  }

  // Raytracer::get_pixel over the image with the reference's ground-truth integrator selected in its GUI: NEEPathIntegrator,
  // "Calculate DI" on, "Calculate GI" off, "MIS Sampler" (P/NEEPathIntegrator.cpp:31-52, P/DirectMISIntegrator.cpp:32-36).
  // sampleBRDF / sampleLightSources are DirectMISIntegrator's two checkboxes. Does not touch the ReSTIR state.
  void produceMisGroundTruth(const Camera& camera, uint32_t frameCtr, float* frame_data, bool sampleBRDF = true,
                             bool sampleLightSources = true) {
    RbParams p = params.toAbi();
    check(rb_set_params(h_, &p), "rb_set_params");
    RbCamera c = camera.toAbi();
    const uint32_t techniques = (sampleBRDF ? RB_MIS_SAMPLE_BRDF : 0u) | (sampleLightSources ? RB_MIS_SAMPLE_LIGHTS : 0u);
    check(rb_render_mis_frame(h_, &c, frameCtr, techniques, frame_data), "rb_render_mis_frame");
  }

  // the rest of SimpleGuiDX11::Producer's loop body after produceRestir (P/simpleguidx11.cpp:246-326): accumulate the
  // frame, fill display_data (width*height float4, may be null: kept on the device), update accumulatorMean /
  // accumulatorVariance. accFrameCtr, tonemap and gammaCorrect are the reference's members of the same names.
  void accumulateAndDisplay(uint32_t accFrameCtr, bool tonemap, bool gammaCorrect, float* display_data) {
    RbImageStats st{};
    check(rb_accumulate_display(h_, accFrameCtr, tonemap ? 1 : 0, gammaCorrect ? 1 : 0, display_data, &st), "rb_accumulate_display");
    accumulatorMean = st.mean;
    accumulatorVariance = st.variance;
  }
  double accumulatorMean{0}, accumulatorVariance{0};  // P/simpleguidx11.h:103-104

  ReSTIRIntegrator params;  // edit like the reference's statics
  RbTimings timings{};      // gBUfferFillDuration ... totalFrameDuration of P/simpleguidx11.h:120-127
  RbHandle handle() const { return h_; }
  int width() const { return width_; }
  int height() const { return height_; }

 private:
  void check(int rc, const char* what) {
    if (rc != RB_OK) throw std::runtime_error(std::string(what) + ": " + rb_last_error(h_));
  }
  RbHandle h_ = nullptr;
  int width_, height_;
};

}  // namespace rb200
