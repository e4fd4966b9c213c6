// rb_host_scene.h — host-side scene preparation shared by the CUDA library and the
// test-only emulation harness: flattening in surface-then-triangle order
// (P/ModelLoader.cpp:227-318), the TriangleCDF table (P/TriangleCDF.cpp:8-34) and the
// Vose alias table (DESIGN.md "alias table": fixed processing order).
#ifndef RB_HOST_SCENE_H_
#define RB_HOST_SCENE_H_

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "rb_common.cuh"

namespace rb {

struct HostScene {
  size_t n = 0;
  std::vector<float> pos;        // [n][9]
  std::vector<F4> nrm;           // [3n]
  std::vector<F4> uv;            // [2n]: {u0,v0,u1,v1} {u2,v2,0,0}; empty when no surface carries texture coordinates
  std::vector<F4> tan;           // [3n] packed like nrm; empty when no surface carries tangents (normal maps only)
  std::vector<U4> info;          // [n]
  std::vector<F4> mat;           // [3 * n_mat]
  std::vector<uint32_t> emissive;
  std::vector<float> cdf, alias_prob;
  std::vector<uint32_t> alias_idx;
  std::vector<uint32_t> alias_pair;  // [2 * n_lights]: {bits(alias_prob[i]), alias_idx[i]} — one 8-byte load per pick
  std::vector<F4> light;  // [6 * n_lights]
  std::vector<F4> light_cull;  // [n_lights]: {bounding sphere centre, radius} for initial_pixel's horizon pre-test; radius = +inf: never
  float maxabs = 0, totalSurface = 0;
};

inline int flatten_scene(const RbSceneDesc* sd, HostScene& hs, std::string& err) {
  size_t n = 0;
  for (uint32_t s = 0; s < sd->n_surfaces; ++s) {
    const RbSurface& sf = sd->surfaces[s];
    if (sf.material >= sd->n_materials || (sf.n_tris && (!sf.pos || !sf.normal))) {
      err = "rb_upload_scene: bad surface " + std::to_string(s);
      return RB_ERR_INVALID_ARGUMENT;
    }
    n += sf.n_tris;
  }
  if (n > 0x7FFFFFF0ull) {
    err = "rb_upload_scene: too many triangles";
    return RB_ERR_UNSUPPORTED;
  }
  // ---- flatten in surface-then-triangle order (P/ModelLoader.cpp:227-318) ----------------
  hs.n = n;
  std::vector<float>& pos = hs.pos;
  std::vector<F4>& nrm = hs.nrm;
  std::vector<U4>& info = hs.info;
  std::vector<F4>& mat = hs.mat;
  pos.assign(9 * n, 0.0f);
  nrm.assign(3 * n, F4{0, 0, 0, 0});
  bool any_uv = false;
  for (uint32_t s = 0; s < sd->n_surfaces; ++s) any_uv = any_uv || (sd->surfaces[s].uv != nullptr && sd->surfaces[s].n_tris > 0);
  hs.uv.clear();
  if (any_uv) hs.uv.assign(2 * n, F4{0, 0, 0, 0});
  bool any_tan = false;
  for (uint32_t s = 0; s < sd->n_surfaces; ++s) any_tan = any_tan || (sd->surfaces[s].tangent != nullptr && sd->surfaces[s].n_tris > 0);
  hs.tan.clear();
  if (any_tan) hs.tan.assign(3 * n, F4{0, 0, 0, 0});
  info.assign(n, U4{0, 0, 0, 0});
  mat.assign(3 * (size_t)sd->n_materials, F4{0, 0, 0, 0});
  for (uint32_t m = 0; m < sd->n_materials; ++m) {
    const RbMaterial& M = sd->materials[m];
    mat[3 * m + 0] = F4{M.diffuse[0], M.diffuse[1], M.diffuse[2], M.shininess};
    mat[3 * m + 1] = F4{M.specular[0], M.specular[1], M.specular[2], u2f(M.type)};
    mat[3 * m + 2] = F4{M.emission[0], M.emission[1], M.emission[2], M.ior};
  }
  std::vector<uint32_t>& emissive = hs.emissive;  // TriangleCDF::tris
  emissive.clear();
  float maxabs = 0;
  {
    size_t t = 0;
    for (uint32_t s = 0; s < sd->n_surfaces; ++s) {
      const RbSurface& sf = sd->surfaces[s];
      const RbMaterial& M = sd->materials[sf.material];
      const bool isEmissive = M.emission[0] + M.emission[1] + M.emission[2] > 0;  // Material::isEmissive, P/material.h:135-137
      memcpy(pos.data() + 9 * t, sf.pos, sizeof(float) * 9 * (size_t)sf.n_tris);
      for (uint32_t i = 0; i < sf.n_tris; ++i, ++t) {
        const float* q = sf.normal + 9 * (size_t)i;
        nrm[3 * t + 0] = F4{q[0], q[1], q[2], q[3]};
        nrm[3 * t + 1] = F4{q[4], q[5], q[6], q[7]};
        nrm[3 * t + 2] = F4{q[8], 0, 0, 0};
        if (any_uv && sf.uv) {
          const float* w = sf.uv + 6 * (size_t)i;
          hs.uv[2 * t + 0] = F4{w[0], w[1], w[2], w[3]};
          hs.uv[2 * t + 1] = F4{w[4], w[5], 0, 0};
        }
        if (any_tan && sf.tangent) {
          const float* w = sf.tangent + 9 * (size_t)i;
          hs.tan[3 * t + 0] = F4{w[0], w[1], w[2], w[3]};
          hs.tan[3 * t + 1] = F4{w[4], w[5], w[6], w[7]};
          hs.tan[3 * t + 2] = F4{w[8], 0, 0, 0};
        }
        int eid = -1;
        if (isEmissive) {
          eid = (int)emissive.size();
          emissive.push_back((uint32_t)t);
        }
        info[t] = U4{s, i, sf.material, (uint32_t)eid};
      }
    }
    for (size_t i = 0; i < pos.size(); ++i) {
      const float a = fabsf(pos[i]);
      if (!(a <= 3.0e37f)) {
        err = "rb_upload_scene: non-finite vertex coordinate";
        return RB_ERR_INVALID_ARGUMENT;
      }
      maxabs = std::max(maxabs, a);
    }
  }
  // ---- light tables: TriangleCDF ctor (P/TriangleCDF.cpp:8-34) + Vose alias table -------------
  const size_t NL = emissive.size();
  std::vector<float> area(NL);
  std::vector<float>&cdf = hs.cdf, &alias_prob = hs.alias_prob;
  std::vector<uint32_t>& alias_idx = hs.alias_idx;
  std::vector<F4>& light = hs.light;
  cdf.assign(NL, 0.0f);
  alias_prob.assign(NL, 1.0f);
  alias_idx.assign(NL, 0u);
  light.assign(6 * NL, F4{0, 0, 0, 0});
  float totalSurface = 0;
  for (size_t i = 0; i < NL; ++i) {
    const float* p = pos.data() + 9 * (size_t)emissive[i];
    const V3 p0 = v3(p[0], p[1], p[2]), p1 = v3(p[3], p[4], p[5]), p2 = v3(p[6], p[7], p[8]);
    area[i] = 0.5f * length(cross(p1 - p0, p2 - p0));  // Triangle ctor, P/triangle.cpp:13-16
    totalSurface += area[i];
  }
  for (size_t i = 0; i < NL; ++i) {
    const float normArea = area[i] / totalSurface;
    cdf[i] = (i == 0 ? 0.0f : cdf[i - 1]) + normArea;
  }
  {
    std::vector<float> q(NL);
    std::vector<uint32_t> small, large;
    for (size_t i = 0; i < NL; ++i) {
      alias_idx[i] = (uint32_t)i;
      q[i] = (area[i] / totalSurface) * (float)NL;
      (q[i] < 1.0f ? small : large).push_back((uint32_t)i);
    }
    while (!small.empty() && !large.empty()) {
      const uint32_t s = small.back();
      small.pop_back();
      const uint32_t l = large.back();
      large.pop_back();
      alias_prob[s] = q[s];
      alias_idx[s] = l;
      q[l] = (q[l] + q[s]) - 1.0f;
      (q[l] < 1.0f ? small : large).push_back(l);
    }
  }
  hs.alias_pair.resize(2 * NL);
  for (size_t i = 0; i < NL; ++i) {
    uint32_t bits;
    memcpy(&bits, &alias_prob[i], 4);
    hs.alias_pair[2 * i] = bits;
    hs.alias_pair[2 * i + 1] = alias_idx[i];
  }
  for (size_t i = 0; i < NL; ++i) {
    const size_t t = emissive[i];
    const float* p = pos.data() + 9 * t;
    const RbMaterial& M = sd->materials[info[t].z];
    light[6 * i + 0] = F4{p[0], p[1], p[2], area[i]};
    light[6 * i + 1] = F4{p[3], p[4], p[5], area[i] / totalSurface};
    light[6 * i + 2] = F4{p[6], p[7], p[8], 1.0f / area[i]};
    light[6 * i + 3] = F4{nrm[3 * t].x, nrm[3 * t].y, nrm[3 * t].z, M.emission[0]};
    light[6 * i + 4] = F4{nrm[3 * t].w, nrm[3 * t + 1].x, nrm[3 * t + 1].y, M.emission[1]};
    light[6 * i + 5] = F4{nrm[3 * t + 1].z, nrm[3 * t + 1].w, nrm[3 * t + 2].x, M.emission[2]};
  }

  // Bounding spheres for the horizon pre-test of the initial pass (rb_passes.cuh: surely_below_horizon). A light is
  // eligible only when everything the exact cull test of initial_pixel asks of the LIGHT holds for every point of it:
  // pick probability, 1/area and emission finite and positive, and vertex normals whose every convex combination can
  // be normalised (all three of length 0.5..2 and within 75 degrees of the first). Everything else gets radius = +inf.
  hs.light_cull.assign(NL, F4{0, 0, 0, std::numeric_limits<float>::infinity()});
  for (size_t i = 0; i < NL; ++i) {
    const F4* L = &light[6 * i];
    auto fin = [](float x) { return std::isfinite(x); };
    bool ok = L[1].w > 0.0f && fin(L[1].w) && L[2].w > 0.0f && fin(L[2].w);
    ok = ok && fin(L[3].w) && fin(L[4].w) && fin(L[5].w) && (L[3].w > 0.0f || L[4].w > 0.0f || L[5].w > 0.0f);
    double nl[3];
    for (int k = 0; k < 3 && ok; ++k) {
      nl[k] = std::sqrt((double)L[3 + k].x * L[3 + k].x + (double)L[3 + k].y * L[3 + k].y + (double)L[3 + k].z * L[3 + k].z);
      ok = nl[k] >= 0.5 && nl[k] <= 2.0;
      const double d0 = (double)L[3 + k].x * L[3].x + (double)L[3 + k].y * L[3].y + (double)L[3 + k].z * L[3].z;
      ok = ok && d0 >= 0.25 * nl[k] * nl[0];
    }
    if (!ok) continue;
    const double cx = ((double)L[0].x + L[1].x + L[2].x) / 3.0, cy = ((double)L[0].y + L[1].y + L[2].y) / 3.0,
                 cz = ((double)L[0].z + L[1].z + L[2].z) / 3.0;
    const float fx = (float)cx, fy = (float)cy, fz = (float)cz;  // the radius is measured from the rounded centre
    double r2 = 0;
    for (int k = 0; k < 3; ++k) {
      const double dx = (double)L[k].x - fx, dy = (double)L[k].y - fy, dz = (double)L[k].z - fz;
      r2 = std::max(r2, dx * dx + dy * dy + dz * dz);
    }
    const float rad = (float)(std::sqrt(r2) * 1.00001) ;
    hs.light_cull[i] = F4{fx, fy, fz, std::nextafter(rad, std::numeric_limits<float>::infinity())};
  }

  hs.maxabs = maxabs;
  hs.totalSurface = totalSurface;
  return RB_OK;
}

}  // namespace rb
#endif
