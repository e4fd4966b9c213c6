// rb_obj_loader.h — Wavefront OBJ / MTL ingestion with the reference's conventions (first half).
//
// Host-only code. The reference loads scenes with ASSIMP (ModelLoader::loadOBJ / loadMaterials,
// P/ModelLoader.cpp:41-216; flags Triangulate | JoinIdenticalVertices | OptimizeGraph | OptimizeMeshes |
// CalcTangentSpace) and hands Embree one non-indexed triangle soup per surface (loadScene, :218-321). ASSIMP is not
// available here (headers only, no binary), so this is a parser of our own that ends in the same RbSceneDesc; what it
// keeps from the reference:
//   * material type from the MTL key `Pc` (ASSIMP maps it to AI_MATKEY_CLEARCOAT_FACTOR): 0 NORMAL, 1 LAMBERT,
//     2 PHONG, 3 MIRROR, 4 DIELECTRIC, 5 DIELECTRIC_TRANSPARENT, anything else UNSUPPORTED (:57-70, P/enums.h);
//   * Kd / Ks / Ka expanded from sRGB to linear when gamma correction is on (Raytracer::gammaCorrect defaults to true;
//     Utils::expand, P/utils.cpp:209-218), Ke / Ns / Ni taken as they are (:100-116); ASSIMP's OBJ defaults for absent
//     keys (Kd 0.6, everything else 0, Ni 1);
//   * materials are indexed in MTL order (the reference skips ASSIMP's default material 0 and indexes with
//     mMaterialIndex - 1, :45, :213): a face without `usemtl` has no material there (std::out_of_range) and is an
//     error here;
//   * one surface per material (what OptimizeMeshes leaves), in order of first use; non-indexed vertices with position,
//     normal and uv; geomID = surface index, primID = triangle index inside it; emitters are found downstream from Ke.
// What cannot be pinned without ASSIMP (stated in DESIGN.md): polygon triangulation (here: a fan from the first
// vertex), mesh order for files that interleave materials, and the generated tangents: when some material names a
// normal map (map_Kn / bump / map_bump), every surface gets per-vertex tangents by the PER-FACE step of ASSIMP's
// CalcTangentSpace (uv-gradient tangent, projected into the plane of each vertex normal, normalised; degenerate uv
// -> the default uv directions) — ASSIMP's later smoothing of tangents across faces that share a vertex is not
// reproduced (Intersection::intersectEmbree re-orthonormalises per hit anyway, so only the in-plane direction matters).
// Texture file names are kept per material for the host's image loader. Faces without normals get the flat face normal
// (the reference dereferences a null mNormals there).
#ifndef RB_OBJ_LOADER_H_
#define RB_OBJ_LOADER_H_

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/restir_b200.h"

namespace rbobj {

struct Surface {
  uint32_t material = 0;
  std::vector<float> pos, normal, uv;  // 3 vertices per triangle
  std::vector<float> tangent;          // 3 vertices per triangle, only when a material of the scene has a normal map
};
struct Scene {
  std::vector<RbMaterial> materials;
  std::vector<std::string> material_names;
  std::vector<std::string> map_kd, map_ks, map_ns, map_kn;  // texture file names as written in the MTL ("" = none)
  std::vector<Surface> surfaces;
  std::vector<RbSurface> abi_surfaces;
  RbSceneDesc desc{};
  uint64_t n_triangles = 0;
};

// The per-face step of ASSIMP's CalcTangentsProcess::ProcessMesh (aiProcess_CalcTangentSpace, P/ModelLoader.cpp:167):
// v = p1 - p0, w = p2 - p0, (sx, sy) = uv1 - uv0, (tx, ty) = uv2 - uv0; tangent = (w * sy - v * ty) * sign(tx*sy - ty*sx)
// (default directions when the uv triangle is degenerate); per vertex: project into the normal's plane, normalise.
inline void calc_tangents(Surface& S) {
  const size_t n = S.pos.size() / 9;
  S.tangent.assign(9 * n, 0.0f);
  for (size_t f = 0; f < n; ++f) {
    const float* p = &S.pos[9 * f];
    const float* t = &S.uv[6 * f];
    const float v[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]}, w[3] = {p[6] - p[0], p[7] - p[1], p[8] - p[2]};
    float sx = t[2] - t[0], sy = t[3] - t[1], tx = t[4] - t[0], ty = t[5] - t[1];
    const float dir = (tx * sy - ty * sx) < 0.0f ? -1.0f : 1.0f;
    if (sx * ty == sy * tx) sx = 0.0f, sy = 1.0f, tx = 1.0f, ty = 0.0f;
    const float tg[3] = {(w[0] * sy - v[0] * ty) * dir, (w[1] * sy - v[1] * ty) * dir, (w[2] * sy - v[2] * ty) * dir};
    for (int c = 0; c < 3; ++c) {
      const float* nn = &S.normal[9 * f + 3 * c];
      const float d = tg[0] * nn[0] + tg[1] * nn[1] + tg[2] * nn[2];
      float l[3] = {tg[0] - nn[0] * d, tg[1] - nn[1] * d, tg[2] - nn[2] * d};
      float len = std::sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
      if (!(len > 0.0f) || !std::isfinite(len)) {  // tangent parallel to the normal / zero-area face: any in-plane direction
        const float a[3] = {std::fabs(nn[0]) < 0.9f ? 1.0f : 0.0f, std::fabs(nn[0]) < 0.9f ? 0.0f : 1.0f, 0.0f};
        const float da = a[0] * nn[0] + a[1] * nn[1] + a[2] * nn[2];
        l[0] = a[0] - nn[0] * da, l[1] = a[1] - nn[1] * da, l[2] = a[2] - nn[2] * da;
        len = std::sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
        if (!(len > 0.0f)) l[0] = 1.0f, l[1] = l[2] = 0.0f, len = 1.0f;
      }
      float* o = &S.tangent[9 * f + 3 * c];
      o[0] = l[0] / len, o[1] = l[1] / len, o[2] = l[2] / len;
    }
  }
}

inline void expand_srgb(float& u) {  // Utils::expand, P/utils.cpp:209-218
  if (u <= 0.0f)
    u = 0.0f;
  else if (u >= 1.0f)
    u = 1.0f;
  else if (u <= 0.04045f)
    u /= 12.92f;
  else
    u = powf((u + 0.055f) / 1.055f, 2.4f);
}

inline std::string dir_of(const std::string& path) {
  const size_t p = path.find_last_of("/\\");
  return p == std::string::npos ? std::string() : path.substr(0, p + 1);
}
inline std::string rest_of_line(std::istringstream& ls) {
  std::string r;
  std::getline(ls, r);
  const size_t a = r.find_first_not_of(" \t");
  const size_t b = r.find_last_not_of(" \t\r");
  return a == std::string::npos ? std::string() : r.substr(a, b - a + 1);
}

inline bool load_mtl(const std::string& path, bool gamma, Scene& sc, std::map<std::string, uint32_t>& index, std::string& err) {
  std::ifstream f(path);
  if (!f) {
    err = "cannot open material library " + path;
    return false;
  }
  std::string line;
  int cur = -1;
  float ka[3] = {0, 0, 0};
  auto finish = [&]() {
    if (cur < 0) return;
    if (gamma) {
      for (int k = 0; k < 3; ++k) expand_srgb(sc.materials[cur].diffuse[k]), expand_srgb(sc.materials[cur].specular[k]);
      for (int k = 0; k < 3; ++k) expand_srgb(ka[k]);
    }
  };
  while (std::getline(f, line)) {
    std::istringstream ls(line);
    std::string key;
    if (!(ls >> key) || key[0] == '#') continue;
    if (key == "newmtl") {
      finish();
      RbMaterial m{};
      m.type = RB_MAT_NORMAL;  // clearcoat == 0 -> MaterialNormal, P/ModelLoader.cpp:57-58
      m.diffuse[0] = m.diffuse[1] = m.diffuse[2] = 0.6f;  // ASSIMP ObjFile::Material defaults
      m.ior = 1.0f;
      const std::string name = rest_of_line(ls);
      cur = (int)sc.materials.size();
      index[name] = (uint32_t)cur;
      sc.materials.push_back(m);
      sc.material_names.push_back(name);
      sc.map_kd.emplace_back(), sc.map_ks.emplace_back(), sc.map_ns.emplace_back(), sc.map_kn.emplace_back();
      ka[0] = ka[1] = ka[2] = 0.0f;
      continue;
    }
    if (cur < 0) continue;
    RbMaterial& m = sc.materials[cur];
    auto read3 = [&](float* d) {
      float a = 0, b = 0, c = 0;
      if (ls >> a) {
        if (!(ls >> b >> c)) b = c = a;  // "Kd g": grey
        d[0] = a, d[1] = b, d[2] = c;
      }
    };
    if (key == "Kd")
      read3(m.diffuse);
    else if (key == "Ks")
      read3(m.specular);
    else if (key == "Ka")
      read3(ka);
    else if (key == "Ke")
      read3(m.emission);
    else if (key == "Ns")
      ls >> m.shininess;
    else if (key == "Ni")
      ls >> m.ior;
    else if (key == "Pc") {
      float pc = 0;
      ls >> pc;
      // `clearcoat == k` float comparisons of the reference, :57-70
      m.type = pc == 0 ? RB_MAT_NORMAL : pc == 1 ? RB_MAT_LAMBERT : pc == 2 ? RB_MAT_PHONG : pc == 3 ? RB_MAT_MIRROR
               : pc == 4 ? RB_MAT_DIELECTRIC : pc == 5 ? RB_MAT_DIELECTRIC_TRANSPARENT : RB_MAT_UNSUPPORTED;
    } else if (key == "map_Kd")
      sc.map_kd[cur] = rest_of_line(ls);
    else if (key == "map_Ks")
      sc.map_ks[cur] = rest_of_line(ls);
    else if (key == "map_Ns")
      sc.map_ns[cur] = rest_of_line(ls);
    else if (key == "map_Kn" || key == "norm" || key == "map_bump" || key == "map_Bump" || key == "bump")
      sc.map_kn[cur] = rest_of_line(ls);
  }
  finish();
  return true;
}

// one "v/vt/vn" reference of a face; indices are 1-based, negative = relative to the end
inline bool parse_ref(const std::string& tok, long nv, long nt, long nn, long* v, long* t, long* n) {
  *v = *t = *n = 0;
  const char* s = tok.c_str();
  char* e = nullptr;
  *v = strtol(s, &e, 10);
  if (e == s) return false;
  if (*e == '/') {
    s = e + 1;
    if (*s != '/') {
      *t = strtol(s, &e, 10);
    } else {
      e = const_cast<char*>(s);
    }
    if (*e == '/') {
      s = e + 1;
      *n = strtol(s, &e, 10);
    }
  }
  if (*v < 0) *v = nv + 1 + *v;
  if (*t < 0) *t = nt + 1 + *t;
  if (*n < 0) *n = nn + 1 + *n;
  return *v >= 1 && *v <= nv && *t >= 0 && *t <= nt && *n >= 0 && *n <= nn;
}

inline bool load_obj(const std::string& path, bool gamma, Scene& sc, std::string& err) {
  std::ifstream f(path);
  if (!f) {
    err = "cannot open " + path;
    return false;
  }
  const std::string dir = dir_of(path);
  std::vector<float> P, N, T;
  std::map<std::string, uint32_t> mat_index;
  std::map<uint32_t, size_t> surface_of;  // material -> surface (order of first use)
  int cur_mat = -1;
  std::string line;
  long line_no = 0;
  while (std::getline(f, line)) {
    ++line_no;
    std::istringstream ls(line);
    std::string key;
    if (!(ls >> key) || key[0] == '#') continue;
    if (key == "v") {
      float x = 0, y = 0, z = 0;
      ls >> x >> y >> z;
      P.insert(P.end(), {x, y, z});
    } else if (key == "vn") {
      float x = 0, y = 0, z = 0;
      ls >> x >> y >> z;
      N.insert(N.end(), {x, y, z});
    } else if (key == "vt") {
      float u = 0, v = 0;
      ls >> u >> v;
      T.insert(T.end(), {u, v});
    } else if (key == "mtllib") {
      if (!load_mtl(dir + rest_of_line(ls), gamma, sc, mat_index, err)) return false;
    } else if (key == "usemtl") {
      const std::string name = rest_of_line(ls);
      auto it = mat_index.find(name);
      if (it == mat_index.end()) {
        err = path + ":" + std::to_string(line_no) + ": unknown material '" + name + "'";
        return false;
      }
      cur_mat = (int)it->second;
    } else if (key == "f") {
      if (cur_mat < 0) {
        err = path + ":" + std::to_string(line_no) + ": face without a material (the reference indexes materials with mMaterialIndex - 1)";
        return false;
      }
      long v[64], t[64], n[64];
      int k = 0;
      std::string tok;
      while (k < 64 && (ls >> tok)) {
        if (!parse_ref(tok, (long)P.size() / 3, (long)T.size() / 2, (long)N.size() / 3, &v[k], &t[k], &n[k])) {
          err = path + ":" + std::to_string(line_no) + ": bad face element '" + tok + "'";
          return false;
        }
        ++k;
      }
      if (k < 3) continue;  // points and lines are dropped (the path renders triangles only)
      auto sit = surface_of.find((uint32_t)cur_mat);
      if (sit == surface_of.end()) {
        sit = surface_of.emplace((uint32_t)cur_mat, sc.surfaces.size()).first;
        sc.surfaces.emplace_back();
        sc.surfaces.back().material = (uint32_t)cur_mat;
      }
      Surface& S = sc.surfaces[sit->second];
      for (int i = 1; i + 1 < k; ++i) {  // fan
        const int idx[3] = {0, i, i + 1};
        const float* p[3] = {&P[3 * (v[0] - 1)], &P[3 * (v[i] - 1)], &P[3 * (v[i + 1] - 1)]};
        float fn[3] = {0, 0, 0};
        if (n[0] == 0 || n[i] == 0 || n[i + 1] == 0) {
          const float ax = p[1][0] - p[0][0], ay = p[1][1] - p[0][1], az = p[1][2] - p[0][2];
          const float bx = p[2][0] - p[0][0], by = p[2][1] - p[0][1], bz = p[2][2] - p[0][2];
          fn[0] = ay * bz - az * by, fn[1] = az * bx - ax * bz, fn[2] = ax * by - ay * bx;
          const float l = std::sqrt(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
          if (l > 0) fn[0] /= l, fn[1] /= l, fn[2] /= l;
        }
        for (int c = 0; c < 3; ++c) {
          const int j = idx[c];
          S.pos.insert(S.pos.end(), p[c], p[c] + 3);
          if (n[j] != 0)
            S.normal.insert(S.normal.end(), &N[3 * (n[j] - 1)], &N[3 * (n[j] - 1)] + 3);
          else
            S.normal.insert(S.normal.end(), fn, fn + 3);
          if (t[j] != 0)
            S.uv.insert(S.uv.end(), &T[2 * (t[j] - 1)], &T[2 * (t[j] - 1)] + 2);
          else
            S.uv.insert(S.uv.end(), {0.0f, 0.0f});
        }
        sc.n_triangles++;
      }
    }
  }
  if (sc.surfaces.empty()) {
    err = path + ": no triangles";
    return false;
  }
  bool any_normal_map = false;
  for (const std::string& s : sc.map_kn) any_normal_map = any_normal_map || !s.empty();
  if (any_normal_map)
    for (Surface& S : sc.surfaces) calc_tangents(S);
  sc.abi_surfaces.resize(sc.surfaces.size());
  for (size_t i = 0; i < sc.surfaces.size(); ++i) {
    RbSurface& a = sc.abi_surfaces[i];
    a.n_tris = (uint32_t)(sc.surfaces[i].pos.size() / 9);
    a.material = sc.surfaces[i].material;
    a.pos = sc.surfaces[i].pos.data();
    a.normal = sc.surfaces[i].normal.data();
    a.uv = sc.surfaces[i].uv.data();
    a.tangent = sc.surfaces[i].tangent.empty() ? nullptr : sc.surfaces[i].tangent.data();
  }
  sc.desc.n_surfaces = (uint32_t)sc.abi_surfaces.size();
  sc.desc.surfaces = sc.abi_surfaces.data();
  sc.desc.n_materials = (uint32_t)sc.materials.size();
  sc.desc.materials = sc.materials.data();
  return true;
}

}  // namespace rbobj
#endif
