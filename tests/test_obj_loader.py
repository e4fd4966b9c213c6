"""SURVEY §8f N3 (first half): Wavefront OBJ / MTL ingestion with the reference's conventions
(ModelLoader::loadOBJ / loadMaterials / loadScene, P/ModelLoader.cpp:41-321), through the C ABI (rb_obj_load).
Unpinned against ASSIMP (no binary in the reference checkout): what is checked is the documented convention set and that
the loaded scene goes through the same boundary as every other scene (kernel bodies = oracle, bit for bit)."""
import os

import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
from restir_embree_b200 import Camera, abi
from restir_embree_b200.renderer import RestirError, load_obj_scene

DATA = os.path.join(os.path.dirname(__file__), "data")
OBJ = os.path.join(DATA, "cornell_like.obj")


def srgb_expand(u):  # Utils::expand, P/utils.cpp:209-218
    u = np.float32(u)
    if u <= 0:
        return np.float32(0)
    if u >= 1:
        return np.float32(1)
    if u <= np.float32(0.04045):
        return np.float32(u / np.float32(12.92))
    return np.float32(np.power(np.float32((u + np.float32(0.055)) / np.float32(1.055)), np.float32(2.4)))


def test_materials_follow_the_reference_conventions(built):
    sc = load_obj_scene(OBJ)
    names = sc.meta["material_names"]
    assert names == ["floor", "matte red", "LAMP_EMITTER_mat", "untouched_defaults"]  # MTL order, names with blanks
    floor, red, lamp, dflt = sc.materials
    assert floor["type"] == abi.MAT_PHONG and red["type"] == abi.MAT_LAMBERT and lamp["type"] == abi.MAT_PHONG
    assert dflt["type"] == 6  # Pc outside 0..5 -> base Material (UNSUPPORTED)
    # Kd / Ks are expanded from sRGB (Raytracer::gammaCorrect defaults to true), Ke / Ns / Ni are not
    assert np.allclose(floor["diffuse"], [srgb_expand(0.735357)] * 3, rtol=1e-6)
    assert np.allclose(floor["specular"], [srgb_expand(0.5)] * 3, rtol=1e-6)
    assert floor["shininess"] == 20.0 and abs(floor["ior"] - 1.45) < 1e-6
    assert np.allclose(lamp["emission"], [100.0, 80.8983, 29.7784]) and lamp["shininess"] == 250.0
    assert np.allclose(red["diffuse"], [srgb_expand(0.8), srgb_expand(0.05), srgb_expand(0.05)], rtol=1e-6)
    # ASSIMP's OBJ defaults for absent keys: Kd 0.6 (expanded), everything else 0, Ni 1
    assert np.allclose(dflt["diffuse"], [srgb_expand(0.6)] * 3, rtol=1e-6) and dflt["specular"] == (0, 0, 0)
    assert dflt["ior"] == 1.0 and dflt["shininess"] == 0.0
    assert sc.meta["texture_names"][0] == ["floor_BaseColor.jpeg", "", "floor_ROUGHNESS.jpeg", "floor_NORMAL.jpeg"]
    raw = load_obj_scene(OBJ, gamma_correct=False)
    assert np.allclose(raw.materials[0]["diffuse"], [0.735357] * 3)


def test_geometry_one_surface_per_material_fan_triangulation_flat_normals(built):
    sc = load_obj_scene(OBJ)
    assert [s[2] for s in sc.surfaces] == [0, 1, 2]  # order of first use; "untouched_defaults" is never used
    floor, red, lamp = (s[0] for s in sc.surfaces)
    assert floor.shape[0] == 2 + 3 and red.shape[0] == 2 + 2 and lamp.shape[0] == 2  # quad = 2, pentagon = 3 triangles
    assert sc.n_triangles == 11 and sc.n_emissive == 2
    # fan from the first vertex: (0,1,2), (0,2,3)
    assert np.array_equal(floor[0], [[-2, -2, 0], [2, -2, 0], [2, 2, 0]]) and np.array_equal(floor[1], [[-2, -2, 0], [2, 2, 0], [-2, 2, 0]])
    # the pentagon (negative indices, no vn): generated flat normal +z, uv zero
    assert np.array_equal(floor[2][0], [3, -1, 0]) and np.allclose(sc.surfaces[0][1][2:], [0, 0, 1])
    assert np.array_equal(sc.meta["uv"][0][0], [[0, 0], [1, 0], [1, 1]]) and not sc.meta["uv"][0][2:].any()
    assert np.allclose(sc.surfaces[2][1], [0, 0, -1])  # lamp faces down (v//vn form)


def test_errors_are_reported_not_thrown_across_the_abi(built, tmp_path):
    with pytest.raises(RestirError, match="cannot open"):
        load_obj_scene(os.path.join(DATA, "missing.obj"))
    p = tmp_path / "nomat.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    with pytest.raises(RestirError, match="face without a material"):
        load_obj_scene(str(p))
    import shutil
    shutil.copy(os.path.join(DATA, "cornell_like.mtl"), tmp_path / "cornell_like.mtl")  # mtllib is relative to the OBJ
    p = tmp_path / "badref.obj"
    p.write_text("mtllib cornell_like.mtl\nusemtl floor\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")
    with pytest.raises(RestirError, match="bad face element"):
        load_obj_scene(str(p))


def test_loaded_scene_renders_identically_in_kernel_bodies_and_oracle(built):
    sc = load_obj_scene(OBJ)
    w, h = 80, 60
    p = abi.default_params(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, lightSampler=abi.LS_ALIAS)
    e = eb.Emu(w, h, seed=3)
    o = ob.Oracle(w, h, seed=3, tracer=ob.TRACER_BRUTE)
    for r in (e, o):
        r.upload_scene(sc)
        r.set_params(p)
    for f in range(2):
        cam = Camera(w, h, 60, (0.5 + 0.1 * f, -6.0, 2.0), (0.5, 0.0, 1.0))
        a, b = e.render_frame(cam, f), o.render_frame(cam, f)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"frame {f}"
        assert a.mean() > 0.01  # the lamp lights the floor
    a, b = e.render_mis_frame(cam, 0), o.render_mis_frame(cam, 0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_tangents_are_generated_when_a_material_names_a_normal_map(built, tmp_path):
    """the per-face step of ASSIMP's CalcTangentSpace: tangent along +u in model space, projected into the normal's
    plane, unit length; degenerate uv (the pentagon has none) -> the default directions, still finite and in-plane"""
    sc = load_obj_scene(OBJ)  # the floor's material has map_Kn
    tans = getattr(sc, "tangents", {})
    assert sorted(tans) == [0, 1, 2]  # every surface, since SceneDev's tangent array spans the scene
    for i, (pos, nrm, _) in enumerate(sc.surfaces):
        t = tans[i]
        assert np.isfinite(t).all()
        assert np.allclose(np.linalg.norm(t, axis=-1), 1, atol=1e-6)
        assert np.allclose((t * nrm).sum(-1), 0, atol=1e-6)
    # floor quad: u runs along +x (vt 0 0 -> 1 0 between v1 and v2)
    assert np.allclose(tans[0][:2], [1, 0, 0])
    # mirrored uv flips nothing in the tangent's defining property: it still points where u increases
    obj = tmp_path / "m.obj"
    (tmp_path / "m.mtl").write_text("newmtl a\nKd 0.5 0.5 0.5\nPc 2\nmap_Kn n.png\nnewmtl lamp\nKe 5 5 5\nPc 2\n")
    obj.write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nvn 0 0 1\nvt 1 0\nvt 0 0\nvt 0 1\nusemtl a\nf 1/1/1 2/2/1 3/3/1\n"
                   "v 0 0 2\nv 1 0 2\nv 1 1 2\nvn 0 0 -1\nusemtl lamp\nf 4//2 6//2 5//2\n")
    m = load_obj_scene(str(obj))
    assert np.allclose(m.tangents[0][0], [-1, 0, 0])
    # without a normal map in the MTL no tangents are made (and none are uploaded)
    (tmp_path / "m.mtl").write_text("newmtl a\nKd 0.5 0.5 0.5\nPc 2\nnewmtl lamp\nKe 5 5 5\nPc 2\n")
    assert not getattr(load_obj_scene(str(obj)), "tangents", {})


def test_loaded_scene_with_its_normal_map_in_kernel_bodies_and_oracle(built):
    """OBJ -> rb_obj_load tangents -> rb_set_textures(normal=...) -> frames: emulated kernel bodies = oracle"""
    sc = load_obj_scene(OBJ)
    rng = np.random.default_rng(2)
    nmap = np.empty((8, 8, 3), dtype=np.float32)
    nmap[..., :2] = 0.5 + (rng.random((8, 8, 2), dtype=np.float32) - 0.5) * 0.6
    nmap[..., 2] = 0.9
    w, h = 80, 60
    p = abi.default_params(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, lightSampler=abi.LS_ALIAS)
    e = eb.Emu(w, h, seed=3)
    o = ob.Oracle(w, h, seed=3, tracer=ob.TRACER_BRUTE)
    for r in (e, o):
        r.upload_scene(sc)
        r.set_params(p)
        r.set_textures([nmap], {0: dict(normal=0)}, len(sc.materials))
    for f in range(2):
        cam = Camera(w, h, 60, (0.5 + 0.1 * f, -6.0, 2.0), (0.5, 0.0, 1.0))
        a, b = e.render_frame(cam, f), o.render_frame(cam, f)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"frame {f}"
    n = e.readback(abi.BUF_GBUF_NORMAL_SHIN)[..., :3].reshape(-1, 3)
    assert len(np.unique(n, axis=0)) > 300  # the floor's normals follow the map


@pytest.mark.gpu
def test_gpu_renders_the_loaded_scene_like_the_oracle(gpu):
    from restir_embree_b200.renderer import Renderer
    sc = load_obj_scene(OBJ)
    w, h = 160, 120
    p = abi.default_params(M_Area=8, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, lightSampler=abi.LS_ALIAS)
    o = ob.Oracle(w, h, seed=3, tracer=ob.TRACER_BRUTE)
    o.upload_scene(sc)
    o.set_params(p)
    with Renderer(w, h, seed=3) as r:
        st = r.upload_scene(sc)
        assert st["n_triangles"] == 11 and st["n_emissive"] == 2
        r.set_params(p)
        for f in range(3):
            cam = Camera(w, h, 60, (0.5 + 0.1 * f, -6.0, 2.0), (0.5, 0.0, 1.0))
            a, b = r.render_frame(cam, f), o.render_frame(cam, f)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"frame {f}"
