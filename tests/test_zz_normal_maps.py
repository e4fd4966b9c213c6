"""SURVEY §8a row a7 / §8f N3 — normal maps: where the hit's material has a texture in Material::kNormalMapSlot,
Intersection::intersectEmbree (P/Intersection.h:25-39) replaces the interpolated (already flipped) normal by
mat3(T, B, n) * (texel * 2 - 1), T = normalize(tangent - dot(tangent, n) * n), B = normalize(cross(n, T)), with the
tangent interpolated from attribute slot 3; the result is neither re-normalised nor flipped again. Every closest-hit
query of the path goes through it: the G-buffer, the BRDF-sampled candidates' emitter hits, the MIS estimator.

Pin: tests/golden/ref_nmap_golden.npz was made by the reference's OWN Intersection.h / Texture.cpp / ReSTIRIntegrator.cpp /
DirectMISIntegrator.cpp (compiled in place by oracle/ref_shim, tests/golden/make_nmap_golden.py); the oracle in the
reference's determinism domain reproduces frames, G-buffers and reservoirs bit for bit. The product's kernel bodies (host
emulation; CUDA in the gpu tier) are bit-identical to the oracle in the counter-RNG / det_math domain.
(File name: sorts after the other test files on purpose, so that the newest GPU tests run last under `-x`.)"""
import ctypes as C
import os

import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
import ref_binding as rb
import tex_fixture as tf
from band_driver import make_bands, render_banded
from restir_embree_b200 import Camera, abi
from test_ref_pin import check_against

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_nmap_golden.npz"), allow_pickle=False)
PARAMS = dict(M_Area=4, M_Brdf=2, doSpatialReuse=1, doTemporalReuse=1)
BUFS = (abi.BUF_GBUF_DIFFUSE_IIM, abi.BUF_GBUF_SPEC_TYPE, abi.BUF_GBUF_NORMAL_SHIN, abi.BUF_RES_LIGHT_IDX)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def cam_from(arr):
    cam = abi.RbCamera()
    C.memmove(C.byref(cam), np.ascontiguousarray(arr, dtype=np.float32).ctypes.data, C.sizeof(cam))
    return cam


def legacy_oracle(w, h, sc, slots):
    o = ob.Oracle(w, h, seed=123, rng=ob.RNG_LEGACY, math=ob.MATH_LIBM, tracer=ob.TRACER_BRUTE, cache_iim=0)
    o.upload_scene(sc)
    o.set_params(abi.default_params(lightSampler=abi.LS_CDF, **PARAMS))
    o.set_textures(tf.normal_map_arrays(), slots, tf.NMAP_N_MATERIALS)
    return o


def test_oracle_reproduces_reference_normal_mapped_golden_frames():
    w, h, n = int(GOLD["W"]), int(GOLD["H"]), int(GOLD["FRAMES"])
    o = legacy_oracle(w, h, tf.normal_mapped_scene(), tf.NMAP_SLOTS)
    for f in range(n):
        img = o.render_frame(cam_from(GOLD[f"f{f}_cam"]), f)
        assert np.array_equal(bits(img), bits(GOLD[f"f{f}_frame"])), f"frame {f}"
        check_against(o, img, GOLD[f"f{f}_res"], GOLD[f"f{f}_gbuf"])
    assert np.array_equal(bits(o.render_mis_frame(cam_from(GOLD[f"f{n - 1}_cam"]), n)), bits(GOLD["mis_frame"]))
    # the maps are really in use: three flat-shaded surfaces carry maps, yet there are many distinct normals, they are
    # not unit length (no re-normalisation) and some face away from the camera (no second flip)
    nrm = GOLD["f0_gbuf"][..., 3:6].reshape(-1, 3)
    hit = GOLD["f0_gbuf"][..., 16].reshape(-1) > 0
    assert len(np.unique(nrm[hit], axis=0)) > 200
    ln = np.linalg.norm(nrm[hit], axis=1)
    assert (np.abs(ln - 1) > 1e-3).mean() > 0.3


def test_normal_map_changes_the_image_and_only_where_mapped():
    """same scene, normal slots cleared: the wall (no normal map) keeps its normals, the floor's differ"""
    w, h = int(GOLD["W"]), int(GOLD["H"])
    plain = {m: {k: v for k, v in d.items() if k != "normal"} for m, d in tf.NMAP_SLOTS.items()}
    o = legacy_oracle(w, h, tf.normal_mapped_scene(), plain)
    o.render_frame(cam_from(GOLD["f0_cam"]), 0)
    n0 = o.readback(abi.BUF_GBUF_NORMAL_SHIN)[..., :3]
    n1 = GOLD["f0_gbuf"][..., 3:6]
    wall = np.all(n0 == np.float32((0, -1, 0)), axis=-1)
    floor = np.all(n0 == np.float32((0, 0, 1)), axis=-1)
    assert wall.sum() > 50 and floor.sum() > 200
    assert np.array_equal(bits(n0[wall]), bits(n1[wall]))
    assert (n0[floor] != n1[floor]).any(-1).mean() > 0.95


@pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (needs the reference checkout at build time)")
def test_live_reference_normal_maps_other_slots():
    """live against the reference with another assignment (float map on the lamp and the plate, 8-bit on the floor)"""
    sc = tf.normal_mapped_scene()
    w, h = 40, 24
    slots = {0: dict(normal=5, diffuse=1), 2: dict(normal=4), 3: dict(normal=4, specular=0)}
    p = abi.default_params(lightSampler=abi.LS_CDF, **PARAMS)
    ref = rb.Reference(w, h, sc)
    ref.set_params(p)
    ref.set_textures(tf.normal_map_arrays(), slots, tf.NMAP_N_MATERIALS)
    o = legacy_oracle(w, h, sc, slots)
    for f in range(2):
        cam = ref.camera(60.0, *tf.camera_path(f))
        a = o.render_frame(cam, f)
        assert np.array_equal(bits(ref.produce_restir()), bits(a)), f"frame {f}"
        check_against(o, a, ref.reservoirs(), ref.gbuffer())


def test_kernel_bodies_match_oracle_with_normal_maps():
    sc = tf.normal_mapped_scene()
    w, h = 96, 64
    for wavefront, params in ((1, PARAMS), (1, dict(PARAMS, M_Brdf=1)), (0, PARAMS)):
        p = abi.default_params(doVisibilityPass=1, lightSampler=abi.LS_ALIAS, wavefront=wavefront, **params)
        e = eb.Emu(w, h, seed=3)
        o = ob.Oracle(w, h, seed=3, tracer=ob.TRACER_BRUTE)
        for r in (e, o):
            r.upload_scene(sc)
            r.set_params(p)
            r.set_textures(tf.normal_map_arrays(), tf.NMAP_SLOTS, tf.NMAP_N_MATERIALS)
        for f in range(3):
            cam = Camera(w, h, 60, *tf.camera_path(f))
            a, b = e.render_frame(cam, f), o.render_frame(cam, f)
            assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
        for buf in BUFS:
            assert np.array_equal(bits(e.readback(buf)), bits(o.readback(buf))), buf
        assert np.array_equal(bits(e.render_mis_frame(cam, 9)), bits(o.render_mis_frame(cam, 9)))
    # dropping the normal slots switches the tangent path off again (sc.tri_tan == nullptr): flat normals
    plain = {m: {k: v for k, v in d.items() if k != "normal"} for m, d in tf.NMAP_SLOTS.items()}
    e.set_textures(tf.normal_map_arrays(), plain, tf.NMAP_N_MATERIALS)
    e.render_frame(cam, 0)
    n = e.readback(abi.BUF_GBUF_NORMAL_SHIN)[..., :3].reshape(-1, 3)
    assert len(np.unique(n, axis=0)) <= 16  # (the slanted plate's normal varies in its last bit with the barycentrics)


def test_bands_with_normal_maps_are_band_count_invariant():
    """the re-derived G-buffer elements of the banded temporal pass go through the same normal-map code"""
    sc = tf.normal_mapped_scene()
    w, h = 64, 48
    p = abi.default_params(doVisibilityPass=1, lightSampler=abi.LS_ALIAS, wavefront=1, **PARAMS)
    full = eb.Emu(w, h, seed=5)
    parts = make_bands(eb.Emu, w, h, 3, seed=5)
    for r in [full] + parts:
        r.upload_scene(sc)
        r.set_params(p)
        r.set_textures(tf.normal_map_arrays(), tf.NMAP_SLOTS, tf.NMAP_N_MATERIALS)
    for f in range(3):
        frm, at = tf.camera_path(f)
        cam = Camera(w, h, 60, (frm[0] + 0.4 * f, frm[1], frm[2] + 0.2 * f), at)  # vertical motion: reprojection crosses the band edge
        a = full.render_frame(cam, f)
        b = render_banded(parts, cam, f, p)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}"


def _gpu_against_oracle(params):
    from restir_embree_b200.renderer import Renderer
    sc = tf.normal_mapped_scene()
    w, h = 160, 96
    p = abi.default_params(doVisibilityPass=1, lightSampler=abi.LS_ALIAS, wavefront=1, **params)
    o = ob.Oracle(w, h, seed=3, tracer=ob.TRACER_BRUTE)
    o.upload_scene(sc)
    o.set_params(p)
    o.set_textures(tf.normal_map_arrays(), tf.NMAP_SLOTS, tf.NMAP_N_MATERIALS)
    with Renderer(w, h, seed=3) as r:
        r.upload_scene(sc)
        r.set_params(p)
        r.set_textures(tf.normal_map_arrays(), tf.NMAP_SLOTS, tf.NMAP_N_MATERIALS)
        for f in range(3):
            cam = Camera(w, h, 60, *tf.camera_path(f))
            a, b = r.render_frame(cam, f), o.render_frame(cam, f)
            assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
        for buf in BUFS:
            assert np.array_equal(bits(r.readback(buf)), bits(o.readback(buf))), buf
        assert np.array_equal(bits(r.render_mis_frame(cam, 9)), bits(o.render_mis_frame(cam, 9)))


@pytest.mark.gpu
def test_gpu_normal_maps_match_oracle_bit_for_bit(gpu):
    from restir_embree_b200.renderer import Renderer, RestirError
    _gpu_against_oracle(dict(PARAMS, M_Brdf=1))
    with Renderer(64, 48, seed=3) as r:  # a scene without tangents cannot take a normal map
        r.upload_scene(tf.textured_scene())
        with pytest.raises(RestirError, match="normal maps"):
            r.set_textures(tf.texel_arrays(), {0: dict(normal=0)}, tf.N_MATERIALS)


@pytest.mark.gpu
def test_gpu_normal_maps_two_brdf_candidates(gpu):
    """M_Brdf = 2 with the visibility pass in the wavefront schedule (k_initial_resolve_nmap with two hit slots per pixel)"""
    _gpu_against_oracle(PARAMS)
