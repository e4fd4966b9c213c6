"""SURVEY §8a rows a5 / a26 — the sky: with RenderParams::useSkybox (the reference's default) a primary ray that misses
writes scene.getSkybox().getTexel(ray.getDir()) into the G-buffer emission (P/ReSTIRIntegrator.cpp:231); SphericalMap
(P/SphericalMap.cpp:10-14) maps the direction to (0.5 + 0.5 atan2f(y, x) / pi, 1 - acos(z) / pi) and reads a BILINEAR,
CLAMP_TO_EDGE Texture. rb_set_sky replaces Scene::setSkybox.

Pin: tests/golden/ref_sky_golden.npz was made by the reference's OWN SphericalMap.cpp / Texture.cpp / ReSTIRIntegrator.cpp
(compiled in place by oracle/ref_shim, tests/golden/make_sky_golden.py); the oracle in the reference's determinism domain
(libm's atan2f / acosf) reproduces frames, G-buffers and reservoirs bit for bit. In the product's domain atan2f / acosf
are det_math.h's (within 1 ulp of libm, tests below) and the kernel bodies (host emulation; CUDA in the gpu tier) are
bit-identical to the oracle.
(File name: sorts after the other test files on purpose, so that the newest GPU tests run last under `-x`.)"""
import ctypes as C
import math
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

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sky_golden.npz"), allow_pickle=False)
PARAMS = dict(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def cam_from(arr):
    cam = abi.RbCamera()
    C.memmove(C.byref(cam), np.ascontiguousarray(arr, dtype=np.float32).ctypes.data, C.sizeof(cam))
    return cam


def _ulps(a, b):
    a = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_det_atan2_acos_against_libm():
    L = ob.lib()
    rng = np.random.default_rng(4)
    y = (rng.random(30000) * 2 - 1).astype(np.float32)
    x = (rng.random(30000) * 2 - 1).astype(np.float32)
    y[::7] *= np.float32(1e-4)
    x[::5] *= np.float32(1e-3)
    got = np.array([L.orc_dm_atan2(float(a), float(b)) for a, b in zip(y, x)], dtype=np.float32)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32)
    u = _ulps(got, ref)
    assert u.max() <= 1 and (u > 0).mean() < 1e-3
    got = np.array([L.orc_dm_acos(float(a)) for a in x], dtype=np.float32)
    ref = np.arccos(x.astype(np.float64)).astype(np.float32)
    u = _ulps(got, ref)
    assert u.max() <= 1 and (u > 0).mean() < 1e-3
    assert L.orc_dm_atan2(0.0, 1.0) == 0.0 and L.orc_dm_atan2(0.0, -1.0) == np.float32(math.pi)
    assert L.orc_dm_atan2(-0.0, -1.0) == -np.float32(math.pi) and L.orc_dm_atan2(1.0, 0.0) == np.float32(math.pi / 2)
    assert L.orc_dm_atan2(-1.0, 0.0) == -np.float32(math.pi / 2) and L.orc_dm_atan2(0.0, 0.0) == 0.0
    assert L.orc_dm_acos(1.0) == 0.0 and L.orc_dm_acos(-1.0) == np.float32(math.pi) and L.orc_dm_acos(0.0) == np.float32(math.pi / 2)
    assert math.isnan(L.orc_dm_acos(1.0000001))


def legacy_oracle(w, h, sc, sky):
    o = ob.Oracle(w, h, seed=123, rng=ob.RNG_LEGACY, math=ob.MATH_LIBM, tracer=ob.TRACER_BRUTE, cache_iim=0)
    o.upload_scene(sc)
    o.set_sky(sky)
    o.set_params(abi.default_params(lightSampler=abi.LS_CDF, useSkybox=1, **PARAMS))
    return o


def test_oracle_reproduces_reference_sky_golden_frames():
    w, h, n = int(GOLD["W"]), int(GOLD["H"]), int(GOLD["FRAMES"])
    o = legacy_oracle(w, h, tf.textured_scene(), tf.sky_arrays()[0])
    for f in range(n):
        img = o.render_frame(cam_from(GOLD[f"f{f}_cam"]), f)
        assert np.array_equal(bits(img), bits(GOLD[f"f{f}_frame"])), f"frame {f}"
        check_against(o, img, GOLD[f"f{f}_res"], GOLD[f"f{f}_gbuf"])
    assert np.array_equal(bits(o.render_mis_frame(cam_from(GOLD[f"f{n - 1}_cam"]), n)), bits(GOLD["mis_frame"]))
    # the sky is really in use: hundreds of miss pixels per frame with many distinct colours, and they are what the
    # frame shows there (no reservoir sample on an emissive G-buffer element)
    for f in range(n):
        g = GOLD[f"f{f}_gbuf"]
        miss = g[..., 16] == 0
        assert miss.sum() > 500
        assert len(np.unique(g[miss][:, 12:15], axis=0)) > 300
        lit = miss & (g[..., 12:15].sum(-1) > 0)
        assert np.array_equal(bits(GOLD[f"f{f}_frame"][lit]), bits(g[lit][:, 12:15]))


@pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (needs the reference checkout at build time)")
def test_live_reference_sky_8bit_and_switching_off():
    sc = tf.textured_scene()
    w, h = 40, 24
    sky8 = tf.sky_arrays()[1]
    p = abi.default_params(lightSampler=abi.LS_CDF, useSkybox=1, **PARAMS)
    ref = rb.Reference(w, h, sc)
    ref.set_sky(sky8)
    ref.set_params(p)
    o = legacy_oracle(w, h, sc, sky8)
    for f in range(3):
        cam = ref.camera(60.0, *tf.sky_camera_path(f))
        a = o.render_frame(cam, f)
        assert np.array_equal(bits(ref.produce_restir()), bits(a)), f"frame {f}"
        check_against(o, a, ref.reservoirs(), ref.gbuffer())
    # useSkybox off again: bgColor, with the sky still attached
    p.useSkybox = 0
    ref.set_params(p)
    o.set_params(p)
    cam = ref.camera(60.0, *tf.sky_camera_path(1))
    assert np.array_equal(bits(ref.produce_restir()), bits(o.render_frame(cam, 3)))


def test_use_skybox_without_a_sky_is_refused():
    e = eb.Emu(32, 24, seed=1)
    o = ob.Oracle(32, 24, seed=1, tracer=ob.TRACER_BRUTE)
    for r in (e, o):
        r.upload_scene(tf.textured_scene())
        with pytest.raises(AssertionError):
            r.set_params(abi.default_params(useSkybox=1))


def test_kernel_bodies_match_oracle_with_sky():
    sc = tf.textured_scene()
    w, h = 96, 64
    for sky in tf.sky_arrays():
        p = abi.default_params(doVisibilityPass=1, lightSampler=abi.LS_ALIAS, wavefront=1, useSkybox=1, **PARAMS)
        e = eb.Emu(w, h, seed=3)
        o = ob.Oracle(w, h, seed=3, tracer=ob.TRACER_BRUTE)
        for r in (e, o):
            r.upload_scene(sc)
            r.set_sky(sky)
            r.set_params(p)
            r.set_textures(tf.texel_arrays(), tf.SLOTS, tf.N_MATERIALS)
        for f in range(3):
            cam = Camera(w, h, 60, *tf.sky_camera_path(f))
            a, b = e.render_frame(cam, f), o.render_frame(cam, f)
            assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
        for buf in (abi.BUF_GBUF_EMISSION, abi.BUF_GBUF_POS_DEPTH, abi.BUF_RES_LIGHT_IDX):
            assert np.array_equal(bits(e.readback(buf)), bits(o.readback(buf))), buf
        assert np.array_equal(bits(e.render_mis_frame(cam, 9)), bits(o.render_mis_frame(cam, 9)))
    # the sky survives a scene upload; removing it needs useSkybox = 0 first
    e.upload_scene(sc)
    e.set_params(p)
    a = e.render_frame(cam, 0)
    em = e.readback(abi.BUF_GBUF_EMISSION)[..., :3]
    miss = e.readback(abi.BUF_GBUF_POS_DEPTH)[..., 3] == 0
    assert miss.sum() > 1000 and len(np.unique(em[miss], axis=0)) > 300


def test_det_and_libm_domains_agree_on_the_sky():
    """same counter RNG, det_math vs libm: the sky texels of the two domains differ by at most a last-bit coordinate
    (this container's glibc atan2f / acosf are not correctly rounded for 16 % / 8 % of arguments; det_math's are)"""
    sc = tf.textured_scene()
    w, h = 64, 40
    p = abi.default_params(lightSampler=abi.LS_ALIAS, useSkybox=1, M_Area=1, M_Brdf=0)
    ems = []
    for m in (ob.MATH_DET, ob.MATH_LIBM):
        o = ob.Oracle(w, h, seed=3, math=m, tracer=ob.TRACER_BRUTE)
        o.upload_scene(sc)
        o.set_sky(tf.sky_arrays()[0])
        o.set_params(p)
        o.render_frame(Camera(w, h, 60, *tf.sky_camera_path(0)), 0)
        ems.append(o.readback(abi.BUF_GBUF_EMISSION)[..., :3])
    assert np.allclose(ems[0], ems[1], rtol=0, atol=2e-5)
    assert (ems[0] != ems[1]).any(-1).mean() < 0.25


def test_bands_with_sky_are_band_count_invariant():
    sc = tf.textured_scene()
    w, h = 64, 48
    p = abi.default_params(doVisibilityPass=1, lightSampler=abi.LS_ALIAS, wavefront=1, useSkybox=1, **PARAMS)
    full = eb.Emu(w, h, seed=5)
    parts = make_bands(eb.Emu, w, h, 3, seed=5)
    for r in [full] + parts:
        r.upload_scene(sc)
        r.set_sky(tf.sky_arrays()[0])
        r.set_params(p)
    for f in range(3):
        cam = Camera(w, h, 60, *tf.sky_camera_path(f))  # large camera jumps: reprojections leave the band
        a = full.render_frame(cam, f)
        b = render_banded(parts, cam, f, p)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}"


@pytest.mark.gpu
def test_gpu_sky_matches_oracle_bit_for_bit(gpu):
    from restir_embree_b200.renderer import Renderer, RestirError
    sc = tf.textured_scene()
    w, h = 160, 96
    p = abi.default_params(doVisibilityPass=1, lightSampler=abi.LS_ALIAS, wavefront=1, useSkybox=1, **PARAMS)
    for sky in tf.sky_arrays():
        o = ob.Oracle(w, h, seed=3, tracer=ob.TRACER_BRUTE)
        o.upload_scene(sc)
        o.set_sky(sky)
        o.set_params(p)
        with Renderer(w, h, seed=3) as r:
            r.upload_scene(sc)
            with pytest.raises(RestirError, match="rb_set_sky"):
                r.set_params(p)
            r.set_sky(sky)
            r.set_params(p)
            for f in range(3):
                cam = Camera(w, h, 60, *tf.sky_camera_path(f))
                a, b = r.render_frame(cam, f), o.render_frame(cam, f)
                assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
            for buf in (abi.BUF_GBUF_EMISSION, abi.BUF_GBUF_POS_DEPTH, abi.BUF_RES_LIGHT_IDX):
                assert np.array_equal(bits(r.readback(buf)), bits(o.readback(buf))), buf
            assert np.array_equal(bits(r.render_mis_frame(cam, 9)), bits(o.render_mis_frame(cam, 9)))
            with pytest.raises(RestirError, match="useSkybox"):
                r.set_sky(None)
            p0 = abi.default_params(**PARAMS)
            r.set_params(p0)
            r.set_sky(None)
