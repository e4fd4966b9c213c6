"""SURVEY §8f N3, second half — textured materials in the G-buffer (rb_set_textures): Material::getDiffuseColor /
getSpecularColor / getShininess (P/material.cpp:105-134) over Texture::get_texel (bilinear, REPEAT, 1 - v flip,
P/Texture.cpp:72-107,170-194) with the hit's interpolated uv (P/Intersection.h:99-100).

Pin: tests/golden/ref_tex_golden.npz was made by the reference's OWN Texture.cpp / material.cpp / ReSTIRIntegrator.cpp
(compiled in place by oracle/ref_shim, tests/golden/make_tex_golden.py); the oracle in the reference's determinism domain
reproduces frames and G-buffer bit for bit. The product's kernel bodies (host emulation; CUDA in the gpu tier) are
bit-identical to the oracle in the counter-RNG / det_math domain."""
import ctypes as C
import os

import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
import ref_binding as rb
import tex_fixture as tf
from restir_embree_b200 import Camera, abi

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_tex_golden.npz"), allow_pickle=False)
PARAMS = dict(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def cam_from(arr):
    cam = abi.RbCamera()
    C.memmove(C.byref(cam), np.ascontiguousarray(arr, dtype=np.float32).ctypes.data, C.sizeof(cam))
    return cam


def check_gbuffer(o, gbuf):
    gd, gs, gn = (o.readback(b) for b in (abi.BUF_GBUF_DIFFUSE_IIM, abi.BUF_GBUF_SPEC_TYPE, abi.BUF_GBUF_NORMAL_SHIN))
    assert np.array_equal(bits(gbuf[..., 6:9]), bits(gd[..., :3]))    # getDiffuseColor(uv)
    assert np.array_equal(bits(gbuf[..., 9:12]), bits(gs[..., :3]))   # getSpecularColor(uv)
    assert np.array_equal(bits(gbuf[..., 15]), bits(gn[..., 3]))      # getShininess(uv): 2 / r^2 - 2


def legacy_oracle(w, h, sc):
    o = ob.Oracle(w, h, seed=123, rng=ob.RNG_LEGACY, math=ob.MATH_LIBM, tracer=ob.TRACER_BRUTE, cache_iim=0)
    o.upload_scene(sc)
    o.set_params(abi.default_params(lightSampler=abi.LS_CDF, **PARAMS))
    o.set_textures(tf.texel_arrays(), tf.SLOTS, tf.N_MATERIALS)
    return o


def test_oracle_reproduces_reference_textured_golden_frames():
    w, h, n = int(GOLD["W"]), int(GOLD["H"]), int(GOLD["FRAMES"])
    o = legacy_oracle(w, h, tf.textured_scene())
    for f in range(n):
        img = o.render_frame(cam_from(GOLD[f"f{f}_cam"]), f)
        assert np.array_equal(bits(img), bits(GOLD[f"f{f}_frame"])), f"frame {f}"
        check_gbuffer(o, GOLD[f"f{f}_gbuf"])
    assert np.array_equal(bits(o.render_mis_frame(cam_from(GOLD[f"f{n - 1}_cam"]), n)), bits(GOLD["mis_frame"]))
    # the maps are really in use: many distinct diffuse colours on a scene with three materials
    d = GOLD["f0_gbuf"][..., 6:9].reshape(-1, 3)
    assert len(np.unique(d, axis=0)) > 100


@pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (needs the reference checkout at build time)")
def test_live_reference_textures_other_slots():
    """live against the reference with another assignment: the 8-bit BGRA texture as specular map, float diffuse"""
    sc = tf.textured_scene()
    w, h = 40, 24
    slots = {0: dict(diffuse=2, specular=1), 1: dict(diffuse=0, shininess=3)}
    p = abi.default_params(lightSampler=abi.LS_CDF, **PARAMS)
    ref = rb.Reference(w, h, sc)
    ref.set_params(p)
    ref.set_textures(tf.texel_arrays(), slots, tf.N_MATERIALS)
    o = ob.Oracle(w, h, seed=123, rng=ob.RNG_LEGACY, math=ob.MATH_LIBM, tracer=ob.TRACER_BRUTE, cache_iim=0)
    o.upload_scene(sc)
    o.set_params(p)
    o.set_textures(tf.texel_arrays(), slots, tf.N_MATERIALS)
    for f in range(2):
        cam = ref.camera(60.0, *tf.camera_path(f))
        assert np.array_equal(bits(ref.produce_restir()), bits(o.render_frame(cam, f))), f"frame {f}"
        check_gbuffer(o, ref.gbuffer())


def test_kernel_bodies_match_oracle_with_textures():
    sc = tf.textured_scene()
    w, h = 96, 64
    p = abi.default_params(doVisibilityPass=1, lightSampler=abi.LS_ALIAS, wavefront=1, **PARAMS)
    e = eb.Emu(w, h, seed=3)
    o = ob.Oracle(w, h, seed=3, tracer=ob.TRACER_BRUTE)
    for r in (e, o):
        r.upload_scene(sc)
        r.set_params(p)
        r.set_textures(tf.texel_arrays(), tf.SLOTS, tf.N_MATERIALS)
    for f in range(3):
        cam = Camera(w, h, 60, *tf.camera_path(f))
        a, b = e.render_frame(cam, f), o.render_frame(cam, f)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
    for buf in (abi.BUF_GBUF_DIFFUSE_IIM, abi.BUF_GBUF_SPEC_TYPE, abi.BUF_GBUF_NORMAL_SHIN):
        assert np.array_equal(bits(e.readback(buf)), bits(o.readback(buf))), buf
    assert np.array_equal(bits(e.render_mis_frame(cam, 9)), bits(o.render_mis_frame(cam, 9)))
    # a scene re-upload drops the textures: constants again
    e.upload_scene(sc)
    e.set_params(p)
    e.render_frame(cam, 0)
    d = e.readback(abi.BUF_GBUF_DIFFUSE_IIM)[..., :3].reshape(-1, 3)
    assert len(np.unique(d, axis=0)) <= 4


@pytest.mark.gpu
def test_gpu_textures_match_oracle_bit_for_bit(gpu):
    from restir_embree_b200.renderer import Renderer
    sc = tf.textured_scene()
    w, h = 160, 96
    p = abi.default_params(doVisibilityPass=1, lightSampler=abi.LS_ALIAS, wavefront=1, **PARAMS)
    o = ob.Oracle(w, h, seed=3, tracer=ob.TRACER_BRUTE)
    o.upload_scene(sc)
    o.set_params(p)
    o.set_textures(tf.texel_arrays(), tf.SLOTS, tf.N_MATERIALS)
    with Renderer(w, h, seed=3) as r:
        r.upload_scene(sc)
        r.set_params(p)
        r.set_textures(tf.texel_arrays(), tf.SLOTS, tf.N_MATERIALS)
        for f in range(3):
            cam = Camera(w, h, 60, *tf.camera_path(f))
            a, b = r.render_frame(cam, f), o.render_frame(cam, f)
            assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
        for buf in (abi.BUF_GBUF_DIFFUSE_IIM, abi.BUF_GBUF_SPEC_TYPE, abi.BUF_GBUF_NORMAL_SHIN):
            assert np.array_equal(bits(r.readback(buf)), bits(o.readback(buf))), buf
        assert np.array_equal(bits(r.render_mis_frame(cam, 9)), bits(o.render_mis_frame(cam, 9)))


@pytest.mark.gpu
def test_gpu_set_textures_argument_checks(gpu):
    from restir_embree_b200.renderer import Renderer, RestirError
    from restir_embree_b200 import scenes
    with Renderer(64, 48, seed=1) as r:
        with pytest.raises(RestirError, match="no scene"):
            r.set_textures(tf.texel_arrays(), tf.SLOTS, tf.N_MATERIALS)
        r.upload_scene(tf.textured_scene())
        with pytest.raises(RestirError, match="per material"):
            r.set_textures(tf.texel_arrays(), tf.SLOTS, 2)
        with pytest.raises(RestirError, match="out of range"):
            r.set_textures(tf.texel_arrays(), {0: dict(diffuse=9)}, tf.N_MATERIALS)
        with pytest.raises(RestirError, match="normal maps"):
            r.set_textures(tf.texel_arrays(), {0: dict(normal=0)}, tf.N_MATERIALS)
        tiny = scenes.scene_config("tiny")  # no texture coordinates
        r.upload_scene(tiny)
        with pytest.raises(RestirError, match="no texture coordinates"):
            r.set_textures(tf.texel_arrays(), {0: dict(diffuse=0)}, len(tiny.materials))
