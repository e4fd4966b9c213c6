"""SURVEY §8f N2 — the reference's ground-truth estimator (one-sample MIS direct lighting: NEEPathIntegrator with DI only
around DirectMISIntegrator) on the same boundary (rb_render_mis_frame).

Pin: tests/golden/ref_mis_golden.npz holds frames made by the reference's OWN DirectMISIntegrator.cpp / MaterialPhong.cpp
(compiled in place by oracle/ref_shim, tests/golden/make_mis_golden.py); the oracle in the reference's determinism domain
(serial mt19937{123}, libm) reproduces them bit for bit. The product's kernel body (host emulation here, the CUDA kernel in
the gpu tier) is then bit-identical to the oracle in the counter-RNG / det_math domain, and — being unbiased — its running
mean is what the ReSTIR image is compared with."""
import ctypes as C
import os

import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
import ref_binding as rb
from restir_embree_b200 import Camera, abi, scenes

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_mis_golden.npz"), allow_pickle=False)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def cam_from(arr):
    cam = abi.RbCamera()
    C.memmove(C.byref(cam), np.ascontiguousarray(arr, dtype=np.float32).ctypes.data, C.sizeof(cam))
    return cam


def test_oracle_reproduces_reference_mis_golden_frames():
    w, h, n = int(GOLD["W"]), int(GOLD["H"]), int(GOLD["FRAMES"])
    sc = scenes.scene_config("tiny")
    o = ob.Oracle(w, h, seed=123, rng=ob.RNG_LEGACY, math=ob.MATH_LIBM, tracer=ob.TRACER_BRUTE, cache_iim=0)
    o.upload_scene(sc)
    o.set_params(abi.default_params(lightSampler=abi.LS_CDF))
    for f in range(n):
        img = o.render_mis_frame(cam_from(GOLD[f"f{f}_cam"]), f)
        ref = GOLD[f"f{f}_frame"]
        assert np.array_equal(bits(img), bits(ref)), f"frame {f}: {(img != ref).any(-1).sum()} px differ"
        assert float(ref.mean()) > 0.1  # a lit image, not a black one


@pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (needs the reference checkout at build time)")
def test_live_reference_mis_on_another_scene():
    sc = scenes.scene_config("small")
    w, h = 40, 24
    ref = rb.Reference(w, h, sc)
    p = abi.default_params(lightSampler=abi.LS_CDF)
    ref.set_params(p)
    o = ob.Oracle(w, h, seed=123, rng=ob.RNG_LEGACY, math=ob.MATH_LIBM, tracer=ob.TRACER_BVH2, cache_iim=0)
    o.upload_scene(sc)
    o.set_params(p)
    for f in range(3):
        cam = ref.camera(55.0, (4.2 + 0.1 * f, -4.4, 1.8), (0, 0, 1.0))
        assert np.array_equal(bits(ref.produce_mis()), bits(o.render_mis_frame(cam, f))), f"frame {f}"


@pytest.mark.parametrize("sampler", [abi.LS_ALIAS, abi.LS_CDF])
def test_kernel_body_matches_oracle_bit_for_bit(sampler):
    """k_gbuffer + k_mis_direct bodies (host emulation) vs the oracle: counter RNG, det_math, all technique subsets;
    a Lambert material among the Phong ones (its evaluateLightingGI divides by pi, draws no lobe selector)."""
    sc = scenes.scene_config("tiny")
    sc.materials[1]["type"] = abi.MAT_LAMBERT
    w, h = 96, 64
    p = abi.default_params(lightSampler=sampler)
    e = eb.Emu(w, h, seed=7)
    e.upload_scene(sc)
    e.set_params(p)
    o = ob.Oracle(w, h, seed=7, tracer=ob.TRACER_BRUTE)
    o.upload_scene(sc)
    o.set_params(p)
    for f in range(2):
        cam = Camera(w, h, 60, (2.2 + 0.05 * f, -2.4, 1.4), (0, 0, 1.0))
        for tech in (3, 1, 2, 0):
            a, b = e.render_mis_frame(cam, f, tech), o.render_mis_frame(cam, f, tech)
            assert np.array_equal(bits(a), bits(b)), f"frame {f} techniques {tech}: {(a != b).any(-1).sum()} px differ"
    # one sample of each technique alone is the full estimator's two halves: MIS weights sum to one
    full = o.render_mis_frame(cam, 5, 3)
    assert np.isfinite(full).all() and (full >= 0).all()


def test_mis_leaves_the_restir_state_alone():
    """A MIS frame between two ReSTIR frames must not change the second one (it renders into the G-buffer the next
    frame overwrites and touches no reservoir)."""
    sc = scenes.scene_config("tiny")
    w, h = 64, 48
    p = abi.default_params(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, lightSampler=abi.LS_ALIAS)
    cams = [Camera(w, h, 60, (2.2 + 0.05 * f, -2.4, 1.4), (0, 0, 1.0)) for f in range(3)]
    plain = eb.Emu(w, h, seed=3)
    mixed = eb.Emu(w, h, seed=3)
    for r in (plain, mixed):
        r.upload_scene(sc)
        r.set_params(p)
    for f in range(3):
        a = plain.render_frame(cams[f], f)
        mixed.render_mis_frame(cams[f], 100 + f)
        b = mixed.render_frame(cams[f], f)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}"


def _one_sided_scene():
    """Floor + two boxes' worth of quads under ONE downward-facing emitter with nothing above it: no receiver sees the
    emitter's back side. (evaluateF weights the emitter cosine with abs(), P/ReSTIRIntegrator.cpp:197, the MIS
    integrator with max(0, .), P/DirectMISIntegrator.cpp:62 — on a scene with visible emitter back sides the two
    estimators integrate different things, as they do in the reference.)"""
    from restir_embree_b200.scenes import SceneArrays, _const_normals, _grid_quads
    sc = SceneArrays()
    floor = sc.add_material(abi.MAT_PHONG, (0.6, 0.5, 0.4), (0.2, 0.2, 0.2), (0, 0, 0), 20.0)
    matte = sc.add_material(abi.MAT_LAMBERT, (0.3, 0.6, 0.7), (0, 0, 0), (0, 0, 0), 1.0)
    lamp = sc.add_material(abi.MAT_PHONG, (0.8, 0.8, 0.8), (0, 0, 0), (30.0, 25.0, 12.0), 10.0)
    t = _grid_quads((-3, -3, 0), (6, 0, 0), (0, 6, 0), 4, 4)
    sc.add_surface(t, _const_normals(t, (0, 0, 1)), floor)
    t = _grid_quads((-1.0, 0.2, 0.0), (1.2, 0, 0), (0, 0, 0.9), 2, 2)  # an upright matte panel casting a shadow
    sc.add_surface(t, _const_normals(t, (0, -1, 0)), matte)
    t = _grid_quads((-0.4, -0.6, 2.5), (0.8, 0, 0), (0, 0.8, 0), 1, 1)
    sc.add_surface(t, _const_normals(t, (0, 0, -1)), lamp)
    sc.meta = dict(center=(0.0, 0.0, 0.5))
    return sc


def test_full_estimator_is_the_sum_of_its_two_techniques():
    """calculateDirectLighting adds the BRDF-sample and the light-sample contributions (P/DirectMISIntegrator.cpp:24-30),
    each already weighted by the power heuristic: with the same draws, techniques 1 and 2 alone add up to 3, exactly."""
    sc = scenes.scene_config("tiny")
    w, h = 64, 40
    o = ob.Oracle(w, h, seed=11, tracer=ob.TRACER_BRUTE)
    o.upload_scene(sc)
    o.set_params(abi.default_params(lightSampler=abi.LS_ALIAS))
    cam = Camera(w, h, 60, (2.2, -2.4, 1.4), (0, 0, 1.0))
    full, brdf, light = (o.render_mis_frame(cam, 0, t) for t in (3, 1, 2))
    # (pixels that show an emitter or the background return that colour whatever the techniques are)
    shaded = ~((full == brdf).all(-1) & (full == light).all(-1) & (full > 0).any(-1))
    assert shaded.sum() > 0.5 * w * h
    assert np.array_equal(bits(full[shaded]), bits((brdf + light)[shaded]))
    assert (brdf[shaded] > 0).any() and (light[shaded] > 0).any()


def test_restir_without_reuse_converges_to_the_mis_ground_truth():
    """Both are unbiased estimators of the same direct-lighting integral (RIS with M candidates and MIS weights is
    unbiased; temporal reuse is what biases the reference, S/temporal_32a_1b_5000it.png.txt): their running means
    approach each other as fast as two seeds of the MIS estimator approach each other."""
    sc = _one_sided_scene()
    w, h, n = 32, 20, 256
    cam = Camera(w, h, 60, (3.0, -3.4, 2.0), (0, 0, 0.3))

    def mean_of(seed, mis):
        o = ob.Oracle(w, h, seed=seed, tracer=ob.TRACER_BRUTE)
        o.upload_scene(sc)
        o.set_params(abi.default_params(M_Area=8, M_Brdf=1, lightSampler=abi.LS_ALIAS))
        s = np.zeros((h, w, 3), dtype=np.float64)
        for f in range(n):
            s += o.render_mis_frame(cam, f) if mis else o.render_frame(cam, f)
        return s / n

    mis_a, mis_b, ris = mean_of(1, True), mean_of(2, True), mean_of(1, False)
    assert mis_a.mean() > 0.05
    assert abs(mis_a.mean() - ris.mean()) / mis_a.mean() < 0.02
    assert ob.relmse(ris, mis_a) < 3 * ob.relmse(mis_b, mis_a) + 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("sampler", [abi.LS_ALIAS, abi.LS_CDF])
def test_gpu_mis_matches_oracle_bit_for_bit(gpu, sampler):
    from restir_embree_b200.renderer import Renderer
    sc = scenes.scene_config("small")
    sc.materials[1]["type"] = abi.MAT_LAMBERT
    w, h = 160, 96
    p = abi.default_params(lightSampler=sampler)
    o = ob.Oracle(w, h, seed=7, tracer=ob.TRACER_BVH2)
    o.upload_scene(sc)
    o.set_params(p)
    with Renderer(w, h, seed=7) as r:
        r.upload_scene(sc)
        r.set_params(p)
        for f in range(2):
            cam = Camera(w, h, 60, (4.2 + 0.1 * f, -4.4, 1.8), (0, 0, 1.0))
            for tech in (3, 1, 2):
                a, b = r.render_mis_frame(cam, f, tech), o.render_mis_frame(cam, f, tech)
                assert np.array_equal(bits(a), bits(b)), f"frame {f} techniques {tech}: {(a != b).any(-1).sum()} px differ"


@pytest.mark.gpu
def test_gpu_mis_between_restir_frames_and_accumulation(gpu):
    """MIS frames interleaved with (pipelined) ReSTIR frames leave those untouched; rb_accumulate_display converges
    the MIS frames like the reference's Producer loop."""
    from restir_embree_b200.renderer import Renderer
    sc = scenes.scene_config("small")
    w, h = 160, 96
    p = abi.default_params(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, lightSampler=abi.LS_ALIAS)
    cams = [Camera(w, h, 60, (4.2 + 0.1 * f, -4.4, 1.8), (0, 0, 1.0)) for f in range(4)]
    with Renderer(w, h, seed=3) as plain, Renderer(w, h, seed=3) as mixed:
        for r in (plain, mixed):
            r.upload_scene(sc)
            r.set_params(p)
        for f in range(4):
            a = plain.render_frame(cams[f], f)
            mixed.render_mis_frame(cams[f], 100 + f, fetch=False)
            b = mixed.render_frame(cams[f], f)
            assert np.array_equal(bits(a), bits(b)), f"frame {f}"
        acc = np.zeros((h, w, 3), dtype=np.float32)
        for f in range(8):
            img = mixed.render_mis_frame(cams[0], f)
            mixed.accumulate_display(f, fetch=False, want_stats=False)
            acc = acc + (img - acc) * np.float32(1.0 / (f + 1))
        got = mixed.readback(abi.BUF_ACCUMULATOR)
        assert np.allclose(got, acc, rtol=1e-5, atol=1e-6)


def test_mis_frame_on_bands_equals_the_full_frame():
    """The estimator is a per-pixel map keyed on the global pixel index: bands need no exchange at all."""
    from band_driver import make_bands
    sc = scenes.scene_config("tiny")
    w, h = 64, 48
    p = abi.default_params(lightSampler=abi.LS_ALIAS)
    cam = Camera(w, h, 60, (2.2, -2.4, 1.4), (0, 0, 1.0))
    one = eb.Emu(w, h, seed=5)
    one.upload_scene(sc)
    one.set_params(p)
    full = one.render_mis_frame(cam, 3)
    bands = make_bands(eb.Emu, w, h, 3, seed=5)
    got = np.zeros_like(full)
    for b in bands:
        b.upload_scene(sc)
        b.set_params(p)
        y0, y1 = b.band
        got[y0:y1] = b.render_mis_frame(cam, 3)[y0:y1]
    assert np.array_equal(bits(full), bits(got))
