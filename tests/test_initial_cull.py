"""The horizon pre-test of the initial pass (rb_passes.cuh: surely_below_horizon / initial_pixel's two loops).

initial_pixel skips a candidate whose light lies, bounding sphere and error margin included, under the pixel's horizon
before it computes the sample point. That is only allowed when the EXACT cull test of the loop (cosThetaI == 0 and
everything finite, so w == +0) would have held as well. The proof is in the source; here the claim is checked
empirically on the host build of the kernel bodies: in check mode a pre-culled candidate is evaluated all the same and
counted as a violation when the exact test disagrees. Frames must be identical with the check on and off, and equal to
the oracle's (which knows nothing of the pre-test) bit for bit — including on scenes made to sit on the margin: emitters
coplanar with the receiver and a hair above / below it, a scene far from the origin, a millimetre-sized and a
kilometre-sized scene, emitters whose vertex normals cannot be interpolated safely."""
import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
from restir_embree_b200 import Camera, abi, scenes
from test_emu_parity import ALL_BUFS, bits

W, H = 64, 40


def horizon_scene(scale=1.0, offset=(0.0, 0.0, 0.0), bad_normals=False):
    """A floor and a slanted plate under a ring of small emitters placed from far above to far below the receivers'
    tangent planes, many of them within a few ulp of the plane."""
    sc = abi.SceneArrays()
    grey = sc.add_material(abi.MAT_PHONG, (0.6, 0.6, 0.6), (0.2, 0.2, 0.2), (0, 0, 0), 20.0)
    lam = sc.add_material(abi.MAT_LAMBERT, (0.7, 0.5, 0.3), (0, 0, 0), (0, 0, 0), 1.0)
    emit = sc.add_material(abi.MAT_PHONG, (0.8, 0.8, 0.8), (0, 0, 0), (50.0, 40.0, 30.0), 10.0)
    off = np.asarray(offset, dtype=np.float64)

    def put(tris, normal, mat):
        tris = np.asarray(tris, dtype=np.float64) * scale + off
        sc.add_surface(tris.astype(np.float32), np.broadcast_to(np.asarray(normal, np.float32), tris.shape).copy(), mat)

    put(scenes._grid_quads((-4, -4, 0), (8, 0, 0), (0, 8, 0), 4, 4), (0, 0, 1), grey)
    # slanted plate: its tangent plane cuts through the emitter ring
    put(scenes._grid_quads((-1, -1, 0.5), (2, 0, 0.7), (0, 2, 0.3), 2, 2),
        np.cross((2, 0, 0.7), (0, 2, 0.3)) / np.linalg.norm(np.cross((2, 0, 0.7), (0, 2, 0.3))), lam)
    rng = np.random.default_rng(3)
    heights = np.concatenate([[0.0, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5, 1e-4, -1e-4, 1e-3, -1e-3, 1e-2, -1e-2],
                              rng.uniform(-1.5, 2.5, 27)])
    quads, normals = [], []
    for i, z in enumerate(heights):
        a = 2 * np.pi * i / len(heights)
        c = np.array([3.0 * np.cos(a), 3.0 * np.sin(a), z])
        s = 0.05 + 0.2 * rng.uniform()
        q = scenes._grid_quads(c - (s, s, 0), (2 * s, 0, 0), (0, 2 * s, 0), 1, 1)
        quads.append(q)
        n = np.broadcast_to(np.array([0.0, 0.0, -1.0]), q.shape).copy()
        if bad_normals and i % 3 == 0:  # opposite vertex normals: an interpolated normal can vanish -> never pre-culled
            n[:, 1] = (0.0, 0.0, 1.0)
        if bad_normals and i % 3 == 1:
            n[:, 2] = (0.0, 0.0, 0.0)
        normals.append(n)
    tris = np.concatenate(quads, 0) * scale + off
    sc.add_surface(tris.astype(np.float32), np.concatenate(normals, 0).astype(np.float32), emit)
    sc.meta = dict(center=tuple(off + np.array([0, 0, 0.5]) * scale))
    return sc


def cam_for(scale, offset, f):
    off = np.asarray(offset, dtype=np.float64)
    return Camera(W, H, 70, tuple(off + np.array([5.5 + 0.1 * f, -4.0, 2.5]) * scale), tuple(off + np.array([0, 0, 0.3]) * scale))


CASES = [
    dict(scale=1.0, offset=(0, 0, 0)),
    dict(scale=1.0, offset=(4000.0, -2500.0, 900.0)),   # far from the origin: coordinates carry only ~4 decimals
    dict(scale=1e-3, offset=(0, 0, 0)),
    dict(scale=1e3, offset=(0, 0, 0)),
    dict(scale=1.0, offset=(0, 0, 0), bad_normals=True),
]
PARAMS = [
    dict(M_Area=32, M_Brdf=1, doVisibilityPass=1, lightSampler=1, wavefront=1),  # the bench line's initial pass
    dict(M_Area=40, M_Brdf=0, lightSampler=0),  # CDF sampler, two chunks of candidates, shadow rays inline (counts rays)
    dict(M_Area=7, M_Brdf=2, doTemporalReuse=1, doSpatialReuse=1, lightSampler=1),
]


def run(sc, p, case, oracle):
    x = ob.Oracle(W, H, seed=11, tracer=ob.TRACER_BVH2) if oracle else eb.Emu(W, H, seed=11)
    x.upload_scene(sc)
    x.set_params(p)
    out = []
    for f in range(2):
        img = x.render_frame(cam_for(case["scale"], case["offset"], f), f)
        out.append([bits(img)] + [bits(x.readback(b)) for b in ALL_BUFS] + [x.counters()])
    return out


@pytest.mark.parametrize("ci", range(len(CASES)))
@pytest.mark.parametrize("pi", range(len(PARAMS)))
def test_pre_test_never_culls_what_the_exact_test_keeps(ci, pi):
    case = CASES[ci]
    sc = horizon_scene(**case)
    p = abi.default_params(**PARAMS[pi])
    eb.horizon_cull_check(0)
    plain = run(sc, p, case, oracle=False)
    seen = eb.horizon_cull_check(1)
    try:
        checked = run(sc, p, case, oracle=False)
    finally:
        stats = eb.horizon_cull_check(0)
    assert stats["violations"] == 0, stats
    assert stats["confirmed"] == stats["pre_culled"] == seen["pre_culled"]
    assert 0.05 * stats["candidates"] < stats["pre_culled"] < stats["candidates"], stats  # it does fire, and not always
    ref = run(sc, p, case, oracle=True)
    for f in range(2):
        for a, b, c in zip(plain[f][:-1], checked[f][:-1], ref[f][:-1]):
            assert np.array_equal(a, b) and np.array_equal(a, c)
        assert plain[f][-1] == checked[f][-1]
        assert plain[f][-1]["any_as_written"] == ref[f][-1]["any_as_written"]


def test_pre_test_on_the_bench_scene_removes_what_the_exact_test_removed():
    """BASELINE configs[1]'s scene (1M triangles, 10k emitters), bench camera, small image: ~60 % of the area candidates
    lie under the pixel's horizon, and the pre-test finds nearly all of them before the sample point is computed"""
    sc = scenes.scene_config("1m")
    p = abi.default_params(M_Area=32, M_Brdf=1, doVisibilityPass=1, lightSampler=1, wavefront=1)
    eb.horizon_cull_check(1)
    try:
        e = eb.Emu(96, 54, seed=7)
        e.upload_scene(sc)
        e.set_params(p)
        c = sc.meta["center"]
        e.render_frame(Camera(96, 54, 60, tuple(scenes.orbit_position(c, 10)), c), 0)
    finally:
        stats = eb.horizon_cull_check(0)
    assert stats["violations"] == 0 and stats["confirmed"] == stats["pre_culled"]
    assert stats["exact_culled"] > 0.5 * stats["candidates"], stats
    assert stats["pre_culled"] > 0.95 * stats["exact_culled"], stats


# ---- GPU tier: the same margin scenes through the C ABI ------------------------------------------------------------
GPU_PARAMS = [
    dict(M_Area=32, M_Brdf=1, doVisibilityPass=1, lightSampler=1, wavefront=1, doTemporalReuse=1, doSpatialReuse=1),
    dict(M_Area=40, M_Brdf=0, lightSampler=0, wavefront=1),  # CDF sampler, two chunks of candidates, inline shadow rays
    dict(M_Area=7, M_Brdf=2, doTemporalReuse=1, doSpatialReuse=1, lightSampler=1),  # inline kernels
]


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(CASES)))
@pytest.mark.parametrize("pi", range(len(GPU_PARAMS)))
def test_gpu_initial_pass_on_margin_scenes_matches_oracle_bit_for_bit(gpu, ci, pi):
    from restir_embree_b200.renderer import Renderer

    case = CASES[ci]
    sc = horizon_scene(**case)
    p = abi.default_params(**GPU_PARAMS[pi])
    o = ob.Oracle(W, H, seed=11, tracer=ob.TRACER_BVH2)
    o.upload_scene(sc)
    o.set_params(p)
    with Renderer(W, H, seed=11) as r:
        r.upload_scene(sc)
        r.set_params(p)
        for f in range(2):
            cam = cam_for(case["scale"], case["offset"], f)
            a = o.render_frame(cam, f)
            b, t = r.render_frame(cam, f, want_timings=True)
            assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
            for buf in ALL_BUFS:
                assert np.array_equal(bits(o.readback(buf)), bits(r.readback(buf))), (f, buf)
            assert o.counters()["any_as_written"] == t["rays_any_as_written"]
