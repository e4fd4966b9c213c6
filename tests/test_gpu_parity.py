"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Bar (BASELINE.json north_star): hit primitive ids and selected-light indices bit-exact; radiance
relMSE <= 1e-3 per frame. Because the kernels and the oracle share one arithmetic contract
(det_math.h, no FMA contraction, fixed ray/triangle operation order) the bar met here is stricter:
every frame, G-buffer plane and reservoir plane is BIT-IDENTICAL at oracle-sized inputs; at the full
BASELINE sizes the checks are sampled rows against the oracle plus size-independent properties."""
import numpy as np
import pytest

import oracle_binding as ob
from restir_embree_b200 import Camera, abi, scenes
from restir_embree_b200.renderer import Renderer, make_rays
from test_emu_parity import ALL_BUFS, CONFIGS, TWO_STEP_PARAMS, bits, coincident_emitter_scene

pytestmark = pytest.mark.gpu

W, H = 128, 72


def test_smoke(gpu):
    import __graft_entry__ as g
    g.smoke()


@pytest.fixture(scope="module")
def small():
    return scenes.scene_config("small")


@pytest.mark.parametrize("ci", range(len(CONFIGS)))
def test_gpu_frames_match_oracle_bit_for_bit(gpu, small, ci):
    p = abi.default_params(**CONFIGS[ci])
    o = ob.Oracle(W, H, seed=7, tracer=ob.TRACER_BVH2)
    o.upload_scene(small)
    o.set_params(p)
    with Renderer(W, H, seed=7) as r:
        r.upload_scene(small)
        r.set_params(p)
        for f in range(3):
            cam = Camera(W, H, 60, (4.2 + 0.15 * f, -4.4, 1.8 + 0.05 * f), (0, 0, 1.0))
            a = o.render_frame(cam, f)
            b, t = r.render_frame(cam, f, want_timings=True)
            assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
            assert ob.relmse(b, a) == 0.0
            for buf in ALL_BUFS:
                assert np.array_equal(bits(o.readback(buf)), bits(r.readback(buf))), (f, buf)
            oc = o.counters()
            assert oc["closest"] == t["rays_closest"] and oc["any_as_written"] == t["rays_any_as_written"]
            assert t["rays_any_traced"] <= t["rays_any_as_written"]


@pytest.mark.parametrize("emitter_first", [True, False])
def test_gpu_two_step_brdf_rays_with_coincident_emitter(gpu, emitter_first):
    """The emissive-only BVH + "does anything precede it" scheme of the BRDF-candidate rays against the oracle's plain
    closest-hit query, on an emitter that coincides with a non-emissive sheet (tie decided by triangle id)."""
    sc = coincident_emitter_scene(emitter_first)
    Wd, Hd = 48, 32
    p = abi.default_params(**TWO_STEP_PARAMS)
    o = ob.Oracle(Wd, Hd, seed=3, tracer=ob.TRACER_BRUTE)
    o.upload_scene(sc)
    o.set_params(p)
    with Renderer(Wd, Hd, seed=3) as r:
        r.upload_scene(sc)
        r.set_params(p)
        for fr in range(2):
            cam = Camera(Wd, Hd, 70, (2.5, -2.6 + 0.1 * fr, 1.2), (0, 0, 0.3))
            a, b = o.render_frame(cam, fr), r.render_frame(cam, fr)
            assert np.array_equal(bits(a), bits(b)), f"frame {fr}: {(a != b).any(-1).sum()} px differ"
            for buf in (abi.BUF_RES_POINT_WSUM, abi.BUF_RES_NORMAL_W, abi.BUF_RES_LIGHT_IDX):
                assert np.array_equal(bits(o.readback(buf)), bits(r.readback(buf))), (fr, buf)


def test_ray_seam_matches_brute_force(gpu):
    sc = scenes.scene_config("tiny")
    o = ob.Oracle(8, 8, tracer=ob.TRACER_BRUTE)
    o.upload_scene(sc)
    rng = np.random.default_rng(5)
    n = 50000
    rays = np.zeros(n, dtype=abi.RAY_DTYPE)
    rays["org"] = rng.uniform((-2.9, -2.9, 0.05), (2.9, 2.9, 2.9), size=(n, 3))
    d = rng.normal(size=(n, 3))
    d[:200, 0] = 0.0
    d[200:400, 1] = 0.0
    d[400:600] = np.eye(3)[rng.integers(0, 3, 200)] * rng.choice([-1.0, 1.0], size=(200, 1))
    rays["dir"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tnear"] = 0.01
    rays["tfar"] = rng.choice([3.4028235e38, 1.0, 2.5], size=n)
    rays["dir"][600:610] = np.nan
    rays["dir"][610:620] = 0.0
    with Renderer(8, 8) as r:
        r.upload_scene(sc)
        ho, hg = o.trace_closest(rays), r.trace_closest(rays)
        assert np.array_equal(ho["primID"], hg["primID"]) and np.array_equal(ho["geomID"], hg["geomID"])
        assert np.array_equal(bits(ho["t"]), bits(hg["t"]))
        assert np.array_equal(bits(ho["u"]), bits(hg["u"])) and np.array_equal(bits(ho["v"]), bits(hg["v"]))
        assert np.array_equal(o.trace_occluded(rays), r.trace_occluded(rays))
        # empty batch is a no-op
        assert r.trace_occluded(rays[:0]).shape == (0,)
        assert r.trace_closest(rays[:0]).shape == (0,)


def test_errors_are_codes_not_crashes(gpu):
    with Renderer(16, 16) as r:
        cam = Camera(16, 16, 60, (1, 1, 1), (0, 0, 0))
        with pytest.raises(Exception, match="no scene"):
            r.render_frame(cam, 0)
        with pytest.raises(Exception, match="useSkybox"):
            r.set_params(abi.default_params(useSkybox=1))
        with pytest.raises(Exception, match="not supported"):
            r.set_params(abi.default_params(spatialReuseNeighborCount=64))


@pytest.fixture(scope="module")
def scene_1m():
    return scenes.scene_config("1m")


def test_one_million_triangles_traversal_vs_oracle_bvh(gpu, scene_1m):
    """Config 4 in miniature: incoherent closest-hit and shadow rays against the 1M-triangle BVH built on
    the GPU; the oracle answers through its own CPU BVH2 (hit set == brute force)."""
    o = ob.Oracle(8, 8, tracer=ob.TRACER_BVH2)
    o.upload_scene(scene_1m)
    rng = np.random.default_rng(11)
    n = 200000
    org = rng.uniform((-9.5, -9.5, 0.1), (9.5, 9.5, 5.9), size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = make_rays(org, direction=d)
    tgt = rng.uniform((-9.5, -9.5, 5.0), (9.5, 9.5, 5.95), size=(n, 3)).astype(np.float32)
    srays = make_rays(org, target=tgt)
    with Renderer(8, 8) as r:
        st = r.upload_scene(scene_1m)
        assert st["n_triangles"] == 1_000_000 and st["n_emissive"] == 10_000 and st["bvh_depth"] < 40
        ho, hg = o.trace_closest(rays), r.trace_closest(rays)
        same = (ho["primID"] == hg["primID"]) & (ho["geomID"] == hg["geomID"])
        # north_star: ids bit-exact except where hit distances differ by < 1e-5 relative
        bad = ~same & ~(np.abs(ho["t"] - hg["t"]) <= 1e-5 * np.abs(ho["t"]))
        assert bad.sum() == 0, f"{bad.sum()} closest hits differ beyond the t-tie tolerance"
        assert same.mean() > 0.99999
        assert np.array_equal(bits(ho["t"][same]), bits(hg["t"][same]))
        oo, og = o.trace_occluded(srays), r.trace_occluded(srays)
        assert (oo != og).sum() == 0
        assert 0.05 < oo.mean() < 0.95


@pytest.mark.parametrize("wavefront", [1, 0])
def test_full_hd_frame_sampled_rows_and_properties(gpu, scene_1m, wavefront):
    """BASELINE config 2 at full size (1920x1080, 1M triangles, 10k emitters, A=32 B=1, temporal + spatial,
    visibility pass), with the wavefront schedule bench.py runs (stream -> trace -> resolve kernels) and with the
    inline kernels. The oracle renders a band of rows of the same frames; rows far enough from the band
    edge (spatial reach, reprojection) must be bit-identical. Plus size-independent properties."""
    Wf, Hf = 1920, 1080
    p = abi.default_params(M_Area=32, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1,
                           lightSampler=abi.LS_ALIAS, wavefront=wavefront)
    y0, y1, margin = 500, 548, 16
    o = ob.Oracle(Wf, Hf, seed=123, tracer=ob.TRACER_BVH2)
    o.upload_scene(scene_1m)
    o.set_params(p)
    o.set_band(y0, y1)
    center = scene_1m.meta["center"]
    with Renderer(Wf, Hf, seed=123) as r, Renderer(Wf, Hf, seed=123) as r2:
        for x in (r, r2):
            x.upload_scene(scene_1m)
            x.set_params(p)
        for f in range(2):
            cam = Camera(Wf, Hf, 55, scenes.orbit_position(center, f, step_deg=0.05), center)
            a = o.render_frame(cam, f)
            b, t = r.render_frame(cam, f, want_timings=True)
            b2 = r2.render_frame(cam, f)
            assert np.array_equal(bits(b), bits(b2)), "two handles, same inputs: frames must be identical"
            assert np.isfinite(b).all() and (b >= 0).all()
            rows = slice(y0 + margin, y1 - margin)
            assert np.array_equal(bits(a[rows]), bits(b[rows])), f"frame {f}: sampled rows differ from the oracle"
            for buf in (abi.BUF_HIT_IDS, abi.BUF_RES_LIGHT_IDX, abi.BUF_RES_NORMAL_W):
                assert np.array_equal(bits(o.readback(buf)[rows]), bits(r.readback(buf)[rows])), (f, buf)
            emissive = (r.readback(abi.BUF_GBUF_SPEC_TYPE)[..., 3].view(np.uint32) & 0x100) != 0
            assert t["rays_closest"] == Wf * Hf + int((~emissive).sum()) * p.M_Brdf
            assert ob.relmse(b[rows], a[rows]) <= 1e-3


def test_full_hd_pipelined_frames_without_sync_match_oracle(gpu, scene_1m):
    """The benchmarked path exactly as bench.py drives it: 1080p / 1M triangles / A=32 B=1 / visibility + temporal +
    spatial, wavefront kernels, EIGHT consecutive rb_render_frame_device calls with no synchronisation in between — the
    front half of frame n+1 (second stream) overlaps the back half of frame n, three G-buffers and four reservoir
    buffers rotate, per-parity queues and counters alternate — on bench.py's camera path (0.5 deg per frame). The last
    frame, its reservoirs (incl. selected-light indices) and hit ids are compared with the oracle, which renders a band
    of the same eight frames; its edges erode by the spatial reach + reprojection shift per frame (measured with the
    oracle: 33 rows after eight frames), so the middle rows are exact."""
    Wf, Hf, n_frames = 1920, 1080, 8
    p = abi.default_params(M_Area=32, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1,
                           lightSampler=abi.LS_ALIAS, wavefront=1)
    y0, y1, margin = 460, 620, 48
    o = ob.Oracle(Wf, Hf, seed=123, tracer=ob.TRACER_BVH2)
    o.upload_scene(scene_1m)
    o.set_params(p)
    o.set_band(y0, y1)
    center = scene_1m.meta["center"]
    cams = [Camera(Wf, Hf, 55, scenes.orbit_position(center, f), center) for f in range(n_frames)]
    for f in range(n_frames):
        a = o.render_frame(cams[f], f)
    with Renderer(Wf, Hf, seed=123, collect_timings=False) as r:
        r.upload_scene(scene_1m)
        r.set_params(p)
        for f in range(n_frames):
            r.render_frame_device(cams[f], f)  # no sync, no readback: frames pipeline
        b = r.readback(abi.BUF_FRAME_RGB)
        rows = slice(y0 + margin, y1 - margin)
        assert np.array_equal(bits(a[rows]), bits(b[rows])), f"{(a[rows] != b[rows]).any(-1).sum()} px of the sampled rows differ"
        for buf in (abi.BUF_HIT_IDS, abi.BUF_RES_LIGHT_IDX, abi.BUF_RES_POINT_WSUM, abi.BUF_RES_NORMAL_W, abi.BUF_RES_LI_CONF):
            assert np.array_equal(bits(o.readback(buf)[rows]), bits(r.readback(buf)[rows])), buf
        assert np.isfinite(b).all() and (b >= 0).all()
        # the same eight frames with a synchronisation after every frame (no overlap in effect): identical image
        with Renderer(Wf, Hf, seed=123) as r2:
            r2.upload_scene(scene_1m)
            r2.set_params(p)
            for f in range(n_frames):
                b2 = r2.render_frame(cams[f], f)
            assert np.array_equal(bits(b), bits(b2)), "pipelined and synchronised frame loops differ"


def test_orbit_64_frames_every_frame_bit_identical_and_converged_relmse(gpu, small):
    """BASELINE configs[4] at oracle size: 64-frame orbit, temporal + spatial reuse, wavefront kernels, frames issued
    without synchronisation (the frame is fetched with rb_readback). Every frame equals the oracle's bit for bit, hence
    the 64-frame running mean (the reference's accumulator, P/simpleguidx11.cpp:246-253) has relMSE 0 <= 1e-4."""
    Wd, Hd, n_frames = 160, 96, 64
    p = abi.default_params(M_Area=8, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1,
                           lightSampler=abi.LS_ALIAS, wavefront=1)
    o = ob.Oracle(Wd, Hd, seed=5, tracer=ob.TRACER_BVH2)
    o.upload_scene(small)
    o.set_params(p)
    acc_o = np.zeros((Hd, Wd, 3), dtype=np.float64)
    acc_g = np.zeros((Hd, Wd, 3), dtype=np.float64)
    with Renderer(Wd, Hd, seed=5, collect_timings=False) as r:
        r.upload_scene(small)
        r.set_params(p)
        for f in range(n_frames):
            cam = Camera(Wd, Hd, 55, scenes.orbit_position((0, 0, 1.0), f, radius=4.5), (0, 0, 1.0))
            a = o.render_frame(cam, f)
            r.render_frame_device(cam, f)
            b = r.readback(abi.BUF_FRAME_RGB)
            assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
            assert ob.relmse(b, a) <= 1e-3
            acc_o += a
            acc_g += b
            _, st = r.accumulate_display(f)
        assert np.array_equal(bits(o.readback(abi.BUF_RES_LIGHT_IDX)), bits(r.readback(abi.BUF_RES_LIGHT_IDX)))
        assert ob.relmse(acc_g / n_frames, acc_o / n_frames) <= 1e-4
        # the library's own accumulator (N1) holds the same running mean (float, glm::mix order) to float accuracy
        acc_lib = r.readback(abi.BUF_ACCUMULATOR)
        assert ob.relmse(acc_lib, acc_o / n_frames) <= 1e-8
    ts = o.temporal_stats()
    assert ts["merged"] > 0.5 * Wd * Hd * (n_frames - 1) * 0.5  # the orbit keeps most pixels reprojectable


def test_async_frames_into_two_host_buffers_equal_blocking_frames(gpu, small):
    """rb_render_frame_async / rb_frame_wait (the double-buffered Producer loop bench.py's e2e leg times): every frame that
    lands in the host buffers equals the frame the blocking rb_render_frame returns, which the oracle checks elsewhere."""
    import torch
    Wd, Hd, n = 256, 144, 10
    p = abi.default_params(M_Area=8, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1,
                           lightSampler=abi.LS_ALIAS, wavefront=1)
    cams = [Camera(Wd, Hd, 55, scenes.orbit_position((0, 0, 1.0), f, radius=4.5), (0, 0, 1.0)) for f in range(n)]
    with Renderer(Wd, Hd, seed=9) as r:
        r.upload_scene(small)
        r.set_params(p)
        want = [r.render_frame(cams[f], f).copy() for f in range(n)]
    bufs = [torch.empty((Hd, Wd, 3), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]
    got = []
    with Renderer(Wd, Hd, seed=9, collect_timings=False) as r:
        r.upload_scene(small)
        r.set_params(p)
        for f in range(n):
            r.render_frame_async(cams[f], f, bufs[f & 1])
            r.frame_wait(1)
            if f > 0:
                got.append(bufs[(f - 1) & 1].copy())
        r.frame_wait(0)
        got.append(bufs[(n - 1) & 1].copy())
    for f in range(n):
        assert np.array_equal(bits(want[f]), bits(got[f])), f"async frame {f} differs"


def test_cpp_host_mirror_example(gpu):
    """The C++ mirror of the reference interface (restir_embree_b200/host/restir_b200.hpp) renders through the C ABI."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "restir_embree_b200", "host", "example_main")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout
