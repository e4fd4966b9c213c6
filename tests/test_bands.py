"""Multi-GPU band decomposition: the image must be BIT-IDENTICAL for any number of bands — counter RNG
keyed on the global pixel, clamps on global bounds, reservoir halo rows exchanged before every spatial pass,
out-of-band G-buffer elements re-derived locally. CPU tier: the product's kernel bodies under the emulation
harness, plus a world_size-2 gloo run of the halo exchange. GPU tier: the same through the C ABI."""
import os
import sys

import numpy as np
import pytest

import emu_binding as eb
from band_driver import assemble, make_bands, move_boundaries, render_banded
from restir_embree_b200 import Camera, abi, scenes
from restir_embree_b200.renderer import band_rows

W, H = 96, 80


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def cams(f):
    # orbit plus a vertical pan: reprojection crosses band boundaries
    return Camera(W, H, 60, (4.2 + 0.1 * f, -4.4, 1.8 + 0.25 * f), (0, 0, 1.0 + 0.2 * f))


CFGS = [
    dict(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, lightSampler=1, wavefront=1),
    dict(M_Area=3, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, spatialPassCount=2, rejectDissimilarNeighbors=1),
    dict(M_Area=2, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=4),
]


def test_band_rows_partition():
    for h in (1080, 2160, 80, 7):
        for n in (1, 2, 3, 4, 8):
            rows = [band_rows(h, n, r) for r in range(n)]
            assert rows[0][0] == 0 and rows[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(rows[:-1], rows[1:]))


@pytest.mark.parametrize("ci", range(len(CFGS)))
@pytest.mark.parametrize("n", [2, 3, 5])
def test_emulated_bands_are_bit_identical_to_one_band(ci, n):
    sc = scenes.scene_config("small")
    p = abi.default_params(**CFGS[ci])
    one = eb.Emu(W, H, seed=5)
    one.upload_scene(sc)
    one.set_params(p)
    bands = make_bands(eb.Emu, W, H, n, seed=5)
    for b in bands:
        b.upload_scene(sc)
        b.set_params(p)
    for f in range(3):
        a = one.render_frame(cams(f), f)
        b = render_banded(bands, cams(f), f, p)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ with {n} bands"
        for buf in (abi.BUF_RES_LIGHT_IDX, abi.BUF_RES_NORMAL_W, abi.BUF_HIT_IDS):
            assert np.array_equal(bits(one.readback(buf)), bits(assemble(bands, buf))), (f, buf)


def test_emulated_bands_fast_pan_uses_the_deferred_rederivation_launch():
    """Thin bands + a fast vertical pan: many reprojections leave the rows a band holds (band + 16-row margin), so the
    second, deferred launch of the banded temporal stream pass (the one with the re-derivation code) has work; the image
    must still be bit-identical to the one-band image."""
    sc = scenes.scene_config("small")
    p = abi.default_params(**CFGS[0])
    fast = lambda f: Camera(W, H, 60, (4.2, -4.4, 1.8 + 0.6 * f), (0, 0, 1.0 + 1.1 * f))
    one = eb.Emu(W, H, seed=5)
    one.upload_scene(sc)
    one.set_params(p)
    bands = make_bands(eb.Emu, W, H, 5, seed=5)
    for b in bands:
        b.upload_scene(sc)
        b.set_params(p)
    for f in range(3):
        a = one.render_frame(fast(f), f)
        b = render_banded(bands, fast(f), f, p)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
    assert sum(b.deferred_total() for b in bands) > 0


def _moving_bounds(n, f):
    """band boundaries that wander by up to 7 rows per frame (inside the 16-row G-buffer margin)"""
    base = [band_rows(H, n, r)[0] for r in range(n)] + [H]
    shift = [0] + [((-1) ** k) * ((3 * f + 2 * k) % 8) for k in range(1, n)] + [0]
    return [b + s for b, s in zip(base, shift)]


@pytest.mark.parametrize("n", [2, 3])
def test_emulated_bands_with_moving_boundaries_are_bit_identical_to_one_band(n):
    """Load balancing: band boundaries move between frames; the rows that change owner take the last frame's
    reservoirs with them, everything else is re-derived (G-buffer margin). The image must not notice."""
    sc = scenes.scene_config("small")
    p = abi.default_params(**CFGS[0])
    one = eb.Emu(W, H, seed=5)
    one.upload_scene(sc)
    one.set_params(p)
    bands = make_bands(eb.Emu, W, H, n, seed=5)
    for b in bands:
        b.upload_scene(sc)
        b.set_params(p)
    for f in range(5):
        if f > 0:
            move_boundaries(bands, _moving_bounds(n, f))
        a = one.render_frame(cams(f), f)
        b = render_banded(bands, cams(f), f, p)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ, bands {[x.band for x in bands]}"
        assert np.array_equal(bits(one.readback(abi.BUF_RES_NORMAL_W)), bits(assemble(bands, abi.BUF_RES_NORMAL_W))), f


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(__file__))
    from dist_bands import DistBandDriver
    sc = scenes.scene_config("small")
    p = abi.default_params(**CFGS[0])
    r = eb.Emu(W, H, seed=5, band=band_rows(H, world, rank))
    r.upload_scene(sc)
    r.set_params(p)
    drv = DistBandDriver(r, rank, world)
    imgs = []
    for f in range(3):
        imgs.append(drv.render(cams(f), f, p))  # gathered on rank 0
    if rank == 0:
        q.put(np.stack(imgs))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_halo_exchange_matches_single_band():
    """world_size-2 on CPU: one process per band, halo rows moved with torch.distributed (gloo) send/recv."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    sc = scenes.scene_config("small")
    p = abi.default_params(**CFGS[0])
    one = eb.Emu(W, H, seed=5)
    one.upload_scene(sc)
    one.set_params(p)
    for f in range(3):
        a = one.render_frame(cams(f), f)
        assert np.array_equal(bits(a), bits(got[f])), f"frame {f}"


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 4])
def test_gpu_bands_are_bit_identical_to_one_band(gpu, n):
    from restir_embree_b200.renderer import Renderer
    sc = scenes.scene_config("small")
    for cfg in CFGS[:2]:
        p = abi.default_params(**cfg)
        one = Renderer(W, H, seed=5)
        one.upload_scene(sc)
        one.set_params(p)
        bands = make_bands(Renderer, W, H, n, seed=5)
        for b in bands:
            b.upload_scene(sc)
            b.set_params(p)
        for f in range(3):
            a = one.render_frame(cams(f), f)
            b = render_banded(bands, cams(f), f, p)
            assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ with {n} bands"
            assert np.array_equal(bits(one.readback(abi.BUF_RES_LIGHT_IDX)), bits(assemble(bands, abi.BUF_RES_LIGHT_IDX)))
        for r in bands + [one]:
            r.close()


@pytest.mark.gpu
def test_gpu_bands_with_moving_boundaries_are_bit_identical_to_one_band(gpu):
    from restir_embree_b200.renderer import Renderer
    sc = scenes.scene_config("small")
    p = abi.default_params(**CFGS[0])
    one = Renderer(W, H, seed=5)
    one.upload_scene(sc)
    one.set_params(p)
    bands = make_bands(Renderer, W, H, 3, seed=5)
    for b in bands:
        b.upload_scene(sc)
        b.set_params(p)
    for f in range(5):
        if f > 0:
            move_boundaries(bands, _moving_bounds(3, f))
        a = one.render_frame(cams(f), f)
        b = render_banded(bands, cams(f), f, p)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ, bands {[x.band for x in bands]}"
    for r in bands + [one]:
        r.close()


@pytest.mark.gpu
def test_gpu_bands_fast_pan_deferred_rederivation(gpu):
    """GPU twin of the fast-pan test: the bulk banded temporal stream kernel defers the pixels whose reprojection
    leaves the held rows to k_temporal_stream_deferred; image and selected lights stay bit-identical to one band."""
    from restir_embree_b200.renderer import Renderer
    sc = scenes.scene_config("small")
    p = abi.default_params(**CFGS[0])
    fast = lambda f: Camera(W, H, 60, (4.2, -4.4, 1.8 + 0.6 * f), (0, 0, 1.0 + 1.1 * f))
    one = Renderer(W, H, seed=5)
    one.upload_scene(sc)
    one.set_params(p)
    bands = make_bands(Renderer, W, H, 5, seed=5)
    for b in bands:
        b.upload_scene(sc)
        b.set_params(p)
    for f in range(3):
        a = one.render_frame(fast(f), f)
        b = render_banded(bands, fast(f), f, p)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
        assert np.array_equal(bits(one.readback(abi.BUF_RES_LIGHT_IDX)), bits(assemble(bands, abi.BUF_RES_LIGHT_IDX)))
    for r in bands + [one]:
        r.close()


def _balance(pairs, height):
    import ctypes as C
    from restir_embree_b200.renderer import load_library
    L = load_library()
    n = len(pairs)
    a = (C.c_float * (2 * n))(*[v for p in pairs for v in p])
    out = (C.c_int32 * (n + 1))()
    L.rb_debug_balance_step.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    assert L.rb_debug_balance_step(a, n, height, out) == 0
    return list(out)


def test_balancer_partition_rule(built):
    """The host arithmetic of the band balancer (every rank runs it on the same all-gathered numbers): equal costs keep
    the bands, an expensive band shrinks, a boundary moves at most 15 rows per period, repeated steps converge to the
    equal-cost partition of a piecewise-constant cost profile, thin bands are left alone."""
    H = 1080
    rows = [135] * 8
    assert _balance([(1.5, r) for r in rows], H) == [135 * i for i in range(9)]
    # rank 0 twice as cheap per row as the others: it must grow, by at most 15 rows
    b = _balance([(0.75, 135)] + [(1.5, 135)] * 7, H)
    assert 135 < b[1] <= 150 and b[0] == 0 and b[8] == H
    assert all(b[i + 1] - b[i] >= 32 for i in range(8))
    # iterate against a synthetic per-row cost profile: ceiling rows cheap, middle rows expensive
    prof = np.concatenate([np.full(300, 0.5), np.full(480, 1.6), np.full(300, 1.0)])
    bounds = [135 * i for i in range(9)]
    for _ in range(40):
        costs = [float(prof[bounds[i]:bounds[i + 1]].sum()) for i in range(8)]
        new = _balance([(c, bounds[i + 1] - bounds[i]) for i, c in enumerate(costs)], H)
        assert all(abs(n - o) <= 15 for n, o in zip(new, bounds))
        bounds = new
    costs = np.array([prof[bounds[i]:bounds[i + 1]].sum() for i in range(8)])
    assert costs.max() / costs.mean() < 1.03, (bounds, costs)
    # bands too thin to move safely: unchanged
    assert _balance([(1.0, 27), (3.0, 27), (1.0, 26)], 80) == [0, 27, 54, 80]


@pytest.mark.gpu
@pytest.mark.parametrize("n_bands", [2, 3])
def test_multi_handle_one_thread_assembles_the_frame_bit_identical(gpu, n_bands):
    """rb_multi_* (one handle, several devices, one host thread — SURVEY 8b): here all bands sit on device 0, which runs the
    same code as N GPUs (peer pointers instead of peer mappings): the frames assembled in ONE host buffer, the reservoir
    planes and the accumulator equal the single-handle ones bit for bit, with blocking and with pipelined calls."""
    import torch
    from restir_embree_b200.renderer import MultiRenderer, Renderer
    sc = scenes.scene_config("small")
    Wd, Hd, n = 160, 96, 6
    p = abi.default_params(M_Area=6, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, spatialPassCount=2,
                           lightSampler=abi.LS_ALIAS, wavefront=1)
    cams = [Camera(Wd, Hd, 55, scenes.orbit_position((0, 0, 1.0), 3 * f, radius=4.5), (0, 0, 1.0)) for f in range(n)]
    with Renderer(Wd, Hd, seed=4) as one:
        one.upload_scene(sc)
        one.set_params(p)
        want = []
        for f in range(n):
            want.append(one.render_frame(cams[f], f).copy())
            one.accumulate_display(f, want_stats=False)
        want_res = [one.readback(b) for b in (abi.BUF_RES_POINT_WSUM, abi.BUF_RES_NORMAL_W, abi.BUF_RES_LIGHT_IDX, abi.BUF_HIT_IDS)]
        want_acc = one.readback(abi.BUF_ACCUMULATOR)
    with MultiRenderer(Wd, Hd, [0] * n_bands, seed=4) as m:
        assert len(m.bands()) == n_bands and m.bands()[0][0] == 0 and m.bands()[-1][1] == Hd
        m.upload_scene(sc)
        m.set_params(p)
        bufs = [torch.empty((Hd, Wd, 3), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]
        for f in range(n):
            if f < 3:  # blocking calls ...
                got = m.render_frame(cams[f], f, out=bufs[f & 1])
            else:      # ... then pipelined ones into two host buffers
                m.render_frame_async(cams[f], f, bufs[f & 1])
                m.frame_wait(0)
                got = bufs[f & 1]
            assert np.array_equal(bits(want[f]), bits(got)), f"frame {f}: {(want[f] != got).any(-1).sum()} px differ"
            st = m.accumulate_display(f)
        for b, w in zip((abi.BUF_RES_POINT_WSUM, abi.BUF_RES_NORMAL_W, abi.BUF_RES_LIGHT_IDX, abi.BUF_HIT_IDS), want_res):
            assert np.array_equal(bits(w), bits(m.readback(b))), b
        assert np.array_equal(bits(want_acc), bits(m.readback(abi.BUF_ACCUMULATOR)))
        assert st["pixels"] == Wd * Hd and abs(st["mean"] - float(want_acc.astype(np.float64).mean())) < 1e-6 * max(1.0, st["mean"])
