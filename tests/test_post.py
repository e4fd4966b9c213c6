"""The step after the path: accumulate + ACES tonemap + sRGB compress + image statistics.

Chain of trust: the reference's own Utils::aces / Utils::compress / glm::mix (compiled in place by oracle/ref_shim,
vectors stored in tests/golden/ref_golden.npz) -> oracle restatement -> kernel body (host emulation here, the CUDA
kernel in the -m gpu test). Pixels are bit-exact; the two double sums are order-dependent on the GPU and carry a
1e-12 relative tolerance."""
import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
from restir_embree_b200 import Camera, abi, scenes
from test_emu_parity import bits
from test_ref_pin import GOLD

W, H = 64, 40
PARAMS = dict(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, wavefront=1)


def test_oracle_post_arithmetic_reproduces_reference_vectors():
    L = ob.lib()
    hdr = np.ascontiguousarray(GOLD["post_hdr"]).copy()
    for i in range(hdr.shape[0]):
        L.orc_aces(hdr[i].ctypes.data)
    assert np.array_equal(bits(hdr), bits(GOLD["post_aces"]))                     # Utils::aces
    comp = np.array([L.orc_compress(ob.MATH_LIBM, float(x)) for x in GOLD["post_u"]], dtype=np.float32)
    assert np.array_equal(bits(comp), bits(GOLD["post_compress"]))                # Utils::compress (libm powf)
    det = np.array([L.orc_compress(ob.MATH_DET, float(x)) for x in GOLD["post_u"]], dtype=np.float32)
    ulp = np.abs(bits(det).astype(np.int64) - bits(GOLD["post_compress"]).astype(np.int64))
    assert ulp.max() <= 2                                                          # det_math pow: what the GPU runs


def test_oracle_accumulator_reproduces_reference_mix_sequence():
    """glm::mix(acc, frame, 1/(k+1)) for k = 0..7 on 16 pixels, through the oracle's frame loop."""
    frames, hist = GOLD["post_frames"], GOLD["post_acc_hist"]
    o = ob.Oracle(4, 4, seed=1, tracer=ob.TRACER_BRUTE)
    for k in range(frames.shape[1]):
        o.set_frame(frames[:, k].reshape(4, 4, 3))
        o.accumulate_display(k, tonemap=False, gamma_correct=False)
        assert np.array_equal(bits(o.readback(abi.BUF_ACCUMULATOR).reshape(16, 3)), bits(hist[:, k]))


def _run(make, frames=4):
    sc = scenes.scene_config("small")
    x = make()
    x.upload_scene(sc)
    x.set_params(abi.default_params(**PARAMS))
    out = []
    for f in range(frames):
        cam = Camera(W, H, 60, (4.2 + 0.1 * f, -4.4, 1.8), (0, 0, 1.0))
        x.render_frame(cam, f)
        r = x.accumulate_display(f, tonemap=True, gamma_correct=True)
        disp, st = (r[0], r[1])
        out.append((x.readback(abi.BUF_ACCUMULATOR), x.readback(abi.BUF_DISPLAY), st))
    return out


def _compare(a, b):
    for f, ((acc_a, dsp_a, st_a), (acc_b, dsp_b, st_b)) in enumerate(zip(a, b)):
        assert np.array_equal(bits(acc_a), bits(acc_b)), f
        assert np.array_equal(bits(dsp_a), bits(dsp_b)), f
        for k in ("sum", "sum_sq", "mean", "variance"):
            assert abs(st_a[k] - st_b[k]) <= 1e-12 * max(1.0, abs(st_a[k])), (f, k, st_a[k], st_b[k])
    assert np.all(a[-1][1][..., 3] == 1.0) and a[-1][1][..., :3].max() <= 1.0 and a[-1][1][..., :3].min() >= 0.0
    assert a[-1][2]["variance"] > 0


def test_emulated_accumulate_display_matches_oracle():
    _compare(_run(lambda: ob.Oracle(W, H, seed=5, tracer=ob.TRACER_BVH2)), _run(lambda: eb.Emu(W, H, seed=5)))


@pytest.mark.gpu
def test_gpu_accumulate_display_matches_oracle(gpu):
    from restir_embree_b200.renderer import Renderer

    class R(Renderer):
        def accumulate_display(self, k, tonemap=True, gamma_correct=True):
            return super().accumulate_display(k, tonemap, gamma_correct, fetch=True)

    a = _run(lambda: ob.Oracle(W, H, seed=5, tracer=ob.TRACER_BVH2))
    r = R(W, H, seed=5)
    try:
        sc = scenes.scene_config("small")
        r.upload_scene(sc)
        r.set_params(abi.default_params(**PARAMS))
        b = []
        for f in range(4):
            cam = Camera(W, H, 60, (4.2 + 0.1 * f, -4.4, 1.8), (0, 0, 1.0))
            r.render_frame(cam, f)
            disp, st = r.accumulate_display(f)
            assert np.array_equal(bits(disp), bits(r.readback(abi.BUF_DISPLAY)))  # host copy == device buffer
            b.append((r.readback(abi.BUF_ACCUMULATOR), r.readback(abi.BUF_DISPLAY), st))
    finally:
        r.close()
    _compare(a, b)
