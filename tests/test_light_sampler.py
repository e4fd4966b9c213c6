"""SURVEY §8c seam 3 — the two light samplers are the same distribution. The reference picks an emissive triangle by CDF
inversion (TriangleCDF: float prefix sums + lower_bound, pdf by adjacent difference, P/TriangleCDF.cpp:8-54); north_star asks
for an alias table (Vose, built on the host with a fixed processing order). Checked here: (1) the probability mass the
alias table assigns to every emitter equals area_i / total — the pdf both samplers report — and the CDF's adjacent
differences; (2) the pick functions of the product's kernel tables and the oracle's are the same arrays; (3) converged
ReSTIR images rendered with either sampler agree (relMSE of the 64-frame means, our definition of SURVEY §8d)."""
import numpy as np

import emu_binding as eb
import oracle_binding as ob
from restir_embree_b200 import Camera, abi, scenes


def emitter_areas(sc):
    out = []
    for pos, _, m in sc.surfaces:
        if sum(sc.materials[m]["emission"]) > 0:
            p = pos.astype(np.float64)
            out.append(0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1))
    return np.concatenate(out)


def alias_pmf(prob, idx):
    n = len(prob)
    pmf = prob.astype(np.float64).copy()
    np.add.at(pmf, idx.astype(np.int64), 1.0 - prob.astype(np.float64))
    return pmf / n


def test_alias_table_and_cdf_describe_the_same_distribution():
    for name in ("tiny", "small"):
        sc = scenes.scene_config(name)
        o = ob.Oracle(16, 16, seed=1, tracer=ob.TRACER_BRUTE)
        o.upload_scene(sc)
        prob, idx, cdf = (o.light_table(b) for b in (abi.BUF_ALIAS_PROB, abi.BUF_ALIAS_IDX, abi.BUF_LIGHT_CDF))
        area = emitter_areas(sc)
        n = len(area)
        assert len(prob) == len(idx) == len(cdf) == n == sc.n_emissive
        want = area / area.sum()
        assert (prob >= 0).all() and (prob <= 1).all() and (idx < n).all()
        # Vose in float32: exact up to the rounding the donations accumulate (an emitter that takes ~100 of them ends
        # 8e-5 relative off on the 200-emitter scene; as a bias of the estimator that is 1e-8 in relMSE)
        assert np.allclose(alias_pmf(prob, idx), want, rtol=2e-4, atol=1e-9)
        assert np.allclose(np.diff(np.concatenate([[0.0], cdf.astype(np.float64)])), want, rtol=0, atol=4e-7 * np.sqrt(n))
        assert abs(float(cdf[-1]) - 1.0) < 1e-4 and (np.diff(cdf) >= 0).all()      # sequential float prefix sum, :15-22
        # the product builds the same tables (shared host code, rb_host_scene.h) — the picks on the GPU are the oracle's
        e = eb.Emu(16, 16, seed=1)
        e.upload_scene(sc)
        for b in (abi.BUF_ALIAS_PROB, abi.BUF_ALIAS_IDX, abi.BUF_LIGHT_CDF):
            assert np.array_equal(e.light_table(b, n), o.light_table(b)), (name, b)


def test_converged_images_agree_between_cdf_and_alias_sampling():
    """the difference between the two samplers' converged means is the noise floor: the relMSE between a CDF run and an
    alias run is what two alias runs with different seeds have between them — without reuse (unbiased RIS, where the
    image means agree too) and with temporal + spatial reuse (static camera: long-lived samples, a higher floor)"""
    sc = scenes.scene_config("tiny")
    w, h, frames = 64, 40, 96
    cam = Camera(w, h, 60, (2.2, -2.4, 1.4), (0, 0, 1.0))

    def mean_image(sampler, seed, **reuse):
        o = ob.Oracle(w, h, seed=seed, tracer=ob.TRACER_BRUTE)
        o.upload_scene(sc)
        o.set_params(abi.default_params(M_Area=16, M_Brdf=1, doVisibilityPass=1, lightSampler=sampler, **reuse))
        acc = np.zeros((h, w, 3), dtype=np.float64)
        for f in range(frames):
            acc += o.render_frame(cam, f)
        return acc / frames

    def relmse(a, b):
        return float(np.mean(((a - b) ** 2).sum(-1) / ((b ** 2).sum(-1) + 1e-2)))

    for reuse in (dict(), dict(doTemporalReuse=1, doSpatialReuse=1)):
        cdf, alias, alias2 = (mean_image(abi.LS_CDF, 7, **reuse), mean_image(abi.LS_ALIAS, 7, **reuse),
                              mean_image(abi.LS_ALIAS, 8, **reuse))
        floor = relmse(alias2, alias)
        assert relmse(cdf, alias) < 1.5 * floor + 1e-5, (reuse, relmse(cdf, alias), floor)
        if not reuse:
            assert floor < 2e-3 and abs(cdf.mean() / alias.mean() - 1) < 0.02
