"""CPU tier: the product's kernel bodies (rb_build/rb_scene/rb_passes .cuh), compiled for the host by
tests/emu, against the oracle — bit-exact frames, reservoirs, selected lights, hit ids and
as-written ray counts. This is the pre-GPU gate for kernel logic; the -m gpu tests repeat it on the
real kernels through the C ABI."""
import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
from restir_embree_b200 import Camera, abi, scenes

W, H = 96, 56

CONFIGS = [
    dict(M_Area=8, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, lightSampler=1),
    dict(M_Area=3, M_Brdf=2, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=1, spatialPassCount=2),
    dict(M_Area=3, M_Brdf=2, doSpatialReuse=1, spatialWeightCalc=2, rejectDissimilarNeighbors=1),
    dict(M_Area=2, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=3, spatialReuseNeighborCount=3),
    dict(M_Area=2, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=4, doVisibilityPass=1),
    dict(M_Area=0, M_Brdf=3, doTemporalReuse=1),
    dict(M_Area=5, M_Brdf=0, doSpatialReuse=1, spatialReuseRadius=100.0, spatialReuseNeighborCount=8),
    dict(),  # reference defaults
    # stream -> trace -> resolve split
    dict(M_Area=8, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, lightSampler=1, wavefront=1),
    dict(M_Area=3, M_Brdf=2, doSpatialReuse=1, doTemporalReuse=1, spatialPassCount=2, wavefront=1),
    dict(M_Area=2, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=4, doVisibilityPass=1, wavefront=1),
    dict(M_Area=4, M_Brdf=0, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, spatialReuseNeighborCount=8,
         rejectDissimilarNeighbors=1, wavefront=1),
    # the other spatial MIS modes in the wavefront schedule (spatial_pixel staged: StagedVis)
    dict(M_Area=3, M_Brdf=2, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=1, spatialPassCount=2, wavefront=1),
    dict(M_Area=3, M_Brdf=2, doSpatialReuse=1, spatialWeightCalc=2, rejectDissimilarNeighbors=1, wavefront=1),
    dict(M_Area=2, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=3, spatialReuseNeighborCount=3, wavefront=1),
    dict(M_Area=2, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=4, spatialPassCount=2, wavefront=1),
    # the repaired temporal fetch (last frame's reservoir of the REPROJECTED pixel; the reference reads the same pixel, :641)
    dict(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, temporalFetchReprojected=1),
    dict(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, temporalFetchReprojected=1, wavefront=1),
    # BALANCE_HEURISTIC with k = 7: 72 slots per pixel, over the staged schedule's limit -> the pass traces inline
    dict(M_Area=2, M_Brdf=1, doSpatialReuse=1, spatialWeightCalc=1, spatialReuseNeighborCount=7, wavefront=1),
]

ALL_BUFS = (abi.BUF_HIT_IDS, abi.BUF_GBUF_POS_DEPTH, abi.BUF_GBUF_NORMAL_SHIN, abi.BUF_GBUF_DIFFUSE_IIM,
            abi.BUF_GBUF_SPEC_TYPE, abi.BUF_GBUF_EMISSION, abi.BUF_RES_POINT_WSUM, abi.BUF_RES_NORMAL_W,
            abi.BUF_RES_LI_CONF, abi.BUF_RES_LIGHT_IDX)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def small():
    return scenes.scene_config("small")


@pytest.mark.parametrize("ci", range(len(CONFIGS)))
def test_emulated_kernels_match_oracle_bit_for_bit(small, ci):
    p = abi.default_params(**CONFIGS[ci])
    o = ob.Oracle(W, H, seed=7, tracer=ob.TRACER_BVH2)
    e = eb.Emu(W, H, seed=7)
    for x in (o, e):
        x.upload_scene(small)
        x.set_params(p)
    assert e.validate_bvh() == 0
    for f in range(3):
        cam = Camera(W, H, 60, (4.2 + 0.15 * f, -4.4, 1.8 + 0.05 * f), (0, 0, 1.0))
        a, b = o.render_frame(cam, f), e.render_frame(cam, f)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
        for buf in ALL_BUFS:
            assert np.array_equal(bits(o.readback(buf)), bits(e.readback(buf))), (f, buf)
        oc, ec = o.counters(), e.counters()
        assert oc["closest"] == ec["closest"] and oc["any_as_written"] == ec["any_as_written"]
        assert ec["any_traced"] <= ec["any_as_written"]


def test_temporal_fetch_reprojected_is_a_different_estimator_under_camera_motion(small):
    """the flag must actually change what is merged (otherwise the two configs above would pass vacuously)"""
    imgs = []
    for flag in (0, 1):
        e = eb.Emu(W, H, seed=7)
        e.upload_scene(small)
        e.set_params(abi.default_params(M_Area=4, M_Brdf=1, doTemporalReuse=1, temporalFetchReprojected=flag))
        for f in range(3):
            img = e.render_frame(Camera(W, H, 60, (4.2 + 0.4 * f, -4.4, 1.8), (0, 0, 1.0)), f)
        imgs.append(img)
    assert (imgs[0] != imgs[1]).any(-1).mean() > 0.05


def test_emulated_traversal_matches_brute_force_on_random_rays():
    sc = scenes.scene_config("tiny")
    o = ob.Oracle(8, 8, tracer=ob.TRACER_BRUTE)
    e = eb.Emu(8, 8)
    o.upload_scene(sc)
    e.upload_scene(sc)
    rng = np.random.default_rng(5)
    n = 20000
    rays = np.zeros(n, dtype=abi.RAY_DTYPE)
    rays["org"] = rng.uniform((-2.9, -2.9, 0.05), (2.9, 2.9, 2.9), size=(n, 3))
    d = rng.normal(size=(n, 3))
    d[:200, 0] = 0.0  # axis-parallel components
    d[200:400, 1] = 0.0
    d[400:600] = np.eye(3)[rng.integers(0, 3, 200)] * rng.choice([-1.0, 1.0], size=(200, 1))
    rays["dir"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tnear"] = 0.01
    rays["tfar"] = rng.choice([3.4028235e38, 1.0, 2.5], size=n)
    rays["dir"][600:610] = np.nan  # from == to in testOcclusion
    rays["dir"][610:620] = 0.0     # empty-reservoir sentinel direction
    ho, he = o.trace_closest(rays), e.trace_closest(rays)
    assert np.array_equal(ho["primID"], he["primID"]) and np.array_equal(ho["geomID"], he["geomID"])
    assert np.array_equal(bits(ho["t"]), bits(he["t"])) and np.array_equal(bits(ho["u"]), bits(he["u"]))
    assert np.array_equal(o.trace_occluded(rays), e.trace_occluded(rays))
    assert (ho["geomID"] != 0xFFFFFFFF).mean() > 0.5


@pytest.mark.parametrize("n_tris", [1, 2, 3, 4, 9])
def test_degenerate_scene_sizes_build_and_trace(n_tris):
    """Scenes at and around the leaf size, with duplicated (coincident) triangles."""
    sc = abi.SceneArrays()
    m = sc.add_material(abi.MAT_PHONG, (0.5, 0.5, 0.5), (0.1, 0.1, 0.1), (0, 0, 0), 20.0)
    em = sc.add_material(abi.MAT_PHONG, (0.5, 0.5, 0.5), (0, 0, 0), (10, 10, 10), 1.0)
    base = np.array([[[-1, -1, 0], [1, -1, 0], [0, 1, 0]]], dtype=np.float32)
    tris = np.concatenate([base + np.array([0, 0, 0.25 * (i // 2)], dtype=np.float32) for i in range(n_tris)], 0)
    sc.add_surface(tris, np.broadcast_to(np.float32([0, 0, 1]), tris.shape).copy(), m)
    light = base * 0.2 + np.float32([0, 0, 3])
    sc.add_surface(light, np.broadcast_to(np.float32([0, 0, -1]), light.shape).copy(), em)
    o = ob.Oracle(8, 8, tracer=ob.TRACER_BRUTE)
    e = eb.Emu(8, 8)
    o.upload_scene(sc)
    e.upload_scene(sc)
    assert e.validate_bvh() == 0
    rng = np.random.default_rng(n_tris)
    n = 2000
    rays = np.zeros(n, dtype=abi.RAY_DTYPE)
    rays["org"] = rng.uniform((-1.5, -1.5, -1), (1.5, 1.5, 4), size=(n, 3))
    d = rng.normal(size=(n, 3))
    rays["dir"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tnear"], rays["tfar"] = 0.01, 3.4028235e38
    ho, he = o.trace_closest(rays), e.trace_closest(rays)
    # coincident triangles: the tie-break picks the smaller (geomID, primID) on both sides
    assert np.array_equal(ho["primID"], he["primID"]) and np.array_equal(bits(ho["t"]), bits(he["t"]))
    assert np.array_equal(o.trace_occluded(rays), e.trace_occluded(rays))


def coincident_emitter_scene(emitter_first):
    """A large emitter lying in the very plane of a non-emissive sheet (overlapping), over a glossy floor, plus a
    small free-standing lamp. Scene order decides which of the coincident surfaces has the smaller ids."""
    sc = abi.SceneArrays()
    floor = sc.add_material(abi.MAT_PHONG, (0.6, 0.5, 0.4), (0.3, 0.3, 0.3), (0, 0, 0), 20.0)
    sheet = sc.add_material(abi.MAT_PHONG, (0.2, 0.2, 0.2), (0.0, 0.0, 0.0), (0, 0, 0), 5.0)
    em = sc.add_material(abi.MAT_PHONG, (0.5, 0.5, 0.5), (0, 0, 0), (30, 25, 20), 1.0)

    def quad(x0, y0, x1, y1, z):
        return np.float32([[[x0, y0, z], [x1, y0, z], [x1, y1, z]], [[x0, y0, z], [x1, y1, z], [x0, y1, z]]])

    f = quad(-3, -3, 3, 3, 0.0)
    sc.add_surface(f, np.broadcast_to(np.float32([0, 0, 1]), f.shape).copy(), floor)
    big = quad(-2, -2, 2, 2, 2.0)
    lamp = quad(-1.5, -1.5, 1.5, 1.5, 2.0)
    order = [(lamp, em), (big, sheet)] if emitter_first else [(big, sheet), (lamp, em)]
    for t, m in order:
        sc.add_surface(t, np.broadcast_to(np.float32([0, 0, -1]), t.shape).copy(), m)
    small_lamp = quad(2.2, 2.2, 2.8, 2.8, 1.0)
    sc.add_surface(small_lamp, np.broadcast_to(np.float32([0, 0, -1]), small_lamp.shape).copy(), em)
    return sc


TWO_STEP_PARAMS = dict(M_Area=2, M_Brdf=3, doTemporalReuse=1, doSpatialReuse=1, doVisibilityPass=1, wavefront=1)


@pytest.mark.parametrize("emitter_first", [True, False])
def test_two_step_brdf_rays_with_coincident_emitter(emitter_first):
    """BRDF-candidate rays go to the emissive-only BVH first and then ask the full BVH whether anything precedes the
    emitter hit (same closest-hit order: smaller t, then smaller triangle id). Checked against the oracle's plain
    closest-hit query with the emitter before and after the coincident sheet in scene order."""
    sc = coincident_emitter_scene(emitter_first)
    Wd, Hd = 48, 32
    p = abi.default_params(**TWO_STEP_PARAMS)
    o = ob.Oracle(Wd, Hd, seed=3, tracer=ob.TRACER_BRUTE)
    e = eb.Emu(Wd, Hd, seed=3)
    for x in (o, e):
        x.upload_scene(sc)
        x.set_params(p)
    for fr in range(2):
        cam = Camera(Wd, Hd, 70, (2.5, -2.6 + 0.1 * fr, 1.2), (0, 0, 0.3))
        a, b = o.render_frame(cam, fr), e.render_frame(cam, fr)
        assert np.array_equal(bits(a), bits(b)), f"frame {fr}: {(a != b).any(-1).sum()} px differ"
        for buf in (abi.BUF_RES_POINT_WSUM, abi.BUF_RES_NORMAL_W, abi.BUF_RES_LIGHT_IDX):
            assert np.array_equal(bits(o.readback(buf)), bits(e.readback(buf))), (fr, buf)
    li = e.readback(abi.BUF_RES_LIGHT_IDX)[..., 0]
    pts = e.readback(abi.BUF_RES_POINT_WSUM)
    assert ((li >= 0) & (pts[..., 2] == 2.0)).any()  # reservoirs do hold samples on the coincident lamp
