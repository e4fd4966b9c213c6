"""Edge cases of the frame loop, product kernel bodies (host emulation; CUDA in the gpu tier) against the oracle, bit for
bit: the limits of every parameter of P/ReSTIRIntegrator.cpp:13-35, image sizes that are not multiples of the 8 x 4 pixel
warp tile, scenes without emitters, cameras that see nothing, and pixels whose G-buffer element is empty but not emissive
(bgColor = 0: the reference runs every pass on them, P/ReSTIRIntegrator.cpp:188,240).
(File name: sorts after the other test files on purpose, so that the newest GPU tests run last under `-x`.)"""
import warnings

import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
import tex_fixture as tf
from restir_embree_b200 import Camera, abi, scenes

BUFS = (abi.BUF_HIT_IDS, abi.BUF_GBUF_POS_DEPTH, abi.BUF_GBUF_EMISSION, abi.BUF_RES_POINT_WSUM, abi.BUF_RES_NORMAL_W,
        abi.BUF_RES_LI_CONF, abi.BUF_RES_LIGHT_IDX)
FULL = dict(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, lightSampler=abi.LS_ALIAS)

# (name, image size, scene, parameter overrides, camera (from, at))
INSIDE = ((2.2, -2.4, 1.4), (0.0, 0.0, 1.0))
CASES = [
    ("max_neighbours", (40, 24), "tiny", dict(FULL, spatialReuseNeighborCount=32, wavefront=1), INSIDE),
    ("max_neighbours_inline", (40, 24), "tiny", dict(FULL, spatialReuseNeighborCount=32, spatialWeightCalc=1), INSIDE),
    ("no_neighbours", (40, 24), "tiny", dict(FULL, spatialReuseNeighborCount=0, wavefront=1), INSIDE),
    ("radius_zero", (40, 24), "tiny", dict(FULL, spatialReuseRadius=0.0, wavefront=1), INSIDE),
    ("radius_below_one", (40, 24), "tiny", dict(FULL, spatialReuseRadius=0.5, wavefront=1), INSIDE),
    ("huge_radius", (40, 24), "tiny", dict(FULL, spatialReuseRadius=1.0e6, wavefront=1), INSIDE),
    ("no_candidates", (40, 24), "tiny", dict(FULL, M_Area=0, M_Brdf=0, wavefront=1), INSIDE),
    ("no_candidates_inline", (40, 24), "tiny", dict(FULL, M_Area=0, M_Brdf=0), INSIDE),
    ("cap_one", (40, 24), "tiny", dict(FULL, confidenceCap=1, wavefront=1), INSIDE),
    ("cap_zero", (40, 24), "tiny", dict(FULL, confidenceCap=0, wavefront=1), INSIDE),
    ("zero_passes", (40, 24), "tiny", dict(FULL, spatialPassCount=0, wavefront=1), INSIDE),
    ("four_passes", (40, 24), "tiny", dict(FULL, spatialPassCount=4, wavefront=1), INSIDE),
    ("reject_everything", (40, 24), "tiny", dict(FULL, rejectDissimilarNeighbors=1, minNormalSimilarity=2.0, wavefront=1), INSIDE),
    ("zero_offsets", (40, 24), "tiny", dict(FULL, tnearOffset=0.0, tfarOffset=0.0, normalOffset=0.0, wavefront=1), INSIDE),
    ("one_pixel", (1, 1), "tiny", dict(FULL, wavefront=1), INSIDE),
    ("one_row", (67, 1), "tiny", dict(FULL, wavefront=1), INSIDE),
    ("one_column", (1, 45), "tiny", dict(FULL, wavefront=1), INSIDE),
    ("ragged_tiles", (37, 23), "tiny", dict(FULL, wavefront=1), INSIDE),
    ("ragged_tiles_inline", (37, 23), "tiny", dict(FULL), INSIDE),
    # open scene seen from outside, black background: most pixels are misses that are NOT emissive
    ("black_background", (48, 32), "open", dict(FULL, bgColor=(0.0, 0.0, 0.0), wavefront=1), ((2.5, -4.0, 2.0), (0.0, 0.5, 2.6))),
    ("black_background_inline", (48, 32), "open", dict(FULL, bgColor=(0.0, 0.0, 0.0)), ((2.5, -4.0, 2.0), (0.0, 0.5, 2.6))),
    ("sees_nothing", (32, 20), "open", dict(FULL, wavefront=1), ((0.0, 0.0, 10.0), (3.0, 0.0, 20.0))),
    # looking along the up vector: glm::lookAt degenerates, every matrix entry and every primary ray is NaN
    ("degenerate_camera", (32, 20), "open", dict(FULL, wavefront=1), ((0.0, 0.0, 10.0), (0.0, 0.0, 20.0))),
    ("no_emitters", (40, 24), "dark", dict(FULL, wavefront=1), ((2.5, -4.0, 2.0), (0.0, 0.5, 0.8))),
    ("no_emitters_inline", (40, 24), "dark", dict(FULL), ((2.5, -4.0, 2.0), (0.0, 0.5, 0.8))),
]


def make_scene(kind):
    if kind == "tiny":
        return scenes.scene_config("tiny")
    sc = tf.textured_scene()  # floor + back wall + lamp, open on four sides
    if kind == "dark":
        for m in sc.materials:
            m["emission"] = (0.0, 0.0, 0.0)
    return sc


def make_params(over):
    over = dict(over)
    bg = over.pop("bgColor", None)
    p = abi.default_params(**over)
    if bg is not None:
        p.bgColor[0], p.bgColor[1], p.bgColor[2] = bg
    return p


def run_case(make_product, case):
    name, (w, h), kind, over, (frm, at) = case
    sc, p = make_scene(kind), make_params(over)
    o = ob.Oracle(w, h, seed=9, tracer=ob.TRACER_BRUTE)
    r = make_product(w, h)
    for x in (o, r):
        x.upload_scene(sc)
        x.set_params(p)
    for f in range(3):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)  # the degenerate camera divides by zero on purpose
            cam = Camera(w, h, 60, (frm[0] + (0.0 if name == "degenerate_camera" else 0.07 * f), frm[1], frm[2] + 0.03 * f), at)
        a, b = r.render_frame(cam, f), o.render_frame(cam, f)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{name}, frame {f}: {(a != b).any(-1).sum()} px differ"
        assert np.isfinite(a).all() and (a >= 0).all(), name  # Integrator::sanitize
    for buf in BUFS:
        assert np.array_equal(r.readback(buf).view(np.uint32), o.readback(buf).view(np.uint32)), (name, buf)
    return a


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_edge_case_kernel_bodies_match_oracle(case):
    img = run_case(lambda w, h: eb.Emu(w, h, seed=9), case)
    if case[0] in ("sees_nothing", "degenerate_camera"):
        assert (img == np.float32(0.5)).all()  # bgColor everywhere
    if case[0].startswith("no_emitters"):
        assert (img[img != np.float32(0.5)] == 0).all()  # hits are black, misses show the background


@pytest.mark.gpu
def test_gpu_edge_cases_match_oracle(gpu):
    from restir_embree_b200.renderer import Renderer
    open_handles = []

    def product(w, h):
        r = Renderer(w, h, seed=9)
        open_handles.append(r)
        return r

    try:
        for case in CASES:
            run_case(product, case)
            open_handles.pop().close()
    finally:
        for r in open_handles:
            r.close()
