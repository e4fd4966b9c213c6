"""The reference's own (disabled) invariant ReSTIRIntegrator::testReprojection (P/ReSTIRIntegrator.cpp:591-622, its inner
loop increments x instead of y — fixed here): projecting a pixel's G-buffer position with the SAME frame's view matrix
and focal length (the arithmetic of reprojectForward, :567-587: glm mat4 * vec4, -x/z * f + w/2, y/z * f + h/2,
glm::round) must give back the pixel's own coordinates. It ties Camera::GenerateRay (pixel corner, P/camera.cpp:20-42),
the hit point org + dir * t and the reprojection of the temporal pass together. Checked on the oracle's and on the
product kernel bodies' G-buffer (host emulation; CUDA in the gpu tier), SURVEY §8c."""
import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
from restir_embree_b200 import Camera, abi, scenes

F = np.float32


def reproject(view_mat, focal, w, h, pos):
    """reprojectForward in float32, glm's operation order; returns (valid, sx, sy)"""
    m = np.asarray(view_mat, dtype=F)  # column-major
    x, y, z = (pos[..., i].astype(F) for i in range(3))
    vx = (m[0] * x + m[4] * y) + (m[8] * z + m[12] * F(1))
    vy = (m[1] * x + m[5] * y) + (m[9] * z + m[13] * F(1))
    vz = (m[2] * x + m[6] * y) + (m[10] * z + m[14] * F(1))
    valid = vz < 0
    with np.errstate(divide="ignore", invalid="ignore"):
        fx = (-vx / vz) * F(focal) + F(w) / F(2)
        fy = (vy / vz) * F(focal) + F(h) / F(2)
    rnd = lambda a: np.sign(a) * np.floor(np.abs(a) + F(0.5))  # glm::round: half away from zero
    sx, sy = rnd(fx), rnd(fy)
    valid &= (sx >= 0) & (sx <= w - 1) & (sy >= 0) & (sy <= h - 1)
    return valid, sx, sy


def check(renderer, w, h, cams):
    ys, xs = np.mgrid[0:h, 0:w]
    for f, cam in enumerate(cams):
        renderer.render_frame(cam, f)
        g = renderer.readback(abi.BUF_GBUF_POS_DEPTH)
        hit = g[..., 3] > 0
        assert hit.mean() > 0.5
        c = cam.to_abi()
        valid, sx, sy = reproject(np.array(c.viewMat[:]), c.focal_px, w, h, g[..., :3])
        assert valid[hit].all(), f"frame {f}: {(~valid[hit]).sum()} hit pixels do not reproject onto the screen"
        assert np.array_equal(sx[hit], xs[hit]) and np.array_equal(sy[hit], ys[hit]), f"frame {f}"


def cameras(w, h):
    # odd sizes, off-centre look-at, near and far: the corner convention (x - w/2, h/2 - y) must hold for all of them
    return [Camera(w, h, 60, (2.2, -2.4, 1.4), (0, 0, 1.0)), Camera(w, h, 35, (0.3, -2.9, 2.6), (0.5, 0.2, 0.1)),
            Camera(w, h, 100, (-1.5, 1.0, 0.4), (1.0, -1.0, 1.5))]


@pytest.mark.parametrize("size", [(96, 64), (77, 51)])
def test_reprojection_invariant_oracle_and_kernel_bodies(size):
    w, h = size
    sc = scenes.scene_config("tiny")
    p = abi.default_params(M_Area=1, M_Brdf=0)
    for r in (ob.Oracle(w, h, seed=1, tracer=ob.TRACER_BRUTE), eb.Emu(w, h, seed=1)):
        r.upload_scene(sc)
        r.set_params(p)
        check(r, w, h, cameras(w, h))


def test_static_camera_every_valid_pixel_merges_with_its_own_history():
    """with an unchanged camera the backward and forward reprojections are that same invariant: no temporal rejections
    among the pixels whose position reprojects at all"""
    w, h = 96, 64
    sc = scenes.scene_config("tiny")
    o = ob.Oracle(w, h, seed=1, tracer=ob.TRACER_BRUTE)
    o.upload_scene(sc)
    o.set_params(abi.default_params(M_Area=2, M_Brdf=1, doTemporalReuse=1))
    cam = cameras(w, h)[0]
    o.render_frame(cam, 0)
    o.temporal_stats()
    o.render_frame(cam, 1)
    st = o.temporal_stats()
    assert st["depth_backward_failed"] == 0 and st["reproject_forward_failed"] == 0 and st["depth_forward_failed"] == 0
    assert st["merged"] > 0.5 * w * h


@pytest.mark.gpu
def test_gpu_reprojection_invariant(gpu):
    from restir_embree_b200.renderer import Renderer
    w, h = 317, 203
    with Renderer(w, h, seed=1) as r:
        r.upload_scene(scenes.scene_config("tiny"))
        r.set_params(abi.default_params(M_Area=1, M_Brdf=0))
        check(r, w, h, cameras(w, h))
