"""Run by tests/test_emu_sanitized.py in a child process whose emulation library was built with
-fsanitize=address,undefined (RB_EMU_LIB) and which has libasan preloaded: every frame mode of the kernel bodies on a
small scene, one band and three bands with halo exchange and a boundary move, the ray seam, textures-free.
Prints SANITIZED-OK when nothing aborted."""
import sys

import numpy as np

import emu_binding as eb
from band_driver import make_bands, move_boundaries, render_banded
from restir_embree_b200 import Camera, abi, scenes
from test_emu_parity import CONFIGS

W, H = 56, 40


def cam(f):
    return Camera(W, H, 60, (4.2 + 0.15 * f, -4.4, 1.8 + 0.2 * f), (0, 0, 1.0 + 0.1 * f))


def main():
    sc = scenes.scene_config("small")
    n = 0
    for cfg in CONFIGS:
        p = abi.default_params(**cfg)
        e = eb.Emu(W, H, seed=3)
        e.upload_scene(sc)
        e.set_params(p)
        assert e.validate_bvh() == 0
        for f in range(2):
            img = e.render_frame(cam(f), f)
            assert np.isfinite(img).all()
        e.close()
        n += 1
    # bands: halo rows, re-derived G-buffer elements, deferred temporal pixels, a boundary move between frames
    for cfg in (CONFIGS[8], CONFIGS[1], CONFIGS[13]):
        p = abi.default_params(**cfg)
        bands = make_bands(eb.Emu, W, H, 3, seed=3)
        for b in bands:
            b.upload_scene(sc)
            b.set_params(p)
        for f in range(3):
            render_banded(bands, cam(f), f, p)
            if f == 0:
                move_boundaries(bands, [0, 10, 30, H])
        n += 1
    # ray seam, incl. degenerate rays
    e = eb.Emu(8, 8)
    e.upload_scene(scenes.scene_config("tiny"))
    rng = np.random.default_rng(1)
    rays = np.zeros(4000, dtype=abi.RAY_DTYPE)
    rays["org"] = rng.uniform(-2.9, 2.9, size=(4000, 3))
    d = rng.normal(size=(4000, 3))
    rays["dir"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["dir"][:10] = np.nan
    rays["dir"][10:20] = 0.0
    rays["tnear"] = 0.01
    rays["tfar"] = 3.4028235e38
    e.trace_closest(rays)
    e.trace_occluded(rays)
    print("SANITIZED-OK", n, flush=True)


if __name__ == "__main__":
    sys.exit(main())
