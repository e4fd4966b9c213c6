#!/usr/bin/env python
"""One process, one host thread, N GPUs (rb_multi_*): end-to-end frames/s of the bench workload with the ASSEMBLED
frame landing in one page-locked host buffer, and a bit-for-bit check of that frame against a single-band handle.

    python tools/bench_multi.py --devices 0,1,2,3 [--steps 20] [--warmup 5] [--config 1m|10m]

Prints one JSON line. (bench.py's N > 1 arm is one process per GPU under torchrun, as the bench contract asks; this is the
same frame loop driven the way the reference's single Producer thread would drive it.)"""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from restir_embree_b200.renderer import MultiRenderer, Renderer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="1m", choices=["1m", "10m"])
    a = ap.parse_args()
    bench.select_config(a.config)
    W, H = bench.WIDTH, bench.HEIGHT
    devs = [int(x) for x in a.devices.split(",")]
    from restir_embree_b200 import scenes
    scene = scenes.scene_config(bench.SCENE)
    p = bench.bench_params()
    bufs = [torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]
    with MultiRenderer(W, H, devs, seed=123) as m:
        m.upload_scene(scene)
        m.set_params(p)
        # parity: frames 0..3 against a single-band handle on the first device
        with Renderer(W, H, device=devs[0], seed=123, collect_timings=False) as one:
            one.upload_scene(scene)
            one.set_params(p)
            for f in range(4):
                ref = one.render_frame(bench.camera_at(scene, f), f)
                got = m.render_frame(bench.camera_at(scene, f), f, out=bufs[0])
        diff = int((ref.view(np.uint32) != got.view(np.uint32)).any(-1).sum())
        frame = 0
        for _ in range(a.warmup):
            m.render_frame_async(bench.camera_at(scene, frame), frame, bufs[frame & 1])
            m.frame_wait(1)
            frame += 1
        m.frame_wait(0)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            m.render_frame_async(bench.camera_at(scene, frame), frame, bufs[frame & 1])
            m.frame_wait(1)
            frame += 1
        m.frame_wait(0)
        dt = time.perf_counter() - t0
        bands = m.bands()
    print(json.dumps({"metric": "frames/sec, one process / one host thread, assembled frame in host memory", "value": a.steps / dt,
                      "unit": "frames/s", "n_gpus": len(devs), "devices": devs, "steps": a.steps, "warmup": a.warmup,
                      "ms_per_step": dt / a.steps * 1e3, "config": {"workload": bench.WORKLOAD, "bands": bands, "balancer": "none (static bands)"},
                      "e2e": {"h2d_bytes_per_step": 144 * len(devs), "d2h_bytes_per_step": W * H * 12,
                              "note": "rb_multi_render_frame_async + rb_multi_frame_wait(1), two page-locked buffers, every band's rows in ONE buffer"},
                      "parity_check": {"n_bands": len(devs), "frames": 4, "bit_identical": diff == 0, "pixels_differing": diff,
                                       "sha256": hashlib.sha256(got.tobytes()).hexdigest()[:16]}}))


if __name__ == "__main__":
    main()
