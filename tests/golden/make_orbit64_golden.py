#!/usr/bin/env python
"""Generates tests/golden/orbit64_1080p.json: BASELINE configs[4] — the 64-frame orbit at 1920x1080 on the 1M-triangle /
10k-emitter scene with bench.py's parameters and camera path — rendered by the CPU ORACLE (full frames, no band edge),
stored as one SHA-256 per frame (frame_data bytes) plus the hashes of the final reservoir planes and of the 64-frame
running mean. The GPU tier renders the same 64 frames and compares hashes: bit-identical frames => relMSE 0 per frame
and on the converged image. Takes ~30-40 min on 8 cores (the oracle is the reference algorithm, one p-hat at a time):

    python tests/golden/make_orbit64_golden.py

The oracle itself is pinned to the reference's own sources (tests/test_ref_pin.py); its counter-RNG / det_math mode is
platform independent (no libm), so the hashes hold on any x86-64 host."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_binding as ob  # noqa: E402
from restir_embree_b200 import Camera, abi, scenes  # noqa: E402

W, H, FRAMES, SEED = 1920, 1080, 64, 123
PARAMS = dict(M_Area=32, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, spatialReuseNeighborCount=5,
              spatialPassCount=1, spatialReuseRadius=30.0, lightSampler=abi.LS_ALIAS)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    sc = scenes.scene_config("1m")
    o = ob.Oracle(W, H, seed=SEED, tracer=ob.TRACER_BVH2)
    o.upload_scene(sc)
    o.set_params(abi.default_params(**PARAMS))
    c = sc.meta["center"]
    out = dict(width=W, height=H, frames=FRAMES, seed=SEED, scene="1m", params=PARAMS, camera="scenes.orbit_position(center, f), fov 55",
               frame_sha256=[], frame_mean=[])
    acc = np.zeros((H, W, 3), dtype=np.float32)
    t0 = time.time()
    for f in range(FRAMES):
        a = o.render_frame(Camera(W, H, 55, scenes.orbit_position(c, f), c), f)
        out["frame_sha256"].append(sha(a))
        out["frame_mean"].append(float(a.astype(np.float64).mean()))
        x = np.float32(1.0) / np.float32(f + 1)   # glm::mix(acc, frame, 1 / (accFrameCtr + 1)), P/simpleguidx11.cpp:251
        acc = acc * (np.float32(1.0) - x) + a * x
        print(f"frame {f}: {time.time() - t0:.0f} s", flush=True)
    out["accumulator_sha256"] = sha(acc)
    out["accumulator_mean"] = float(acc.astype(np.float64).mean())
    for name, buf in (("res_light_idx", abi.BUF_RES_LIGHT_IDX), ("res_point_wsum", abi.BUF_RES_POINT_WSUM),
                      ("res_normal_W", abi.BUF_RES_NORMAL_W), ("res_Li_conf", abi.BUF_RES_LI_CONF), ("hit_ids", abi.BUF_HIT_IDS)):
        out[name + "_sha256"] = sha(o.readback(buf))
    out["temporal_stats"] = o.temporal_stats()
    json.dump(out, open(os.path.join(HERE, "orbit64_1080p.json"), "w"), indent=1)
    print("wrote orbit64_1080p.json")


if __name__ == "__main__":
    main()
