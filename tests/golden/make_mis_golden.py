#!/usr/bin/env python
"""Generates tests/golden/ref_mis_golden.npz from the REFERENCE'S OWN CODE (oracle/_ref/libref.so): its
DirectMISIntegrator::calculateDirectLighting (P/DirectMISIntegrator.cpp) and material virtuals (MaterialPhong.cpp,
MaterialLambert.cpp) compiled where they lie, driven like Raytracer::get_pixel + NEEPathIntegrator (DI only) by
oracle/ref_shim (ref_produce_mis). SURVEY §8f N2: the ground-truth estimator.

Run in the build container (needs /root/reference):   python tests/golden/make_mis_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_binding as rb  # noqa: E402
from restir_embree_b200 import abi, scenes  # noqa: E402

W, H, FRAMES = 48, 32, 3


def main():
    sc = scenes.scene_config("tiny")
    out = {"W": W, "H": H, "FRAMES": FRAMES}
    ref = rb.Reference(W, H, sc)
    ref.set_params(abi.default_params(lightSampler=abi.LS_CDF))
    for f in range(FRAMES):
        cam = ref.camera(60.0, (2.2 + 0.05 * f, -2.4, 1.4 + 0.1 * f), (0.0, 0.0, 1.0))
        out[f"f{f}_cam"] = np.frombuffer(bytes(cam), dtype=np.float32).copy()
        out[f"f{f}_frame"] = ref.produce_mis()
    np.savez_compressed(os.path.join(HERE, "ref_mis_golden.npz"), **out)
    print("wrote ref_mis_golden.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
