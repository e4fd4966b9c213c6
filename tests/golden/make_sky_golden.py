#!/usr/bin/env python
"""Generates tests/golden/ref_sky_golden.npz from the REFERENCE'S OWN CODE (oracle/_ref/libref.so): P/SphericalMap.cpp over
P/Texture.cpp (BILINEAR, CLAMP_TO_EDGE), compiled where they lie and reached through ReSTIRIntegrator::gBufferFillPass
(:231, useSkybox) and the MIS estimator's miss branch (P/NEEPathIntegrator.cpp:131) on an open scene
(tests/tex_fixture.py). Run in the build container:   python tests/golden/make_sky_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_binding as rb  # noqa: E402
import tex_fixture as tf  # noqa: E402
from restir_embree_b200 import abi  # noqa: E402

W, H, FRAMES = 48, 32, 3


def main():
    sc = tf.textured_scene()
    ref = rb.Reference(W, H, sc)
    ref.set_sky(tf.sky_arrays()[0])
    ref.set_params(abi.default_params(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, lightSampler=abi.LS_CDF, useSkybox=1))
    out = {"W": W, "H": H, "FRAMES": FRAMES}
    for f in range(FRAMES):
        frm, at = tf.sky_camera_path(f)
        cam = ref.camera(60.0, frm, at)
        out[f"f{f}_cam"] = np.frombuffer(bytes(cam), dtype=np.float32).copy()
        out[f"f{f}_frame"] = ref.produce_restir()
        out[f"f{f}_gbuf"] = ref.gbuffer()
        out[f"f{f}_res"] = ref.reservoirs()
    out["mis_frame"] = ref.produce_mis()
    np.savez_compressed(os.path.join(HERE, "ref_sky_golden.npz"), **out)
    g = out["f0_gbuf"]
    print("wrote ref_sky_golden.npz; miss pixels per frame:", [int((out[f"f{f}_gbuf"][..., 16] == 0).sum()) for f in range(FRAMES)])


if __name__ == "__main__":
    main()
