#!/usr/bin/env python
"""Generates tests/golden/ref_tex_golden.npz from the REFERENCE'S OWN CODE (oracle/_ref/libref.so): P/Texture.cpp
(get_texel, bilinear, REPEAT) and Material::getDiffuseColor / getSpecularColor / getShininess (P/material.cpp:105-134)
compiled where they lie, called by ReSTIRIntegrator::gBufferFillPass on a textured scene (tests/tex_fixture.py).
SURVEY §8f N3, second half.   Run in the build container:   python tests/golden/make_tex_golden.py"""
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
    ref.set_params(abi.default_params(M_Area=4, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, lightSampler=abi.LS_CDF))
    ref.set_textures(tf.texel_arrays(), tf.SLOTS, tf.N_MATERIALS)
    out = {"W": W, "H": H, "FRAMES": FRAMES}
    for f in range(FRAMES):
        frm, at = tf.camera_path(f)
        cam = ref.camera(60.0, frm, at)
        out[f"f{f}_cam"] = np.frombuffer(bytes(cam), dtype=np.float32).copy()
        out[f"f{f}_frame"] = ref.produce_restir()
        out[f"f{f}_gbuf"] = ref.gbuffer()
    out["mis_frame"] = ref.produce_mis()
    np.savez_compressed(os.path.join(HERE, "ref_tex_golden.npz"), **out)
    print("wrote ref_tex_golden.npz")


if __name__ == "__main__":
    main()
