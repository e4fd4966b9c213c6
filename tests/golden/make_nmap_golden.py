#!/usr/bin/env python
"""Generates tests/golden/ref_nmap_golden.npz from the REFERENCE'S OWN CODE (oracle/_ref/libref.so): the normal-map branch
of Intersection::intersectEmbree (P/Intersection.h:25-39: Gram-Schmidt tangent, bitangent, TBN * (texel * 2 - 1)) over
P/Texture.cpp, compiled where they lie and reached through ReSTIRIntegrator::gBufferFillPass / brdfSampleLight and
DirectMISIntegrator on a scene with per-vertex tangents (tests/tex_fixture.py: normal_mapped_scene).
Run in the build container:   python tests/golden/make_nmap_golden.py"""
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
    sc = tf.normal_mapped_scene()
    ref = rb.Reference(W, H, sc)
    ref.set_params(abi.default_params(M_Area=4, M_Brdf=2, doSpatialReuse=1, doTemporalReuse=1, lightSampler=abi.LS_CDF))
    ref.set_textures(tf.normal_map_arrays(), tf.NMAP_SLOTS, tf.NMAP_N_MATERIALS)
    out = {"W": W, "H": H, "FRAMES": FRAMES}
    for f in range(FRAMES):
        frm, at = tf.camera_path(f)
        cam = ref.camera(60.0, frm, at)
        out[f"f{f}_cam"] = np.frombuffer(bytes(cam), dtype=np.float32).copy()
        out[f"f{f}_frame"] = ref.produce_restir()
        out[f"f{f}_gbuf"] = ref.gbuffer()
        out[f"f{f}_res"] = ref.reservoirs()
    out["mis_frame"] = ref.produce_mis()
    np.savez_compressed(os.path.join(HERE, "ref_nmap_golden.npz"), **out)
    print("wrote ref_nmap_golden.npz")


if __name__ == "__main__":
    main()
