#!/usr/bin/env python
"""Generates tests/golden/ref_golden.npz from the REFERENCE'S OWN CODE (oracle/_ref/libref.so: the files
P/ReSTIRIntegrator.cpp, MaterialPhong.cpp, MaterialLambert.cpp, Sampling.cpp, TriangleCDF.cpp, camera.cpp,
Reservoir.h, GBufferElement.h ... compiled where they lie, Embree replaced by the oracle's tracer).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures are small on purpose (48x32 frames): they pin the oracle's restatement, which then scales."""
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
CONFIGS = [
    dict(M_Area=4, M_Brdf=2, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=0),
    dict(M_Area=4, M_Brdf=2, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=1),
    dict(M_Area=4, M_Brdf=2, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=2, rejectDissimilarNeighbors=1),
    dict(M_Area=3, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=3, spatialReuseNeighborCount=3),
    dict(M_Area=3, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, spatialWeightCalc=4, doVisibilityPass=1),
    dict(M_Area=32, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1),  # BASELINE settings
    dict(),  # reference defaults (A=1, B=1, no reuse)
]


def camera_path(f):
    return (2.2 + 0.05 * f, -2.4, 1.4), (0.0, 0.0, 1.0)


def main():
    sc = scenes.scene_config("tiny")
    out = {"W": W, "H": H, "FRAMES": FRAMES, "n_configs": len(CONFIGS)}
    for ci, cfg in enumerate(CONFIGS):
        ref = rb.Reference(W, H, sc)
        ref.set_params(abi.default_params(**cfg))
        for f in range(FRAMES):
            frm, at = camera_path(f)
            cam = ref.camera(60.0, frm, at)
            img = ref.produce_restir()
            out[f"c{ci}_f{f}_cam"] = np.frombuffer(bytes(cam), dtype=np.float32).copy()
            out[f"c{ci}_f{f}_frame"] = img
            out[f"c{ci}_f{f}_res"] = ref.reservoirs()
            if f == 0:
                out[f"c{ci}_f{f}_gbuf"] = ref.gbuffer()
        out[f"c{ci}_cfg"] = np.array(sorted(cfg.items()), dtype=object) if cfg else np.array([], dtype=object)
    # leaf functions with the reference's shared mt19937 stream
    R = rb.lib()
    R.ref_seed(123)
    out["mt_floats"] = np.array([R.ref_random() for _ in range(16)], dtype=np.float32)
    R.ref_seed(7)
    disk = np.zeros((64, 2), dtype=np.float32)
    for i in range(64):
        R.ref_sampleDiskUniform(30.0, disk[i].ctypes.data)
    out["disk_r30_seed7"] = disk
    tri = np.array([0.1, 0.2, 0.3, 1.5, 0.1, 0.2, 0.3, 1.1, 0.9, 0, 0, 1, 0, 0.6, 0.8, 0.6, 0, 0.8], dtype=np.float32)
    R.ref_seed(9)
    ts = np.zeros((64, 7), dtype=np.float32)
    for i in range(64):
        R.ref_sampleTriangle(tri.ctypes.data, ts[i].ctypes.data)
    out["tri"] = tri
    out["tri_samples_seed9"] = ts
    rng = np.random.default_rng(0)
    n = 256
    elems = np.zeros((n, 13), dtype=np.float32)
    elems[:, 0:3] = rng.uniform(-2, 2, (n, 3))
    nn = rng.normal(size=(n, 3))
    elems[:, 3:6] = nn / np.linalg.norm(nn, axis=1, keepdims=True)
    elems[:, 6:9] = rng.uniform(0.2, 0.8, (n, 3))
    elems[:, 9:12] = rng.uniform(0.04, 0.5, (n, 1))
    elems[:, 12] = rng.choice([1.0, 5.0, 20.0, 80.0, 250.0, 1000.0], n)
    cams = (elems[:, 0:3] + elems[:, 3:6] * 2 + rng.normal(size=(n, 3)) * 0.5).astype(np.float32)
    wi = rng.normal(size=(n, 3))
    wi = (wi / np.linalg.norm(wi, axis=1, keepdims=True)).astype(np.float32)
    brdf = np.zeros((n, 3), dtype=np.float32)
    pdf = np.zeros(n, dtype=np.float32)
    smp = np.zeros((n, 4), dtype=np.float32)
    R.ref_seed(11)
    for i in range(n):
        R.ref_phong_evalBRDF(elems[i].ctypes.data, cams[i].ctypes.data, wi[i].ctypes.data, brdf[i].ctypes.data)
        pdf[i] = R.ref_phong_evalPdf(elems[i].ctypes.data, cams[i].ctypes.data, wi[i].ctypes.data)
        R.ref_phong_sampleBRDF(elems[i].ctypes.data, cams[i].ctypes.data, smp[i].ctypes.data)
    out.update(phong_elems=elems, phong_cams=cams, phong_wi=wi, phong_brdf=brdf, phong_pdf=pdf, phong_samples_seed11=smp)
    # post-path arithmetic: Utils::aces / Utils::compress / the accumulator's glm::mix
    rng = np.random.default_rng(21)
    hdr = np.concatenate([rng.random((200, 3)) * np.float32(4.0), rng.random((40, 3)) * np.float32(0.01),
                          np.float32([[0, 0, 0], [1, 1, 1], [100, 50, 3], [0.0031308, 0.0031309, 0.003]])]).astype(np.float32)
    aces = hdr.copy()
    for i in range(aces.shape[0]):
        R.ref_aces(aces[i].ctypes.data)
    u = np.concatenate([hdr.reshape(-1), np.float32([-1.0, 0.0, 1.0, 2.0, 0.0031308, 0.5])]).astype(np.float32)
    comp = np.array([R.ref_compress(float(x)) for x in u], dtype=np.float32)
    acc = np.zeros((16, 3), dtype=np.float32)
    frames = (rng.random((16, 8, 3)) * 3).astype(np.float32)
    hist = np.zeros((16, 8, 3), dtype=np.float32)
    for p in range(16):
        for k in range(8):
            R.ref_accumulate_mix(acc[p].ctypes.data, frames[p, k].ctypes.data, k)
            hist[p, k] = acc[p]
    out.update(post_hdr=hdr, post_aces=aces, post_u=u, post_compress=comp, post_frames=frames, post_acc_hist=hist)
    cfgs = np.array([repr(c) for c in CONFIGS])
    out = {k: v for k, v in out.items() if not (isinstance(v, np.ndarray) and v.dtype == object)}
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), configs=cfgs, **out)
    print("wrote", os.path.join(HERE, "ref_golden.npz"))


if __name__ == "__main__":
    main()
