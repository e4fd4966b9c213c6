#!/usr/bin/env python
"""BVH-quality probe (CPU, test-only emulation of the device build + traversal): average node visits and
triangle tests per ray on the bench scene, for primary rays and for shadow rays G-buffer point -> random
point on a random emitter. Used to iterate on the builder without GPU time.

  python tools/bvh_quality.py [scene=1m] [width=240]
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import emu_binding as eb  # noqa: E402
from restir_embree_b200 import Camera, abi, scenes  # noqa: E402
from restir_embree_b200.renderer import make_rays  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "1m"
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 240
    H = W * 9 // 16
    sc = scenes.scene_config(name)
    e = eb.Emu(W, H, seed=1)
    t0 = time.time()
    e.upload_scene(sc)
    st = e.scene_stats()
    print("scene", name, st, "build %.1fs" % (time.time() - t0))
    L = e.L
    L.emu_trace_steps.restype = C.c_double
    L.emu_trace_steps.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
    L.emu_last_tri_tests.restype = C.c_double
    c = sc.meta["center"]
    cam = Camera(W, H, 55, scenes.orbit_position(c, 0), c)
    p = abi.default_params(M_Area=1, M_Brdf=0)
    e.set_params(p)
    e.render_frame(cam, 0)
    pos = e.readback(abi.BUF_GBUF_POS_DEPTH).reshape(-1, 4)
    ids = e.readback(abi.BUF_HIT_IDS).reshape(-1, 2)
    hit = ids[:, 0] != 0xFFFFFFFF
    # primary rays
    px = np.stack(np.meshgrid(np.arange(W), np.arange(H)), -1).reshape(-1, 2)
    ca = cam.to_abi()
    inv = np.array(ca.invViewMat, dtype=np.float32).reshape(4, 4).T  # column-major
    d_c = np.stack([px[:, 0] - W / 2, H / 2 - px[:, 1], np.full(len(px), -ca.focal_px)], -1).astype(np.float32)
    d_w = d_c @ inv[:3, :3].T
    d_w /= np.linalg.norm(d_w, axis=1, keepdims=True)
    org = np.tile(np.array(ca.pos, dtype=np.float32), (len(px), 1))
    rays = make_rays(org, direction=d_w, tnear=0.01, tfar=np.full(len(px), 3e38, dtype=np.float32))
    s = L.emu_trace_steps(e.h, rays.ctypes.data, len(rays), 0)
    print("primary  closest: %.2f node visits/ray, %.2f tri tests/ray" % (s, L.emu_last_tri_tests()))
    # shadow rays to random emitter points
    rng = np.random.default_rng(5)
    em = np.concatenate([pos_ for pos_, _, m in sc.surfaces if sum(sc.materials[m]["emission"]) > 0], 0)  # [n, 3, 3]
    P = pos[hit, :3]
    k = rng.integers(0, len(em), len(P))
    b = rng.random((len(P), 2)).astype(np.float32)
    sq = np.sqrt(b[:, :1])
    w0, w1, w2 = 1 - sq, sq * (1 - b[:, 1:]), sq * b[:, 1:]
    T = em[k, 0] * w0 + em[k, 1] * w1 + em[k, 2] * w2
    rays = make_rays(P, target=T.astype(np.float32))
    s = L.emu_trace_steps(e.h, rays.ctypes.data, len(rays), 1)
    tt = L.emu_last_tri_tests()
    occ = e.trace_occluded(rays)
    print("shadow   any-hit: %.2f node visits/ray, %.2f tri tests/ray, occluded %.1f%%" % (s, tt, 100 * occ.mean()))
    s = L.emu_trace_steps(e.h, rays.ctypes.data, len(rays), 0)
    print("shadow   closest: %.2f node visits/ray, %.2f tri tests/ray" % (s, L.emu_last_tri_tests()))
    # how much of the any-hit cost is the visiting ORDER: the same segments traced from the light towards the surface
    # (front-to-back from the other end), and split into occluded / unoccluded rays
    back = make_rays(T.astype(np.float32), target=P)
    s = L.emu_trace_steps(e.h, back.ctypes.data, len(back), 1)
    print("shadow   any-hit, light -> surface: %.2f node visits/ray, %.2f tri tests/ray" % (s, L.emu_last_tri_tests()))
    for name, sel in (("occluded", occ != 0), ("unoccluded", occ == 0)):
        sub = np.ascontiguousarray(rays[sel])
        a = L.emu_trace_steps(e.h, sub.ctypes.data, len(sub), 1)
        ta = L.emu_last_tri_tests()
        c = L.emu_trace_steps(e.h, sub.ctypes.data, len(sub), 0)
        print("shadow   %-10s any-hit %.2f visits + %.2f tri tests; closest %.2f + %.2f" % (name, a, ta, c, L.emu_last_tri_tests()))


if __name__ == "__main__":
    main()
