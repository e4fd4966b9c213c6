#!/usr/bin/env python
"""How many of the shadow rays the wavefront frame really traces are occluded, per queue (CPU, host emulation of the
kernel bodies — the queues are the device's, ray for ray): decides whether an any-hit traversal order of its own is
worth building (an occluded ray stops at its first blocker, an unoccluded one visits everything either way).

  python tools/queue_stats.py [scene=1m] [width=480] [frames=3]
Queues of a frame, in order: closest_emissive + any_precedes (BRDF candidates), any (visibility pass), any (temporal), any (spatial)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(name, W, n):
    import emu_binding as eb
    from restir_embree_b200 import Camera, abi, scenes
    H = W * 9 // 16
    sc = scenes.scene_config(name)
    e = eb.Emu(W, H, seed=123)
    e.upload_scene(sc)
    e.set_params(abi.default_params(M_Area=32, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1,
                                    spatialReuseNeighborCount=5, spatialPassCount=1, spatialReuseRadius=30.0,
                                    lightSampler=abi.LS_ALIAS, wavefront=1))
    c = sc.meta["center"]
    for f in range(n):
        sys.stderr.write(f"[emu queue] frame {f}\n")
        sys.stderr.flush()
        e.render_frame(Camera(W, H, 55, scenes.orbit_position(c, f), c), f)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "1m"
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 480
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    env = dict(os.environ, EMU_QUEUE_STATS="1")
    out = subprocess.run([sys.executable, __file__, "--child", name, str(W), str(n)], env=env, stderr=subprocess.PIPE, text=True).stderr
    px = W * (W * 9 // 16)
    for line in out.splitlines():
        m = re.match(r"\[emu queue\] (\w+) rays (\d+) occluded (\d+)", line)
        if m:
            q, r, o = m.group(1), int(m.group(2)), int(m.group(3))
            extra = "" if q.startswith("closest") else "  occluded %5.1f %%" % (100.0 * o / max(r, 1))
            print("  %-17s %9d rays = %5.2f per pixel%s" % (q, r, r / px, extra))
        elif line.startswith("[emu queue] frame"):
            print(line[12:])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    else:
        main()
