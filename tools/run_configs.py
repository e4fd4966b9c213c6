#!/usr/bin/env python
"""Runs the BASELINE.json configs other than the bench line through the C ABI on one GPU and prints one JSON line
each (they are parity-test / microbench cases, not the headline):

  python tools/run_configs.py room      configs[0] stand-in: ~200 k triangles, 1280x720, A=32 B=1, temporal + 1 spatial pass
  python tools/run_configs.py 10m       configs[2] on ONE GPU: 10 M triangles, 100 k emitters, 3840x2160, 3 spatial passes k=5
  python tools/run_configs.py rays      configs[3]: 33 M shadow rays against the 10 M-triangle BVH — the rays the path really traces
                                        (visibility / temporal / spatial queues of consecutive 1080p frames), as queued, shuffled
                                        and sorted, through rb_trace_occluded_device; CPU leg: oracle BVH2 on all host cores
  python tools/run_configs.py temporal-branches   CPU only (oracle, a 48-row band of the 1080p orbit): fraction of pixels taking each
                                        temporal-reject branch (P/ReSTIRIntegrator.cpp:644,660,671,686) over the 64-frame orbit
  python tools/run_configs.py bias      static camera, 256 frames each: running mean of ReSTIR (as benchmarked, and without temporal
                                        reuse) against the running mean of the MIS ground-truth estimator (rb_render_mis_frame, N2)
  python tools/run_configs.py textured  the bench scene (1 M triangles, 1080p) four times: plain / + diffuse, specular and roughness
                                        maps / + normal maps / + sky (open ceiling is not modelled: the sky shows through the
                                        camera's misses only) — what rb_set_textures and rb_set_sky cost per frame
  python tools/run_configs.py orbit64   configs[4]: 64-frame orbit at 1080p on the 1 M scene: fps, temporal-reuse statistics,
                                        per-frame relMSE of sampled rows against the CPU oracle (bit-exact => 0)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from restir_embree_b200 import Camera, abi, scenes  # noqa: E402
from restir_embree_b200.renderer import Renderer, make_rays  # noqa: E402  (the library loads without a GPU)


def params(**kw):
    base = dict(M_Area=32, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, spatialReuseNeighborCount=5,
                spatialPassCount=1, spatialReuseRadius=30.0, lightSampler=abi.LS_ALIAS, wavefront=1)
    base.update(kw)
    return abi.default_params(**base)


def frames(r, scene, W, H, n, warm=3, fov=55, cam_fn=None):
    c = scene.meta["center"]
    cam_fn = cam_fn or (lambda t: Camera(W, H, fov, scenes.orbit_position(c, t), c))
    for f in range(warm):
        r.render_frame_device(cam_fn(f), f)
    r.synchronize()
    r.timer_begin()
    for f in range(warm, warm + n):
        r.render_frame_device(cam_fn(f), f)
    ms = r.timer_end()
    t = r.render_frame_device(cam_fn(warm + n), warm + n, want_timings=True)
    return ms / n, t


def run_frames(name, scene_name, W, H, n, **pk):
    t0 = time.time()
    sc = scenes.scene_config(scene_name)
    gen_s = time.time() - t0
    with Renderer(W, H, seed=123) as r:
        st = r.upload_scene(sc)
        r.set_params(params(**pk))
        ms, t = frames(r, sc, W, H, n)
        rays = t["rays_closest"] + t["rays_any_as_written"]
        print(json.dumps({"config": name, "width": W, "height": H, "ms_per_frame": ms, "fps": 1e3 / ms,
                          "mrays_s_as_written": rays / (ms * 1e-3) / 1e6, "rays_traced": t["rays_closest"] + t["rays_any_traced"],
                          "per_pass_ms": {k[3:]: t[k] for k in ("ms_gbuffer", "ms_initial", "ms_visibility", "ms_temporal", "ms_spatial", "ms_shade")},
                          "scene": {k: st[k] for k in ("n_triangles", "n_emissive", "n_bvh_nodes", "bvh_depth", "build_ms")},
                          "scene_generation_s": gen_s}))


def run_rays(scene_name="10m", target=33177600, cpu_sample=400000):
    """configs[3]: shadow-ray-only throughput on the 10M-triangle BVH with the rays the path REALLY traces — the
    visibility / temporal / spatial queues of consecutive 1080p frames (captured with rb_debug_ray_queue between the
    phases of the frame), about 33 M of them = 16 as-written rays per pixel. Orders: as queued (pixel order: coherent
    origins, a pixel's rays adjacent), shuffled (incoherent), and sorted by direction octant + origin Morton cell (what a
    ray-reordering pass could achieve at best). CPU leg beside it: the oracle's BVH2 any-hit traversal
    (Intersection::testOcclusion's arithmetic, P/Intersection.h:43-60) on all host cores over a random sample."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    sc = scenes.scene_config(scene_name)
    W, H = 1920, 1080
    out = {"config": "rays (configs[3])"}
    with Renderer(W, H, seed=123) as r:
        st = r.upload_scene(sc)
        out["scene"] = {k: st[k] for k in ("n_triangles", "n_emissive", "n_bvh_nodes", "bvh_depth", "build_ms")}
        r.set_params(params())
        c = sc.meta["center"]
        parts, mix, f = [], {"visibility": 0, "temporal": 0, "spatial": 0}, 0
        while sum(len(p) for p in parts) < target and f < 8:
            r.frame_begin(Camera(W, H, 55, scenes.orbit_position(c, f), c), f)
            qv, qt = r.debug_ray_queue(1), (r.debug_ray_queue(0) if f > 0 else None)
            r.frame_spatial(0)
            qs = r.debug_ray_queue(0)
            r.frame_end()
            for name, q in (("visibility", qv), ("temporal", qt), ("spatial", qs)):
                if q is not None and len(q):
                    parts.append(q)
                    mix[name] += len(q)
            f += 1
        rays = np.concatenate(parts)[:target]
        n = rays.shape[0]
        out.update(rays=int(n), frames_captured=f, queue_mix=mix)
        rng = np.random.default_rng(3)
        # direction octant (3 bits) + 30-bit Morton code of the origin cell
        lo = rays["org"].min(0)
        ext = np.maximum(rays["org"].max(0) - lo, 1e-6)
        g = np.minimum(((rays["org"] - lo) / ext * 1024).astype(np.uint64), 1023)

        def spread(v):
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            return (v | (v << 2)) & 0x09249249
        morton = spread(g[:, 0]) | (spread(g[:, 1]) << 1) | (spread(g[:, 2]) << 2)
        octant = ((rays["dir"][:, 0] < 0).astype(np.uint64) | ((rays["dir"][:, 1] < 0).astype(np.uint64) << 1) |
                  ((rays["dir"][:, 2] < 0).astype(np.uint64) << 2))
        orders = (("as_queued", None), ("shuffled", rng.permutation(n)), ("sorted_octant_morton", np.argsort((octant << 30) | morton, kind="stable")))
        for label, order in orders:
            rr = rays if order is None else rays[order]
            d_rays = torch.from_numpy(rr.view(np.uint8).reshape(-1)).cuda()
            d_occ = torch.empty(n, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            ms = [r.trace_device(d_rays.data_ptr(), d_occ.data_ptr(), n, True) for _ in range(3)]
            out[label] = {"ms": float(min(ms)), "mrays_s": n / (min(ms) * 1e-3) / 1e6, "occluded_fraction": float(d_occ.float().mean().item())}
            if order is None:
                occ_gpu = d_occ.cpu().numpy()
            del d_rays, d_occ
    # CPU leg: the oracle's BVH2 any-hit query on all host cores, random sample of the same rays
    ob.set_num_threads(len(os.sched_getaffinity(0)))
    o = ob.Oracle(8, 8, tracer=ob.TRACER_BVH2)
    t0 = time.time()
    o.upload_scene(sc)
    build_s = time.time() - t0
    idx = np.sort(rng.choice(n, size=min(cpu_sample, n), replace=False))
    sample = np.ascontiguousarray(rays[idx])
    o.trace_occluded(sample[:20000])  # warm the caches / thread pool
    t0 = time.time()
    occ_cpu = o.trace_occluded(sample)
    dt = time.time() - t0
    out["cpu"] = {"kind": "port (oracle BVH2, OpenMP)", "cores": ob.max_threads(), "rays": int(len(sample)), "ms": dt * 1e3,
                  "mrays_s": len(sample) / dt / 1e6, "occluded_fraction": float(occ_cpu.mean()), "bvh2_build_s": build_s,
                  "agrees_with_gpu": bool(np.array_equal(occ_cpu, occ_gpu[idx]))}
    out["gpu_over_cpu_as_queued"] = out["as_queued"]["mrays_s"] / out["cpu"]["mrays_s"]
    print(json.dumps(out))


def run_modes():
    """the five spatial MIS modes (P/ReSTIRIntegrator.h:19-25) at the bench settings (1080p, 1M triangles, A=32 B=1,
    visibility + temporal + 1 spatial pass k=5), each in the wavefront schedule (CONSTANT: stream / trace / resolve kernels; the
    others: spatial_pixel staged, StagedVis) and with the spatial pass traced inline (RB_STAGED_SPATIAL_OFF=1, round 1's path).
    The reference's own timings of the O(k^2) and O(k) modes: S/s_k_gris_5.png.txt 1.156 s, S/s_k_pair_5.png.txt 0.851 s (640x480,
    A=1 B=1, spatial only, unknown CPU)."""
    sc = scenes.scene_config("1m")
    W, H = 1920, 1080
    names = ["CONSTANT", "BALANCE_HEURISTIC", "PAIRWISE_MIS", "CONSTANT_DEBIAS_Z_TERM", "CONSTANT_DEBIAS_CONTRIB"]
    out = {"config": "spatial MIS modes (N4)", "width": W, "height": H, "modes": {}}
    for staged in (True, False):
        if staged:
            os.environ.pop("RB_STAGED_SPATIAL_OFF", None)
        else:
            os.environ["RB_STAGED_SPATIAL_OFF"] = "1"
        with Renderer(W, H, seed=123) as r:
            r.upload_scene(sc)
            for m, name in enumerate(names):
                if m == 0 and not staged:
                    continue
                r.set_params(params(spatialWeightCalc=m))
                ms, t = frames(r, sc, W, H, 12)
                d = out["modes"].setdefault(name, {})
                key = "wavefront" if staged else "inline_spatial"
                d[key] = {"ms_per_frame": ms, "fps": 1e3 / ms, "ms_spatial_pass": t["ms_spatial"],
                          "rays_any_traced": t["rays_any_traced"], "rays_any_as_written": t["rays_any_as_written"]}
    for name, d in out["modes"].items():
        if "inline_spatial" in d:
            d["speedup_spatial_pass"] = d["inline_spatial"]["ms_spatial_pass"] / d["wavefront"]["ms_spatial_pass"]
    out["reference_seconds_640x480_A1_B1_spatial_only_k5"] = {"CONSTANT (st_k_spatial)": 0.954, "BALANCE_HEURISTIC (s_k_gris_5)": 1.156,
                                                             "PAIRWISE_MIS (s_k_pair_5)": 0.851}
    print(json.dumps(out))


def run_orbit64():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    sc = scenes.scene_config("1m")
    W, H = 1920, 1080
    band, margin = (500, 548), 16  # the oracle renders a band; rows `margin` inside it are valid for the first frames
    p = params()
    c = sc.meta["center"]
    o = ob.Oracle(W, H, seed=123, tracer=ob.TRACER_BVH2)
    o.upload_scene(sc)
    o.set_params(p)
    o.set_band(*band)
    rows = (band[0] + margin, band[1] - margin)
    with Renderer(W, H, seed=123) as r:
        r.upload_scene(sc)
        r.set_params(p)
        worst, ms_all, keep = 0.0, [], []
        acc = None
        for f in range(64):
            cam = Camera(W, H, 55, scenes.orbit_position(c, f), c)
            img, t = r.render_frame(cam, f, want_timings=True)
            ms_all.append(t["ms_total"])
            r.accumulate_display(f, fetch=False, want_stats=False)
            if f < 2:  # (band-edge effects move inwards by the reuse reach every frame)
                ref = o.render_frame(cam, f)
                a, b = img[rows[0]:rows[1]], ref[rows[0]:rows[1]]
                worst = max(worst, float(ob.relmse(a, b)))
        accum = r.readback(abi.BUF_ACCUMULATOR)
        print(json.dumps({"config": "orbit64", "frames": 64, "fps_median": 1e3 / float(np.median(ms_all[3:])),
                          "relmse_vs_oracle_rows_516_532_first_2_frames_max": worst,
                          "accumulator_mean": float(accum.mean()), "accumulator_finite": bool(np.isfinite(accum).all())}))


def run_temporal_branches(frames=64):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    sc = scenes.scene_config("1m")
    W, H = 1920, 1080
    band = (516, 564)
    o = ob.Oracle(W, H, seed=123, tracer=ob.TRACER_BVH2)
    o.upload_scene(sc)
    o.set_params(params())
    o.set_band(*band)
    c = sc.meta["center"]
    o.temporal_stats()
    for f in range(frames):
        o.render_frame(Camera(W, H, 55, scenes.orbit_position(c, f), c), f)
    st = o.temporal_stats()
    total = sum(st.values())
    print(json.dumps({"config": "configs[4] temporal branches (oracle, rows %d-%d, %d frames, 0.5 deg/frame orbit)" % (band[0], band[1], frames),
                      "pixels": total, "fractions": {k: v / total for k, v in st.items()}}))


def run_bias(n=256):
    """What the author did with S/mis_reference*.png.txt vs S/temporal_*.png.txt: image means and relMSE of converged ReSTIR
    images against the converged one-sample-MIS image (unbiased). Pixels that show emitter BACK sides differ by construction
    (evaluateF weights the emitter cosine with abs(), the MIS integrator with max(0, .)); they are few in this scene."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    sc = scenes.scene_config("1m")
    W, H = 1920, 1080
    c = sc.meta["center"]
    cam = Camera(W, H, 55, scenes.orbit_position(c, 0), c)
    out = {"config": "bias", "frames": n}
    with Renderer(W, H, seed=123) as r:
        r.upload_scene(sc)

        def converge(render):
            t0 = time.time()
            for f in range(n):
                render(f)
                r.accumulate_display(f, fetch=False, want_stats=False)
            r.synchronize()
            return r.readback(abi.BUF_ACCUMULATOR).astype(np.float64), (time.time() - t0) / n * 1e3

        r.set_params(params())
        mis, ms = converge(lambda f: r.render_mis_frame(cam, f, fetch=False))
        out["mis_ms_per_frame"] = ms
        mis2, _ = converge(lambda f: r.render_mis_frame(cam, 10000 + f, fetch=False))
        out["mis_mean"] = float(mis.mean())
        out["relmse_mis_vs_mis_other_frames"] = float(ob.relmse(mis2, mis))
        for name, kw in (("restir_bench_settings", {}), ("restir_no_temporal", dict(doTemporalReuse=0)),
                         ("restir_no_reuse", dict(doTemporalReuse=0, doSpatialReuse=0))):
            r.set_params(params(**kw))
            img, ms = converge(lambda f: r.render_frame_device(cam, f))
            out[name] = {"mean": float(img.mean()), "mean_vs_mis": float(img.mean() / mis.mean()), "relmse_vs_mis": float(ob.relmse(img, mis)),
                         "ms_per_frame_incl_accumulate": ms}
    print(json.dumps(out))


def with_uv_and_tangents(sc):
    """the procedural scene with planar texture coordinates (xy, 2 repeats per metre) and one tangent direction that is
    parallel to no face normal of the generator's axis-aligned room and icosphere blobs"""
    out = abi.SceneArrays()
    out.materials = list(sc.materials)
    out.meta = dict(sc.meta)
    t = np.array((1.0, 0.31, 0.17), dtype=np.float32)
    for pos, nrm, m in sc.surfaces:
        out.add_surface(pos, nrm, m, uv=(pos[:, :, :2] * 2.0).astype(np.float32), tangent=np.broadcast_to(t, pos.shape))
    return out


def run_textured(n=20):
    W, H = 1920, 1080
    sc = with_uv_and_tangents(scenes.scene_config("1m"))
    rng = np.random.default_rng(1)
    tex = [rng.integers(0, 256, (512, 512, 3), dtype=np.uint8), rng.random((512, 512, 3), dtype=np.float32),
           (rng.random((256, 256, 3), dtype=np.float32) * 0.5 + 0.3).astype(np.float32)]
    nmap = np.empty((512, 512, 3), dtype=np.float32)
    nmap[..., :2] = 0.5 + (rng.random((512, 512, 2), dtype=np.float32) - 0.5) * 0.4
    nmap[..., 2] = 0.9
    sky = (rng.random((512, 1024, 3), dtype=np.float32) * 2.0).astype(np.float32)
    receivers = [i for i, m in enumerate(sc.materials) if sum(m["emission"]) == 0]
    maps = {i: dict(diffuse=0, specular=1, shininess=2) for i in receivers}
    maps_n = {i: dict(diffuse=0, specular=1, shininess=2, normal=3) for i in receivers}
    out = {"config": "textured (1M triangles, 1080p, bench parameters)", "frames": n}
    with Renderer(W, H, seed=123) as r:
        r.upload_scene(sc)
        r.set_params(params())
        out["plain_ms"] = frames(r, sc, W, H, n)[0]
        r.set_textures(tex + [nmap], maps, len(sc.materials))
        out["maps_ms"] = frames(r, sc, W, H, n)[0]
        r.set_textures(tex + [nmap], maps_n, len(sc.materials))
        out["maps_and_normal_maps_ms"] = frames(r, sc, W, H, n)[0]
        r.set_sky(sky)
        r.set_params(params(useSkybox=1))
        out["maps_normal_maps_sky_ms"] = frames(r, sc, W, H, n)[0]
    print(json.dumps(out))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "room"
    if what == "room":
        run_frames("configs[0] room stand-in", "room", 1280, 720, 20)
    elif what == "10m":
        run_frames("configs[2] on one GPU", "10m", 3840, 2160, 6, spatialPassCount=3)
    elif what == "rays":
        run_rays()
    elif what == "rays-1m":  # the same measurement on the bench scene
        run_rays("1m")
    elif what == "rays-small":  # quick functional check of the measurement itself
        run_rays("small", target=2000000, cpu_sample=100000)
    elif what == "orbit64":
        run_orbit64()
    elif what == "modes":
        run_modes()
    elif what == "textured":
        run_textured()
    elif what == "bias":
        run_bias()
    elif what == "temporal-branches":
        run_temporal_branches()
    else:
        raise SystemExit(__doc__)
