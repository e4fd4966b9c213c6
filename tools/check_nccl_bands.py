#!/usr/bin/env python
"""Multi-GPU check, run under torchrun with N >= 2 ranks on one node:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_nccl_bands.py
Every rank renders its band through rb_render_frame with the in-library NCCL halo exchange; rank 0 also renders the
whole frame on a single handle. The gathered bands must be bit-identical to it (SURVEY §8e invariance)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from restir_embree_b200 import Camera, abi, scenes  # noqa: E402
from restir_embree_b200.renderer import Renderer, band_rows, comm_unique_id  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = 640, 360
    sc = scenes.scene_config("small")
    p = abi.default_params(M_Area=8, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1,
                           spatialPassCount=2, lightSampler=abi.LS_ALIAS, wavefront=1)
    r = Renderer(W, H, device=local, seed=9, band=band_rows(H, world, rank))
    r.upload_scene(sc)
    r.set_params(p)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    os.environ.setdefault("RB_BAL_PERIOD", "2")  # move the boundaries often: this is a correctness check
    r.comm_init(rank, world, idt.cpu().numpy().tobytes())
    full = None
    if rank == 0:
        print("halo transport:", r.comm_transport(), flush=True)
        full = Renderer(W, H, device=local, seed=9)
        full.upload_scene(sc)
        full.set_params(p)
    ok = True
    bands_seen = []
    for f in range(int(os.environ.get("CHECK_FRAMES", "16"))):
        cam = Camera(W, H, 60, (4.2 + 0.1 * f, -4.4, 1.8 + 0.2 * f), (0, 0, 1.0 + 0.15 * f))
        band = torch.tensor(r.get_band(), dtype=torch.int32, device="cuda")  # the library balances the bands: they move
        allb = [torch.zeros_like(band) for _ in range(world)]
        dist.all_gather(allb, band)
        bands_seen.append([tuple(int(v) for v in b.tolist()) for b in allb])
        img = torch.from_numpy(r.render_frame(cam, f)).cuda()
        r.accumulate_display(f, want_stats=False)  # the running mean must follow the rows when boundaries move
        dist.all_reduce(img)  # bands are disjoint, zero elsewhere
        if rank == 0:
            ref = full.render_frame(cam, f)
            full.accumulate_display(f, want_stats=False)
            got = img.cpu().numpy()
            same = np.array_equal(ref.view(np.uint32), got.view(np.uint32)) or np.array_equal(ref, got)
            bb = bands_seen[-1]
            contiguous = bb[0][0] == 0 and bb[-1][1] == H and all(a[1] == b[0] for a, b in zip(bb[:-1], bb[1:]))
            print(f"frame {f}: {world} bands {bb} over NCCL vs one band: {'bit-identical' if same else 'DIFFERENT'} "
                  f"({(ref != got).any(-1).sum()} px differ){'' if contiguous else ' BANDS NOT CONTIGUOUS'}", flush=True)
            same = same and contiguous
            ok &= same
    # the accumulator (rb_accumulate_display, N1): every rank contributes the rows it owns NOW
    y0, y1 = r.get_band()
    acc = np.zeros((H, W, 3), dtype=np.float32)
    acc[y0:y1] = r.readback(abi.BUF_ACCUMULATOR)[y0:y1]
    acc_t = torch.from_numpy(acc).cuda()
    dist.all_reduce(acc_t)
    if rank == 0:
        ref_acc = full.readback(abi.BUF_ACCUMULATOR)
        got_acc = acc_t.cpu().numpy()
        same = np.array_equal(ref_acc.view(np.uint32), got_acc.view(np.uint32)) or np.array_equal(ref_acc, got_acc)
        moved = len({tuple(b) for b in bands_seen})
        print(f"accumulator after {len(bands_seen)} frames ({moved} different band layouts): "
              f"{'bit-identical' if same else 'DIFFERENT'} ({(ref_acc != got_acc).any(-1).sum()} px differ)", flush=True)
        ok &= same
    dist.barrier()
    r.close()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
