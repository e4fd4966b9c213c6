"""One-process-per-band driver over torch.distributed (gloo on CPU in the tests): the halo rows of the phase API
are sent to / received from rank-1 and rank+1 with isend/irecv, exactly the pattern rb_render_frame runs with
ncclSend/ncclRecv on the GPUs. Works with any handle exposing the phase surface (Emu on CPU, Renderer on GPU)."""
import numpy as np
import torch
import torch.distributed as dist


class DistBandDriver:
    def __init__(self, handle, rank, world):
        self.r, self.rank, self.world = handle, rank, world

    def exchange_halos(self):
        r, R = self.r, self.r.halo_rows()
        y0, y1 = r.band
        reqs, recvs = [], []
        rows = min(R, y1 - y0)
        if self.rank > 0:
            up = torch.from_numpy(r.halo_export(y0, rows))
            reqs.append(dist.isend(up, self.rank - 1))
            buf = torch.empty(min(R, y0) * r.width * 52, dtype=torch.uint8)
            reqs.append(dist.irecv(buf, self.rank - 1))
            recvs.append((y0 - min(R, y0), min(R, y0), buf))
        if self.rank + 1 < self.world:
            dn = torch.from_numpy(r.halo_export(y1 - rows, rows))
            reqs.append(dist.isend(dn, self.rank + 1))
            n = min(R, r.height - y1)
            buf = torch.empty(n * r.width * 52, dtype=torch.uint8)
            reqs.append(dist.irecv(buf, self.rank + 1))
            recvs.append((y1, n, buf))
        for q in reqs:
            q.wait()
        for y, n, buf in recvs:
            r.halo_import(y, n, buf.numpy())

    def render(self, cam, frame_idx, params):
        r = self.r
        r.frame_begin(cam, frame_idx)
        if params.doSpatialReuse:
            for i in range(params.spatialPassCount):
                self.exchange_halos()
                r.frame_spatial(i)
        out = np.zeros((r.height, r.width, 3), dtype=np.float32)
        r.frame_end(out)
        # gather the bands on rank 0 (display / metrics only; off the critical path)
        t = torch.from_numpy(out)
        dist.reduce(t, 0, op=dist.ReduceOp.SUM)  # bands are disjoint and zero elsewhere
        return t.numpy()
