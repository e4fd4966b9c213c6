"""Host-side band driver used by the tests: N handles (CUDA `Renderer`s or emulated `Emu`s — same phase
surface), one per horizontal band, stepped in lockstep with halo rows moved through host memory. This is the
single-process stand-in for what rb_render_frame does with NCCL send/recv once rb_comm_init was called."""
import numpy as np

from restir_embree_b200.renderer import band_rows, exchange_halos


def make_bands(factory, width, height, n, **kw):
    return [factory(width, height, band=band_rows(height, n, r), **kw) for r in range(n)]


def render_banded(handles, cam, frame_idx, params):
    """One frame over all bands; returns the assembled image."""
    h, w = handles[0].height, handles[0].width
    for r in handles:
        r.frame_begin(cam, frame_idx)
    if params.doSpatialReuse:
        for i in range(params.spatialPassCount):
            exchange_halos(handles)
            for r in handles:
                r.frame_spatial(i)
    out = np.zeros((h, w, 3), dtype=np.float32)
    tmp = np.zeros((h, w, 3), dtype=np.float32)
    for r in handles:
        r.frame_end(tmp)
        y0, y1 = r.band
        out[y0:y1] = tmp[y0:y1]
    return out


def assemble(handles, buf):
    a = None
    for r in handles:
        x = r.readback(buf)
        if a is None:
            a = np.zeros_like(x)
        y0, y1 = r.band
        a[y0:y1] = x[y0:y1]
    return a


def move_boundaries(handles, new_bounds):
    """Load balancing between frames: ship the LAST frame's reservoirs of the rows that change owner (halo_export /
    halo_import outside a frame act on them), then move the bands. new_bounds = [0, b1, ..., height]."""
    for k in range(len(handles) - 1):
        up, dn = handles[k], handles[k + 1]
        old, new = up.band[1], new_bounds[k + 1]
        if new > old:    # the upper band grows downwards: rows [old, new) come from the lower band
            up.halo_import(old, new - old, dn.halo_export(old, new - old))
        elif new < old:  # the lower band grows upwards
            dn.halo_import(new, old - new, up.halo_export(new, old - new))
    for k, r in enumerate(handles):
        r.set_band(new_bounds[k], new_bounds[k + 1])
