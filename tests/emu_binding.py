"""ctypes wrapper of tests/emu/libemu.so — TEST-ONLY host emulation of the CUDA kernel bodies
(see tests/emu/emu.cpp). Never imported by the product package or bench.py."""
import ctypes as C
import os
import subprocess

import numpy as np

from restir_embree_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "emu.cpp")
LIB_PATH = os.path.join(ROOT, "tests", "emu", "libemu.so")
CSRC = os.path.join(ROOT, "restir_embree_b200", "csrc")


def build_emu_sanitized():
    """the same translation unit with AddressSanitizer + UndefinedBehaviorSanitizer (tests/test_emu_sanitized.py)"""
    out = os.path.join(ROOT, "tests", "emu", "libemu_asan.so")
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    subprocess.check_call(["/usr/bin/g++", "-O1", "-g", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC",
                           "-std=c++17", "-DRB_TRAV_STATS", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-x", "c++", "-shared", "-o", out, SRC])
    return out


def build_emu(force=False):
    if os.environ.get("RB_EMU_LIB"):  # an alternative build of the same source (sanitized), made by the caller
        return
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC",
                           "-std=c++17", "-DRB_TRAV_STATS", "-x", "c++", "-shared", "-o", LIB_PATH, SRC])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_emu()
        L = C.CDLL(os.environ.get("RB_EMU_LIB", LIB_PATH))
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_upload_scene.argtypes = [C.c_void_p, C.POINTER(abi.RbSceneDesc)]
        L.emu_set_params.argtypes = [C.c_void_p, C.POINTER(abi.RbParams)]
        L.emu_scene_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_render_frame.argtypes = [C.c_void_p, C.POINTER(abi.RbCamera), C.c_uint32, C.c_void_p]
        L.emu_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_set_textures.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.emu_set_sky.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_render_mis_frame.argtypes = [C.c_void_p, C.POINTER(abi.RbCamera), C.c_uint32, C.c_uint32, C.c_void_p]
        L.emu_deferred_total.argtypes = [C.c_void_p]
        L.emu_deferred_total.restype = C.c_uint64
        L.emu_readback.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.emu_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.emu_trace_occluded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.emu_validate_bvh.argtypes = [C.c_void_p]
        L.emu_accumulate_display.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.emu_frame_begin.argtypes = [C.c_void_p, C.POINTER(abi.RbCamera), C.c_uint32]
        L.emu_frame_spatial.argtypes = [C.c_void_p, C.c_int]
        L.emu_frame_end.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_halo_rows.argtypes = [C.c_void_p]
        L.emu_horizon_cull_check.argtypes = [C.c_int, C.c_void_p]
        L.emu_calc_I_M.argtypes = [C.c_float, C.c_float, C.c_int]
        L.emu_calc_I_M.restype = C.c_float
        L.emu_set_band.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.emu_halo_export.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.emu_halo_import.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def horizon_cull_check(mode):
    """switch the check mode of initial_pixel's horizon pre-test (process-wide); returns and resets the counts so far"""
    c = np.zeros(5, dtype=np.uint64)
    lib().emu_horizon_cull_check(int(mode), c.ctypes.data)
    return dict(pre_culled=int(c[0]), confirmed=int(c[1]), violations=int(c[2]), candidates=int(c[3]), exact_culled=int(c[4]))


class Emu:
    def __init__(self, width, height, seed=123, band=None):
        self.L = lib()
        self.width, self.height = width, height
        y0, y1 = band if band else (0, height)
        self.band = (y0, y1)
        self.h = self.L.emu_create(width, height, seed, y0, y1)
        self._keep = None

    def close(self):
        if self.h:
            self.L.emu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_scene(self, scene):
        d, keep = scene.desc()
        rc = self.L.emu_upload_scene(self.h, C.byref(d))
        assert rc == 0, rc
        self._keep = keep

    def set_params(self, p):
        assert self.L.emu_set_params(self.h, C.byref(p)) == 0

    def scene_stats(self):
        a = np.zeros(4, dtype=np.uint32)
        self.L.emu_scene_stats(self.h, a.ctypes.data)
        return dict(n_triangles=int(a[0]), n_emissive=int(a[1]), n_bvh_nodes=int(a[2]), bvh_depth=int(a[3]))

    def validate_bvh(self):
        return self.L.emu_validate_bvh(self.h)

    def render_frame(self, cam, frame_idx):
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        out = np.empty((self.height, self.width, 3), dtype=np.float32)
        rc = self.L.emu_render_frame(self.h, C.byref(c), frame_idx, out.ctypes.data)
        assert rc == 0, rc
        return out

    def set_textures(self, textures, slots, n_materials):
        tex, n, per, keep = abi.texture_tables(textures, slots, n_materials)
        assert self.L.emu_set_textures(self.h, tex, n, per, n_materials) == 0

    def set_sky(self, sky):
        if sky is None:
            assert self.L.emu_set_sky(self.h, None) == 0
            return
        tex, keep = abi.sky_table(sky)
        assert self.L.emu_set_sky(self.h, C.byref(tex)) == 0

    def render_mis_frame(self, cam, frame_idx, techniques=3):
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        out = np.empty((self.height, self.width, 3), dtype=np.float32)
        rc = self.L.emu_render_mis_frame(self.h, C.byref(c), frame_idx, techniques, out.ctypes.data)
        assert rc == 0, rc
        return out

    # frame in phases + halo rows: same surface as restir_embree_b200.renderer.Renderer
    def frame_begin(self, cam, frame_idx):
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        assert self.L.emu_frame_begin(self.h, C.byref(c), frame_idx) == 0

    def frame_spatial(self, i):
        assert self.L.emu_frame_spatial(self.h, i) == 0

    def frame_end(self, out=None):
        assert self.L.emu_frame_end(self.h, out.ctypes.data if out is not None else None) == 0

    def set_band(self, y0, y1):
        rc = self.L.emu_set_band(self.h, int(y0), int(y1))
        assert rc == 0, rc
        self.band = (int(y0), int(y1))

    def halo_rows(self):
        return int(self.L.emu_halo_rows(self.h))

    def halo_export(self, y, rows):
        buf = np.empty(rows * self.width * 52, dtype=np.uint8)
        assert self.L.emu_halo_export(self.h, y, rows, buf.ctypes.data) == 0
        return buf

    def halo_import(self, y, rows, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        assert buf.nbytes == rows * self.width * 52
        assert self.L.emu_halo_import(self.h, y, rows, buf.ctypes.data) == 0

    def accumulate_display(self, acc_frame_ctr, tonemap=True, gamma_correct=True):
        out = np.zeros((self.height, self.width, 4), dtype=np.float32)
        st = np.zeros(4, dtype=np.float64)
        assert self.L.emu_accumulate_display(self.h, int(acc_frame_ctr), int(bool(tonemap)), int(bool(gamma_correct)),
                                             out.ctypes.data, st.ctypes.data) == 0
        return out, dict(sum=st[0], sum_sq=st[1], mean=st[2], variance=st[3])

    def deferred_total(self):
        """pixels the banded temporal stream pass handed to its second (re-derivation) launch, over all frames"""
        return int(self.L.emu_deferred_total(self.h))

    def counters(self):
        c = np.zeros(3, dtype=np.uint64)
        self.L.emu_counters(self.h, c.ctypes.data)
        return dict(closest=int(c[0]), any_as_written=int(c[1]), any_traced=int(c[2]))

    def readback(self, buf):
        dt, ch = abi.BUFFER_LAYOUT[buf]
        a = np.empty((self.height, self.width, ch), dtype=dt)
        assert self.L.emu_readback(self.h, buf, a.ctypes.data, a.nbytes) == 0
        return a

    def light_table(self, buf, n):
        a = np.empty(n, dtype=np.uint32 if buf == abi.BUF_ALIAS_IDX else np.float32)
        assert self.L.emu_readback(self.h, buf, a.ctypes.data, a.nbytes) == 0
        return a

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=abi.HIT_DTYPE)
        assert self.L.emu_trace_closest(self.h, rays.ctypes.data, hits.ctypes.data, rays.shape[0]) == 0
        return hits

    def trace_occluded(self, rays):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        occ = np.empty(rays.shape[0], dtype=np.uint8)
        assert self.L.emu_trace_occluded(self.h, rays.ctypes.data, occ.ctypes.data, rays.shape[0]) == 0
        return occ
