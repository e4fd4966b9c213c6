"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module. The product package (restir_embree_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from restir_embree_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "liboracle.so")

RNG_COUNTER, RNG_LEGACY = 0, 1
MATH_DET, MATH_LIBM = 0, 1
TRACER_BRUTE, TRACER_BVH2 = 0, 1


def build_oracle():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build_oracle()
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_band.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_set_row_segments.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_upload_scene.argtypes = [C.c_void_p, C.POINTER(abi.RbSceneDesc)]
        L.orc_set_params.argtypes = [C.c_void_p, C.POINTER(abi.RbParams)]
        L.orc_render_frame.argtypes = [C.c_void_p, C.POINTER(abi.RbCamera), C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_set_textures.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_set_sky.argtypes = [C.c_void_p, C.c_void_p]
        for name in ("orc_dm_atan2", "orc_dm_acos"):
            getattr(L, name).restype = C.c_float
        L.orc_dm_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_dm_acos.argtypes = [C.c_float]
        L.orc_temporal_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_render_mis_frame.argtypes = [C.c_void_p, C.POINTER(abi.RbCamera), C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_readback.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_accumulate_display.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_frame.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_aces.argtypes = [C.c_void_p]
        L.orc_compress.restype = C.c_float
        L.orc_compress.argtypes = [C.c_int, C.c_float]
        L.orc_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_num_emissive.argtypes = [C.c_void_p]
        L.orc_num_emissive.restype = C.c_uint32
        L.orc_num_triangles.argtypes = [C.c_void_p]
        L.orc_num_triangles.restype = C.c_uint32
        L.orc_total_emissive_area.argtypes = [C.c_void_p]
        L.orc_total_emissive_area.restype = C.c_float
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_trace_occluded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        for name in ("orc_dm_sin", "orc_dm_cos", "orc_dm_exp", "orc_dm_lgamma"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_float]
        L.orc_dm_pow.restype = C.c_float
        L.orc_dm_pow.argtypes = [C.c_float, C.c_float]
        L.orc_dm_ibeta.restype = C.c_float
        L.orc_dm_ibeta.argtypes = [C.c_float, C.c_float, C.c_float]
        L.orc_rng_bits.restype = C.c_uint32
        L.orc_rng_bits.argtypes = [C.c_uint32] * 6
        L.orc_calc_I_M.restype = C.c_float
        L.orc_calc_I_M.argtypes = [C.c_int, C.c_float, C.c_float]
        L.orc_phong_evalBRDF.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_phong_evalPdf.restype = C.c_float
        L.orc_phong_evalPdf.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_legacy_phong_sampleBRDF.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_legacy_sampleDiskUniform.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_float, C.c_void_p]
        L.orc_legacy_sampleTriangle.argtypes = [C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_legacy_cdf_pick.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.orc_legacy_floats.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        L.orc_have_boost.restype = C.c_int
        _lib = L
    return _lib


class Oracle:
    def __init__(self, width, height, seed=123, rng=RNG_COUNTER, math=MATH_DET, tracer=TRACER_BVH2, cache_iim=1):
        self.L = lib()
        self.width, self.height = width, height
        self.h = self.L.orc_create(width, height, seed, rng, math, tracer, cache_iim)
        self._keep = None

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_band(self, y0, y1):
        assert self.L.orc_set_band(self.h, y0, y1) == 0

    def set_row_segments(self, segments):
        """bench sampling: render only these [y0, y1) row segments of every frame"""
        a = np.ascontiguousarray(segments, dtype=np.int32).reshape(-1, 2)
        assert self.L.orc_set_row_segments(self.h, a.ctypes.data, a.shape[0]) == 0

    def upload_scene(self, scene):
        d, keep = scene.desc()
        rc = self.L.orc_upload_scene(self.h, C.byref(d))
        assert rc == 0, rc
        self._keep = keep

    def set_params(self, p):
        rc = self.L.orc_set_params(self.h, C.byref(p))
        assert rc == 0, rc

    def render_frame(self, cam, frame_idx, want_times=False):
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        out = np.empty((self.height, self.width, 3), dtype=np.float32)
        times = np.zeros(8, dtype=np.float64)
        rc = self.L.orc_render_frame(self.h, C.byref(c), frame_idx, out.ctypes.data, times.ctypes.data)
        assert rc == 0, rc
        return (out, times) if want_times else out

    def render_mis_frame(self, cam, frame_idx, techniques=3):
        """N2: one frame of the reference's one-sample MIS direct-lighting estimator (NEEPathIntegrator, DI only)"""
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        out = np.empty((self.height, self.width, 3), dtype=np.float32)
        rc = self.L.orc_render_mis_frame(self.h, C.byref(c), frame_idx, techniques, out.ctypes.data)
        assert rc == 0, rc
        return out

    def set_textures(self, textures, slots, n_materials):
        tex, n, per, keep = abi.texture_tables(textures, slots, n_materials)
        assert self.L.orc_set_textures(self.h, tex, n, per, n_materials) == 0

    def set_sky(self, sky):
        if sky is None:
            assert self.L.orc_set_sky(self.h, None) == 0
            return
        tex, keep = abi.sky_table(sky)
        assert self.L.orc_set_sky(self.h, C.byref(tex)) == 0

    def temporal_stats(self, reset=True):
        """pixels per outcome of the temporal pass since the last reset: backward reprojection failed, depth test at the
        reprojected pixel failed, forward reprojection failed, depth test at the forward-reprojected pixel failed, merged"""
        c = np.zeros(5, dtype=np.uint64)
        self.L.orc_temporal_stats(self.h, c.ctypes.data, int(bool(reset)))
        keys = ("reproject_backward_failed", "depth_backward_failed", "reproject_forward_failed", "depth_forward_failed", "merged")
        return dict(zip(keys, (int(v) for v in c)))

    def set_frame(self, rgb):
        a = np.ascontiguousarray(rgb, dtype=np.float32)
        assert a.size == self.width * self.height * 3
        assert self.L.orc_set_frame(self.h, a.ctypes.data) == 0

    def accumulate_display(self, acc_frame_ctr, tonemap=True, gamma_correct=True):
        out = np.zeros((self.height, self.width, 4), dtype=np.float32)
        st = np.zeros(4, dtype=np.float64)
        rc = self.L.orc_accumulate_display(self.h, int(acc_frame_ctr), int(bool(tonemap)), int(bool(gamma_correct)),
                                           out.ctypes.data, st.ctypes.data)
        assert rc == 0, rc
        return out, dict(sum=st[0], sum_sq=st[1], mean=st[2], variance=st[3])

    def readback(self, buf):
        dt, ch = abi.BUFFER_LAYOUT[buf]
        a = np.empty((self.height, self.width, ch), dtype=dt)
        rc = self.L.orc_readback(self.h, buf, a.ctypes.data, a.nbytes)
        assert rc == 0, rc
        return a

    def light_table(self, buf):
        n = self.L.orc_num_emissive(self.h)
        a = np.empty(n, dtype=np.uint32 if buf == abi.BUF_ALIAS_IDX else np.float32)
        assert self.L.orc_readback(self.h, buf, a.ctypes.data, a.nbytes) == 0
        return a

    def counters(self, reset=True):
        c = np.zeros(3, dtype=np.uint64)
        self.L.orc_counters(self.h, c.ctypes.data, int(reset))
        return dict(closest=int(c[0]), any_as_written=int(c[1]), any_traced=int(c[2]))

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=abi.HIT_DTYPE)
        assert self.L.orc_trace_closest(self.h, rays.ctypes.data, hits.ctypes.data, rays.shape[0]) == 0
        return hits

    def trace_occluded(self, rays):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        occ = np.empty(rays.shape[0], dtype=np.uint8)
        assert self.L.orc_trace_occluded(self.h, rays.ctypes.data, occ.ctypes.data, rays.shape[0]) == 0
        return occ


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def max_threads():
    return int(lib().orc_max_threads())


def relmse(img, ref):
    """mean_px( |I-R|^2 / (|R|^2 + 1e-2) ) on linear HDR (our definition, the reference has none)."""
    img = np.asarray(img, dtype=np.float64).reshape(-1, 3)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1, 3)
    num = ((img - ref) ** 2).sum(1)
    den = (ref ** 2).sum(1) + 1e-2
    return float((num / den).mean())
