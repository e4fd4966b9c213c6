"""Host-side mirror of the reference's frame driver for the ReSTIR path.

`Renderer` plays the role SimpleGuiDX11 plays around produceRestir
(P/simpleguidx11.cpp:359-487): it owns the handle whose device state persists across
frames (reservoir ping-pong + last frame, G-buffer + last frame), takes the scene the way
ModelLoader::loadScene produces it, the ReSTIRIntegrator statics as RbParams, and one Camera
per frame, and returns `frame_data` (linear HDR, float3 per pixel).

All compute happens in restir_embree_b200/librestir_b200.so (hand-written CUDA behind the
C ABI of include/restir_b200.h). There is no CPU fallback: a missing library or a missing
GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librestir_b200.so")

_lib = None


class RestirError(RuntimeError):
    """Raised where the reference's Embree error callback would throw (P/tutorials.cpp:6-24)."""


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    lib_path = os.environ.get("RB_LIB", LIB_PATH)  # RB_LIB: an experimental build of the same CUDA sources
    if not os.path.exists(lib_path):
        raise RestirError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). This package has no CPU fallback.")
    L = C.CDLL(lib_path)
    H = C.c_void_p
    L.rb_abi_version.restype = C.c_uint32
    L.rb_last_error.restype = C.c_char_p
    L.rb_last_error.argtypes = [H]
    L.rb_default_params.argtypes = [C.POINTER(abi.RbParams)]
    L.rb_create.argtypes = [C.POINTER(abi.RbCreateInfo), C.POINTER(H)]
    L.rb_destroy.argtypes = [H]
    L.rb_upload_scene.argtypes = [H, C.POINTER(abi.RbSceneDesc)]
    L.rb_set_params.argtypes = [H, C.POINTER(abi.RbParams)]
    L.rb_set_textures.argtypes = [H, C.POINTER(abi.RbTexture), C.c_uint32, C.POINTER(abi.RbMaterialTextures), C.c_uint32]
    L.rb_set_sky.argtypes = [H, C.POINTER(abi.RbTexture)]
    L.rb_render_frame.argtypes = [H, C.POINTER(abi.RbCamera), C.c_uint32, C.c_void_p, C.POINTER(abi.RbTimings)]
    L.rb_render_frame_device.argtypes = [H, C.POINTER(abi.RbCamera), C.c_uint32, C.c_void_p, C.POINTER(abi.RbTimings)]
    L.rb_render_frame_async.argtypes = [H, C.POINTER(abi.RbCamera), C.c_uint32, C.c_void_p]
    L.rb_frame_wait.argtypes = [H, C.c_uint32]
    L.rb_render_mis_frame.argtypes = [H, C.POINTER(abi.RbCamera), C.c_uint32, C.c_uint32, C.c_void_p]
    L.rb_readback.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t]
    L.rb_synchronize.argtypes = [H]
    L.rb_timer_begin.argtypes = [H]
    L.rb_timer_end.argtypes = [H, C.POINTER(C.c_float)]
    L.rb_trace_closest.argtypes = [H, C.c_void_p, C.c_void_p, C.c_uint32]
    L.rb_trace_occluded.argtypes = [H, C.c_void_p, C.c_void_p, C.c_uint32]
    L.rb_trace_closest_device.argtypes = [H, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
    L.rb_trace_occluded_device.argtypes = [H, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
    L.rb_scene_stats.argtypes = [H, C.POINTER(abi.RbSceneStats)]
    L.rb_comm_init.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t]
    L.rb_comm_unique_id.argtypes = [C.c_void_p, C.c_size_t]
    L.rb_halo_bytes.restype = C.c_size_t
    L.rb_halo_bytes.argtypes = [H, C.c_int32]
    L.rb_halo_export.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p]
    L.rb_halo_import.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p]
    L.rb_comm_transport.restype = C.c_int32
    L.rb_comm_transport.argtypes = [H]
    L.rb_halo_rows.restype = C.c_int32
    L.rb_halo_rows.argtypes = [H]
    L.rb_frame_begin.argtypes = [H, C.POINTER(abi.RbCamera), C.c_uint32]
    L.rb_frame_spatial.argtypes = [H, C.c_int32]
    L.rb_set_band.argtypes = [H, C.c_int32, C.c_int32]
    L.rb_get_band.argtypes = [H, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.rb_frame_end.argtypes = [H, C.c_void_p, C.POINTER(abi.RbTimings)]
    L.rb_debug_ray_queue.argtypes = [H, C.c_int32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.rb_accumulate_display.argtypes = [H, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(abi.RbImageStats)]
    if L.rb_abi_version() != 2:
        raise RestirError("librestir_b200.so ABI version mismatch")
    _lib = L
    return L


def _timings_dict(t):
    d = {}
    for k, _ in abi.RbTimings._fields_:
        if k == "reserved":
            continue
        v = getattr(t, k)
        d[k] = list(v) if hasattr(v, "__len__") else v
    return d


class Renderer:
    def __init__(self, width, height, device=0, seed=123, band=None, collect_timings=True):
        self.L = load_library()
        self.width, self.height = int(width), int(height)
        info = abi.RbCreateInfo()
        info.width, info.height, info.device, info.seed = self.width, self.height, int(device), int(seed)
        info.band_y0, info.band_y1 = band if band is not None else (0, self.height)
        info.collect_timings = 1 if collect_timings else 0
        self.band = (info.band_y0, info.band_y1)
        self.h = C.c_void_p()
        rc = self.L.rb_create(C.byref(info), C.byref(self.h))
        if rc != abi.RB_OK:
            msg = self.L.rb_last_error(None)
            self.h = None
            raise RestirError(f"rb_create failed ({rc}): {msg.decode() if msg else ''}")
        self._scene_keep = None
        self.params = abi.default_params()

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.rb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != abi.RB_OK:
            msg = self.L.rb_last_error(self.h)
            raise RestirError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    # -- scene / params ---------------------------------------------------------------
    def upload_scene(self, scene):
        d, keep = scene.desc()
        self._check(self.L.rb_upload_scene(self.h, C.byref(d)), "rb_upload_scene")
        self._scene_keep = None  # arrays are copied by the library
        del keep
        return self.scene_stats()

    def scene_stats(self):
        s = abi.RbSceneStats()
        self._check(self.L.rb_scene_stats(self.h, C.byref(s)), "rb_scene_stats")
        return dict(n_triangles=s.n_triangles, n_emissive=s.n_emissive, n_bvh_nodes=s.n_bvh_nodes,
                    bvh_depth=s.bvh_depth, build_ms=s.build_ms, total_emissive_area=s.total_emissive_area,
                    bounds_lo=tuple(s.bounds_lo), bounds_hi=tuple(s.bounds_hi))

    def set_params(self, p):
        self._check(self.L.rb_set_params(self.h, C.byref(p)), "rb_set_params")
        self.params = p

    def set_textures(self, textures, slots, n_materials):
        """Material::set_texture for the uploaded scene: `textures` = texel arrays ([h, w, 3|4] uint8 B,G,R[,A] or float32
        R,G,B[,A]), `slots` = {material index: dict(diffuse=, specular=, shininess=)} (abi.texture_tables)."""
        tex, n, per, keep = abi.texture_tables(textures, slots, n_materials)
        self._check(self.L.rb_set_textures(self.h, tex, n, per, n_materials), "rb_set_textures")

    def set_sky(self, sky):
        """Scene::setSkybox (P/Scene.cpp:47-50): `sky` = the decoded image's texel array ([h, w, 3|4] float32 R,G,B[,A] or
        uint8 B,G,R[,A]) or None to remove it. Needed before set_params(useSkybox=1)."""
        if sky is None:
            self._check(self.L.rb_set_sky(self.h, None), "rb_set_sky")
            return
        tex, keep = abi.sky_table(sky)
        self._check(self.L.rb_set_sky(self.h, C.byref(tex)), "rb_set_sky")

    # -- frames -------------------------------------------------------------------------
    def render_frame(self, cam, frame_idx, out=None, want_timings=False, fetch=True):
        """produceRestir: returns frame_data [h, w, 3] float32 (host). `out` may be a pinned numpy view."""
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        t = abi.RbTimings() if want_timings else None
        ptr = None
        if fetch:
            if out is None:
                out = np.zeros((self.height, self.width, 3), dtype=np.float32)
            assert out.dtype == np.float32 and out.size == self.width * self.height * 3 and out.flags["C_CONTIGUOUS"]
            ptr = out.ctypes.data
        rc = self.L.rb_render_frame(self.h, C.byref(c), int(frame_idx), ptr, C.byref(t) if t is not None else None)
        self._check(rc, "rb_render_frame")
        if want_timings:
            return out, _timings_dict(t)
        return out

    def render_frame_device(self, cam, frame_idx, dev_ptr=None, want_timings=False):
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        t = abi.RbTimings() if want_timings else None
        rc = self.L.rb_render_frame_device(self.h, C.byref(c), int(frame_idx), dev_ptr,
                                           C.byref(t) if t is not None else None)
        self._check(rc, "rb_render_frame_device")
        if want_timings:
            return _timings_dict(t)
        return None

    def render_frame_async(self, cam, frame_idx, out):
        """rb_render_frame_async: issue the frame and the copy of its rows into `out` (host array, pinned for a truly
        asynchronous copy); returns at once. frame_wait(k) blocks until at most k such frames are in flight."""
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        assert out.dtype == np.float32 and out.size == self.width * self.height * 3 and out.flags["C_CONTIGUOUS"]
        self._check(self.L.rb_render_frame_async(self.h, C.byref(c), int(frame_idx), out.ctypes.data), "rb_render_frame_async")

    def frame_wait(self, frames_in_flight=0):
        self._check(self.L.rb_frame_wait(self.h, int(frames_in_flight)), "rb_frame_wait")

    def render_mis_frame(self, cam, frame_idx, techniques=3, fetch=True):
        """One frame of the reference's ground-truth estimator (one-sample MIS direct lighting: NEEPathIntegrator with DI
        only around DirectMISIntegrator, P/DirectMISIntegrator.cpp:18-144). techniques: bit 0 = sample the BRDF, bit 1 =
        sample the light sources. Leaves the ReSTIR state alone; accumulate_display() converges it."""
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        out = np.zeros((self.height, self.width, 3), dtype=np.float32) if fetch else None
        rc = self.L.rb_render_mis_frame(self.h, C.byref(c), int(frame_idx), int(techniques), out.ctypes.data if fetch else None)
        self._check(rc, "rb_render_mis_frame")
        return out

    def accumulate_display(self, acc_frame_ctr, tonemap=True, gamma_correct=True, fetch=False, want_stats=True):
        """The Producer loop's step after produceRestir (P/simpleguidx11.cpp:246-326): accumulate the frame just
        rendered, tonemap + gamma-compress into display_data, image mean / variance. Returns (display or None, stats)."""
        out = np.zeros((self.height, self.width, 4), dtype=np.float32) if fetch else None
        st = abi.RbImageStats() if want_stats else None
        rc = self.L.rb_accumulate_display(self.h, int(acc_frame_ctr), int(bool(tonemap)), int(bool(gamma_correct)),
                                          out.ctypes.data if fetch else None, C.byref(st) if want_stats else None)
        self._check(rc, "rb_accumulate_display")
        stats = dict(sum=st.sum, sum_sq=st.sum_sq, mean=st.mean, variance=st.variance, pixels=st.pixels) if want_stats else None
        return out, stats

    def synchronize(self):
        self._check(self.L.rb_synchronize(self.h), "rb_synchronize")

    def timer_begin(self):
        self._check(self.L.rb_timer_begin(self.h), "rb_timer_begin")

    def timer_end(self):
        ms = C.c_float(0)
        self._check(self.L.rb_timer_end(self.h, C.byref(ms)), "rb_timer_end")
        return ms.value

    def readback(self, buf):
        dt, ch = abi.BUFFER_LAYOUT[buf]
        a = np.empty((self.height, self.width, ch), dtype=dt)
        self._check(self.L.rb_readback(self.h, buf, a.ctypes.data, a.nbytes), "rb_readback")
        return a

    def light_table(self, buf):
        n = self.scene_stats()["n_emissive"]
        a = np.empty(n, dtype=np.uint32 if buf == abi.BUF_ALIAS_IDX else np.float32)
        self._check(self.L.rb_readback(self.h, buf, a.ctypes.data, a.nbytes), "rb_readback")
        return a

    # -- ray seam -------------------------------------------------------------------------
    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=abi.HIT_DTYPE)
        self._check(self.L.rb_trace_closest(self.h, rays.ctypes.data, hits.ctypes.data, rays.shape[0]), "rb_trace_closest")
        return hits

    def trace_occluded(self, rays):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        occ = np.empty(rays.shape[0], dtype=np.uint8)
        self._check(self.L.rb_trace_occluded(self.h, rays.ctypes.data, occ.ctypes.data, rays.shape[0]), "rb_trace_occluded")
        return occ

    def trace_device(self, rays_ptr, out_ptr, n, any_hit):
        ms = C.c_float(0)
        fn = self.L.rb_trace_occluded_device if any_hit else self.L.rb_trace_closest_device
        self._check(fn(self.h, rays_ptr, out_ptr, int(n), C.byref(ms)), "rb_trace_*_device")
        return ms.value

    # -- frame in phases (hosts that move the halo rows themselves) ------------------------------------
    def frame_begin(self, cam, frame_idx):
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        self._check(self.L.rb_frame_begin(self.h, C.byref(c), int(frame_idx)), "rb_frame_begin")

    def frame_spatial(self, i):
        self._check(self.L.rb_frame_spatial(self.h, int(i)), "rb_frame_spatial")

    def frame_end(self, out=None):
        ptr = out.ctypes.data if out is not None else None
        self._check(self.L.rb_frame_end(self.h, ptr, None), "rb_frame_end")

    def debug_ray_queue(self, which):
        """The rays of the open frame's current queue (which=0: temporal after frame_begin / spatial after frame_spatial)
        or of its visibility pass (which=1), RTCRay layout, queue order (rb_debug_ray_queue)."""
        n = C.c_uint32(0)
        self._check(self.L.rb_debug_ray_queue(self.h, int(which), None, 0, C.byref(n)), "rb_debug_ray_queue")
        rays = np.zeros(n.value, dtype=abi.RAY_DTYPE)
        if n.value:
            self._check(self.L.rb_debug_ray_queue(self.h, int(which), rays.ctypes.data, n.value, C.byref(n)), "rb_debug_ray_queue")
        return rays

    def set_band(self, y0, y1):
        self._check(self.L.rb_set_band(self.h, int(y0), int(y1)), "rb_set_band")
        self.band = (int(y0), int(y1))

    def get_band(self):
        """Rows the next frame renders (they move when the library balances the bands, rb_comm_init)."""
        a, b = C.c_int32(0), C.c_int32(0)
        self._check(self.L.rb_get_band(self.h, C.byref(a), C.byref(b)), "rb_get_band")
        return a.value, b.value

    def halo_rows(self):
        return int(self.L.rb_halo_rows(self.h))

    def comm_init(self, rank, nranks, unique_id_bytes):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id_bytes)[:128])
        self._check(self.L.rb_comm_init(self.h, int(rank), int(nranks), buf, 128), "rb_comm_init")

    def comm_transport(self):
        """How rb_render_frame moves the halo rows: 'peer-memory' (CUDA IPC over NVLink), 'nccl', or None."""
        return {0: None, 1: "peer-memory", 2: "nccl"}[int(self.L.rb_comm_transport(self.h))]

    # -- band halos -------------------------------------------------------------------------
    def halo_export(self, y, rows):
        buf = np.empty(self.L.rb_halo_bytes(self.h, rows), dtype=np.uint8)
        self._check(self.L.rb_halo_export(self.h, y, rows, buf.ctypes.data), "rb_halo_export")
        return buf

    def halo_import(self, y, rows, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        assert buf.nbytes == self.L.rb_halo_bytes(self.h, rows)
        self._check(self.L.rb_halo_import(self.h, y, rows, buf.ctypes.data), "rb_halo_import")


class MultiRenderer:
    """One handle, several GPUs, one host thread (rb_multi_*): the frame driver of `Renderer` over N devices. The scene
    goes to every device, a frame is issued band by band from the calling thread, halo rows travel over peer memory,
    and render_frame returns the ASSEMBLED frame_data. `devices` may name the same ordinal several times (several
    bands on one GPU: what the single-GPU tests use)."""

    def __init__(self, width, height, devices, seed=123):
        self.L = L = load_library()
        M = C.c_void_p
        L.rb_multi_create.argtypes = [C.POINTER(abi.RbCreateInfo), C.POINTER(C.c_int32), C.c_int32, C.POINTER(M)]
        L.rb_multi_destroy.argtypes = [M]
        L.rb_multi_last_error.restype = C.c_char_p
        L.rb_multi_last_error.argtypes = [M]
        L.rb_multi_device_count.argtypes = [M]
        L.rb_multi_member.restype = C.c_void_p
        L.rb_multi_member.argtypes = [M, C.c_int32]
        L.rb_multi_upload_scene.argtypes = [M, C.POINTER(abi.RbSceneDesc)]
        L.rb_multi_set_params.argtypes = [M, C.POINTER(abi.RbParams)]
        L.rb_multi_render_frame.argtypes = [M, C.POINTER(abi.RbCamera), C.c_uint32, C.c_void_p]
        L.rb_multi_render_frame_async.argtypes = [M, C.POINTER(abi.RbCamera), C.c_uint32, C.c_void_p]
        L.rb_multi_frame_wait.argtypes = [M, C.c_uint32]
        L.rb_multi_synchronize.argtypes = [M]
        L.rb_multi_readback.argtypes = [M, C.c_int, C.c_void_p, C.c_size_t]
        L.rb_multi_accumulate_display.argtypes = [M, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(abi.RbImageStats)]
        self.width, self.height = int(width), int(height)
        info = abi.RbCreateInfo()
        info.width, info.height, info.seed, info.collect_timings = self.width, self.height, int(seed), 0
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        self.m = C.c_void_p()
        rc = L.rb_multi_create(C.byref(info), devs, len(devices), C.byref(self.m))
        if rc != abi.RB_OK:
            msg = L.rb_multi_last_error(None)
            self.m = None
            raise RestirError(f"rb_multi_create failed ({rc}): {msg.decode() if msg else ''}")
        self.n = len(devices)

    def close(self):
        if getattr(self, "m", None):
            self.L.rb_multi_destroy(self.m)
            self.m = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != abi.RB_OK:
            msg = self.L.rb_multi_last_error(self.m)
            raise RestirError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def upload_scene(self, scene):
        d, keep = scene.desc()
        self._check(self.L.rb_multi_upload_scene(self.m, C.byref(d)), "rb_multi_upload_scene")
        del keep

    def set_params(self, p):
        self._check(self.L.rb_multi_set_params(self.m, C.byref(p)), "rb_multi_set_params")

    def bands(self):
        out = []
        for i in range(self.n):
            a, b = C.c_int32(0), C.c_int32(0)
            self.L.rb_get_band(C.c_void_p(self.L.rb_multi_member(self.m, i)), C.byref(a), C.byref(b))
            out.append((a.value, b.value))
        return out

    def render_frame(self, cam, frame_idx, out=None):
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        if out is None:
            out = np.zeros((self.height, self.width, 3), dtype=np.float32)
        self._check(self.L.rb_multi_render_frame(self.m, C.byref(c), int(frame_idx), out.ctypes.data), "rb_multi_render_frame")
        return out

    def render_frame_async(self, cam, frame_idx, out):
        c = cam.to_abi() if hasattr(cam, "to_abi") else cam
        self._check(self.L.rb_multi_render_frame_async(self.m, C.byref(c), int(frame_idx), out.ctypes.data), "rb_multi_render_frame_async")

    def frame_wait(self, frames_in_flight=0):
        self._check(self.L.rb_multi_frame_wait(self.m, int(frames_in_flight)), "rb_multi_frame_wait")

    def synchronize(self):
        self._check(self.L.rb_multi_synchronize(self.m), "rb_multi_synchronize")

    def readback(self, buf):
        dt, ch = abi.BUFFER_LAYOUT[buf]
        a = np.empty((self.height, self.width, ch), dtype=dt)
        self._check(self.L.rb_multi_readback(self.m, buf, a.ctypes.data, a.nbytes), "rb_multi_readback")
        return a

    def accumulate_display(self, acc_frame_ctr, tonemap=True, gamma_correct=True):
        st = abi.RbImageStats()
        self._check(self.L.rb_multi_accumulate_display(self.m, int(acc_frame_ctr), int(bool(tonemap)), int(bool(gamma_correct)), None,
                                                       C.byref(st)), "rb_multi_accumulate_display")
        return dict(sum=st.sum, sum_sq=st.sum_sq, mean=st.mean, variance=st.variance, pixels=st.pixels)


def load_obj_scene(path, gamma_correct=True):
    """Wavefront OBJ + MTL -> SceneArrays through the library's own parser (rb_obj_load: the conventions of the
    reference's ModelLoader, P/ModelLoader.cpp:41-321). Host-only: works without a GPU. meta carries the material names
    and the texture file names of the MTL (the host decodes them and calls set_textures); surfaces carry uv, and tangents when a
    material names a normal map."""
    L = load_library()
    L.rb_obj_load.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    L.rb_obj_scene_desc.restype = C.POINTER(abi.RbSceneDesc)
    L.rb_obj_scene_desc.argtypes = [C.c_void_p]
    L.rb_obj_material_name.restype = C.c_char_p
    L.rb_obj_material_name.argtypes = [C.c_void_p, C.c_uint32]
    L.rb_obj_texture_name.restype = C.c_char_p
    L.rb_obj_texture_name.argtypes = [C.c_void_p, C.c_uint32, C.c_int32]
    L.rb_obj_free.argtypes = [C.c_void_p]
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = L.rb_obj_load(os.fsencode(path), int(bool(gamma_correct)), C.byref(h), err, len(err))
    if rc != abi.RB_OK:
        raise RestirError(f"rb_obj_load failed ({rc}): {err.value.decode(errors='replace')}")
    try:
        d = L.rb_obj_scene_desc(h).contents
        sc = abi.SceneArrays()
        names, textures = [], []
        for i in range(d.n_materials):
            m = d.materials[i]
            sc.add_material(int(m.type), tuple(m.diffuse), tuple(m.specular), tuple(m.emission), float(m.shininess), float(m.ior))
            names.append(L.rb_obj_material_name(h, i).decode())
            textures.append([L.rb_obj_texture_name(h, i, s).decode() for s in range(4)])
        uvs = []
        for i in range(d.n_surfaces):
            sf = d.surfaces[i]
            n = int(sf.n_tris)
            pos = np.ctypeslib.as_array(sf.pos, shape=(n, 3, 3)).copy()
            nrm = np.ctypeslib.as_array(sf.normal, shape=(n, 3, 3)).copy()
            uvs.append(np.ctypeslib.as_array(sf.uv, shape=(n, 3, 2)).copy())
            tan = np.ctypeslib.as_array(sf.tangent, shape=(n, 3, 3)).copy() if sf.tangent else None
            sc.add_surface(pos, nrm, int(sf.material), uv=uvs[-1], tangent=tan)
        lo = np.min([s[0].reshape(-1, 3).min(0) for s in sc.surfaces], axis=0)
        hi = np.max([s[0].reshape(-1, 3).max(0) for s in sc.surfaces], axis=0)
        sc.meta = dict(kind="obj", path=str(path), material_names=names, texture_names=textures, uv=uvs,
                       center=tuple(float(v) for v in (lo + hi) / 2), bounds=(lo.tolist(), hi.tolist()))
        return sc
    finally:
        L.rb_obj_free(h)


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0); 128 bytes to hand to every rank's Renderer.comm_init."""
    L = load_library()
    buf = (C.c_char * 128)()
    rc = L.rb_comm_unique_id(buf, 128)
    if rc != abi.RB_OK:
        raise RestirError(f"rb_comm_unique_id failed ({rc})")
    return bytes(buf)


def band_rows(height, nranks, rank):
    """Rows [y0, y1) of band `rank` of `nranks` contiguous horizontal bands of ceil(height/nranks) rows."""
    rows = (height + nranks - 1) // nranks
    return min(height, rank * rows), min(height, (rank + 1) * rows)


def exchange_halos(renderers):
    """Single-process band emulation: move the halo rows between neighbouring handles through host memory
    (what rb_render_frame does with NCCL send/recv when rb_comm_init was called). `renderers` top to bottom."""
    if len(renderers) < 2:
        return
    R = renderers[0].halo_rows()
    for up, dn in zip(renderers[:-1], renderers[1:]):
        y = up.band[1]
        assert dn.band[0] == y
        r_up = min(R, up.band[1] - up.band[0])
        r_dn = min(R, dn.band[1] - dn.band[0])
        a = up.halo_export(y - r_up, r_up)      # bottom rows of the upper band -> lower band's top halo
        b = dn.halo_export(y, r_dn)             # top rows of the lower band -> upper band's bottom halo
        dn.halo_import(y - r_up, r_up, a)
        up.halo_import(y, r_dn, b)


def make_rays(org, target=None, direction=None, tnear=0.01, tfar=None, tfar_offset=0.001):
    """Rays in the RTCRay layout. With `target`: shadow rays as Intersection::testOcclusion builds them
    (P/Intersection.h:43-60): dir = normalize(to-from), tnear = FLT_MIN + 0.01, tfar = dist - 0.001."""
    org = np.asarray(org, dtype=np.float32).reshape(-1, 3)
    n = org.shape[0]
    rays = np.zeros(n, dtype=abi.RAY_DTYPE)
    rays["org"] = org
    if target is not None:
        d = np.asarray(target, dtype=np.float32).reshape(-1, 3) - org
        dist = np.sqrt((d * d).sum(1, dtype=np.float32), dtype=np.float32)
        rays["dir"] = d * (np.float32(1.0) / dist)[:, None]
        rays["tfar"] = dist - np.float32(tfar_offset)
    else:
        rays["dir"] = np.asarray(direction, dtype=np.float32).reshape(-1, 3)
        rays["tfar"] = np.float32(3.4028235e38) if tfar is None else np.asarray(tfar, dtype=np.float32)
    rays["tnear"] = np.float32(1.17549435e-38) + np.float32(tnear)
    return rays
