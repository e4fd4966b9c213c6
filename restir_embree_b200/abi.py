"""ctypes mirror of include/restir_b200.h (the C ABI of the hot path).

Every structure here has the same field order and types as the C header; the
layout is asserted against the shared library at load time (rb_abi_version and
sizeof checks in tests/test_abi.py).
"""
import ctypes as C

import numpy as np

RB_OK = 0

# RbMaterialType — P/enums.h:3-11
MAT_NORMAL, MAT_LAMBERT, MAT_PHONG, MAT_MIRROR, MAT_DIELECTRIC, MAT_DIELECTRIC_TRANSPARENT, MAT_UNSUPPORTED = range(7)
# RbSpatialWeightCalc — P/ReSTIRIntegrator.h:19-25
SW_CONSTANT, SW_CONSTANT_DEBIAS_CONTRIB, SW_CONSTANT_DEBIAS_Z_TERM, SW_BALANCE_HEURISTIC, SW_PAIRWISE_MIS = range(5)
LS_CDF, LS_ALIAS = 0, 1

(BUF_GBUF_POS_DEPTH, BUF_GBUF_NORMAL_SHIN, BUF_GBUF_DIFFUSE_IIM, BUF_GBUF_SPEC_TYPE, BUF_GBUF_EMISSION, BUF_HIT_IDS,
 BUF_RES_POINT_WSUM, BUF_RES_NORMAL_W, BUF_RES_LI_CONF, BUF_RES_LIGHT_IDX, BUF_FRAME_RGB, BUF_ALIAS_PROB, BUF_ALIAS_IDX,
 BUF_LIGHT_CDF, BUF_ACCUMULATOR, BUF_DISPLAY) = range(16)

# (dtype, channels) of each readback buffer, per pixel unless noted
BUFFER_LAYOUT = {
    BUF_GBUF_POS_DEPTH: (np.float32, 4),
    BUF_GBUF_NORMAL_SHIN: (np.float32, 4),
    BUF_GBUF_DIFFUSE_IIM: (np.float32, 4),
    BUF_GBUF_SPEC_TYPE: (np.float32, 4),
    BUF_GBUF_EMISSION: (np.float32, 4),
    BUF_HIT_IDS: (np.uint32, 2),
    BUF_RES_POINT_WSUM: (np.float32, 4),
    BUF_RES_NORMAL_W: (np.float32, 4),
    BUF_RES_LI_CONF: (np.float32, 4),
    BUF_RES_LIGHT_IDX: (np.int32, 1),
    BUF_FRAME_RGB: (np.float32, 3),
    BUF_ACCUMULATOR: (np.float32, 3),
    BUF_DISPLAY: (np.float32, 4),
}


class RbMaterial(C.Structure):
    _fields_ = [("type", C.c_uint32), ("diffuse", C.c_float * 3), ("specular", C.c_float * 3),
                ("emission", C.c_float * 3), ("shininess", C.c_float), ("ior", C.c_float)]


class RbSurface(C.Structure):
    _fields_ = [("n_tris", C.c_uint32), ("material", C.c_uint32), ("pos", C.POINTER(C.c_float)),
                ("normal", C.POINTER(C.c_float)), ("uv", C.POINTER(C.c_float)), ("tangent", C.POINTER(C.c_float))]


class RbTexture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("scan_width", C.c_int32), ("pixel_size", C.c_int32),
                ("data", C.c_void_p)]


class RbMaterialTextures(C.Structure):
    _fields_ = [("diffuse", C.c_int32), ("specular", C.c_int32), ("shininess", C.c_int32), ("normal", C.c_int32)]


def texture_tables(textures, slots, n_materials):
    """textures: list of numpy arrays [h, w, 3|4] uint8 (B,G,R[,A] as FreeImage delivers them) or float32 (R,G,B[,A]);
    slots: {material index: dict(diffuse=i, specular=i, shininess=i)}. Returns (RbTexture[], n, RbMaterialTextures[], keep)."""
    import numpy as _np
    keep = [_np.ascontiguousarray(t) for t in textures]
    tex = (RbTexture * max(len(keep), 1))()
    for i, a in enumerate(keep):
        assert a.ndim == 3 and a.dtype in (_np.uint8, _np.float32) and a.shape[2] in (3, 4)
        tex[i].height, tex[i].width = a.shape[0], a.shape[1]
        tex[i].pixel_size = a.shape[2] * a.dtype.itemsize
        tex[i].scan_width = a.strides[0]
        tex[i].data = a.ctypes.data
    per = (RbMaterialTextures * n_materials)()
    for m in range(n_materials):
        d = slots.get(m, {})
        per[m].diffuse, per[m].specular = d.get("diffuse", -1), d.get("specular", -1)
        per[m].shininess, per[m].normal = d.get("shininess", -1), d.get("normal", -1)
    return tex, len(keep), per, keep


def sky_table(sky):
    """sky: numpy array [h, w, 3|4] float32 (R,G,B[,A], the usual .hdr / .exr case) or uint8 (B,G,R[,A]). Returns
    (RbTexture, keep)."""
    tex, _, _, keep = texture_tables([sky], {}, 0)
    return tex[0], (tex, keep)


class RbSceneDesc(C.Structure):
    _fields_ = [("n_surfaces", C.c_uint32), ("surfaces", C.POINTER(RbSurface)), ("n_materials", C.c_uint32),
                ("materials", C.POINTER(RbMaterial))]


class RbParams(C.Structure):
    _fields_ = [("M_Area", C.c_int32), ("M_Brdf", C.c_int32), ("spatialReuseNeighborCount", C.c_int32),
                ("spatialPassCount", C.c_int32), ("confidenceCap", C.c_int32), ("spatialReuseRadius", C.c_float),
                ("minNormalSimilarity", C.c_float), ("maxDepthDifference", C.c_float), ("doSpatialReuse", C.c_int32),
                ("doTemporalReuse", C.c_int32), ("doVisibilityPass", C.c_int32),
                ("rejectDissimilarNeighbors", C.c_int32), ("spatialWeightCalc", C.c_int32),
                ("tnearOffset", C.c_float), ("tfarOffset", C.c_float), ("normalOffset", C.c_float),
                ("bgColor", C.c_float * 3), ("useSkybox", C.c_int32), ("lightSampler", C.c_int32),
                ("wavefront", C.c_int32), ("temporalFetchReprojected", C.c_int32)]


def default_params(**overrides):
    """Reference defaults (P/ReSTIRIntegrator.cpp:13-35, P/RenderParams.h:8-17) except useSkybox=0
    (it needs a sky texture first: rb_set_sky)."""
    p = RbParams()
    p.M_Area, p.M_Brdf = 1, 1
    p.spatialReuseNeighborCount, p.spatialPassCount, p.confidenceCap = 5, 1, 20
    p.spatialReuseRadius, p.minNormalSimilarity, p.maxDepthDifference = 30.0, 0.85, 0.2
    p.doSpatialReuse = p.doTemporalReuse = p.doVisibilityPass = p.rejectDissimilarNeighbors = 0
    p.spatialWeightCalc = SW_CONSTANT
    p.tnearOffset, p.tfarOffset, p.normalOffset = 0.01, 0.001, 0.001
    p.bgColor[0] = p.bgColor[1] = p.bgColor[2] = 0.5
    p.useSkybox = 0
    p.lightSampler = LS_CDF
    p.wavefront = 0
    p.temporalFetchReprojected = 0
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def copy_params(p, **overrides):
    q = RbParams()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(RbParams))
    for k, v in overrides.items():
        if not hasattr(q, k):
            raise AttributeError(k)
        setattr(q, k, v)
    return q


class RbCamera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("focal_px", C.c_float), ("viewMat", C.c_float * 16),
                ("invViewMat", C.c_float * 16)]


class RbTimings(C.Structure):
    _fields_ = [("ms_gbuffer", C.c_float), ("ms_initial", C.c_float), ("ms_visibility", C.c_float),
                ("ms_temporal", C.c_float), ("ms_spatial", C.c_float), ("ms_shade", C.c_float),
                ("ms_total", C.c_float), ("ms_trace_any", C.c_float), ("ms_stream", C.c_float * 6),
                ("ms_trace", C.c_float * 6), ("rays_closest", C.c_uint64),
                ("rays_any_as_written", C.c_uint64), ("rays_any_traced", C.c_uint64), ("kernel_launches", C.c_uint32),
                ("ms_halo", C.c_float)]


class RbImageStats(C.Structure):
    _fields_ = [("sum", C.c_double), ("sum_sq", C.c_double), ("mean", C.c_double), ("variance", C.c_double),
                ("pixels", C.c_uint64)]


class RbCreateInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint32),
                ("band_y0", C.c_int32), ("band_y1", C.c_int32), ("collect_timings", C.c_int32),
                ("reserved", C.c_int32)]


class RbRay(C.Structure):
    _fields_ = [("org_x", C.c_float), ("org_y", C.c_float), ("org_z", C.c_float), ("tnear", C.c_float),
                ("dir_x", C.c_float), ("dir_y", C.c_float), ("dir_z", C.c_float), ("time", C.c_float),
                ("tfar", C.c_float), ("mask", C.c_uint32), ("id", C.c_uint32), ("flags", C.c_uint32)]


class RbHit(C.Structure):
    _fields_ = [("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("primID", C.c_uint32), ("geomID", C.c_uint32)]


class RbSceneStats(C.Structure):
    _fields_ = [("n_triangles", C.c_uint32), ("n_emissive", C.c_uint32), ("n_bvh_nodes", C.c_uint32),
                ("bvh_depth", C.c_uint32), ("build_ms", C.c_float), ("total_emissive_area", C.c_float),
                ("bounds_lo", C.c_float * 3), ("bounds_hi", C.c_float * 3)]


RAY_DTYPE = np.dtype([("org", np.float32, 3), ("tnear", np.float32), ("dir", np.float32, 3), ("time", np.float32),
                      ("tfar", np.float32), ("mask", np.uint32), ("id", np.uint32), ("flags", np.uint32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("primID", np.uint32),
                      ("geomID", np.uint32)])
assert RAY_DTYPE.itemsize == C.sizeof(RbRay) == 48
assert HIT_DTYPE.itemsize == C.sizeof(RbHit) == 20

# Every symbol include/restir_b200.h declares (checked by tests/test_abi.py).
EXPORTED_SYMBOLS = [
    "rb_abi_version", "rb_last_error", "rb_default_params", "rb_create", "rb_destroy", "rb_upload_scene",
    "rb_set_params", "rb_set_textures", "rb_set_sky", "rb_render_frame", "rb_render_frame_device", "rb_render_frame_async", "rb_frame_wait", "rb_render_mis_frame", "rb_readback", "rb_synchronize", "rb_timer_begin", "rb_timer_end",
    "rb_trace_closest",
    "rb_trace_occluded", "rb_trace_closest_device", "rb_trace_occluded_device", "rb_scene_stats", "rb_comm_init",
    "rb_comm_unique_id", "rb_comm_transport", "rb_debug_balance_step", "rb_debug_ray_queue", "rb_multi_create", "rb_multi_destroy", "rb_multi_last_error", "rb_multi_device_count", "rb_multi_member", "rb_multi_upload_scene", "rb_multi_set_params", "rb_multi_set_textures", "rb_multi_set_sky", "rb_multi_render_frame", "rb_multi_render_frame_async", "rb_multi_frame_wait", "rb_multi_synchronize", "rb_multi_readback", "rb_multi_accumulate_display", "rb_obj_load", "rb_obj_scene_desc",
    "rb_obj_material_name", "rb_obj_texture_name", "rb_obj_free", "rb_halo_bytes", "rb_halo_export", "rb_halo_import", "rb_halo_rows", "rb_frame_begin",
    "rb_frame_spatial", "rb_frame_end", "rb_accumulate_display", "rb_set_band", "rb_get_band",
]


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class SceneArrays:
    """Host-side scene in the layout ModelLoader::loadScene hands Embree (P/ModelLoader.cpp:227-318):
    a list of surfaces, each a non-indexed triangle soup with per-vertex normals and one material."""

    def __init__(self):
        self.materials = []  # dicts: type, diffuse, specular, emission, shininess, ior
        self.surfaces = []   # (pos[n,3,3] f32, normal[n,3,3] f32, material_index)
        self.uvs = {}        # surface index -> uv[n,3,2] f32 (attribute slot 1; textured materials)
        self.tangents = {}   # surface index -> tangent[n,3,3] f32 (attribute slot 3; normal-mapped materials)
        self.meta = {}

    def add_material(self, type=MAT_PHONG, diffuse=(0.5, 0.5, 0.5), specular=(0.0, 0.0, 0.0), emission=(0, 0, 0),
                     shininess=10.0, ior=1.0):
        self.materials.append(dict(type=type, diffuse=tuple(map(float, diffuse)), specular=tuple(map(float, specular)),
                                   emission=tuple(map(float, emission)), shininess=float(shininess), ior=float(ior)))
        return len(self.materials) - 1

    def add_surface(self, pos, normal, material, uv=None, tangent=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3, 3)
        normal = np.ascontiguousarray(normal, dtype=np.float32).reshape(-1, 3, 3)
        assert pos.shape == normal.shape and pos.shape[0] > 0
        assert 0 <= material < len(self.materials)
        self.surfaces.append((pos, normal, int(material)))
        if uv is not None:  # texture coordinates [n, 3, 2] (attribute slot 1 of the reference's Embree geometry)
            uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 3, 2)
            assert uv.shape[0] == pos.shape[0]
            self.uvs[len(self.surfaces) - 1] = uv
        if tangent is not None:  # per-vertex tangents [n, 3, 3] (attribute slot 3; read by normal-mapped materials)
            tangent = np.ascontiguousarray(tangent, dtype=np.float32).reshape(-1, 3, 3)
            assert tangent.shape == pos.shape
            self.tangents[len(self.surfaces) - 1] = tangent
        return len(self.surfaces) - 1

    @property
    def n_triangles(self):
        return sum(s[0].shape[0] for s in self.surfaces)

    @property
    def n_emissive(self):
        return sum(s[0].shape[0] for s in self.surfaces if sum(self.materials[s[2]]["emission"]) > 0)

    def desc(self):
        """Returns (RbSceneDesc, keepalive) — keepalive owns the ctypes arrays."""
        mats = (RbMaterial * len(self.materials))()
        for i, m in enumerate(self.materials):
            mats[i].type = m["type"]
            for c in range(3):
                mats[i].diffuse[c] = m["diffuse"][c]
                mats[i].specular[c] = m["specular"][c]
                mats[i].emission[c] = m["emission"][c]
            mats[i].shininess = m["shininess"]
            mats[i].ior = m["ior"]
        surfs = (RbSurface * len(self.surfaces))()
        for i, (pos, nrm, mat) in enumerate(self.surfaces):
            surfs[i].n_tris = pos.shape[0]
            surfs[i].material = mat
            surfs[i].pos = fptr(pos)
            surfs[i].normal = fptr(nrm)
            uv = self.uvs.get(i)
            surfs[i].uv = fptr(uv) if uv is not None else None
            tg = self.tangents.get(i)
            surfs[i].tangent = fptr(tg) if tg is not None else None
        d = RbSceneDesc()
        d.n_surfaces = len(self.surfaces)
        d.surfaces = surfs
        d.n_materials = len(self.materials)
        d.materials = mats
        return d, (mats, surfs, self)
