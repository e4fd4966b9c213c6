"""ctypes wrapper of oracle/_ref/libref.so — the reference's OWN ReSTIR sources compiled in place
(oracle/ref_shim). TEST INFRASTRUCTURE; exists only where /root/reference was available at build time
(the built .so travels to the GPU box, the sources do not)."""
import ctypes as C
import os

import numpy as np

from restir_embree_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libref.so")
ORACLE_BOOST_PATH = os.path.join(ROOT, "oracle", "_ref", "liboracle_boost.so")


def available():
    return os.path.exists(LIB_PATH) and os.path.exists(ORACLE_BOOST_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        R = C.CDLL(LIB_PATH)
        R.ref_create.restype = C.c_void_p
        R.ref_create.argtypes = [C.c_int, C.c_int, C.POINTER(abi.RbSceneDesc)]
        R.ref_set_params.argtypes = [C.c_void_p, C.POINTER(abi.RbParams)]
        R.ref_camera.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(abi.RbCamera)]
        R.ref_produce_restir.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_produce_mis.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_set_textures.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        R.ref_set_sky.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_reservoirs.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_gbuffer.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_seed.argtypes = [C.c_uint32]
        R.ref_random.restype = C.c_float
        R.ref_sampleDiskUniform.argtypes = [C.c_float, C.c_void_p]
        R.ref_sampleTriangle.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_phong_evalBRDF.argtypes = [C.c_void_p] * 4
        R.ref_phong_evalPdf.restype = C.c_float
        R.ref_phong_evalPdf.argtypes = [C.c_void_p] * 3
        R.ref_phong_sampleBRDF.argtypes = [C.c_void_p] * 3
        R.ref_cdf_pick.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_sanitize.argtypes = [C.c_void_p]
        R.ref_aces.argtypes = [C.c_void_p]
        R.ref_compress.restype = C.c_float
        R.ref_compress.argtypes = [C.c_float]
        R.ref_accumulate_mix.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib = R
    return _lib


class Reference:
    """The reference's ReSTIRIntegrator run in its deterministic (serial, _DEBUG-build) order."""

    def __init__(self, width, height, scene):
        self.R = lib()
        self.width, self.height = width, height
        d, self._keep = scene.desc()
        self.h = self.R.ref_create(width, height, C.byref(d))

    def set_params(self, p):
        self.R.ref_set_params(self.h, C.byref(p))

    def camera(self, fov_deg, view_from, view_at):
        f = np.asarray(view_from, dtype=np.float32)
        a = np.asarray(view_at, dtype=np.float32)
        cam = abi.RbCamera()
        self.R.ref_camera(self.h, float(fov_deg), f.ctypes.data, a.ctypes.data, C.byref(cam))
        return cam

    def produce_restir(self):
        out = np.zeros((self.height, self.width, 3), dtype=np.float32)
        self.R.ref_produce_restir(self.h, out.ctypes.data)
        return out

    def set_textures(self, textures, slots, n_materials):
        """Material::set_texture with the reference's own Texture objects (P/Texture.cpp compiled in place)"""
        tex, n, per, keep = abi.texture_tables(textures, slots, n_materials)
        self.R.ref_set_textures(self.h, tex, n, per, n_materials)

    def set_sky(self, sky):
        """Scene::skybox = SphericalMap over the reference's own Texture (BILINEAR, CLAMP_TO_EDGE), members filled from `sky`"""
        tex, keep = abi.sky_table(sky)
        self.R.ref_set_sky(self.h, C.byref(tex))

    def produce_mis(self):
        """N2: Raytracer::get_pixel over the image with NEEPathIntegrator (DI only) + the reference's DirectMISIntegrator"""
        out = np.zeros((self.height, self.width, 3), dtype=np.float32)
        self.R.ref_produce_mis(self.h, out.ctypes.data)
        return out

    def reservoirs(self):
        out = np.zeros((self.height, self.width, 12), dtype=np.float32)
        self.R.ref_reservoirs(self.h, out.ctypes.data)
        return out

    def gbuffer(self):
        out = np.zeros((self.height, self.width, 18), dtype=np.float32)
        self.R.ref_gbuffer(self.h, out.ctypes.data)
        return out
