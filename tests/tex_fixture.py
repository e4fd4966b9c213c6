"""A small textured scene + texel arrays shared by the texture tests and the golden generator."""
import numpy as np

from restir_embree_b200 import abi
from restir_embree_b200.scenes import SceneArrays, _const_normals, _grid_quads


def textured_scene():
    """Floor (Phong: diffuse, specular and roughness maps), back wall (Lambert: diffuse map), a lamp. The texture
    coordinates run outside [0, 1] and below 0: REPEAT with abs(x % w) and the 1 - v flip are exercised."""
    sc = SceneArrays()
    floor = sc.add_material(abi.MAT_PHONG, (0.6, 0.5, 0.4), (0.2, 0.2, 0.2), (0, 0, 0), 20.0)
    wall = sc.add_material(abi.MAT_LAMBERT, (0.3, 0.6, 0.7), (0, 0, 0), (0, 0, 0), 1.0)
    lamp = sc.add_material(abi.MAT_PHONG, (0.8, 0.8, 0.8), (0, 0, 0), (30.0, 25.0, 12.0), 10.0)
    t = _grid_quads((-3, -3, 0), (6, 0, 0), (0, 6, 0), 3, 3)
    sc.add_surface(t, _const_normals(t, (0, 0, 1)), floor, uv=(t[:, :, :2] * 0.7).astype(np.float32))
    t = _grid_quads((-3, 3, 0), (6, 0, 0), (0, 0, 3), 2, 2)
    sc.add_surface(t, _const_normals(t, (0, -1, 0)), wall, uv=(t[:, :, [0, 2]] * 0.5).astype(np.float32))
    t = _grid_quads((-0.5, -0.5, 2.8), (1, 0, 0), (0, 1, 0), 1, 1)
    sc.add_surface(t, _const_normals(t, (0, 0, -1)), lamp)
    sc.meta = dict(center=(0.0, 0.0, 1.0))
    return sc


def texel_arrays():
    """8-bit B,G,R; 8-bit B,G,R,A; float R,G,B; a float roughness map (0.2 .. 0.8)"""
    rng = np.random.default_rng(5)
    return [rng.integers(0, 256, (7, 5, 3), dtype=np.uint8), rng.integers(0, 256, (4, 6, 4), dtype=np.uint8),
            rng.random((3, 4, 3), dtype=np.float32), (rng.random((5, 5, 3), dtype=np.float32) * 0.6 + 0.2).astype(np.float32)]


SLOTS = {0: dict(diffuse=0, specular=2, shininess=3), 1: dict(diffuse=1)}
N_MATERIALS = 3


def camera_path(f):
    return (2.5 + 0.1 * f, -4.0, 2.0), (0.0, 0.5, 0.8)


# ---- normal maps (Material::kNormalMapSlot, TBN in Intersection::intersectEmbree, P/Intersection.h:25-39) ----------------
def _const_tangents(tris, t):
    return np.broadcast_to(np.asarray(t, dtype=np.float32), tris.shape).copy()


def normal_mapped_scene():
    """textured_scene() with per-vertex tangents: the floor's are neither unit length nor orthogonal to the normal and
    vary per vertex (Gram-Schmidt and the interpolation of attribute slot 3 are exercised), the wall has tangents but no
    normal map, the lamp (an emitter, reached by the BRDF-sampled candidate rays) has a normal map too."""
    sc = SceneArrays()
    floor = sc.add_material(abi.MAT_PHONG, (0.6, 0.5, 0.4), (0.2, 0.2, 0.2), (0, 0, 0), 20.0)
    wall = sc.add_material(abi.MAT_LAMBERT, (0.3, 0.6, 0.7), (0, 0, 0), (0, 0, 0), 1.0)
    lamp = sc.add_material(abi.MAT_PHONG, (0.8, 0.8, 0.8), (0, 0, 0), (30.0, 25.0, 12.0), 10.0)
    box = sc.add_material(abi.MAT_PHONG, (0.5, 0.5, 0.5), (0.3, 0.3, 0.3), (0, 0, 0), 40.0)
    rng = np.random.default_rng(11)
    t = _grid_quads((-3, -3, 0), (6, 0, 0), (0, 6, 0), 3, 3)
    tg = _const_tangents(t, (1.3, 0.2, 0.4)) + (rng.random(t.shape, dtype=np.float32) - 0.5) * 0.3
    sc.add_surface(t, _const_normals(t, (0, 0, 1)), floor, uv=(t[:, :, :2] * 0.7).astype(np.float32), tangent=tg)
    t = _grid_quads((-3, 3, 0), (6, 0, 0), (0, 0, 3), 2, 2)
    sc.add_surface(t, _const_normals(t, (0, -1, 0)), wall, uv=(t[:, :, [0, 2]] * 0.5).astype(np.float32),
                   tangent=_const_tangents(t, (1, 0, 0)))
    t = _grid_quads((-0.5, -0.5, 2.8), (1, 0, 0), (0, 1, 0), 1, 1)
    sc.add_surface(t, _const_normals(t, (0, 0, -1)), lamp, uv=(t[:, :, :2] + 0.5).astype(np.float32),
                   tangent=_const_tangents(t, (0, 1, 0)))
    # a slanted plate (normal-mapped Phong, 8-bit map) seen at a grazing angle: the mapped normal is NOT re-flipped
    # towards the ray, so some of its pixels keep normals facing away from the camera
    t = _grid_quads((1.0, -1.0, 0.0), (1.2, 0.3, 0.0), (0.0, 0.4, 1.1), 2, 2)
    nrm = np.cross(np.array((1.2, 0.3, 0.0)), np.array((0.0, 0.4, 1.1)))
    nrm = (nrm / np.linalg.norm(nrm)).astype(np.float32)
    sc.add_surface(t, _const_normals(t, -nrm), box, uv=(t[:, :, [0, 2]] * 1.3).astype(np.float32),
                   tangent=_const_tangents(t, (1.2, 0.3, 0.0)))
    sc.meta = dict(center=(0.0, 0.0, 1.0))
    return sc


def normal_map_arrays():
    """texel_arrays() + two normal maps: float RGB around (0.5, 0.5, 1) and an 8-bit B,G,R one"""
    rng = np.random.default_rng(17)
    nf = np.empty((6, 7, 3), dtype=np.float32)
    nf[..., :2] = 0.5 + (rng.random((6, 7, 2), dtype=np.float32) - 0.5) * 0.5
    nf[..., 2] = 0.85 + rng.random((6, 7), dtype=np.float32) * 0.15
    n8 = np.empty((5, 4, 3), dtype=np.uint8)
    n8[..., 0] = rng.integers(215, 256, (5, 4))       # B = z
    n8[..., 1:] = rng.integers(96, 160, (5, 4, 2))    # G = y, R = x
    return texel_arrays() + [nf, n8]


NMAP_SLOTS = {0: dict(diffuse=0, specular=2, shininess=3, normal=4), 2: dict(normal=5), 3: dict(normal=5, diffuse=1)}
NMAP_N_MATERIALS = 4


# ---- sky (SphericalMap over a CLAMP_TO_EDGE bilinear Texture, P/SphericalMap.cpp:10-14) ----------------------------------
def sky_arrays():
    """a float R,G,B latitude-longitude map (the .hdr case) and an 8-bit B,G,R,A one"""
    rng = np.random.default_rng(23)
    hdr = (rng.random((9, 16, 3), dtype=np.float32) * np.float32(3.0)).astype(np.float32)
    hdr[0] = 0.0  # a black row at v = 0: sky pixels whose emission is exactly (0, 0, 0) run the light passes like hits do
    ldr = rng.integers(0, 256, (6, 11, 4), dtype=np.uint8)
    return hdr, ldr


def sky_camera_path(f):
    """looks over the floor's edge, up and then down: zenith, horizon and nadir directions, all four atan2 quadrants"""
    return (2.5 + 0.1 * f, -4.0, 2.0), [(0.0, 0.5, 2.6), (-4.0, 3.0, 0.2), (0.5, -3.5, -3.0)][f % 3]
