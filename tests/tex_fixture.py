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
