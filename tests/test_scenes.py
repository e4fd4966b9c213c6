import numpy as np

from restir_embree_b200 import Camera, scenes


def test_counts_are_exact_and_deterministic():
    a = scenes.scene_config("small")
    b = scenes.scene_config("small")
    assert a.n_triangles == 20000 and a.n_emissive == 200
    for (p, n, m), (q, o, k) in zip(a.surfaces, b.surfaces):
        assert m == k and np.array_equal(p, q) and np.array_equal(n, o)
    t = scenes.scene_config("tiny")
    assert t.n_emissive == 8 and t.n_triangles == 280


def test_one_million_triangle_config():
    s = scenes.scene_config("1m")
    assert s.n_triangles == 1_000_000 and s.n_emissive == 10_000
    for p, n, _ in s.surfaces:
        assert np.isfinite(p).all() and np.isfinite(n).all()


def test_camera_matches_glm_lookat_conventions():
    c = Camera(1280, 720, 55, (1.877986, -7.724095, 1.602229), (0, 0, 0))  # P/tutorials.cpp:35
    V = c.viewMat.T.astype(np.float64)  # math (row-major) view of the column-major glm matrix
    eye = np.append(c.view_from.astype(np.float64), 1.0)
    assert np.allclose(V @ eye, [0, 0, 0, 1], atol=1e-5)          # eye maps to the origin
    at = V @ np.array([0, 0, 0, 1.0])
    assert abs(at[0]) < 1e-5 and abs(at[1]) < 1e-5 and at[2] < 0   # looks down -z
    assert np.allclose(c.invViewMat.T.astype(np.float64) @ V, np.eye(4), atol=1e-5)
    assert abs(float(c.f_y) - 720 / (2 * np.tan(np.radians(55) / 2))) < 1e-2
