"""Seeded procedural scenes in the reference's scene layout.

The reference's own geometry (room.obj / living_room.obj) is not in its checkout
(git-ignored, .MISSING_LARGE_BLOBS), so every BASELINE config runs on a synthetic
stand-in: a closed z-up room with tessellated walls, a field of displaced
icosphere "blobs" with smooth normals, small emissive quads under the ceiling and
small emissive spheres. Materials follow D/room/room.mtl conventions: receivers
are Phong (Pc 2) with kd in U(0.2,0.8)^3, ks in [0.04,0.5], n in {5,20,80,250};
emitters use Ke (100, 80.9, 29.8) and (100,100,100) (room.mtl:51,93).

Triangle and emitter counts are hit EXACTLY so "1M triangles / 10k emitters" is
literal.
"""
import numpy as np

from .abi import MAT_LAMBERT, MAT_PHONG, SceneArrays

f32 = np.float32


def _icosphere(level):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
        ab, bc, ca = (a + b) / 2, (b + c) / 2, (c + a) / 2
        tri = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1),
                              np.stack([ab, bc, ca], 1)], 0)
        v = tri.reshape(-1, 3)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        f = np.arange(v.shape[0]).reshape(-1, 3)
    return v[f]  # [n_tris, 3, 3] unit-sphere positions (soup)


_ICO_CACHE = {}


def icosphere(level):
    if level not in _ICO_CACHE:
        _ICO_CACHE[level] = _icosphere(level)
    return _ICO_CACHE[level]


def _grid_quads(origin, du, dv, nu, nv):
    """nu x nv quads spanning origin + [0,1] du + [0,1] dv -> [2*nu*nv, 3, 3]."""
    origin, du, dv = (np.asarray(a, dtype=np.float64) for a in (origin, du, dv))
    iu, iv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    iu, iv = iu.reshape(-1, 1), iv.reshape(-1, 1)
    p00 = origin + (iu / nu) * du + (iv / nv) * dv
    p10 = origin + ((iu + 1) / nu) * du + (iv / nv) * dv
    p01 = origin + (iu / nu) * du + ((iv + 1) / nv) * dv
    p11 = origin + ((iu + 1) / nu) * du + ((iv + 1) / nv) * dv
    t0 = np.stack([p00, p10, p11], 1)
    t1 = np.stack([p00, p11, p01], 1)
    return np.concatenate([t0, t1], 0)


def _const_normals(tris, n):
    return np.broadcast_to(np.asarray(n, dtype=f32), tris.shape).copy()


def _blobs(rng, level, count, lo, hi, rmin, rmax, lumpy=0.15):
    """count displaced icospheres; returns positions [count*nt,3,3], normals (sphere normals)."""
    base = icosphere(level)  # [nt,3,3]
    nt = base.shape[0]
    centers = rng.uniform(lo, hi, size=(count, 3))
    radii = np.exp(rng.uniform(np.log(rmin), np.log(rmax), size=count))
    freq = rng.uniform(2.0, 6.0, size=(count, 3))
    phase = rng.uniform(0, 2 * np.pi, size=(count, 3))
    d = base[None, :, :, :]  # unit directions
    wob = 1.0 + lumpy * np.sin(d[..., 0] * freq[:, None, None, 0] + phase[:, None, None, 0]) * np.sin(
        d[..., 1] * freq[:, None, None, 1] + phase[:, None, None, 1]) * np.sin(
            d[..., 2] * freq[:, None, None, 2] + phase[:, None, None, 2])
    pos = centers[:, None, None, :] + radii[:, None, None, None] * wob[..., None] * d
    nrm = np.broadcast_to(d, pos.shape)
    return pos.reshape(count * nt, 3, 3).astype(f32), nrm.reshape(count * nt, 3, 3).astype(f32)


ICO_TRIS = {5: 20480, 4: 5120, 3: 1280, 2: 320, 1: 80, 0: 20}


def shell_tri_count(g, hx, H):
    gz = max(1, int(round(g * H / (2 * hx))))
    return 2 * (g * g * 2) + 4 * (g * gz * 2)


def make_room_scene(n_triangles, n_emissive, seed=0xB200, half_extent=(10.0, 10.0), height=6.0, wall_grid=None,
                    lambert_fraction=0.0, max_level=4):
    """Closed room + blob field + emitters with exactly n_triangles triangles, n_emissive of them emissive."""
    assert n_emissive % 2 == 0 and n_emissive >= 2 and n_triangles > n_emissive + 12
    rng = np.random.default_rng(seed)
    sc = SceneArrays()
    hx, hy = half_extent
    H = height

    # ---- materials -----------------------------------------------------------------
    shin = [5.0, 20.0, 80.0, 250.0]
    room_mats = []
    for i in range(6):
        kd = rng.uniform(0.35, 0.8, 3)
        ks = rng.uniform(0.04, 0.2)
        room_mats.append(sc.add_material(MAT_PHONG, kd, (ks, ks, ks), (0, 0, 0), shin[i % 4]))
    blob_mats = []
    for i in range(12):
        kd = rng.uniform(0.2, 0.8, 3)
        ks = rng.uniform(0.04, 0.5)
        is_lambert = (i / 12.0) < lambert_fraction
        blob_mats.append(sc.add_material(MAT_LAMBERT if is_lambert else MAT_PHONG, kd, (ks, ks, ks), (0, 0, 0),
                                         shin[i % 4]))
    emit_a = sc.add_material(MAT_PHONG, (0.8, 0.8, 0.8), (0, 0, 0), (100.0, 80.9, 29.8), 10.0)  # room.mtl:51
    emit_b = sc.add_material(MAT_PHONG, (0.8, 0.8, 0.8), (0, 0, 0), (100.0, 100.0, 100.0), 10.0)  # room.mtl:93

    # ---- emitters ------------------------------------------------------------------
    # half of the emissive triangles: small downward quads just under the ceiling;
    # the rest: level-1 icospheres (80 tris) hovering + filler quads.
    n_sphere = (n_emissive // 2) // 80
    n_sphere_tris = n_sphere * 80
    n_quad_tris = n_emissive - n_sphere_tris
    n_quads = n_quad_tris // 2
    side = max(0.02, min(0.25, 0.4 * (2 * hx) / max(1.0, np.sqrt(n_quads))))
    qc = np.stack([rng.uniform(-hx + side, hx - side, n_quads), rng.uniform(-hy + side, hy - side, n_quads),
                   rng.uniform(H - 0.30, H - 0.05, n_quads)], 1)
    s = side * rng.uniform(0.5, 1.0, size=(n_quads, 1))
    ex = np.concatenate([s, np.zeros_like(s), np.zeros_like(s)], 1)
    ey = np.concatenate([np.zeros_like(s), s, np.zeros_like(s)], 1)
    p00, p10, p11, p01 = qc - ex - ey, qc + ex - ey, qc + ex + ey, qc - ex + ey
    quads = np.concatenate([np.stack([p00, p11, p10], 1), np.stack([p00, p01, p11], 1)], 0)
    sc.add_surface(quads, _const_normals(quads, (0, 0, -1)), emit_a)
    if n_sphere > 0:
        pos, nrm = _blobs(rng, 1, n_sphere, (-hx * 0.6, -hy * 0.6, 0.8), (hx * 0.6, hy * 0.6, H * 0.7), 0.03, 0.08, 0.0)
        sc.add_surface(pos, nrm, emit_b)

    # ---- room shell ------------------------------------------------------------------
    remaining = n_triangles - n_emissive
    if wall_grid is None:
        # ~5 % of the budget on the shell, capped so quads stay >= ~5 cm
        g = int(np.clip(np.sqrt(0.05 * remaining / 12.0), 1, 200))
        wall_grid = g
    g = int(wall_grid)
    gz = max(1, int(round(g * H / (2 * hx))))
    shell = [
        (_grid_quads((-hx, -hy, 0), (2 * hx, 0, 0), (0, 2 * hy, 0), g, g), (0, 0, 1)),  # floor
        (_grid_quads((-hx, -hy, H), (2 * hx, 0, 0), (0, 2 * hy, 0), g, g), (0, 0, -1)),  # ceiling
        (_grid_quads((-hx, -hy, 0), (2 * hx, 0, 0), (0, 0, H), g, gz), (0, 1, 0)),
        (_grid_quads((-hx, hy, 0), (2 * hx, 0, 0), (0, 0, H), g, gz), (0, -1, 0)),
        (_grid_quads((-hx, -hy, 0), (0, 2 * hy, 0), (0, 0, H), g, gz), (1, 0, 0)),
        (_grid_quads((hx, -hy, 0), (0, 2 * hy, 0), (0, 0, H), g, gz), (-1, 0, 0)),
    ]
    shell_tris = sum(t.shape[0] for t, _ in shell)
    assert shell_tris < remaining, "wall_grid too fine for this triangle budget"
    for i, (t, n) in enumerate(shell):
        sc.add_surface(t, _const_normals(t, n), room_mats[i])
    remaining -= shell_tris

    # ---- blob field: greedy fill, largest icosphere level first -------------------------
    per_mat_pos = [[] for _ in blob_mats]
    per_mat_nrm = [[] for _ in blob_mats]
    mat_cursor = 0
    size_by_level = {5: (0.6, 1.4), 4: (0.25, 0.9), 3: (0.12, 0.45), 2: (0.06, 0.25), 1: (0.04, 0.12), 0: (0.02, 0.06)}
    for level in range(max_level, -1, -1):
        nt = ICO_TRIS[level]
        count = remaining // nt
        if count == 0:
            continue
        rmin, rmax = size_by_level[level]
        # spread the blobs of this level over the palette
        chunks = np.array_split(np.arange(count), min(count, len(blob_mats)))
        for ch in chunks:
            if len(ch) == 0:
                continue
            pos, nrm = _blobs(rng, level, len(ch), (-hx * 0.65, -hy * 0.65, rmin), (hx * 0.65, hy * 0.65, H * 0.55), rmin,
                              rmax)
            per_mat_pos[mat_cursor % len(blob_mats)].append(pos)
            per_mat_nrm[mat_cursor % len(blob_mats)].append(nrm)
            mat_cursor += 1
        remaining -= count * nt
    # remainder < 20: tiny floating tiles (2 tris) and at most one single triangle
    if remaining > 0:
        n_tiles = remaining // 2
        tiles = []
        for _ in range(n_tiles):
            c = rng.uniform((-hx * 0.8, -hy * 0.8, 0.3), (hx * 0.8, hy * 0.8, 2.0))
            tiles.append(_grid_quads(c, (0.1, 0, 0), (0, 0.1, 0), 1, 1))
        if remaining % 2 == 1:
            c = rng.uniform((-hx * 0.8, -hy * 0.8, 0.3), (hx * 0.8, hy * 0.8, 2.0))
            tiles.append(np.array([[c, c + (0.1, 0, 0), c + (0, 0.1, 0)]]))
        t = np.concatenate(tiles, 0)
        per_mat_pos[0].append(t.astype(f32))
        per_mat_nrm[0].append(_const_normals(t, (0, 0, 1)))
        remaining = 0
    for i, m in enumerate(blob_mats):
        if per_mat_pos[i]:
            sc.add_surface(np.concatenate(per_mat_pos[i], 0), np.concatenate(per_mat_nrm[i], 0), m)

    assert sc.n_triangles == n_triangles, (sc.n_triangles, n_triangles)
    assert sc.n_emissive == n_emissive, (sc.n_emissive, n_emissive)
    sc.meta = dict(kind="room", n_triangles=n_triangles, n_emissive=n_emissive, seed=int(seed), half_extent=(hx, hy),
                   height=H, center=(0.0, 0.0, 1.5))
    return sc


def make_tiny_scene(seed=1, n_blob_level=1, n_blobs=3, n_emissive=8):
    """A few hundred triangles: small enough for the brute-force oracle tracer."""
    shell = shell_tri_count(2, 3.0, 3.0)
    sc = make_room_scene(n_triangles=shell + n_blobs * ICO_TRIS[n_blob_level] + n_emissive, n_emissive=n_emissive,
                         seed=seed, half_extent=(3.0, 3.0), height=3.0, wall_grid=2, lambert_fraction=0.34,
                         max_level=n_blob_level)
    sc.meta["center"] = (0.0, 0.0, 1.0)
    return sc


# ---- the BASELINE.json configs -----------------------------------------
def scene_config(name):
    if name == "tiny":
        return make_tiny_scene()
    if name == "small":  # CPU-oracle friendly: 20k triangles, 200 emitters
        return make_room_scene(20000, 200, seed=0xB200, half_extent=(6.0, 6.0), height=4.0)
    if name == "room":  # config 1 stand-in: ~200k triangles, 600 emissive
        return make_room_scene(200000, 600, seed=0xB201, half_extent=(3.0, 4.0), height=3.0)
    if name == "1m":  # config 2
        return make_room_scene(1000000, 10000, seed=0xB200)
    if name == "10m":  # config 3 / 4
        return make_room_scene(10000000, 100000, seed=0xB203, half_extent=(16.0, 16.0), height=8.0, max_level=5)
    raise KeyError(name)


def orbit_position(center, t, radius=8.0, theta0_deg=20.0, step_deg=0.5):
    """Config-5 camera path: from_t = c + r (cos th, sin th, 0.2), th = th0 + 0.5 deg * t."""
    th = np.radians(theta0_deg + step_deg * t)
    c = np.asarray(center, dtype=np.float64)
    return (c + radius * np.array([np.cos(th), np.sin(th), 0.2])).astype(f32)
