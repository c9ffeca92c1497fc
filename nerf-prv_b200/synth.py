"""Deterministic synthetic workloads for the PRV hot path (BASELINE.json configs).

Geometry: analytic closed surfaces sampled with a counter-based splitmix64 generator (bit-stable across
numpy versions), quantised to the 1024^3 integer lattice and de-duplicated exactly like the reference's
ShapeNet pre-processing (ShapeNet_scripts/mesh_sampling_geo_color_shapenet.py:246-255), coloured with a
smooth function of position that never yields (255,255,255) (main.cpp:3538-3540 rule), then pushed through
the reference's own normalisation (toward pose 4, centring, scale to `predicted_size`; main.cpp:674-1010)
via the host shim, and inserted into the ground-truth map at `resolution`.

Views: the reference's Hemisphere/<N>.txt sets (fixture tests/golden/hemisphere_sets.json) or, for N the
reference does not ship (1024), a Fibonacci hemisphere lattice in the same format with row 0 = (0,0,1).
"""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_FIXTURE = os.path.join(os.path.dirname(_HERE), "tests", "golden", "hemisphere_sets.json")

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(seed, stream, n):
    """n doubles in [0,1): splitmix64(counter) with (seed, stream) folded into the counter."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([seed], dtype=np.uint64) * np.uint64(0x100000001B3) + np.uint64(stream))[0]
        ctr = np.arange(n, dtype=np.uint64) + base
    return (_splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _torus(seed, n, R, r, center=(0, 0, 0), axis=2):
    m = int(n * 1.6) + 64
    u = uniform(seed, 1, m) * 2 * np.pi
    v = uniform(seed, 2, m) * 2 * np.pi
    keep = uniform(seed, 3, m) * (R + r) <= (R + r * np.cos(v))  # area-uniform rejection
    u, v = u[keep][:n], v[keep][:n]
    x = (R + r * np.cos(v)) * np.cos(u)
    y = (R + r * np.cos(v)) * np.sin(u)
    z = r * np.sin(v)
    p = np.stack([x, y, z], axis=1)
    if axis == 0:
        p = p[:, [2, 0, 1]]
    elif axis == 1:
        p = p[:, [0, 2, 1]]
    return p + np.asarray(center, dtype=np.float64)


def _box_frame(seed, n, half, bar):
    """12 square bars along the edges of a cube of half-size `half`, bar half-width `bar` (surface samples)."""
    e = uniform(seed, 11, n)
    edge = np.minimum((e * 12).astype(np.int64), 11)
    t = uniform(seed, 12, n) * 2 - 1           # along the edge
    side = np.minimum((uniform(seed, 13, n) * 4).astype(np.int64), 3)
    s = (uniform(seed, 14, n) * 2 - 1) * bar   # across the face
    axis = edge // 4
    corner = edge % 4
    a = np.where(corner & 1, half, -half).astype(np.float64)
    b = np.where(corner & 2, half, -half).astype(np.float64)
    da = np.where(side == 0, bar, np.where(side == 1, -bar, s))
    db = np.where(side == 2, bar, np.where(side == 3, -bar, s))
    p = np.zeros((n, 3))
    for ax in range(3):
        sel = axis == ax
        o1, o2 = (ax + 1) % 3, (ax + 2) % 3
        p[sel, ax] = t[sel] * half
        p[sel, o1] = a[sel] + da[sel]
        p[sel, o2] = b[sel] + db[sel]
    return p


def _superquadric(seed, n, e1, e2, scale):
    eta = (uniform(seed, 21, n) - 0.5) * np.pi
    omg = (uniform(seed, 22, n) * 2 - 1) * np.pi

    def sp(x, e):
        return np.sign(x) * np.abs(x) ** e
    x = scale[0] * sp(np.cos(eta), e1) * sp(np.cos(omg), e2)
    y = scale[1] * sp(np.cos(eta), e1) * sp(np.sin(omg), e2)
    z = scale[2] * sp(np.sin(eta), e1)
    return np.stack([x, y, z], axis=1)


def raw_surface(kind, seed, n_points, obj_index=0):
    if kind == "torus":
        return _torus(seed, n_points, 0.6, 0.25)
    if kind == "tori_frame":  # two interlocked tori + a box frame (self-occlusion)
        n1 = int(n_points * 0.4)
        n2 = int(n_points * 0.4)
        a = _torus(seed, n1, 0.5, 0.16, center=(-0.25, 0, 0), axis=2)
        b = _torus(seed + 101, n2, 0.5, 0.16, center=(0.25, 0, 0), axis=1)
        c = _box_frame(seed + 202, n_points - n1 - n2, 0.95, 0.05)
        return np.concatenate([a, b, c], axis=0)
    if kind == "superquadric":
        k = obj_index
        e1 = 0.3 + 1.7 * ((k * 7) % 16) / 15.0
        e2 = 0.3 + 1.7 * ((k * 11 + 5) % 16) / 15.0
        scale = (1.0, 0.55 + 0.45 * ((k * 3) % 8) / 7.0, 0.45 + 0.55 * ((k * 5 + 2) % 8) / 7.0)
        return _superquadric(seed + 1000 * k, n_points, e1, e2, scale)
    raise ValueError(kind)


def lattice_cloud(raw):
    """1024^3 lattice quantisation + de-duplication; returns (xyz float32 lattice coords, rgb uint8)."""
    lo = raw.min(axis=0)
    ext = (raw.max(axis=0) - lo).max()
    q = np.rint((raw - lo) / ext * 1023.0).astype(np.int64)
    q = np.unique(q, axis=0)
    f = q.astype(np.float64) / 1023.0
    r = 40 + 170 * (0.5 + 0.5 * np.sin(6.0 * f[:, 0] + 1.0))
    g = 40 + 170 * (0.5 + 0.5 * np.sin(5.0 * f[:, 1] + 2.0))
    b = 40 + 170 * (0.5 + 0.5 * np.sin(7.0 * f[:, 2] + 3.0))
    rgb = np.stack([r, g, b], axis=1).astype(np.uint8)
    assert not np.any(np.all(rgb == 255, axis=1))
    return q.astype(np.float32), rgb


def hemisphere_set(n, fixture=_FIXTURE):
    """pt_sphere rows for N views: the reference's file if shipped in the fixture, else a Fibonacci hemisphere."""
    if os.path.exists(fixture):
        sets = json.load(open(fixture))["sets"]
        if str(n) in sets:
            return np.array([[float(t) for t in row] for row in sets[str(n)]], dtype=np.float64)
    return fibonacci_hemisphere(n)


def fibonacci_hemisphere(n):
    """z >= 0 Fibonacci lattice, row 0 forced to the pole (0,0,1) like every reference set contains the pole."""
    i = np.arange(n, dtype=np.float64)
    z = 1.0 - i / max(n - 1, 1) * 0.98  # stay slightly above the horizon
    rad = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = i * (np.pi * (3.0 - np.sqrt(5.0)))
    p = np.stack([rad * np.cos(phi), rad * np.sin(phi), z], axis=1)
    p[0] = (0.0, 0.0, 1.0)
    return p


def intrinsics_for(make_intrinsics, W, H):
    """DefaultConfiguration.yaml:38-49 scaled to W x H; model 2 and coefficients unchanged."""
    f32 = np.float32
    return make_intrinsics(W, H,
                           fx=float(f32(9.1560668945312500e+02) * f32(W) / f32(1280)),
                           fy=float(f32(9.1332666015625000e+02) * f32(W) / f32(1280)),
                           ppx=float(f32(6.4714532470703125e+02) * f32(W) / f32(1280)),
                           ppy=float(f32(3.7251531982421875e+02) * f32(H) / f32(720)),
                           model=2,
                           coeffs=(1.2042199820280075e-01, -2.1373499929904938e-01, 5.3860000334680080e-03, -2.1210000850260258e-03, 0.0))


# name -> (surface kind, config id, raw sample count, resolution, views, W, H)
CONFIGS = {
    "C1": ("torus", 1, 50000, 0.002, 32, 640, 480),
    "C2": ("tori_frame", 2, 200000, 0.001, 100, 640, 480),
    "C3": ("tori_frame", 2, 200000, 0.002, 1024, 1280, 960),
    "C4": ("superquadric", 4, 50000, 0.002, 100, 640, 480),
    "C5": ("tori_frame", 2, 200000, 0.002, 100, 800, 800),
}


def build_workload(prv, name, obj_index=0, n_views=None, size=None, n_points=None, target_size=0.10, view_space_radius=0.3):
    """Assemble one object + view set through the product's host shim (`prv` = the nerf_prv_b200 module)."""
    kind, cid, npts, res, nv, W, H = CONFIGS[name]
    if n_views is not None:
        nv = n_views
    if size is not None:
        W, H = size
    if n_points is not None:
        npts = n_points
    raw = raw_surface(kind, 20240 + cid, npts, obj_index)
    lat, rgb = lattice_cloud(raw)
    cloud, _ = prv.host_normalize_cloud(lat, target_size)
    keys, map_rgb = prv.host_build_map(cloud, rgb, res)
    sphere = hemisphere_set(nv)
    center, predicted_size, init_pos = prv.host_view_space(cloud, sphere, view_space_radius)
    pose_world = prv.view_poses(init_pos, center)
    intr = intrinsics_for(prv.make_intrinsics, W, H)
    return {"name": name, "cloud": cloud, "cloud_rgb": rgb, "keys": keys, "map_rgb": map_rgb, "resolution": res, "sphere": sphere,
            "center": center, "predicted_size": predicted_size, "init_pos": init_pos, "pose_world": pose_world, "intr": intr,
            "W": W, "H": H, "n_views": len(init_pos)}
