"""Known-answer tests that pin the CPU oracle (the reference ships none: SURVEY.md section 8(c) lists what to author).

Analytic castRay scenes, key maths, projection quirks, leaf order, frozen greedy and splat definitions.
"""
import os

import numpy as np
import pytest

RES = 0.002


def key_of(c):
    return int(np.floor(c / RES)) + 32768


def centre(k):
    return (k - 32768 + 0.5) * RES


def one_voxel_map(orc, kx, ky, kz, rgb=(10, 20, 30)):
    return orc.Map.from_keys(np.array([[kx, ky, kz]], dtype=np.uint16), np.array([rgb], dtype=np.uint8), RES)


def test_key_maths(orc):
    L = orc.lib()
    import ctypes as C
    k = C.c_uint16(0)
    assert L.orc_coord_to_key(0.0, RES, C.byref(k)) == 1 and k.value == 32768
    assert L.orc_coord_to_key(-1e-9, RES, C.byref(k)) == 1 and k.value == 32767
    assert L.orc_coord_to_key(0.0019999, RES, C.byref(k)) == 1 and k.value == 32768
    assert L.orc_coord_to_key(65.5359, RES, C.byref(k)) == 1 and k.value == 65535
    assert L.orc_coord_to_key(65.5361, RES, C.byref(k)) == 0   # scaled key 65536: out of range
    assert L.orc_coord_to_key(-65.5361, RES, C.byref(k)) == 0
    assert L.orc_key_to_coord(32768, RES) == 0.5 * RES
    assert L.orc_key_to_coord(32767, RES) == -0.5 * RES


def test_leaf_order_is_morton_z_major(orc):
    keys = np.array([[32769, 32768, 32768], [32768, 32769, 32768], [32768, 32768, 32769], [32768, 32768, 32768], [32767, 32767, 32767],
                     [32769, 32769, 32768]], dtype=np.uint16)
    m = orc.Map.from_keys(keys, None, RES)
    # below 32768 first (bit 15 clear), then x, y, x+y, z within the (32768..) octant
    assert m.keys.tolist() == [[32767, 32767, 32767], [32768, 32768, 32768], [32769, 32768, 32768], [32768, 32769, 32768],
                               [32769, 32769, 32768], [32768, 32768, 32769]]


def test_first_colour_wins_and_dedupe(orc):
    pts = np.array([[0.0001, 0.0001, 0.0001], [0.0011, 0.0011, 0.0011], [0.0051, 0.0, 0.0]], dtype=np.float32)
    rgb = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], dtype=np.uint8)
    m = orc.Map.from_points(pts, rgb, RES)
    assert m.n == 2
    assert m.rgb.tolist() == [[1, 2, 3], [7, 8, 9]]  # second point falls in the first voxel and is ignored (main.cpp:1018)


@pytest.mark.parametrize("axis,sign", [(0, 1), (0, -1), (1, 1), (1, -1), (2, 1), (2, -1)])
def test_castray_axis_aligned(orc, axis, sign):
    k = [32768, 32768, 32768]
    k[axis] += sign * 40
    m = one_voxel_map(orc, *k)
    origin = np.array([centre(32768)] * 3, dtype=np.float32)
    d = np.zeros(3, dtype=np.float32)
    d[axis] = sign * 0.37
    st = orc.CastStats()
    found, end, rank = m.cast_ray(origin, d, True, 1.0, st)
    assert found and rank == 0
    assert end.tolist() == [np.float32(centre(kk)) for kk in k]
    assert st.steps == 40 and st.probes_in == 1
    # opposite direction: marches to max range and misses
    found, end, rank = m.cast_ray(origin, -d, True, 1.0)
    assert not found and rank == orc.NONE


def test_castray_diagonal_tie_order(orc):
    """Exact diagonal from a voxel centre: all three tMax are equal at every step; the strict '<' chain picks dim 2,
    then 1, then 0.  So from (0,0,0) the visited keys are (0,0,1),(0,1,1),(1,1,1),(1,1,2)...: voxel (1,1,0) and (1,0,0)
    are never visited, (0,0,1) and (0,1,1) are."""
    o = 32768
    origin = np.array([centre(o)] * 3, dtype=np.float32)
    d = np.array([1, 1, 1], dtype=np.float32)
    for off, expect in (((0, 0, 1), True), ((0, 1, 1), True), ((1, 1, 1), True), ((1, 0, 0), False), ((1, 1, 0), False), ((0, 1, 0), False)):
        m = one_voxel_map(orc, o + off[0], o + off[1], o + off[2])
        found, end, rank = m.cast_ray(origin, d, True, 0.05)
        assert found == expect, off


def test_castray_origin_inside_occupied_and_zero_direction(orc):
    m = one_voxel_map(orc, 32768, 32768, 32768)
    origin = np.array([0.0003, 0.0011, 0.0019], dtype=np.float32)  # not the centre
    found, end, rank = m.cast_ray(origin, [0, 0, 1], True, 1.0)
    assert found and rank == 0 and end.tolist() == [np.float32(centre(32768))] * 3   # end = voxel centre, not origin
    m2 = one_voxel_map(orc, 32800, 32768, 32768)
    found, end, rank = m2.cast_ray(np.array([centre(32768)] * 3, dtype=np.float32), [0, 0, 0], True, 1.0)
    assert not found  # "Raycasting in direction (0,0,0) is not possible!"
    found, _, _ = m2.cast_ray(np.array([100.0, 0, 0], dtype=np.float32), [1, 0, 0], True, 1.0)
    assert not found  # origin out of the key range


def test_castray_max_range_is_on_centre_distance(orc):
    """The range test compares the distance between the ORIGIN and the reached voxel's CENTRE (float terms, double sum)
    with maxRange, strictly greater => miss."""
    o = 32768
    m = one_voxel_map(orc, o + 50, o, o)
    origin = np.array([centre(o)] * 3, dtype=np.float32)
    d50 = float(np.float32(centre(o + 50)) - np.float32(centre(o)))
    found, _, _ = m.cast_ray(origin, [1, 0, 0], True, d50 * 1.0001)
    assert found
    found, _, _ = m.cast_ray(origin, [1, 0, 0], True, d50 * 0.9999)
    assert not found
    found, _, _ = m.cast_ray(origin, [1, 0, 0], True, 0.0)   # maxRange <= 0: no limit
    assert found
    # ignoreUnknown = False stops at the first unknown voxel
    found, _, _ = m.cast_ray(origin, [1, 0, 0], False, 1.0)
    assert not found


def test_slow_lookup_agrees_with_bitmap(orc):
    rng = np.random.default_rng(3)
    keys = (32768 + rng.integers(-12, 12, size=(400, 3))).astype(np.uint16)
    m = orc.Map.from_keys(keys, None, RES)
    origin = np.array([centre(32768 + 40), centre(32768 + 3), centre(32768 - 2)], dtype=np.float32)
    dirs = rng.normal(size=(300, 3)).astype(np.float32)
    dirs[:, 0] = -np.abs(dirs[:, 0]) * 4
    res_a = [m.cast_ray(origin, d, True, 1.0) for d in dirs]
    m.set_slow_lookup(True)
    res_b = [m.cast_ray(origin, d, True, 1.0) for d in dirs]
    assert [(a[0], a[2]) for a in res_a] == [(b[0], b[2]) for b in res_b]
    assert sum(a[0] for a in res_a) > 20


def test_projection_quirks(orc):
    it = orc.make_intrinsics(640, 480, 457.8, 456.6, 323.5, 248.3, 2, (0.12, -0.21, 0.0054, -0.0021, 0.0))
    # pin-hole centre
    p = orc.project_point_to_pixel(it, [0, 0, 1])
    assert p.tolist() == [np.float32(323.5), np.float32(248.3)]
    # coefficient index mapping: coeffs[2] and [3] are the tangential terms, coeffs[4] the r^6 radial term
    x, y = np.float32(0.3), np.float32(-0.2)
    f32 = np.float32
    r2 = f32(x * x + y * y)
    c = [f32(v) for v in (0.12, -0.21, 0.0054, -0.0021, 0.0)]
    f = f32(f32(f32(1) + f32(c[0] * r2)) + f32(f32(c[1] * r2) * r2)) + f32(f32(f32(c[4] * r2) * r2) * r2)
    f = f32(f)
    xd, yd = f32(x * f), f32(y * f)
    dx = f32(f32(xd + f32(f32(f32(f32(2) * c[2]) * xd) * yd)) + f32(c[3] * f32(r2 + f32(f32(f32(2) * xd) * xd))))
    dy = f32(f32(yd + f32(f32(f32(f32(2) * c[3]) * xd) * yd)) + f32(c[2] * f32(r2 + f32(f32(f32(2) * yd) * yd))))
    exp = [f32(f32(dx * f32(457.8)) + f32(323.5)), f32(f32(dy * f32(456.6)) + f32(248.3))]
    got = orc.project_point_to_pixel(it, [0.3, -0.2, 1.0])
    assert got.tolist() == exp
    # deproject(project(p)) is NOT the identity for model 2 (the same polynomial is applied both ways) -- keep the quirk
    back = orc.deproject_pixel_to_point(it, got, 1.0)
    assert abs(float(back[0]) - 0.3) > 1e-4
    # project_pixel_to_ray_end takes ints: identity pose, pixel (ppx-ish) -> z = 1
    e = orc.project_pixel_to_ray_end(100, 50, it, np.eye(4), 1.0)
    assert e[2] == 1.0


def test_precept_accepts_pixel_equal_to_width(orc):
    """main.cpp:248 rejects pixel > width, so a voxel projecting to u in [W, W+1) ... only u == W exactly passes; a voxel
    at u slightly below W (truncates to W-1) is kept, one beyond W is dropped."""
    it = orc.make_intrinsics(64, 48, 50.0, 50.0, 32.0, 24.0, 0)
    pose = np.eye(4)
    # camera at the origin voxel looking along +z; voxel at x such that u = W exactly: x/z*50+32 = 64 -> x/z = 0.64
    o = 32768
    m = one_voxel_map(orc, o + 64, o, o + 100)
    ok, pts, ranks = m.precept(it, pose, [centre(o)] * 3)
    assert ok
    # the ray through the truncated pixel may or may not hit the voxel; what matters here is that the oracle ran the cast
    st = orc.CastStats()
    m.precept(it, pose, [centre(o)] * 3, stats=st)
    u = (centre(o + 64)) / (centre(o + 100)) * 50 + 32
    assert (st.rays == 1) == (0 <= u <= 64)


def test_view_pose_properties(orc):
    """get_next_camera_pos: the camera looks at the object (camera +Z through the centre), pose is rigid."""
    c = np.array([1e-9, -2e-9, 3e-9])
    for ip in ([0.1, 0.2, 0.2], [0.3, 0.0, 1e-3], [-0.12, 0.05, 0.27]):
        p = orc.view_pose(np.array(ip), c)
        pw = orc.view_pose_world(p)
        R, t = pw[:3, :3], pw[:3, 3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-9) and abs(np.linalg.det(R) - 1) < 1e-9
        assert np.allclose(t, ip, atol=1e-12)                       # camera sits at the view position
        z_axis = R[:, 2]
        to_obj = (c - ip) / np.linalg.norm(c - ip)
        assert np.allclose(z_axis, to_obj, atol=1e-9)
        assert np.allclose(orc.mat4_inverse(pw) @ pw, np.eye(4), atol=1e-12)


def test_greedy_definition(orc):
    vis = np.zeros((5, 2), dtype=np.uint64)
    vis[0, 0] = 0b00001111
    vis[1, 0] = 0b11110000
    vis[2, 0] = 0b11110000          # same as view 1: tie -> lowest id (1) wins
    vis[3, 0] = 0b100000011
    vis[4, 1] = 0b1
    seq, gain, cov, scored = orc.greedy(vis, 0, 64)
    assert seq.tolist() == [0, 1, 3, 4] and gain.tolist() == [4, 4, 1, 1]
    assert int(cov[0]) == 0b111111111 and int(cov[1]) == 1
    assert scored == 4 + 3 + 2 + 1   # unchosen views scored per iteration (the last one finds gain 0 and stops)
    seq, gain, _, _ = orc.greedy(vis, 3, 1)
    assert seq.tolist() == [3, 1] and gain.tolist() == [3, 4]
    seq, gain, _, _ = orc.greedy(np.zeros((3, 2), dtype=np.uint64), 2, 8)
    assert seq.tolist() == [2] and gain.tolist() == [0]


def test_splat_definition(orc):
    it = orc.make_intrinsics(40, 30, 50.0, 60.0, 20.7, 15.3, 0)
    f = orc.splat_focal(it)
    assert f == np.float32(30 * np.float32(60.0) / (2.0 * 15))      # H*fy / (2*(int)ppy)
    pts = np.array([[0, 0, 1.0], [0, 0, 2.0], [0.2, 0.1, 1.0], [0, 0, 0.005], [0.05, 0, 1.0]], dtype=np.float32)
    rgb = np.array([[1, 2, 3], [9, 9, 9], [255, 255, 255], [7, 7, 7], [50, 60, 70]], dtype=np.uint8)
    rgba, depth, index = orc.splat(pts, rgb, it, np.eye(4), 5)
    # point 0 lands at (20,15): covers 18..22 x 13..17, in front of point 1
    assert index[15, 20] == 0 and rgba[15, 20].tolist() == [1, 2, 3, 255] and depth[15, 20] == 1.0
    assert index[13, 18] == 0 and index[17, 22] in (0, 4) and index[12, 20] == orc.NONE
    # point 3 is in front of the near plane (z <= 0.01): clipped
    assert 3 not in index
    # pure white point: alpha 0 but it still owns the pixel (depth written)
    u = int(np.floor(0.2 / 1.0 * f + 20)); v = int(np.floor(0.1 / 1.0 * f + 15))
    assert index[v, u] == 2 and rgba[v, u].tolist() == [255, 255, 255, 0] and depth[v, u] == 1.0
    # background
    assert rgba[0, 0].tolist() == [255, 255, 255, 0] and depth[0, 0] == 0.0
    # equal depth overlap: the lower point index wins (points 0 and 4 both at z = 1)
    assert index[15, 22] == 0


def test_reference_thread_structure_equals_plain_precept(prv, orc, synth):
    """main.cpp:124-130: one std::thread per voxel in batches of num_of_thread; same cloud as the plain loop."""
    w = synth.build_workload(prv, "C1", n_views=2, size=(160, 120), n_points=6000)
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, 2, list(w["intr"].coeffs))
    ok, a, ra = m.precept(it, w["pose_world"][1], w["init_pos"][1])
    ok2, b, rb = m.precept_threads(it, w["pose_world"][1], w["init_pos"][1], num_of_thread=20)
    assert ok and ok2 and np.array_equal(ra, rb) and np.array_equal(a, b) and (ra != orc.NONE).sum() > 50


def test_exact_skip_ahead_of_the_tmax_recurrence_matches_the_literal_loop(tmp_path):
    """tests/cpp/test_advance.cpp: the binade-wise exact skip-ahead of `t = fl(t + d)` (a measured dead end, DESIGN.md
    section 7) reproduces `for k < n: t = t + d` bit for bit, ties included."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "test_advance"
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), os.path.join(root, "tests", "cpp", "test_advance.cpp")], check=True)
    r = subprocess.run([str(exe), "2000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout[-2000:]


def test_float_sqrt_equals_double_sqrt_rounded_to_float_for_every_float(tmp_path):
    """tests/cpp/test_sqrt_rounding.cpp: octomath's norm() is `(float)sqrt((double)norm_sq)`; the march takes it with one
    correctly rounded float square root.  Exhaustive over all non-negative floats."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "test_sqrt_rounding"
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-fopenmp", "-ffp-contract=off", "-o", str(exe), os.path.join(root, "tests", "cpp", "test_sqrt_rounding.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mismatches: 0 of 2139095041" in r.stdout, r.stdout[-500:]


def _py_cast_ray(occ, origin, dirp, res, max_range, ignore_unknown=True):
    """castRay of OctoMap 1.9.x written a second time, independently of oracle/prv_oracle.cpp, straight from the frozen
    spec in SURVEY.md section 8(c), with numpy scalars standing in for the C types (np.float32 = float, python float =
    double).  Slow: small scenes only.  Returns (found, key or None, steps)."""
    f32 = np.float32
    rf = 1.0 / res
    key = []
    for i in range(3):
        k = int(np.floor(rf * float(origin[i]))) + 32768
        if k < 0 or k >= 65536:
            return False, None, 0
        key.append(k)
    if tuple(key) in occ:
        return True, tuple(key), 0
    if not ignore_unknown:
        return False, None, 0
    d = [f32(dirp[0]), f32(dirp[1]), f32(dirp[2])]
    nsq = f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2]))
    ln = float(np.sqrt(np.float64(nsq)))
    if ln > 0:
        d = [f32(x / f32(ln)) for x in d]
    step, tmax, tdelta = [0] * 3, [0.0] * 3, [0.0] * 3
    big = float(np.finfo(np.float64).max)
    for i in range(3):
        step[i] = 1 if d[i] > 0 else (-1 if d[i] < 0 else 0)
        if step[i] != 0:
            border = (float(key[i] - 32768) + 0.5) * res
            border += float(step[i] * res * 0.5)
            tmax[i] = (border - float(origin[i])) / float(d[i])
            tdelta[i] = res / abs(float(d[i]))
        else:
            tmax[i] = tdelta[i] = big
    if step == [0, 0, 0]:
        return False, None, 0
    steps = 0
    while True:
        if tmax[0] < tmax[1]:
            dim = 0 if tmax[0] < tmax[2] else 2
        else:
            dim = 1 if tmax[1] < tmax[2] else 2
        if (step[dim] < 0 and key[dim] == 0) or (step[dim] > 0 and key[dim] == 65535):
            return False, None, steps
        key[dim] += step[dim]
        tmax[dim] += tdelta[dim]
        steps += 1
        if max_range > 0:
            d2 = 0.0
            for j in range(3):
                end = f32((float(key[j] - 32768) + 0.5) * res)
                df = f32(end - f32(origin[j]))
                d2 += float(f32(df * df))
            if d2 > max_range * max_range:
                return False, None, steps
        if tuple(key) in occ:
            return True, tuple(key), steps
        if not ignore_unknown:
            return False, None, steps


@pytest.mark.parametrize("res,seed", [(0.002, 1), (0.001, 2), (0.0025, 3)])
def test_castray_agrees_with_an_independent_python_restatement(orc, res, seed):
    """Random small scenes: the C++ oracle and a line-by-line Python restatement of the same frozen spec must return the
    same hit voxel and the same number of DDA steps for every ray (generic directions, axis-parallel and exactly diagonal
    ones, origins on and off voxel centres, short and unlimited max range)."""
    rng = np.random.default_rng(seed)
    o = 32768
    keys = np.unique(rng.integers(o - 14, o + 15, size=(260, 3)), axis=0)
    keys = keys[np.any(np.abs(keys - o) > 2, axis=1)]  # keep the origin neighbourhood free
    # leaf (Morton) order is what Map.from_keys expects: sort by interleaved code, z most significant within a triple
    def morton(k):
        c = 0
        for b in range(16):
            c |= ((int(k[0]) >> b) & 1) << (3 * b) | ((int(k[1]) >> b) & 1) << (3 * b + 1) | ((int(k[2]) >> b) & 1) << (3 * b + 2)
        return c
    keys = np.array(sorted(keys.tolist(), key=morton), dtype=np.uint16)
    m = orc.Map.from_keys(keys, np.full((len(keys), 3), 7, dtype=np.uint8), res)
    occ = {tuple(int(x) for x in k) for k in keys}
    n_hit = 0
    for i in range(160):
        origin = np.array([(rng.integers(-2, 3) + (0.5 if i % 3 else rng.random())) * res for _ in range(3)], dtype=np.float32)
        kind = i % 8
        if kind == 0:
            d = np.zeros(3, dtype=np.float32); d[rng.integers(0, 3)] = rng.choice([-1.0, 1.0])
        elif kind == 1:
            d = rng.choice([-1.0, 1.0], size=3).astype(np.float32)                      # exact diagonal: ties at every step
        elif kind == 2:
            d = np.array([rng.choice([-2.0, 2.0]), rng.choice([-1.0, 1.0]), 0.0], dtype=np.float32)
        else:
            d = rng.normal(size=3).astype(np.float32)
        max_range = [1.0, 0.02, 0.0, 0.011][i % 4]
        st = orc.CastStats()
        found, end, rank = m.cast_ray(origin, d, True, max_range, st)
        p_found, p_key, p_steps = _py_cast_ray(occ, origin, d, res, max_range)
        assert found == p_found, (i, origin, d, max_range)
        assert st.steps == p_steps, (i, st.steps, p_steps)
        if found:
            n_hit += 1
            assert tuple(int(x) for x in keys[rank]) == p_key
            assert end.tolist() == [np.float32((p_key[j] - 32768 + 0.5) * res) for j in range(3)]
    assert n_hit > 10


@pytest.mark.parametrize("model", [0, 1, 2, 3, 4, 5])
def test_camera_maths_pinned_against_the_reference_code(orc, model):
    """oracle/_ref/librs2_ref.so is the reference's OWN rs2_project_point_to_pixel / rs2_deproject_pixel_to_point
    (Share_Data.hpp:92-196) compiled from /root/reference by `make -C oracle ref`; the oracle's restatement must agree
    with it bit for bit -- D435 intrinsics of DefaultConfiguration.yaml:38-49 and stronger distortion, every model."""
    if orc.ref_rs2() is None:
        pytest.skip("oracle/_ref/librs2_ref.so not built and /root/reference not present")
    rng = np.random.default_rng(100 + model)
    yaml_coeffs = (0.12042199820280075, -0.21373499929904938, -0.0021210000850260258, 0.0053860000334680080, 0.0)
    strong = (0.31, -0.47, 0.013, -0.009, 0.12)
    n_cmp = 0
    for coeffs in (yaml_coeffs, strong, (0, 0, 0, 0, 0)):
        it = orc.make_intrinsics(1280, 720, 915.60669, 913.32666, 647.14532, 372.51532, model, coeffs)
        for _ in range(400):
            pt = np.array([rng.normal() * 0.2, rng.normal() * 0.2, 0.05 + rng.random()], dtype=np.float32)
            a, b = orc.project_point_to_pixel(it, pt), orc.ref_project_point_to_pixel(it, pt)
            assert a.tobytes() == b.tobytes(), (model, coeffs, pt, a, b)
            if model != 1:  # the reference asserts on deprojecting a forward-distorted image
                px = np.array([rng.random() * 1281, rng.random() * 721], dtype=np.float32)
                if rng.random() < 0.3:
                    px = np.floor(px)  # integer pixels: what project_pixel_to_ray_end passes (main.cpp:253)
                depth = 1.0 if rng.random() < 0.7 else float(np.float32(rng.random() + 0.1))
                a, b = orc.deproject_pixel_to_point(it, px, depth), orc.ref_deproject_pixel_to_point(it, px, depth)
                assert a.tobytes() == b.tobytes(), (model, coeffs, px, depth, a, b)
            n_cmp += 1
    assert n_cmp == 1200


def test_view_pose_logic_pinned_against_the_reference_view_class(prv, orc):
    """oracle/_ref/libview_ref.so is the reference's OWN `class View` (View_Space.hpp:40-199: look-at frame, 71-step roll
    search with its 1e-6 tie rule, final pose) compiled from /root/reference against oracle/ref_eigen_shim.hpp.  The
    oracle's orc_view_pose and the host mirror's prv_host_view_pose must reproduce its pose bit for bit for every view of
    the shipped hemisphere sets (pole view included) and for random camera / object placements."""
    if orc.ref_view() is None:
        pytest.skip("oracle/_ref/libview_ref.so not built and /root/reference not present")
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sets = json.load(open(os.path.join(root, "tests", "golden", "hemisphere_sets.json")))["sets"]
    rng = np.random.default_rng(7)
    n = 0
    for name in ("3", "32", "100"):
        pts = np.array([[float(c) for c in row] for row in sets[name]])
        centre = np.array([0.013, -0.021, 0.004])
        for p in pts:
            init_pos = p * 0.3 + centre
            ref = orc.ref_view_pose(init_pos, centre)
            assert orc.view_pose(init_pos, centre).tobytes() == ref.tobytes(), (name, p)
            assert prv.host_view_pose(init_pos, centre).tobytes() == ref.tobytes(), (name, p)
            n += 1
    for _ in range(60):  # a moved camera frame (now_camera_pose_world != I) and arbitrary positions
        ang = rng.random(3) * 6.0
        cz, sz, cy, sy, cx, sx = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
        R = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        now = np.eye(4)
        now[:3, :3] = R
        now[:3, 3] = rng.normal(size=3) * 0.2
        centre = rng.normal(size=3) * 0.05
        init_pos = centre + rng.normal(size=3) * 0.3
        ref = orc.ref_view_pose(init_pos, centre, now)
        assert orc.view_pose(init_pos, centre, now).tobytes() == ref.tobytes()
        assert prv.host_view_pose(init_pos, centre, now).tobytes() == ref.tobytes()
        n += 1
    assert n == 3 + 32 + 100 + 60


def test_per_voxel_logic_pinned_against_the_reference_precept_thread_process(prv, orc, synth):
    """oracle/_ref/libprecept_ref.so runs the reference's OWN Perception_3D::precept_thread_process (main.cpp:238-284:
    voxel -> camera frame -> distorting projection -> '>' bounds test -> int-truncated pixel -> project_pixel_to_ray_end
    -> direction -> castRay -> colour lookup) for every voxel, compiled from /root/reference against Eigen / OctoMap / PCL
    shims (castRay answered by the oracle).  orc_precept must give the identical cloud->points image."""
    if orc.ref_precept_lib() is None:
        pytest.skip("oracle/_ref/libprecept_ref.so not built and /root/reference not present")
    w = synth.build_workload(prv, "C1", n_views=6, size=(640, 480))
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                             list(w["intr"].coeffs))
    seen = 0
    for v in range(w["n_views"]):
        ok_ref, ref = m.ref_precept(it, w["pose_world"][v], w["init_pos"][v])
        ok, pts, _ = m.precept(it, w["pose_world"][v], w["init_pos"][v])
        assert ok == ok_ref
        for f in ("x", "y", "z", "r", "g", "b"):
            assert np.array_equal(pts[f], ref[f]), (v, f)
        seen += int(np.count_nonzero(ref["x"] != 0))
    assert seen > 1000
    # a view whose origin is outside the key range: "View out of map" -> all-zero cloud on both sides
    ok_ref, ref = m.ref_precept(it, w["pose_world"][0], np.array([1.0e6, 0.0, 0.0]))
    ok, pts, _ = m.precept(it, w["pose_world"][0], np.array([1.0e6, 0.0, 0.0]))
    assert not ok_ref and not ok and not ref["x"].any() and not pts["x"].any()


def test_view_space_pinned_against_the_reference_get_view_space(prv, orc, synth):
    """The reference's OWN View_Space::get_view_space (View_Space.hpp:517-558: sequential double centroid sums, farthest
    point * 17/16, hemisphere rows scaled by view_space_radius / pt_norm with pt_norm taken from row 0, z < 0 rows skipped),
    compiled from /root/reference over the Eigen shim, against the oracle and the host mirror: bit for bit."""
    if orc.ref_view() is None:
        pytest.skip("oracle/_ref/libview_ref.so not built and /root/reference not present")
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sets = json.load(open(os.path.join(root, "tests", "golden", "hemisphere_sets.json")))["sets"]
    w = synth.build_workload(prv, "C1", n_views=4, size=(160, 120))
    cloud = np.ascontiguousarray(w["cloud"], dtype=np.float32)
    for name in ("3", "32", "100"):
        sph = np.array([[float(c) for c in row] for row in sets[name]])
        sph2 = sph.copy()
        sph2[1, 2] = -0.25  # a row below the horizon is skipped
        for s in (sph, sph2):
            pt_norm = float(np.sqrt(s[0, 0] * s[0, 0] + (s[0, 1] * s[0, 1] + s[0, 2] * s[0, 2])))
            rc, rs, ri = orc.ref_view_space(cloud, s, 0.3, pt_norm)
            oc, os_, oi = orc.view_space(cloud, s, 0.3, pt_norm)
            hc, hs, hi = prv.host_view_space(cloud, s, 0.3, pt_norm)
            assert rc.tobytes() == oc.tobytes() == hc.tobytes()
            assert rs == os_ == hs
            assert ri.shape == oi.shape == hi.shape and ri.tobytes() == oi.tobytes() == hi.tobytes()
            assert len(ri) == int((s[:, 2] >= 0).sum())


# ---------------------------------------------------------------- exact rational DDA: castRay without floating point
def _rational_cast(occ, origin_key, dirp, res, max_range, margin):
    """The voxel sequence of castRay in EXACT arithmetic (fractions.Fraction), for an origin at a voxel centre.

    Not a restatement of the floating-point code: the DDA visits the voxels in the order in which the ray
    origin + t * dirp crosses the faces of the voxel lattice (OccupancyOcTreeBase::castRay is Amanatides-Woo: tMax_i is the
    ray parameter of the next face crossing on axis i, tDelta_i the parameter distance between crossings), so the k-th crossing
    on axis i happens at t_i(k) = (k + 1/2) * res / |dirp_i| (the origin is a voxel centre; the common normalisation factor
    does not change the order).  Every floating-point implementation of that algorithm -- OctoMap's, the oracle's, the
    kernels' -- perturbs those parameters by at most ~2e-7 relative (float origin, float-normalised direction, double
    accumulation), so wherever all competing crossings are further apart than `margin` (relative) the order, hence the hit
    voxel and the step count, is PROVABLY what the exact arithmetic gives.  Returns (provable, found, key, steps)."""
    from fractions import Fraction as F
    d = [F(float(x)) for x in dirp]
    step = [1 if x > 0 else (-1 if x < 0 else 0) for x in d]
    if step == [0, 0, 0]:
        return True, False, None, 0
    r = F(float(res))
    key = list(origin_key)
    k = [0, 0, 0]  # crossings taken per axis

    def t_next(i):
        return (F(k[i]) + F(1, 2)) * r / abs(d[i])
    steps = 0
    provable = True
    mr2 = F(float(max_range)) ** 2
    while True:
        cand = [(t_next(i), i) for i in range(3) if step[i] != 0]
        cand.sort()
        t0, dim = cand[0]
        if len(cand) > 1 and (cand[1][0] - t0) <= margin * cand[1][0]:
            provable = False  # two crossings closer than the floating-point perturbations: the order is not provable
            # (castRay itself resolves exact ties towards the higher axis; keep walking the exact order for the caller's statistics)
            tied = [c for c in cand if (c[0] - t0) <= margin * cand[1][0]]
            dim = max(c[1] for c in tied)
        if (step[dim] < 0 and key[dim] == 0) or (step[dim] > 0 and key[dim] == 65535):
            return provable, False, None, steps
        key[dim] += step[dim]
        k[dim] += 1
        steps += 1
        if max_range > 0:
            d2 = sum(((F(key[j] - origin_key[j])) * r) ** 2 for j in range(3))  # centre-to-centre distance
            if abs(d2 - mr2) <= F(1, 10 ** 6) * mr2:
                provable = False
            if d2 > mr2:
                return provable, False, None, steps
        if tuple(key) in occ:
            return provable, True, tuple(key), steps
        if steps > 4000:
            return False, False, None, steps


@pytest.mark.parametrize("res,seed", [(0.002, 11), (0.001, 12), (0.0025, 13), (0.002, 14)])
def test_castray_agrees_with_exact_rational_arithmetic(orc, res, seed):
    """known answers that do not depend on any floating-point restatement.  For random scenes the voxel
    sequence is computed with exact rational arithmetic from the geometry alone; for every ray whose face crossings are
    separated by more than 1e-5 relative (two orders of magnitude above any float / double perturbation of the algorithm's
    parameters) the oracle must report exactly that hit voxel and that number of DDA steps.  OctoMap 1.9.6 itself is still
    not available here (parity of the last-bit behaviour at ties stays UNPINNED); this pins everything that is not a tie."""
    rng = np.random.default_rng(seed)
    o = 32768
    keys = np.unique(rng.integers(o - 20, o + 21, size=(4000, 3)), axis=0)
    keys = keys[np.any(np.abs(keys - o) > 3, axis=1)]  # keep the origin neighbourhood free

    def morton(k):
        c = 0
        for b in range(16):
            c |= ((int(k[0]) >> b) & 1) << (3 * b) | ((int(k[1]) >> b) & 1) << (3 * b + 1) | ((int(k[2]) >> b) & 1) << (3 * b + 2)
        return c
    keys = np.array(sorted(keys.tolist(), key=morton), dtype=np.uint16)
    m = orc.Map.from_keys(keys, np.full((len(keys), 3), 7, dtype=np.uint8), res)
    occ = {tuple(int(x) for x in k) for k in keys}
    n_provable = n_hit = n_range = 0
    for i in range(400):
        ok_key = [o + int(rng.integers(-3, 4)) for _ in range(3)]
        origin = np.array([(k - o + 0.5) * res for k in ok_key], dtype=np.float32)  # a voxel centre, narrowed to float like point3d
        d = rng.normal(size=3).astype(np.float32)
        if i % 5 == 0:
            d[rng.integers(0, 3)] = 0.0  # rays inside a lattice plane
        max_range = [1.0, 0.0, 0.03, 0.017][i % 4]
        provable, r_found, r_key, r_steps = _rational_cast(occ, ok_key, d, res, max_range, 1e-5)
        if not provable:
            continue
        n_provable += 1
        st = orc.CastStats()
        found, end, rank = m.cast_ray(origin, d, True, max_range, st)
        assert found == r_found, (i, found, r_found)
        assert st.as_dict()["steps"] == r_steps, (i, st.as_dict()["steps"], r_steps)
        if found:
            assert tuple(int(x) for x in keys[rank]) == r_key, (i, keys[rank], r_key)
            np.testing.assert_array_equal(end, np.array([(k - o + 0.5) * res for k in r_key], dtype=np.float32))
            n_hit += 1
        elif max_range > 0 and r_steps < 4000:
            n_range += 1
    assert n_provable >= 250 and n_hit >= 150 and n_range >= 20, (n_provable, n_hit, n_range)
