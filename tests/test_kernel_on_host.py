"""The per-ray device code of the ray-cast kernels (nerf-prv_b200/csrc/prv_kernels.cuh, unchanged) compiled with g++ and
run on the CPU against the oracle (tests/cpp/kernel_on_host.cpp): exactness of the march variants, of the per-view
fast-path proof and of the three conservative culls, without a GPU.  A checker only: the product has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATS = ("rays", "region_culled", "loose_culled", "coarse_culled", "marched", "probes", "steps", "hits", "flags", "region_ok", "box_entries", "box_fallbacks")


@pytest.fixture(scope="module")
def koh(tmp_path_factory):
    out = tmp_path_factory.mktemp("koh") / "libkernel_on_host.so"
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-I/usr/local/cuda/include",
                    "-o", str(out), os.path.join(ROOT, "tests", "cpp", "kernel_on_host.cpp")], check=True)
    lib = C.CDLL(str(out))
    assert lib.koh_num_stats() == len(STATS)
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def cast_dense(koh, w, v, variant, max_range=1.0, force_region_cull=-1, intr=None, brick=8, entry=True):
    it = intr if intr is not None else w["intr"]
    keys = np.ascontiguousarray(w["keys"], dtype=np.uint16)
    rgb = np.ascontiguousarray(w["map_rgb"], dtype=np.uint8)
    pw = np.ascontiguousarray(w["pose_world"][v], dtype=np.float64)
    ip = np.ascontiguousarray(w["init_pos"][v], dtype=np.float64)
    hit = np.zeros((it.height, it.width), dtype=np.uint32)
    depth = np.zeros((it.height, it.width), dtype=np.float32)
    st = np.zeros(len(STATS), dtype=np.uint64)
    rc = koh.koh_cast_view_dense(_p(keys, C.c_uint16), _p(rgb, C.c_uint8), C.c_uint32(len(keys)), C.c_double(w["resolution"]), C.byref(it),
                                 C.c_double(max_range), _p(pw, C.c_double), _p(ip, C.c_double), variant, force_region_cull, brick, 1 if entry else 0,
                                 _p(hit, C.c_uint32), _p(depth, C.c_float), _p(st, C.c_uint64))
    assert rc == 0
    return hit, depth, dict(zip(STATS, (int(x) for x in st)))


def oracle_view(orc, w, v, max_range=1.0, intr=None):
    it = intr if intr is not None else w["intr"]
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    ointr = orc.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, it.model, list(it.coeffs))
    st = orc.CastStats()
    ok, r, d = m.cast_view_dense(ointr, w["pose_world"][v], w["init_pos"][v], max_range=max_range, stats=st)
    return m, ointr, r, d, st.as_dict()


@pytest.mark.parametrize("name,n_views,size", [("C1", 6, (320, 240)), ("C2", 4, (256, 192)), ("C1", 2, (640, 480))])
def test_march_variants_and_culls_match_oracle(koh, prv, orc, synth, name, n_views, size):
    w = synth.build_workload(prv, name, n_views=n_views, size=size)
    region_culled = marched = 0
    for v in range(n_views):
        _, _, o_rank, o_depth, o_st = oracle_view(orc, w, v)
        for variant in (0, 1, 2):
            hit, depth, st = cast_dense(koh, w, v, variant)
            assert np.array_equal(hit, o_rank), "%s view %d variant %d: first-hit ranks differ" % (name, v, variant)
            assert np.array_equal(depth, o_depth)
            assert st["hits"] == o_st["hits"] and st["rays"] == size[0] * size[1]
            assert st["flags"] == 1 | 4  # in map, not in the object, fast-path proof holds at max_range 1.0
            if variant == 1:  # FAST executes every in-AABB probe of the literal algorithm: S_in of the roofline
                assert st["probes"] == o_st["probes_in"]
            if variant == 2:
                assert st["region_culled"] + st["loose_culled"] + st["coarse_culled"] + st["marched"] == st["rays"]
                assert st["probes"] <= o_st["probes_in"]
                region_culled += st["region_culled"]
                marched += st["marched"]
    assert marched > 0
    if size[0] >= 640:  # at the bench resolution the shipped intrinsics pass the region-cull validation
        assert region_culled > 0


def test_full_size_view_of_the_bench_workload(koh, prv, orc, synth):
    """One 640x480 view of C2 (BASELINE configs[1]) through the whole AXIS pipeline, and the stage split the DESIGN quotes."""
    w = synth.build_workload(prv, "C2", n_views=100)
    v = 37
    _, _, o_rank, o_depth, o_st = oracle_view(orc, w, v)
    hit, depth, st = cast_dense(koh, w, v, 2)
    assert np.array_equal(hit, o_rank) and np.array_equal(depth, o_depth)
    assert st["hits"] == o_st["hits"] > 10000
    assert st["marched"] < 0.25 * st["rays"]  # the culls remove most of the image (3.55 M of 30.7 M rays marched)
    assert st["marched"] >= st["hits"]


def test_region_cull_never_removes_a_hit(koh, prv, orc, synth):
    """The region-level cull (whole 32x32-pixel regions dismissed by four plane tests) on cameras it was validated for --
    pin-hole, the shipped Brown-Conrady coefficients, an off-centre principal point, odd image sizes: a culled region may
    only contain misses.  Where prv_set_camera's validation rejects the camera (coarse images under the same distortion:
    a region then spans too much of the lens) the cull must stay off."""
    total_culled = 0
    for size, model, ppx_off, ppy_off, expect_ok in (((640, 480), 2, 0.0, 0.0, 1), ((640, 480), 0, 0.0, 0.0, 1), ((652, 470), 2, 41.0, -23.0, 1),
                                                     ((640, 480), 4, -30.0, 17.0, 1), ((200, 136), 2, 0.0, 0.0, 0)):
        w = synth.build_workload(prv, "C1", n_views=8, size=size)
        it = w["intr"]
        it.model = model
        it.ppx += ppx_off
        it.ppy += ppy_off
        for v in (0, 5):
            _, _, o_rank, o_depth, _ = oracle_view(orc, w, v)
            hit, depth, st = cast_dense(koh, w, v, 2)
            assert st["region_ok"] == expect_ok, (size, model)
            assert np.array_equal(hit, o_rank) and np.array_equal(depth, o_depth), (size, model, v)
            if expect_ok:
                assert st["region_culled"] > 0
                hit_off, _, st_off = cast_dense(koh, w, v, 2, force_region_cull=0)
                assert np.array_equal(hit_off, hit) and st_off["region_culled"] == 0
                assert st_off["marched"] == st["marched"]  # the region cull only removes rays the slab test would remove too
            else:
                assert st["region_culled"] == 0
            total_culled += st["region_culled"]
    assert total_culled > 0


@pytest.mark.parametrize("name,size,views", [("C1", (640, 480), (0, 11, 31)), ("C2", (640, 480), (3, 50, 99)), ("C2", (200, 152), (0, 1, 2, 3, 4, 5))])
def test_brick_entry_is_exact_for_every_brick_size(koh, prv, orc, synth, name, size, views):
    """prv_set_brick_cull: bricks of 16 / 8 / 4 voxels, the exact march starting at the AABB face or at the first set brick --
    identical results; smaller bricks prove more misses, brick entry probes less at the same DDA step count."""
    w = synth.build_workload(prv, name, n_views=max(views) + 1, size=size)
    for v in views:
        _, _, o_rank, o_depth, _ = oracle_view(orc, w, v)
        prev = None
        for brick in (16, 8, 4):
            hit, depth, st = cast_dense(koh, w, v, 2, brick=brick, entry=False)
            assert np.array_equal(hit, o_rank) and np.array_equal(depth, o_depth), (name, v, brick)
            assert st["marched"] >= st["hits"] and st["box_entries"] == 0
            if prev is not None:
                assert st["hits"] == prev["hits"] and st["region_culled"] == prev["region_culled"] and st["loose_culled"] == prev["loose_culled"]
            prev = st
            hit_e, depth_e, st_e = cast_dense(koh, w, v, 2, brick=brick, entry=True)
            assert np.array_equal(hit_e, o_rank) and np.array_equal(depth_e, o_depth), (name, v, brick, "entry")
            assert st_e["marched"] == st["marched"] and st_e["steps"] == st["steps"] and st_e["hits"] == st["hits"]
            assert st_e["probes"] <= st["probes"] and st_e["box_fallbacks"] == 0
            if brick <= 8:  # (C1's flat torus sets every 16-voxel brick: the first brick a ray meets is set, nothing to skip)
                assert st_e["probes"] < st["probes"] and st_e["box_entries"] > 0


def test_max_range_disables_the_fast_path(koh, prv, orc, synth):
    """maxRange inside the scene: the per-view proof must fail (flags without kViewFastOk) and the literal march must agree."""
    w = synth.build_workload(prv, "C1", n_views=3, size=(96, 72))
    for mr in (0.29, 0.33):
        for v in range(3):
            _, _, o_rank, o_depth, _ = oracle_view(orc, w, v, max_range=mr)
            for variant in (0, 1, 2):
                hit, depth, st = cast_dense(koh, w, v, variant, max_range=mr)
                assert st["flags"] == 1, "fast-path proof must fail when maxRange can cut a ray short"
                assert np.array_equal(hit, o_rank) and np.array_equal(depth, o_depth)
                if variant == 2:
                    assert st["marched"] == st["rays"]  # no cull may run without the proof


def test_view_in_object_and_out_of_map(koh, prv, orc, synth):
    w = synth.build_workload(prv, "C1", n_views=2, size=(64, 48))
    res = w["resolution"]
    k = w["keys"][len(w["keys"]) // 2].astype(np.float64)
    inside = (k - 32768 + 0.5) * res  # centre of an occupied voxel: "view in the object" (main.cpp:263-267)
    w["init_pos"][0] = inside
    w["init_pos"][1] = np.array([1.0e6, 0.0, 0.0])  # coordToKeyChecked fails: "View out of map" (main.cpp:139)
    for v, flags in ((0, 1 | 2), (1, 0)):
        hit, depth, st = cast_dense(koh, w, v, 2)
        assert st["flags"] & 3 == flags & 3
        assert np.all(hit == 0xFFFFFFFF) and np.all(depth == 0) and st["rays"] == 0


def test_precept_matches_oracle(koh, prv, orc, synth):
    """Voxel-driven mode (Perception_3D::precept): projection, truncated pixel (== W / == H pass), per-voxel gather."""
    w = synth.build_workload(prv, "C1", n_views=4, size=(320, 240))
    it = w["intr"]
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    ointr = orc.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, it.model, list(it.coeffs))
    keys = np.ascontiguousarray(w["keys"], dtype=np.uint16)
    rgb = np.ascontiguousarray(w["map_rgb"], dtype=np.uint8)
    for v in range(4):
        ok, o_pts, o_ranks = m.precept(ointr, w["pose_world"][v], w["init_pos"][v])
        pts = np.zeros(len(keys), dtype=prv.POINT_DTYPE)  # pcl::PointXYZRGB image, 32 B
        ranks = np.zeros(len(keys), dtype=np.uint32)
        in_map = C.c_int(0)
        pw = np.ascontiguousarray(w["pose_world"][v], dtype=np.float64)
        ip = np.ascontiguousarray(w["init_pos"][v], dtype=np.float64)
        rc = koh.koh_precept(_p(keys, C.c_uint16), _p(rgb, C.c_uint8), C.c_uint32(len(keys)), C.c_double(w["resolution"]), C.byref(it),
                             C.c_double(1.0), _p(pw, C.c_double), _p(ip, C.c_double), pts.ctypes.data_as(C.c_void_p), _p(ranks, C.c_uint32),
                             C.byref(in_map))
        assert rc == 0 and in_map.value == 1 and ok
        assert np.array_equal(ranks, o_ranks)
        for fld in ("x", "y", "z", "r", "g", "b"):
            assert np.array_equal(pts[fld], o_pts[fld]), fld
        assert np.all(pts["w"] == 1.0) and np.all(pts["a"] == 255)
        assert (ranks != orc.NONE).sum() > 100


@pytest.mark.parametrize("name,n_views", [("C1", 32), ("C2", 100)])
def test_reproduces_the_counters_the_b200_recorded(koh, prv, synth, name, n_views):
    """profiles/r1_bench_C{1,2}_n1.json hold the cast counters of the round-1 B200 runs of workloads C1 (32 views) and C2
    (BASELINE configs[1], 100 views, the bench default), 640x480.  The same per-ray code run here must count exactly the same
    rays through each stage: marched rays (what the three culls let through -- on the device they use approximate
    intrinsics), in-AABB probes, DDA steps and hits."""
    import json
    rec = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_%s_n1.json" % name)))["cast_stats"]
    w = synth.build_workload(prv, name)
    assert w["n_views"] == n_views
    tot = dict(rays=0, marched=0, probes=0, steps=0, hits=0)
    for v in range(n_views):
        _, _, st = cast_dense(koh, w, v, 2, entry=False)  # round 1 marched from the AABB face (prv_set_brick_cull(8, 0) today)
        for k in tot:
            tot[k] += st[k]
    assert tot["rays"] == rec["rays"] and tot["hits"] == rec["hits"]
    # The culls run on approximate intrinsics on the device (__fdividef, rsqrtf: <= 2 ulp) and on IEEE operations here, so a
    # ray sitting exactly on a cull margin -- a miss either way -- may be marched on one side and proven a miss on the other:
    # C1 agrees to the last step; on C2 the B200 counted 24 of 193 474 057 probes and 34 of 1 498 024 850 steps more (one
    # grazing ray; moving the reciprocal by one ulp here shifts the totals by as much).  Results never depend on it.
    assert abs(tot["marched"] - rec["marched"]) <= 4
    assert abs(tot["probes"] - rec["probes_in"]) <= 2e-6 * rec["probes_in"]
    assert abs(tot["steps"] - rec["steps"]) <= 2e-6 * rec["steps"]


def test_roofline_numerator_s_in_of_the_bench_workload(koh, prv, synth):
    """bench.py's roofline numerator is 4 B x S_in (+ 8 B per ray, + per-view terms), S_in = in-AABB probes of the literal
    algorithm, counted on the device with variant FAST.  The same code on the CPU counts the same S_in for C2."""
    import json
    rec = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_C2_n1.json")))
    w = synth.build_workload(prv, "C2")
    s_in = sum(cast_dense(koh, w, v, 1)[2]["probes"] for v in range(w["n_views"]))
    assert s_in == rec["roofline"]["s_in_probes"] == 313642529
    per_ray = 4 * s_in + 8 * rec["cast_stats"]["rays"]
    assert per_ray < rec["roofline"]["cast_pipeline"]["algorithmic_bytes"] < per_ray * 1.05  # + bitmap and row per view


# ---- random small scenes, tie-prone on purpose ------------------------------------------------------------------------------
def _morton_sorted(keys):
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0xFFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FF0000FF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00F00F00F00F)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0C30C30C30C3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x249249249249)
        return v
    code = spread(keys[:, 0]) | (spread(keys[:, 1]) << np.uint64(1)) | (spread(keys[:, 2]) << np.uint64(2))
    return keys[np.argsort(code, kind="stable")]


def _signed_permutations():
    import itertools
    out = []
    for p in itertools.permutations(range(3)):
        for sg in itertools.product((1.0, -1.0), repeat=3):
            m = np.zeros((3, 3))
            for r in range(3):
                m[r, p[r]] = sg[r]
            out.append(m)
    return out


def _random_scene(rng, prv, perms):
    """A random occupancy box of up to 21^3 voxels near the key-space centre and one camera.  Kinds 0 / 1: pose = signed
    permutation matrix, pin-hole camera with a power-of-two focal length and integer principal point, camera on a voxel
    centre outside / inside the AABB -- ray directions with exact zeros and exact equal tMax values (castRay's tie rule:
    the higher axis steps first).  Kinds 2 / 3: random rotation / look-at pose, Brown-Conrady or pin-hole, any position."""
    n = rng.integers(1, 22, size=3)
    lo = 32768 + rng.integers(-50, 50, size=3)
    occ = rng.random(tuple(n)) < rng.choice([0.01, 0.05, 0.3, 0.9])
    if not occ.any():
        occ[tuple(rng.integers(0, n))] = True
    keys = _morton_sorted((lo + np.argwhere(occ)).astype(np.uint16))
    rgb = rng.integers(0, 255, size=keys.shape).astype(np.uint8)
    res = float(rng.choice([0.002, 0.001, 0.01, 0.05]))
    kind = int(rng.integers(0, 4))
    W, H = int(rng.integers(9, 40)), int(rng.integers(7, 36))
    pose = np.eye(4)
    if kind in (0, 1):
        pose[:3, :3] = perms[rng.integers(len(perms))]
        f = float(rng.choice([4.0, 8.0, 16.0, 32.0]))
        intr = prv.make_intrinsics(W, H, f, f, float(W // 2), float(H // 2), 0, [0] * 5)
        ck = lo + rng.integers(-30, 30 + n.max(), size=3) if kind == 0 else lo + rng.integers(0, n)
        pos = (ck.astype(np.float64) - 32768 + 0.5) * res
    else:
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        pose[:3, :3] = q
        f = float(rng.uniform(10, 60))
        intr = prv.make_intrinsics(W, H, f, f * rng.uniform(0.9, 1.1), W / 2 + rng.uniform(-3, 3), H / 2 + rng.uniform(-3, 3), int(rng.choice([0, 2, 4])),
                                   [0.12, -0.21, 0.005, -0.002, 0.0])
        c = (lo + n / 2.0 - 32768) * res
        d = rng.normal(size=3)
        pos = c + d / np.linalg.norm(d) * rng.uniform(0, 3.0) * n.max() * res
        if kind == 3:
            z = c - pos
            z /= max(np.linalg.norm(z), 1e-12)
            x = np.cross(z, rng.normal(size=3))
            x /= np.linalg.norm(x)
            pose[:3, 0], pose[:3, 1], pose[:3, 2] = x, np.cross(z, x), z
    pose[:3, 3] = pos
    max_range = float(rng.choice([1.0, 1.0, 1.0, 0.9 * n.max() * res, -1.0, 100.0]))  # sometimes inside the scene, sometimes unlimited
    return dict(keys=keys, map_rgb=rgb, resolution=res, intr=intr, pose_world=pose[None], init_pos=pos[None]), max_range


def test_random_tie_prone_scenes(koh, prv, orc):
    """200 random scenes (4 000 more were run once, all exact): every march variant and every brick-cull setting against the
    oracle.  This found the one place where the header relied on a device-only float->int conversion (NaN -> 0)."""
    rng = np.random.default_rng(20240)
    perms = _signed_permutations()
    hits = fast = in_object = 0
    for case in range(200):
        w, max_range = _random_scene(rng, prv, perms)
        _, _, o_rank, o_depth, _ = oracle_view(orc, w, 0, max_range=max_range)
        for variant, brick, entry in ((0, 8, 0), (1, 8, 0), (2, 8, 0), (2, 16, 0), (2, 4, 0), (2, 16, 1), (2, 8, 1), (2, 4, 1)):
            hit, depth, st = cast_dense(koh, w, 0, variant, max_range=max_range, brick=brick, entry=bool(entry))
            assert np.array_equal(hit, o_rank) and np.array_equal(depth, o_depth), "scene %d variant %d brick %d entry %d" % (case, variant, brick, entry)
            assert st["box_fallbacks"] == 0
        hits += int((o_rank != 0xFFFFFFFF).sum())
        fast += 1 if st["flags"] & 4 else 0
        in_object += 1 if st["flags"] & 2 else 0
    assert hits > 5000 and 100 < fast < 200 and in_object > 5  # hits, both march paths and the in-object case all occurred


def test_full_size_golden_views(koh, prv, synth):
    """A few views of the full-size oracle vectors (tests/golden/golden_full.json) through the host-compiled per-ray code, with
    and without the opt-in cull level / later march start."""
    import hashlib
    import json
    cases = {c["name"]: c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "golden_full.json")))["cases"]}

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    for name, views in (("C1", (0, 17)), ("C2", (0, 42, 99))):
        w = synth.build_workload(prv, name)
        for v in views:
            for brick, entry in ((8, False), (8, True), (4, True)):
                hit, depth, st = cast_dense(koh, w, v, 2, brick=brick, entry=entry)
                assert sha(hit) == cases[name]["hit_sha"][v] and sha(depth) == cases[name]["depth_sha"][v], (name, v, brick, entry)
    c3 = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_full.json")))["samples"][0]
    w = synth.build_workload(prv, "C3")
    for k, v in enumerate(c3["views"][:3]):
        for brick, entry in ((8, False), (8, True)):
            hit, depth, st = cast_dense(koh, w, v, 2, brick=brick, entry=entry)
            assert sha(hit) == c3["hit_sha"][k] and sha(depth) == c3["depth_sha"][k], ("C3", v, brick, entry)


def test_fast_path_proof_at_its_threshold(koh, prv, orc):
    """maxRange swept across the exact point where the per-view proof (prv_view_const.hpp) flips: just below the distance to
    the farthest voxel centre of the AABB grown by one voxel the proof must fail (literal march with the range test), from
    there on it may hold (range test dropped) -- the results must equal the oracle's either way.  (6 480 more cases run once.)"""
    rng = np.random.default_rng(77)
    perms = _signed_permutations()
    held = failed = 0
    for case in range(40):
        w, _ = _random_scene(rng, prv, perms)
        res = w["resolution"]
        keys = w["keys"].astype(np.float64)
        lo = (keys.min(0) - 1 - 32768 + 0.5) * res
        hi = (keys.max(0) + 1 - 32768 + 0.5) * res
        pos = (np.floor(w["init_pos"][0] / res) + 0.5) * res  # the snapped origin (main.cpp:112-114)
        far = float(np.sqrt(sum(max(abs(pos[a] - lo[a]), abs(pos[a] - hi[a])) ** 2 for a in range(3))))
        for f in (0.999, 0.9999999, 1.0, 1.0000001, 1.001):
            mr = far * f
            _, _, o_rank, o_depth, _ = oracle_view(orc, w, 0, max_range=mr)
            for variant, brick, entry in ((1, 8, False), (2, 8, False), (2, 4, True)):
                hit, depth, st = cast_dense(koh, w, 0, variant, max_range=mr, brick=brick, entry=entry)
                assert np.array_equal(hit, o_rank) and np.array_equal(depth, o_depth), (case, f, variant)
            if st["flags"] & 1 and not st["flags"] & 2:
                held += 1 if st["flags"] & 4 else 0
                failed += 0 if st["flags"] & 4 else 1
    assert held > 30 and failed > 30  # the sweep really straddles the threshold


# ---- barrier-free kernels run as kernels on the CPU (tests/cpp/map_kernels_on_host.cpp) ---------------------------------------
@pytest.fixture(scope="module")
def mkh(tmp_path_factory):
    out = tmp_path_factory.mktemp("mkh") / "libmap_kernels_on_host.so"
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-I/usr/local/cuda/include",
                    "-o", str(out), os.path.join(ROOT, "tests", "cpp", "map_kernels_on_host.cpp")], check=True)
    return C.CDLL(str(out))


def test_map_build_kernels_write_the_documented_tables(mkh, prv, synth):
    """map_scatter_kernel / map_shell_kernel / map_rank_kernel -- the kernel source, threads executed one
    after the other -- must leave exactly the occupancy bitmap, shell-padded bitmap, brick grid (every brick size) and rank table that
    the per-ray checks above run on (built on the host from the documented layout)."""
    rng = np.random.default_rng(3)
    perms = _signed_permutations()
    tables = [synth.build_workload(prv, name, n_views=1, size=(32, 24)) for name in ("C1", "C2")]
    tables += [_random_scene(rng, prv, perms)[0] for _ in range(40)]
    for w in tables:
        keys = np.ascontiguousarray(w["keys"], dtype=np.uint16)
        for brick in (16, 8, 4):
            rc = mkh.mkh_check_map_kernels(_p(keys, C.c_uint16), C.c_uint32(len(keys)), C.c_double(w["resolution"]), brick)
            assert rc == 0, "table %d differs (%d voxels, bricks of %d)" % (rc, len(keys), brick)


def test_voxel_mode_kernels_match_oracle(mkh, prv, orc, synth):
    """project_voxels_kernel, gather_voxel_hits_kernel and precept_points_kernel run as kernels: Perception_3D::precept's
    cloud->points image against the oracle."""
    w = synth.build_workload(prv, "C1", n_views=3, size=(320, 240))
    it = w["intr"]
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    ointr = orc.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, it.model, list(it.coeffs))
    keys = np.ascontiguousarray(w["keys"], dtype=np.uint16)
    rgb = np.ascontiguousarray(w["map_rgb"], dtype=np.uint8)
    for v in range(3):
        ok, o_pts, o_ranks = m.precept(ointr, w["pose_world"][v], w["init_pos"][v])
        pts = np.zeros(len(keys), dtype=prv.POINT_DTYPE)
        ranks = np.zeros(len(keys), dtype=np.uint32)
        pw = np.ascontiguousarray(w["pose_world"][v], dtype=np.float64)
        ip = np.ascontiguousarray(w["init_pos"][v], dtype=np.float64)
        rc = mkh.mkh_precept(_p(keys, C.c_uint16), _p(rgb, C.c_uint8), C.c_uint32(len(keys)), C.c_double(w["resolution"]), C.byref(it), C.c_double(1.0),
                             _p(pw, C.c_double), _p(ip, C.c_double), pts.ctypes.data_as(C.c_void_p), _p(ranks, C.c_uint32))
        assert rc == 0 and ok
        assert np.array_equal(ranks, o_ranks)
        for fld in ("x", "y", "z", "r", "g", "b"):
            assert np.array_equal(pts[fld], o_pts[fld]), fld
        assert np.all(pts["w"] == 1.0) and np.all(pts["a"] == 255) and (ranks != orc.NONE).sum() > 100


def test_ingest_kernels_match_the_host_map_build(mkh, prv, synth):
    """prv_set_map_from_cloud's kernels (point -> key, voxel heads, compaction with the first point's colour) run as kernels,
    cub's stable sort and scan replaced by their definitions: same leaf-ordered keys and colours as prv_host_build_map
    (main.cpp:1005-1036), including points outside the key range and many points per voxel."""
    w = synth.build_workload(prv, "C1", n_views=1, size=(32, 24))
    rng = np.random.default_rng(11)
    clouds = [(np.ascontiguousarray(w["cloud"], dtype=np.float32), np.ascontiguousarray(w["cloud_rgb"], dtype=np.uint8), w["resolution"])]
    pts = rng.normal(scale=0.02, size=(20000, 3)).astype(np.float32)       # ~40 points per voxel at 5 mm
    pts[::97] = 1.0e6                                                        # invalid keys (coordToKeyChecked fails)
    clouds.append((pts, rng.integers(0, 256, size=(20000, 3)).astype(np.uint8), 0.005))
    for xyz, rgb, res in clouds:
        k_ref, c_ref = prv.host_build_map(xyz, rgb, res)
        keys = np.zeros((len(xyz), 3), dtype=np.uint16)
        col = np.zeros((len(xyz), 3), dtype=np.uint8)
        n = mkh.mkh_ingest(_p(xyz, C.c_float), _p(rgb, C.c_uint8), C.c_uint32(len(xyz)), C.c_double(res), _p(keys, C.c_uint16), _p(col, C.c_uint8))
        assert n == len(k_ref) > 100
        assert np.array_equal(keys[:n], k_ref) and np.array_equal(col[:n], c_ref)


# ---- the cast KERNELS themselves on a CPU SIMT emulator (tests/cpp/pipeline_on_host.cpp, tests/cpp/simt_on_host.hpp) -------------
@pytest.fixture(scope="module")
def poh(tmp_path_factory):
    out = tmp_path_factory.mktemp("poh") / "libpipeline_on_host.so"
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++20", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-pthread", "-I/usr/local/cuda/include",
                    "-o", str(out), os.path.join(ROOT, "tests", "cpp", "pipeline_on_host.cpp")], check=True)
    return C.CDLL(str(out))


FORCE_LARGE_CAST_PATHS = ["-DPRV_COARSE_WARP_TILES=0", "-DPRV_TICKET_SPREAD=1"]  # warps on whole tiles, region tickets, multi-chunk march tickets


@pytest.fixture(scope="module")
def poh_large(tmp_path_factory):
    """The same kernels with the thresholds that switch to the large-cast paths (kernels_cast.cuh) set to "always"."""
    out = tmp_path_factory.mktemp("pohl") / "libpipeline_on_host_large.so"
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++20", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-pthread", "-I/usr/local/cuda/include"] + FORCE_LARGE_CAST_PATHS +
                   ["-o", str(out), os.path.join(ROOT, "tests", "cpp", "pipeline_on_host.cpp")], check=True)
    return C.CDLL(str(out))


CONFIGS = ((8, 1), (8, 0), (4, 1))  # (brick edge, enter at brick): the default pipeline, the march from the AABB face, smaller bricks


def run_kernels(poh, w, views, mode, brick=8, entry=1, grid=2, max_range=1.0, intr=None):
    """cull -> coarse -> march kernels (and the voxel-mode kernels) as cast_impl launches them, on the emulator."""
    it = intr if intr is not None else w["intr"]
    keys = np.ascontiguousarray(w["keys"], dtype=np.uint16)
    rgb = np.ascontiguousarray(w["map_rgb"], dtype=np.uint8)
    n, V = len(keys), len(views)
    words = poh.poh_words(n)
    gw, gh = (it.width, it.height) if mode == 1 else (it.width + 1, it.height + 1)
    pw = np.ascontiguousarray(np.asarray(w["pose_world"])[list(views)], dtype=np.float64)
    ip = np.ascontiguousarray(np.asarray(w["init_pos"])[list(views)], dtype=np.float64)
    out = dict(bits=np.zeros((V, words), dtype=np.uint64), hit=np.zeros((V, gh, gw), dtype=np.uint32), depth=np.zeros((V, gh, gw), dtype=np.float32),
               stats=np.zeros((V, 4), dtype=np.uint64), marched=np.zeros(V, dtype=np.uint32), voxel_hit=np.zeros((V, n), dtype=np.uint32))
    rc = poh.poh_cast_views(_p(keys, C.c_uint16), _p(rgb, C.c_uint8), C.c_uint32(n), C.c_double(w["resolution"]), C.byref(it), C.c_double(max_range),
                            _p(pw, C.c_double), _p(ip, C.c_double), C.c_uint32(V), mode, brick, entry, grid, _p(out["bits"], C.c_uint64),
                            _p(out["hit"], C.c_uint32), _p(out["depth"], C.c_float), _p(out["stats"], C.c_ulonglong), _p(out["marched"], C.c_uint32),
                            _p(out["voxel_hit"], C.c_uint32))
    assert rc == 0
    return out


@pytest.mark.parametrize("name,size,grid", [("C1", (96, 72), 2), ("C2", (64, 48), 3), ("C1", (97, 61), 1)])
def test_cast_kernels_on_the_emulator_match_oracle(poh, koh, prv, orc, synth, name, size, grid):
    """cull_kernel, coarse_kernel, march_kernel -- kernel source unchanged, one OS
    thread per CUDA thread -- against the oracle: per-pixel ranks and depths (every pixel written by exactly the kernel that
    owns it), coverage rows (the atomicOr scatter), per-view counters (the warp-reduced statistics).  Odd image size: the
    scalar no-hit store path."""
    w = synth.build_workload(prv, name, n_views=3, size=size)
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    words = orc.bitset_words(m.n)
    for brick, entry in CONFIGS:
        out = run_kernels(poh, w, range(3), 1, brick, entry, grid)
        for v in range(3):
            _, _, o_rank, o_depth, o_st = oracle_view(orc, w, v)
            assert np.array_equal(out["hit"][v], o_rank) and np.array_equal(out["depth"][v], o_depth), (name, v, brick, entry)
            assert np.array_equal(out["bits"][v], orc.bitset_from_ranks(o_rank, words))
            _, _, st = cast_dense(koh, w, v, 2, brick=brick, entry=bool(entry))  # the per-ray check counts the same work
            assert out["stats"][v].tolist() == [st["rays"], st["probes"], st["hits"], st["steps"]] and out["marched"][v] == st["marched"]
            assert st["hits"] == o_st["hits"]


@pytest.mark.parametrize("name,size,grid,mode", [("C1", (96, 72), 1, 1), ("C2", (97, 61), 2, 1), ("C1", (160, 120), 1, 0)])
def test_large_cast_paths_on_the_emulator_match_the_default_paths(poh, poh_large, prv, synth, name, size, grid, mode):
    """coarse_kernel with every warp walking whole tiles on its own (region tickets) and march_kernel with tickets of several
    chunks -- the forms the 1024-view workload runs -- give what the small-cast forms give, counters included."""
    w = synth.build_workload(prv, name, n_views=3, size=size)
    for brick, entry in ((8, 1), (4, 0)):
        a = run_kernels(poh, w, range(3), mode, brick, entry, grid)
        b = run_kernels(poh_large, w, range(3), mode, brick, entry, grid)
        for k in ("hit", "depth", "bits", "stats", "marched", "voxel_hit"):
            assert np.array_equal(a[k], b[k]), (k, brick, entry)
        assert int(b["marched"].sum()) > 0


def test_cast_kernels_on_the_emulator_full_size_view(poh, prv, synth):
    """One 640x480 view of the bench workload through the kernels: the region cull inside cull_kernel (whole regions dismissed,
    128-bit no-hit stores) against the frozen oracle vectors."""
    import hashlib
    import json
    case = [c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "golden_full.json")))["cases"] if c["name"] == "C2"][0]
    w = synth.build_workload(prv, "C2")
    v = 37
    for brick, entry in ((8, 1), (4, 0)):
        out = run_kernels(poh, w, [v], 1, brick, entry, grid=4)
        assert hashlib.sha256(out["hit"][0].tobytes()).hexdigest() == case["hit_sha"][v]
        assert hashlib.sha256(out["depth"][0].tobytes()).hexdigest() == case["depth_sha"][v]
        assert hashlib.sha256(out["bits"][0].tobytes()).hexdigest() == case["row_sha"][v]
        assert out["marched"][0] < 0.15 * 640 * 480


def test_voxel_mode_kernels_on_the_emulator(poh, prv, orc, synth):
    """Voxel-driven mode end to end as kernels: project_voxels_kernel -> cull_kernel<MASKED> -> coarse -> march ->
    gather_voxel_hits_kernel, against Perception_3D::precept of the oracle."""
    w = synth.build_workload(prv, "C1", n_views=2, size=(160, 120))
    it = w["intr"]
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    ointr = orc.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, it.model, list(it.coeffs))
    words = orc.bitset_words(m.n)
    for brick, entry in ((8, 1), (4, 0)):
        out = run_kernels(poh, w, range(2), 0, brick, entry)
        for v in range(2):
            ok, _, o_ranks = m.precept(ointr, w["pose_world"][v], w["init_pos"][v])
            assert ok and np.array_equal(out["voxel_hit"][v], o_ranks), (v, brick, entry)
            assert np.array_equal(out["bits"][v], orc.bitset_from_ranks(o_ranks, words))
            assert (o_ranks != orc.NONE).sum() > 100


def test_special_views_on_the_emulator(poh, prv, orc, synth):
    """Camera inside an occupied voxel, camera outside the key range, and maxRange inside the scene (views without the
    fast-path proof go through the same queues but are marched literally) in one launch with an ordinary view."""
    w = synth.build_workload(prv, "C1", n_views=4, size=(64, 48))
    res = w["resolution"]
    k = w["keys"][len(w["keys"]) // 2].astype(np.float64)
    w["init_pos"][1] = (k - 32768 + 0.5) * res           # "view in the object" (main.cpp:263-267)
    w["init_pos"][2] = np.array([1.0e6, 0.0, 0.0])       # "View out of map" (main.cpp:139)
    for max_range in (1.0, 0.3):
        for brick, entry in ((8, 1), (4, 0)):
            out = run_kernels(poh, w, range(4), 1, brick, entry, max_range=max_range)
            for v in range(4):
                _, _, o_rank, o_depth, _ = oracle_view(orc, w, v, max_range=max_range)
                if v in (1, 2):  # such views launch no rays; cull_kernel still records "no hit" for every pixel
                    assert np.all(o_rank == 0xFFFFFFFF) and int(out["stats"][v][0]) == 0 and not out["bits"][v].any()
                assert np.array_equal(out["hit"][v], o_rank) and np.array_equal(out["depth"][v], o_depth), (v, max_range, brick, entry)


def test_splat_kernels_on_the_emulator(poh, prv, orc, synth):
    """splat_points_kernel (64-bit atomicMin on the corner cell) + splat_resolve_kernel (shared-memory tile) as kernels:
    RGBA bit-exact and depth identical to the oracle's frozen splat definition (north-star tolerance: 1e-5 relative)."""
    w = synth.build_workload(prv, "C1", n_views=3, size=(96, 72))
    it = w["intr"]
    ointr = orc.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, it.model, list(it.coeffs))
    xyz = np.ascontiguousarray(w["cloud"][::3], dtype=np.float32)  # a third of the points: one OS thread per point and view
    rgb = np.ascontiguousarray(w["cloud_rgb"][::3], dtype=np.uint8)
    pw = np.ascontiguousarray(w["pose_world"], dtype=np.float64)
    for point_size in (5, 2):
        rgba = np.zeros((3, it.height, it.width, 4), dtype=np.uint8)
        depth = np.zeros((3, it.height, it.width), dtype=np.float32)
        rc = poh.poh_render_views(_p(xyz, C.c_float), _p(rgb, C.c_uint8), C.c_uint64(len(xyz)), C.byref(it), _p(pw, C.c_double), C.c_uint32(3), point_size,
                                  _p(rgba, C.c_uint8), _p(depth, C.c_float))
        assert rc == 0
        for v in range(3):
            o_rgba, o_depth, _ = orc.splat(xyz, rgb, ointr, w["pose_world"][v], point_size)
            assert np.array_equal(rgba[v], o_rgba), (v, point_size)
            np.testing.assert_allclose(depth[v], o_depth, rtol=1e-5, atol=0)
            assert np.array_equal(depth[v], o_depth) and (o_rgba[..., 3] > 0).sum() > 50


@pytest.mark.parametrize("method,E", [(3, 5), (3, 2), (2, 2)])
def test_ensemble_kernels_on_the_emulator(poh, orc, method, E):
    """ensemble_terms_kernel + ensemble_sum_kernel (nbv_loop cases 2 / 3, main.cpp:2039-2161) as kernels: bit-exact scores for
    method 3 and for method 2 at the reference's ensemble of two."""
    rng = np.random.default_rng(9)
    V, H, W = 7, 45, 80
    images = rng.integers(0, 256, size=(V, E, H, W, 4)).astype(np.uint8)
    images[2, :, :10] = images[2, 0:1, :10]  # zero-variance pixels (method 2 skips them)
    scores = np.zeros(V)
    rc = poh.poh_score_ensemble(_p(images, C.c_uint8), C.c_uint32(V), C.c_uint32(E), W, H, method, _p(scores, C.c_double))
    assert rc == 0
    _, o_scores = orc.score_ensemble(images, method)
    assert np.array_equal(scores, o_scores)


@pytest.mark.parametrize("flags", [[], FORCE_LARGE_CAST_PATHS], ids=["small-cast paths", "large-cast paths"])
def test_kernels_are_race_free_under_threadsanitizer(tmp_path, flags):
    """The emulator runs CUDA threads as real concurrent OS threads synchronised only by the kernels' own barriers, warp
    collectives and atomics -- so ThreadSanitizer sees what compute-sanitizer's racecheck sees on the device, and more (global
    memory too).  The cull / coarse / march / voxel-mode / splat kernels must run without a single report; the one suppressed
    site is write_hit's deliberate test-before-atomicOr (tests/tsan/suppressions.txt)."""
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    tsan = subprocess.run([gxx, "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(tsan) or not os.path.exists(tsan):
        pytest.skip("libtsan not installed")
    lib = tmp_path / "libpipeline_on_host_tsan.so"
    # (both forms of coarse_kernel / of the tickets: the block barriers of the small casts, the shared-memory stage and per-warp tickets of the large ones)
    subprocess.run([gxx, "-O1", "-g", "-std=c++20", "-ffp-contract=off", "-fsanitize=thread", "-shared", "-fPIC", "-pthread", "-I/usr/local/cuda/include"] +
                   flags + ["-o", str(lib), os.path.join(ROOT, "tests", "cpp", "pipeline_on_host.cpp")], check=True)
    env = dict(os.environ, LD_PRELOAD=tsan,
               TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=2 suppressions=" + os.path.join(ROOT, "tests", "tsan", "suppressions.txt"))
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tsan", "run_kernels_under_tsan.py"), str(lib)], env=env, capture_output=True, text=True,
                       timeout=600)
    if "KERNELS-RAN-UNDER-TSAN" not in r.stdout and "ThreadSanitizer" not in r.stderr and "Traceback" not in r.stderr:
        pytest.skip("the interpreter does not run under libtsan here: " + r.stderr[-300:])  # (a Python error in the child is a failure, not a skip)
    assert "KERNELS-RAN-UNDER-TSAN" in r.stdout, r.stderr[-2000:]
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[:4000]


def test_count_and_iterative_greedy_kernels_on_the_emulator(poh, orc):
    """popcount_rows_kernel and the launch-per-iteration greedy (greedy_init_kernel / greedy_iter_kernel, 64-bit atomicMax of
    (gain << 32 | ~view id)) as kernels against the oracle's greedy: ties go to the lowest view id, empty and duplicate rows,
    permuted global view ids, early stop at gain 0 and the max_iter cut."""
    rng = np.random.default_rng(21)
    for trial in range(4):
        V, words = int(rng.integers(3, 24)), 2 * int(rng.integers(1, 4))
        rows = rng.integers(0, 2 ** 63, size=(V, words), dtype=np.uint64) & rng.integers(0, 2 ** 63, size=(V, words), dtype=np.uint64)
        if trial % 2 == 0:
            rows &= rng.integers(0, 2 ** 63, size=(V, words), dtype=np.uint64)  # sparser: more ties
        rows[V // 2] = 0                      # an empty row
        rows[V - 1] = rows[0]                 # a duplicate: tie on every gain
        ids = np.arange(V, dtype=np.uint32) if trial < 2 else rng.permutation(V).astype(np.uint32)
        max_iter = 24 if trial != 1 else 3
        first = int(ids[int(rng.integers(0, V))])
        counts = np.zeros(V, dtype=np.uint32)
        seq = np.zeros(max_iter + 1, dtype=np.uint32)
        gains = np.zeros(max_iter + 1, dtype=np.uint32)
        cov = np.zeros(words, dtype=np.uint64)
        n = poh.poh_counts_and_greedy(_p(rows, C.c_uint64), C.c_uint32(V), C.c_uint32(words), _p(ids, C.c_uint32), C.c_uint32(first), C.c_uint32(max_iter),
                                      _p(counts, C.c_uint32), _p(seq, C.c_uint32), _p(gains, C.c_uint32), _p(cov, C.c_uint64))
        assert n > 0
        assert counts.tolist() == [int(np.unpackbits(r.view(np.uint8)).sum()) for r in rows]
        # oracle greedy works on rows indexed by view id
        order = np.argsort(ids, kind="stable")
        o_seq, o_gain, o_cov, _ = orc.greedy(rows[order], first, max_iter)
        assert seq[:n].tolist() == o_seq.tolist() and gains[:n].tolist() == o_gain.tolist(), trial
        assert np.array_equal(cov, o_cov)
