"""Edge cases of the CUDA path against the oracle: odd image sizes, pin-hole / Brown-Conrady(4) cameras, camera inside
the occupancy AABB, thin and single-voxel maps, coarse voxels, unlimited range, more views than one launch batch, and
the error behaviour of the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def oracle_rows(orc, keys, rgb, res, intr, pose_world, init_pos, max_range=1.0):
    m = orc.Map.from_keys(keys, rgb, res)
    it = orc.make_intrinsics(intr.width, intr.height, intr.fx, intr.fy, intr.ppx, intr.ppy, intr.model, list(intr.coeffs))
    words = orc.bitset_words(m.n)
    hits, rows = [], []
    for v in range(len(init_pos)):
        ok, r, d = m.cast_view_dense(it, pose_world[v], init_pos[v], max_range=max_range, want_depth=False)
        hits.append(r)
        rows.append(orc.bitset_from_ranks(r, words))
    return m, it, np.stack(hits), np.stack(rows)


def check_dense(prv, orc, ctx, keys, rgb, res, intr, pose_world, init_pos, max_range=1.0, variants=(0, 1, 2)):
    m, it, o_hit, o_rows = oracle_rows(orc, keys, rgb, res, intr, pose_world, init_pos, max_range)
    for variant in variants:
        ctx.set_variant(variant)
        ctx.set_map(keys, rgb, res)
        ctx.set_camera(intr, max_range)
        bits, counts, hit, _ = ctx.cast_views(pose_world, init_pos, mode=prv.MODE_DENSE, want_hit_rank=True)
        assert np.array_equal(hit, o_hit), "variant %d" % variant
        assert np.array_equal(bits, o_rows)
    ctx.set_variant(2)
    return o_hit


@pytest.mark.parametrize("size", [(97, 61), (33, 9), (31, 7), (257, 130)])
def test_odd_image_sizes(prv, orc, synth, ctx, size):
    w = synth.build_workload(prv, "C1", n_views=3, size=size)
    hit = check_dense(prv, orc, ctx, w["keys"], w["map_rgb"], w["resolution"], w["intr"], w["pose_world"], w["init_pos"])
    assert (hit != orc.NONE).sum() > 10


@pytest.mark.parametrize("model", [0, 4])
def test_pinhole_camera_models(prv, orc, synth, ctx, model):
    w = synth.build_workload(prv, "C1", n_views=3, size=(128, 96))
    it = w["intr"]
    intr = prv.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, model, list(it.coeffs))
    check_dense(prv, orc, ctx, w["keys"], w["map_rgb"], w["resolution"], intr, w["pose_world"], w["init_pos"])
    # voxel mode too (projection without distortion)
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    oit = orc.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, model, list(it.coeffs))
    bits, counts, hit, _ = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_VOXEL, want_hit_rank=True)
    for v in range(3):
        ok, pts, ranks = m.precept(oit, w["pose_world"][v], w["init_pos"][v])
        assert np.array_equal(hit[v], ranks)


def test_strong_distortion_disables_region_cull_but_stays_exact(prv, orc, synth, ctx):
    w = synth.build_workload(prv, "C1", n_views=2, size=(160, 120))
    it = w["intr"]
    intr = prv.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, 2, (0.9, -1.5, 0.05, -0.04, 0.8))
    check_dense(prv, orc, ctx, w["keys"], w["map_rgb"], w["resolution"], intr, w["pose_world"], w["init_pos"], variants=(1, 2))


def test_camera_inside_the_aabb(prv, orc, synth, ctx):
    """view_space_radius smaller than the object: the origin key lies inside the occupancy AABB (approach phase empty)."""
    w = synth.build_workload(prv, "C1", n_views=6, size=(96, 72), view_space_radius=0.03)
    lo, hi = w["keys"].min(0).astype(int), w["keys"].max(0).astype(int)
    inside = 0
    for ip in w["init_pos"]:
        k = np.floor(ip / w["resolution"]).astype(int) + 32768
        inside += int(np.all(k >= lo) and np.all(k <= hi))
    assert inside >= 3
    hit = check_dense(prv, orc, ctx, w["keys"], w["map_rgb"], w["resolution"], w["intr"], w["pose_world"], w["init_pos"])
    assert (hit != orc.NONE).sum() > 100


def test_single_voxel_and_thin_maps(prv, orc, synth, ctx):
    w = synth.build_workload(prv, "C1", n_views=4, size=(64, 48))
    one = np.array([[32768, 32770, 32765]], dtype=np.uint16)
    hit = check_dense(prv, orc, ctx, one, np.array([[9, 8, 7]], dtype=np.uint8), 0.002, w["intr"], w["pose_world"], w["init_pos"])
    assert (hit == 0).sum() >= 1 and ctx.full_voxels == 1 and ctx.words == 2
    # a one-voxel-thick slab, built in leaf order through the host shim
    xs, ys = np.meshgrid(np.arange(-20, 21), np.arange(-15, 16))
    pts = np.stack([xs.ravel() * 0.002 + 0.001, ys.ravel() * 0.002 + 0.001, np.full(xs.size, 0.0031)], axis=1).astype(np.float32)
    keys, rgb = prv.host_build_map(pts, np.full((len(pts), 3), 100, dtype=np.uint8), 0.002)
    assert len(keys) == 41 * 31 and len(set(keys[:, 2].tolist())) == 1
    hit = check_dense(prv, orc, ctx, keys, rgb, 0.002, w["intr"], w["pose_world"], w["init_pos"])
    assert (hit != orc.NONE).sum() > 50


def test_coarse_resolution_and_unlimited_range(prv, orc, synth, ctx):
    w = synth.build_workload(prv, "C1", n_views=3, size=(96, 72))
    keys, rgb = prv.host_build_map(w["cloud"], w["cloud_rgb"], 0.006)
    check_dense(prv, orc, ctx, keys, rgb, 0.006, w["intr"], w["pose_world"], w["init_pos"])
    # maxRange <= 0: castRay has no range limit; the march ends at the hit, the AABB exit or the key-space border
    check_dense(prv, orc, ctx, w["keys"], w["map_rgb"], w["resolution"], w["intr"], w["pose_world"], w["init_pos"], max_range=0.0, variants=(1, 2))


def test_more_views_than_one_launch_batch(prv, orc, synth, ctx):
    """2100 views > kMaxViewsPerLaunch (1024): the persistent kernels run in three view batches."""
    w = synth.build_workload(prv, "C1", n_views=100, size=(24, 16))
    reps = 21
    pw = np.concatenate([w["pose_world"]] * reps)
    ip = np.concatenate([w["init_pos"]] * reps)
    ctx.set_variant(2)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    bits, counts, hit, _ = ctx.cast_views(pw, ip, mode=prv.MODE_DENSE, want_hit_rank=True)
    m, it, o_hit, o_rows = oracle_rows(orc, w["keys"], w["map_rgb"], w["resolution"], w["intr"], w["pose_world"], w["init_pos"])
    for r in range(reps):
        assert np.array_equal(hit[r * 100:(r + 1) * 100], o_hit) and np.array_equal(bits[r * 100:(r + 1) * 100], o_rows)
    seq, gain = ctx.greedy(2099, 3)
    o_seq, o_gain, _, _ = orc.greedy(bits, 2099, 3)
    assert seq.tolist() == o_seq.tolist() and gain.tolist() == o_gain.tolist()
    seq, gain = ctx.greedy(5, 0)   # max_iter = 0: only the start view
    assert seq.tolist() == [5] and gain.tolist() == [int(counts[5])]


def test_error_behaviour(prv, synth, ctx):
    L = prv.lib()
    w = synth.build_workload(prv, "C1", n_views=2, size=(32, 24))
    c = prv.Context(0)
    # call order
    with pytest.raises(prv.PrvError) as e:
        c.set_views(w["pose_world"], w["init_pos"])
    assert e.value.code == prv.ERR_INVALID and "prv_set_map" in str(e.value)
    with pytest.raises(prv.PrvError):
        c.cast_async(prv.MODE_DENSE)
    # keys not in leaf order / duplicates
    bad = w["keys"][::-1].copy()
    with pytest.raises(prv.PrvError) as e:
        c.set_map(bad, None, 0.002)
    assert "leaf (Morton) order" in str(e.value)
    with pytest.raises(prv.PrvError):
        c.set_map(np.zeros((0, 3), dtype=np.uint16), None, 0.002)
    with pytest.raises(prv.PrvError):
        c.set_map(w["keys"], None, -1.0)
    c.set_map(w["keys"], None, 0.002)
    # unsupported / invalid cameras
    it = w["intr"]
    for model, code in ((1, prv.ERR_UNSUPPORTED), (7, prv.ERR_INVALID)):
        with pytest.raises(prv.PrvError) as e:
            c.set_camera(prv.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, model), 1.0)
        assert e.value.code == code
    with pytest.raises(prv.PrvError):
        c.set_camera(prv.make_intrinsics(0, 10, 1, 1, 0, 0, 0), 1.0)
    c.set_camera(it, 1.0)
    c.set_views(w["pose_world"], w["init_pos"])
    with pytest.raises(prv.PrvError):
        c.cast_async(7)
    with pytest.raises(prv.PrvError):
        c.greedy_async(0, 4)          # nothing cast yet
    with pytest.raises(prv.PrvError):
        c.get_bitsets()
    c.cast_async(prv.MODE_DENSE, False)
    with pytest.raises(prv.PrvError):
        c.get_hit_rank(prv.MODE_DENSE)  # per-pixel results were not kept
    with pytest.raises(prv.PrvError):
        c.greedy_async(99, 4)         # not a resident view id
    with pytest.raises(prv.PrvError):
        c.render_async(2, 5)          # no cloud
    with pytest.raises(prv.PrvError):
        c.set_variant(9)
    # NULL context never crashes
    assert L.prv_sync(None) == prv.ERR_INVALID and L.prv_cast_async(None, 1, 0) == prv.ERR_INVALID
    assert L.prv_set_map(None, None, None, 0, 0.002) == prv.ERR_INVALID
    h = C.c_void_p()
    assert L.prv_create(C.byref(h), 4096) == prv.ERR_INVALID and b"out of range" in L.prv_last_error(None)
    # the context is still usable after errors
    c.greedy_async(1, 4)
    seq, gain, _ = c.get_greedy(4)
    assert seq[0] == 1
    c.close()


@pytest.mark.parametrize("method,E", [(2, 2), (3, 5), (2, 5), (3, 2)])
def test_ensemble_uncertainty_scoring(prv, orc, ctx, method, E):
    """nbv_loop cases 2/3 (main.cpp:2039-2161) at the reference's render size (W/16 x H/16 = 80 x 45)."""
    rng = np.random.default_rng(100 + method * 10 + E)
    V, H, W = 37, 45, 80
    base = rng.integers(0, 256, size=(V, 1, H, W, 4), dtype=np.int64)
    noise = rng.integers(-6, 7, size=(V, E, H, W, 4))
    images = np.clip(base + noise, 0, 255).astype(np.uint8)
    images[:, :, :5] = images[:, :1, :5]          # identical rows across the ensemble: zero variance (skipped by method 2)
    images[3] = images[3, :1]                    # a view with no uncertainty at all
    chosen = np.zeros(V, dtype=np.uint8)
    chosen[[0, 11]] = 1
    o_best, o_scores = orc.score_ensemble(images, method, chosen)
    g_best, g_scores = ctx.score_ensemble(images, method, chosen)
    assert g_best == o_best and o_best >= 0
    if method == 3 or E == 2:
        assert np.array_equal(g_scores, o_scores)     # same sequence of double operations (log terms from the host table)
    else:
        np.testing.assert_allclose(g_scores, o_scores, rtol=1e-12, atol=0)
    assert g_scores[0] == 0 and g_scores[11] == 0
    # everything chosen -> no candidate
    b, _ = ctx.score_ensemble(images, method, np.ones(V, dtype=np.uint8))
    assert b == -1 and orc.score_ensemble(images, method, np.ones(V, dtype=np.uint8))[0] == -1


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_gpu_ingest_matches_host_map_build(prv, orc, synth, name):
    """prv_set_map_from_cloud (GPU: key per point, stable Morton sort, first colour wins) == prv_host_build_map == oracle."""
    w = synth.build_workload(prv, name, n_views=2, size=(64, 48))
    cloud = w["cloud"].copy()
    rgb = w["cloud_rgb"].copy()
    # a few points outside the key range (coordToKeyChecked fails -> skipped), non-finite points (the clouds are is_dense = false:
    # (int)floor(NaN) is INT_MIN on the reference's x86 -> rejected; 0 in CUDA, hence the range test on the double) and duplicates
    # with different colours
    cloud = np.concatenate([cloud, [[70.0, 0, 0], [0, -70.0, 0], [np.nan, 0, 0], [0.01, np.inf, 0], [0, 0, -np.inf], [3.0e9, 0, 0]], cloud[:50]]).astype(np.float32)
    rgb = np.concatenate([rgb, [[1, 2, 3], [4, 5, 6], [7, 8, 9], [9, 8, 7], [6, 5, 4], [3, 2, 1]], 255 - rgb[:50]]).astype(np.uint8)
    c = prv.Context(0)
    c.set_map_from_cloud(cloud, rgb, w["resolution"])
    keys, col = c.get_map()
    h_keys, h_col = prv.host_build_map(cloud, rgb, w["resolution"])
    assert np.array_equal(keys, h_keys) and np.array_equal(col, h_col)
    m = orc.Map.from_points(cloud, rgb, w["resolution"])
    assert np.array_equal(keys, m.keys) and np.array_equal(col, m.rgb)
    # and the map works: cast equals the cast on the host-built map
    c.set_camera(w["intr"], 1.0)
    b1, n1, _, _ = c.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE)
    c.set_map(h_keys, h_col, w["resolution"])
    b2, n2, _, _ = c.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE)
    assert np.array_equal(b1, b2) and n1.sum() > 100
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("model,coeffs", [(3, (0.9, 0, 0, 0, 0)), (5, (0.05, -0.01, 0.002, -0.0005, 0)), (4, (0.1, -0.2, 0.001, 0.002, 0))])
def test_transcendental_distortion_models_match_the_oracle(prv, orc, synth, model, coeffs):
    """rs2 models 3 (F-Theta) and 5 (Kannala-Brandt4), Share_Data.hpp:109-136, 156-190 (round 1 refused them): tan / atan are the
    host libm's float overloads, so prv_set_camera tabulates the deprojection of every integer pixel on the host and the voxel-mode
    projection runs on the host too -- the kernels never evaluate a transcendental and the results are the oracle's bit for bit
    (dense cast, voxel-driven cast, precept cloud).  Model 4 (no distortion code in rs2) rides along."""
    w = synth.build_workload(prv, "C1", n_views=4, size=(160, 120))
    it0 = w["intr"]
    it = prv.make_intrinsics(it0.width, it0.height, it0.fx, it0.fy, it0.ppx, it0.ppy, model, coeffs)
    oit = orc.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, model, list(coeffs))
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    c = prv.Context(0)
    try:
        c.set_map(w["keys"], w["map_rgb"], w["resolution"])
        c.set_camera(it, 1.0)
        for variant in (prv.VARIANT_AXIS, prv.VARIANT_FAST):
            c.set_variant(variant)
            bits, counts, hit, depth = c.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE, want_hit_rank=True, want_depth=True)
            n_hit = 0
            for v in range(w["n_views"]):
                ok, r, d = m.cast_view_dense(oit, w["pose_world"][v], w["init_pos"][v])
                assert np.array_equal(hit[v], r) and np.array_equal(depth[v], d), (model, variant, v)
                n_hit += int((r != orc.NONE).sum())
            assert n_hit > 1000
        c.set_variant(prv.VARIANT_AXIS)
        bv, cv, hv, _ = c.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_VOXEL, want_hit_rank=True)
        for v in range(w["n_views"]):
            ok, o_pts, o_ranks = m.precept(oit, w["pose_world"][v], w["init_pos"][v])
            assert np.array_equal(hv[v], o_ranks), (model, v)
        pts, in_map = c.precept(w["pose_world"][1], w["init_pos"][1])
        ok, o_pts, _ = m.precept(oit, w["pose_world"][1], w["init_pos"][1])
        for f in ("x", "y", "z", "r", "g", "b"):
            assert np.array_equal(pts[f], o_pts[f]), (model, f)
    finally:
        c.close()


@pytest.mark.gpu
def test_abi_guards_of_round_2(prv, synth):
    """result capacity of prv_get_greedy, validated view ids, V mismatch, cast statistics of the LATEST cast."""
    import ctypes as C
    w = synth.build_workload(prv, "C1", n_views=6, size=(96, 72))
    c = prv.Context(0)
    try:
        L = prv.lib()
        c.set_map(w["keys"], w["map_rgb"], w["resolution"])
        c.set_camera(w["intr"], 1.0)
        c.set_views(w["pose_world"], w["init_pos"])
        c.cast_async(prv.MODE_DENSE, False)
        c.greedy_async(0, 64)
        # a buffer sized for a smaller max_iter than the async call ran with: nothing is written, the needed size is reported
        seq = np.full(2, 0xABABABAB, dtype=np.uint32)
        gains = np.full(2, 0xABABABAB, dtype=np.uint32)
        n = C.c_uint32(0)
        rc = L.prv_get_greedy(c._h, seq.ctypes.data_as(C.POINTER(C.c_uint32)), gains.ctypes.data_as(C.POINTER(C.c_uint32)), 2, C.byref(n), None)
        full_seq, full_gain, _ = c.get_greedy()
        assert len(full_seq) > 2
        assert rc == prv.ERR_INVALID and n.value == len(full_seq) and np.all(seq == 0xABABABAB) and np.all(gains == 0xABABABAB)
        assert b"hold 2" in L.prv_last_error(c._h)
        # get_greedy(10) after greedy_async(0, 64) (the advisor's heap-corruption case) sizes from the remembered max_iter
        s10, g10, _ = c.get_greedy(10)
        assert s10.tolist() == full_seq.tolist()
        # view ids: duplicates, sentinels, V mismatch
        V = w["n_views"]
        for bad in ([0, 1, 2, 2, 4, 5], [0, 1, 2, 3, 4, 0xFFFFFFF0], [0, 1, 2, 3, 4, 0xFFFFFFFF]):
            with pytest.raises(prv.PrvError) as e:
                c.set_views(w["pose_world"], w["init_pos"], view_ids=np.array(bad, dtype=np.uint32))
            assert e.value.code == prv.ERR_INVALID
        ids = np.arange(V, dtype=np.uint32) * 3 + 1  # (monotone: ties would break the same way)
        assert L.prv_set_view_ids(c._h, ids.ctypes.data_as(C.POINTER(C.c_uint32)), V) == 0
        pw = np.ascontiguousarray(w["pose_world"][:4].reshape(-1, 16))
        ip = np.ascontiguousarray(w["init_pos"][:4])
        rc = L.prv_set_views(c._h, pw.ctypes.data_as(C.POINTER(C.c_double)), ip.ctypes.data_as(C.POINTER(C.c_double)), 4)
        assert rc == prv.ERR_INVALID and b"prv_set_view_ids gave 6 ids" in L.prv_last_error(c._h)
        c.set_views(w["pose_world"], w["init_pos"], view_ids=ids)  # same V: accepted, ties break on these ids
        c.cast_async(prv.MODE_DENSE, False)
        seq_ids, _ = c.greedy(int(ids[0]), 64)
        assert seq_ids[0] == ids[0] and set(seq_ids.tolist()) <= set(ids.tolist())
        assert [int(np.where(ids == s)[0][0]) for s in seq_ids] == full_seq.tolist()  # the same views under other names
        # statistics belong to the latest cast, whatever variant is selected afterwards
        c.set_views(w["pose_world"], w["init_pos"])
        c.set_variant(prv.VARIANT_FAST)
        c.cast_async(prv.MODE_DENSE, False)
        c.set_variant(prv.VARIANT_AXIS)
        st = c.get_cast_stats()
        assert st["rays"] == V * 96 * 72 and st["marched"] == st["rays"]
        c.cast_async(prv.MODE_DENSE, False)
        st2 = c.get_cast_stats()
        assert st2["rays"] == st["rays"] and st2["hits"] == st["hits"] and st2["marched"] < st["marched"]
        t = c.get_timing()
        assert t["dropped"] == 0
    finally:
        c.close()


@pytest.mark.gpu
def test_shared_reciprocal_division_equals_ddiv_rn_bit_for_bit(tmp_path):
    """ddiv_pair (prv_kernels.cuh): the two divisions of a castRay axis with one refined reciprocal, against __ddiv_rn on the GPU
    -- 2^28 operand triples of the march's own ranges, 2^26 raw bit patterns (NaNs, infinities, denormals, zeros) and 2^26 with
    exponents near the ends of the range (tests/cuda/test_ddiv_pair.cu, compiled here with nvcc)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "test_ddiv_pair"
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-I", os.path.join(root, "nerf-prv_b200", "csrc"), "-I", os.path.join(root, "include"),
                    "-o", str(exe), os.path.join(root, "tests", "cuda", "test_ddiv_pair.cu")], check=True, capture_output=True)
    r = subprocess.run([str(exe), "26"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mismatches: 0" in r.stdout, r.stdout[-2000:] + r.stderr[-500:]
