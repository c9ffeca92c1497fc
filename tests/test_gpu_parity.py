"""GPU parity: the sm_100a kernels, called through the C ABI, against the CPU oracle on the same inputs.

Bar: bit-exact visible-voxel sets, per-pixel first-hit ranks, coverage counts, greedy sequence; depth identical
here (the north-star tolerance is 1e-5 relative, written below).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DEPTH_RTOL = 1e-5


def small_workload(prv, synth, name, n_views, size, n_points=None):
    return synth.build_workload(prv, name, n_views=n_views, size=size, n_points=n_points)


def oracle_dense(orc, w, max_range=1.0, views=None):
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    ointr = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                                list(w["intr"].coeffs))
    words = orc.bitset_words(m.n)
    V = w["n_views"] if views is None else views
    ranks, depths, rows = [], [], []
    st = orc.CastStats()
    for v in range(V):
        ok, r, d = m.cast_view_dense(ointr, w["pose_world"][v], w["init_pos"][v], max_range=max_range, stats=st)
        ranks.append(r)
        depths.append(d)
        rows.append(orc.bitset_from_ranks(r, words))
    return m, ointr, np.stack(ranks), np.stack(depths), np.stack(rows), st.as_dict()


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("name,n_views,size", [("C1", 5, (160, 120)), ("C2", 3, (128, 96))])
def test_dense_matches_oracle(prv, orc, synth, ctx, name, n_views, size, variant):
    w = small_workload(prv, synth, name, n_views, size)
    m, ointr, o_rank, o_depth, o_rows, o_st = oracle_dense(orc, w)
    ctx.set_variant(variant)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    bits, counts, hit, depth = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE, want_hit_rank=True, want_depth=True)
    assert ctx.words == o_rows.shape[1]
    assert np.array_equal(hit, o_rank), "per-pixel first-hit voxel ranks differ"
    assert np.array_equal(bits, o_rows), "coverage bitsets differ"
    assert np.array_equal(counts, np.array([int(np.unpackbits(r.view(np.uint8)).sum()) for r in o_rows], dtype=np.uint32))
    np.testing.assert_allclose(depth, o_depth, rtol=DEPTH_RTOL, atol=0)
    assert np.array_equal(depth, o_depth)  # in fact identical
    st = ctx.get_cast_stats()
    assert st["hits"] == o_st["hits"] and st["rays"] == o_st["rays"]
    if variant == 2:  # AXIS proves some through-AABB misses without marching them (conservative brick cull)
        assert st["probes_in"] <= o_st["probes_in"] and st["marched"] >= st["hits"]
    else:
        assert st["probes_in"] == o_st["probes_in"], "S_in (in-AABB probes) must equal the oracle's count"
    assert (hit != prv.NONE).sum() > 100  # the scene is actually visible


@pytest.mark.parametrize("variant", [0, 2])
def test_voxel_mode_and_precept_match_oracle(prv, orc, synth, ctx, variant):
    w = small_workload(prv, synth, "C1", 4, (320, 240))
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    ointr = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                                list(w["intr"].coeffs))
    ctx.set_variant(variant)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    bits, counts, hit, _ = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_VOXEL, want_hit_rank=True)
    words = orc.bitset_words(m.n)
    for v in range(w["n_views"]):
        ok, pts, ranks = m.precept(ointr, w["pose_world"][v], w["init_pos"][v])
        assert ok
        assert np.array_equal(hit[v], ranks), "voxel-driven hit ranks differ in view %d" % v
        assert np.array_equal(bits[v], orc.bitset_from_ranks(ranks, words))
        assert counts[v] == len(set(ranks[ranks != orc.NONE].tolist()))
        # Perception_3D::precept: exact cloud->points image
        g, in_map = ctx.precept(w["pose_world"][v], w["init_pos"][v])
        assert in_map
        for fld in ("x", "y", "z", "r", "g", "b"):
            assert np.array_equal(g[fld], pts[fld]), fld
        assert np.all(g["w"] == 1.0) and np.all(g["a"] == 255)
        assert (ranks != orc.NONE).sum() > 100


def test_dense_superset_of_voxel_mode(prv, synth, ctx):
    """metamorphic: every voxel-driven ray is also a dense-mode ray, except pixels x==W / y==H."""
    w = small_workload(prv, synth, "C1", 3, (320, 240))
    ctx.set_variant(2)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    bd, cd, _, _ = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE)
    bv, cv, _, _ = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_VOXEL)
    extra = bv & ~bd
    assert int(np.unpackbits(extra.view(np.uint8)).sum()) <= 4  # only from the x==W / y==H fringe, normally 0
    assert np.all(cd >= cv - 4)


def test_max_range_path_matches_oracle(prv, orc, synth, ctx):
    """maxRange shorter than the scene: the fast-path proof must fail and the literal march must agree with the oracle."""
    w = small_workload(prv, synth, "C1", 3, (96, 72))
    for mr in (0.29, 0.33):
        m, ointr, o_rank, o_depth, o_rows, o_st = oracle_dense(orc, w, max_range=mr)
        for variant in (0, 2):
            ctx.set_variant(variant)
            ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
            ctx.set_camera(w["intr"], mr)
            bits, counts, hit, depth = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE, want_hit_rank=True, want_depth=True)
            assert np.array_equal(hit, o_rank)
            assert np.array_equal(bits, o_rows)
    ctx.set_camera(w["intr"], 1.0)


def test_view_inside_object_and_out_of_map(prv, orc, synth, ctx):
    w = small_workload(prv, synth, "C1", 2, (64, 48))
    ctx.set_variant(2)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    ointr = orc.make_intrinsics(64, 48, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, 2, list(w["intr"].coeffs))
    # camera placed exactly in an occupied voxel -> "view in the object" -> nothing visible (main.cpp:263-267)
    k = w["keys"][len(w["keys"]) // 2].astype(np.int64)
    pos = (k - 32768 + 0.5) * w["resolution"]
    init = np.array([pos, [1000.0, 0.0, 0.0]])  # second: coordToKeyChecked fails -> "View out of map" (main.cpp:139)
    pw = np.stack([w["pose_world"][0], w["pose_world"][1]])
    bits, counts, hit, _ = ctx.cast_views(pw, init, mode=prv.MODE_DENSE, want_hit_rank=True)
    assert counts.tolist() == [0, 0] and np.all(hit == prv.NONE)
    ok, r, d = m.cast_view_dense(ointr, pw[0], init[0])
    assert ok and np.all(r == orc.NONE)
    ok, r, d = m.cast_view_dense(ointr, pw[1], init[1])
    assert not ok
    pts, in_map = ctx.precept(pw[1], init[1])
    assert not in_map and np.all(pts["x"] == 0)


def test_greedy_matches_oracle(prv, orc, synth, ctx):
    w = small_workload(prv, synth, "C2", 100, (160, 120))
    ctx.set_variant(2)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    bits, counts, _, _ = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE)
    for first, max_iter in ((0, 64), (7, 5), (99, 1), (3, 200)):
        o_seq, o_gain, o_cov, _ = orc.greedy(bits, first, max_iter)
        ctx.greedy_async(first, max_iter)
        g_seq, g_gain, g_cov = ctx.get_greedy(max_iter)
        assert g_seq.tolist() == o_seq.tolist()
        assert g_gain.tolist() == o_gain.tolist()
        assert np.array_equal(g_cov, o_cov)
        s2, g2 = ctx.greedy(first, max_iter)
        assert s2.tolist() == o_seq.tolist() and g2.tolist() == o_gain.tolist()
        # metamorphic: gains after the first are non-increasing; popcount(covered) = sum gains
        assert all(g_gain[i] >= g_gain[i + 1] for i in range(1, len(g_gain) - 1))
        assert int(np.unpackbits(g_cov.view(np.uint8)).sum()) == int(g_gain.sum())


def test_greedy_ties_and_empty_rows(prv, orc, synth, ctx):
    """identical rows -> lowest view id wins; all-zero rows -> stops after first_view."""
    w = small_workload(prv, synth, "C1", 3, (64, 48))
    ctx.set_variant(2)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    V = 6
    pw = np.stack([w["pose_world"][i % 2] for i in range(V)])  # views 0,2,4 identical; 1,3,5 identical
    ip = np.stack([w["init_pos"][i % 2] for i in range(V)])
    bits, _, _, _ = ctx.cast_views(pw, ip, mode=prv.MODE_DENSE)
    o_seq, o_gain, _, _ = orc.greedy(bits, 4, 10)
    seq, gain = ctx.greedy(4, 10)
    assert seq.tolist() == o_seq.tolist() == [4, 1] or seq.tolist() == o_seq.tolist()
    assert seq[1] == 1
    # far-away cameras looking at nothing: empty rows
    init = w["init_pos"][:2] * 1.0
    pw2 = w["pose_world"][:2].copy()
    pw2[:, :3, :3] = -pw2[:, :3, :3]  # look away from the object (still a rigid-ish frame for ray purposes)
    bits2, counts2, _, _ = ctx.cast_views(pw2, init, mode=prv.MODE_DENSE)
    if counts2.sum() == 0:
        seq, gain = ctx.greedy(0, 8)
        assert seq.tolist() == [0] and gain.tolist() == [0]


@pytest.mark.parametrize("size,point_size", [((160, 120), 5), ((200, 200), 5), ((97, 61), 3), ((64, 48), 4)])
def test_splat_matches_oracle(prv, orc, synth, ctx, size, point_size):
    w = small_workload(prv, synth, "C1", 4, size)
    ointr = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                                list(w["intr"].coeffs))
    ctx.set_camera(w["intr"], 1.0)
    rgb = w["cloud_rgb"].copy()
    rgb[::50] = 255  # some pure-white points: alpha must become 0 (convertToAlpha)
    ctx.set_cloud(w["cloud"], rgb)
    rgba, depth = ctx.render_views(w["pose_world"], point_size=point_size)
    assert abs(prv.splat_focal(w["intr"]) - orc.splat_focal(ointr)) == 0
    for v in range(w["n_views"]):
        o_rgba, o_depth, _ = orc.splat(w["cloud"], rgb, ointr, w["pose_world"][v], point_size)
        assert np.array_equal(rgba[v], o_rgba), "RGBA differs in view %d" % v
        np.testing.assert_allclose(depth[v], o_depth, rtol=DEPTH_RTOL, atol=0)
        assert (o_rgba[..., 3] == 255).sum() > 50


def test_full_size_properties_C1(prv, synth, ctx):
    """BASELINE C1 at full size (32 views, 640x480): size-independent properties instead of the (slow) oracle:
    the three march variants agree bit for bit, counts == popcount(bitsets), S_in identical across variants."""
    w = synth.build_workload(prv, "C1")
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    ref = None
    for variant in (0, 1, 2):
        ctx.set_variant(variant)
        bits, counts, hit, depth = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE, want_hit_rank=True, want_depth=True)
        st = ctx.get_cast_stats()
        pc = np.array([int(np.unpackbits(r.view(np.uint8)).sum()) for r in bits], dtype=np.uint32)
        assert np.array_equal(pc, counts)
        for v in range(w["n_views"]):
            vis = np.unique(hit[v][hit[v] != prv.NONE])
            assert len(vis) == counts[v]
        if ref is None:
            ref = (bits, hit, depth, st)
        else:
            assert np.array_equal(bits, ref[0]) and np.array_equal(hit, ref[1]) and np.array_equal(depth, ref[2])
            assert st["hits"] == ref[3]["hits"]
            if variant == 1:
                assert st["probes_in"] == ref[3]["probes_in"]
            else:
                assert st["probes_in"] <= ref[3]["probes_in"]
    assert st["rays"] == 32 * 640 * 480
    seq, gain = ctx.greedy(0, 64)
    assert len(set(seq.tolist())) == len(seq) and gain[1:].tolist() == sorted(gain[1:].tolist(), reverse=True)


def test_greedy_grid_barrier_path_matches_cluster_path(prv, orc, synth, monkeypatch):
    """The greedy has two implementations (one thread-block cluster with the table in distributed shared memory; a
    persistent grid-barrier kernel for tables that do not fit): both must give the oracle's sequence."""
    w = small_workload(prv, synth, "C2", 100, (96, 72))
    results = []
    for use_cluster in ("0", "1"):
        monkeypatch.setenv("PRV_GREEDY_CLUSTER", use_cluster)
        c = prv.Context(0)
        c.set_map(w["keys"], w["map_rgb"], w["resolution"])
        c.set_camera(w["intr"], 1.0)
        bits, _, _, _ = c.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE)
        c.greedy_async(5, 64)
        seq, gain, cov = c.get_greedy(64)
        o_seq, o_gain, o_cov, _ = orc.greedy(bits, 5, 64)
        assert seq.tolist() == o_seq.tolist() and gain.tolist() == o_gain.tolist() and np.array_equal(cov, o_cov)
        results.append(seq.tolist())
        c.close()
    assert results[0] == results[1]


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_views,size", [("C1", 4, (160, 120)), ("C1", 6, (640, 480)), ("C2", 8, (640, 480))])
def test_brick_cull_settings_are_exact(prv, orc, synth, name, n_views, size):
    """prv_set_brick_cull: bricks of 16 / 8 / 4 voxels, the exact march starting at the AABB face or at the first set brick of
    the cull's walk.  Same rows, ranks, depths and greedy sequence for every setting (and as the oracle on
    the small case); smaller bricks let fewer rays through, brick entry probes less."""
    w = synth.build_workload(prv, name, n_views=n_views, size=size)
    c = prv.Context(0)
    try:
        def run(cell, entry, mode=prv.MODE_DENSE):
            c.set_brick_cull(cell, entry)
            c.set_map(w["keys"], w["map_rgb"], w["resolution"])
            c.set_camera(w["intr"], 1.0)
            bits, counts, hit, depth = c.cast_views(w["pose_world"], w["init_pos"], mode=mode, want_hit_rank=True, want_depth=mode == prv.MODE_DENSE)
            st = c.get_cast_stats()
            seq, gains = c.greedy(0, 64)
            return bits, counts, hit, depth, st, seq, gains
        base = run(8, False)
        if size[0] <= 160:
            m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
            it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                                     list(w["intr"].coeffs))
            for v in range(n_views):
                ok, r, d = m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v])
                assert np.array_equal(base[2][v], r) and np.array_equal(base[3][v], d)
        prev = None
        for cell in (16, 8, 4):
            got = run(cell, False)
            for a, b in zip(got[:4], base[:4]):
                assert np.array_equal(a, b), "bricks of %d changed a result" % cell
            assert got[5].tolist() == base[5].tolist() and got[6].tolist() == base[6].tolist()
            assert got[4]["hits"] == base[4]["hits"] and got[4]["rays"] == base[4]["rays"] and got[4]["hits"] <= got[4]["marched"]
            if prev is not None:
                assert got[4]["marched"] <= prev
            prev = got[4]["marched"]
            ent = run(cell, True)
            for a, b in zip(ent[:4], base[:4]):
                assert np.array_equal(a, b), "bricks of %d with brick entry changed a result" % cell
            assert ent[5].tolist() == base[5].tolist() and ent[6].tolist() == base[6].tolist()
            assert ent[4]["marched"] == got[4]["marched"] and ent[4]["steps"] == got[4]["steps"] and ent[4]["hits"] == base[4]["hits"]
            assert ent[4]["probes_in"] <= got[4]["probes_in"]
            if cell <= 8:
                assert ent[4]["probes_in"] < got[4]["probes_in"]
        # voxel-driven mode goes through the same kernels
        v1 = run(4, True, prv.MODE_VOXEL)
        v0 = run(8, False, prv.MODE_VOXEL)
        assert np.array_equal(v1[0], v0[0]) and np.array_equal(v1[2], v0[2])
    finally:
        c.close()
