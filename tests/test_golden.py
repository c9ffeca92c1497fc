"""Golden vectors (tests/golden/golden_small.json, frozen from the CPU oracle by tests/golden/make_golden.py).

CPU: the oracle still reproduces them.  GPU: the CUDA path reproduces them without the oracle in the loop.
"""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_small.json")))["cases"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_oracle_reproduces_golden(case, prv, orc, synth):
    w = synth.build_workload(prv, case["name"], n_views=case["n_views"], size=tuple(case["size"]))
    assert sha(w["keys"]) == case["keys_sha"] and sha(w["pose_world"]) == case["pose_world_sha"] and sha(w["init_pos"]) == case["init_pos_sha"]
    assert [float.hex(float(x)) for x in w["pose_world"][0].reshape(-1)] == case["pose_world_view0"]
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                             list(w["intr"].coeffs))
    words = orc.bitset_words(m.n)
    assert m.n == case["full_voxels"] and words == case["words"]
    rows, hits, depths = [], [], []
    st = orc.CastStats()
    for v in range(w["n_views"]):
        ok, r, d = m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v], stats=st)
        rows.append(orc.bitset_from_ranks(r, words))
        hits.append(r)
        depths.append(d)
    rows = np.stack(rows)
    assert sha(rows) == case["dense_rows_sha"] and sha(np.stack(hits)) == case["dense_hit_sha"] and sha(np.stack(depths)) == case["dense_depth_sha"]
    assert st.as_dict() == case["stats"]
    seq, gain, _, scored = orc.greedy(rows, 0, 64)
    assert seq.tolist() == case["greedy_seq"] and gain.tolist() == case["greedy_gain"] and scored == case["greedy_scored"]
    rgba, sdepth, _ = orc.splat(w["cloud"], w["cloud_rgb"], it, w["pose_world"][1], 5)
    assert sha(rgba) == case["splat_rgba_sha"] and sha(sdepth) == case["splat_depth_sha"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_gpu_reproduces_golden(case, prv, synth, ctx):
    w = synth.build_workload(prv, case["name"], n_views=case["n_views"], size=tuple(case["size"]))
    ctx.set_variant(prv.VARIANT_AXIS)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    assert ctx.full_voxels == case["full_voxels"] and ctx.words == case["words"]
    bits, counts, hit, depth = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE, want_hit_rank=True, want_depth=True)
    assert counts.tolist() == case["dense_counts"]
    assert sha(bits) == case["dense_rows_sha"] and sha(hit) == case["dense_hit_sha"] and sha(depth) == case["dense_depth_sha"]
    seq, gain = ctx.greedy(0, 64)
    assert seq.tolist() == case["greedy_seq"] and gain.tolist() == case["greedy_gain"]
    bits_v, counts_v, hit_v, _ = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_VOXEL, want_hit_rank=True)
    assert counts_v.tolist() == case["voxel_counts"] and sha(bits_v) == case["voxel_rows_sha"] and sha(hit_v) == case["voxel_hit_sha"]
    ctx.set_cloud(w["cloud"], w["cloud_rgb"])
    rgba, sdepth = ctx.render_views(w["pose_world"][1:2], 5)
    assert sha(rgba[0]) == case["splat_rgba_sha"] and sha(sdepth[0]) == case["splat_depth_sha"]


# ---- full-size BASELINE workloads (tests/golden/golden_full.json, frozen from the CPU oracle by make_golden_full.py) ----------
FULL = json.load(open(os.path.join(HERE, "golden", "golden_full.json")))["cases"]
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("case", FULL, ids=[c["name"] for c in FULL])
def test_recorded_b200_run_matches_the_full_size_oracle_vectors(case):
    """profiles/r1_bench_C{1,2}_n1.json are the bench lines the round-1 B200 runs printed.  What they record of the results --
    hit count, greedy sequence, coverage rate, S_in -- must be what the oracle computes for the whole workload."""
    rec = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_%s_n1.json" % case["name"])))
    assert rec["cast_stats"]["rays"] == case["rays"] and rec["cast_stats"]["hits"] == case["hits"]
    assert rec["greedy_seq"] == case["greedy_seq"] and rec["greedy_len"] == len(case["greedy_seq"])
    assert rec["coverage_rate"] == sum(case["greedy_gain"]) / case["full_voxels"]
    assert rec["roofline"]["s_in_probes"] == case["s_in"]
    assert rec["config"]["voxels"] == case["full_voxels"] and rec["config"]["views"] == case["n_views"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", FULL, ids=[c["name"] for c in FULL])
def test_gpu_reproduces_full_size_golden(case, prv, synth, ctx):
    """BASELINE C1 / C2 at full size through the C ABI: every view's first-hit ranks, depths and coverage row, the counts and
    the greedy sequence against the frozen oracle vectors -- the region cull included, which only full-size images exercise."""
    w = synth.build_workload(prv, case["name"])
    assert sha(w["keys"]) == case["keys_sha"] and sha(w["pose_world"]) == case["pose_world_sha"] and sha(w["init_pos"]) == case["init_pos_sha"]
    ctx.set_variant(prv.VARIANT_AXIS)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    assert ctx.full_voxels == case["full_voxels"] and ctx.words == case["words"]
    bits, counts, hit, depth = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE, want_hit_rank=True, want_depth=True)
    assert counts.tolist() == case["counts"]
    for v in range(case["n_views"]):
        assert sha(hit[v]) == case["hit_sha"][v], "first-hit ranks of view %d" % v
        assert sha(depth[v]) == case["depth_sha"][v], "depths of view %d" % v
        assert sha(bits[v]) == case["row_sha"][v], "coverage row of view %d" % v
    st = ctx.get_cast_stats()
    assert st["rays"] == case["rays"] and st["hits"] == case["hits"]
    ctx.greedy_async(0, 64)
    seq, gain, cov = ctx.get_greedy(64)
    assert seq.tolist() == case["greedy_seq"] and gain.tolist() == case["greedy_gain"] and sha(cov) == case["covered_sha"]


SAMPLES = json.load(open(os.path.join(HERE, "golden", "golden_full.json")))["samples"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", SAMPLES, ids=[c["name"] for c in SAMPLES])
def test_gpu_reproduces_sampled_views_of_the_1024_view_workload(case, prv, synth, ctx):
    """C3 (1024 Fibonacci-hemisphere views, 1280x960, the strong-scaling workload): a sample of its views against the oracle."""
    w = synth.build_workload(prv, case["name"])
    assert sha(w["keys"]) == case["keys_sha"] and sha(w["pose_world"]) == case["pose_world_sha"] and sha(w["init_pos"]) == case["init_pos_sha"]
    ids = case["views"]
    ctx.set_variant(prv.VARIANT_AXIS)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    bits, counts, hit, depth = ctx.cast_views(w["pose_world"][ids], w["init_pos"][ids], mode=prv.MODE_DENSE, want_hit_rank=True, want_depth=True)
    assert counts.tolist() == case["counts"]
    for k, v in enumerate(ids):
        assert sha(hit[k]) == case["hit_sha"][k] and sha(depth[k]) == case["depth_sha"][k] and sha(bits[k]) == case["row_sha"][k], "view %d" % v
    st = ctx.get_cast_stats()
    assert st["rays"] == case["rays"] and st["hits"] == case["hits"]


RENDERS = json.load(open(os.path.join(HERE, "golden", "golden_full.json")))["renders"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", RENDERS, ids=[c["name"] for c in RENDERS])
def test_gpu_reproduces_full_size_renders_and_voxel_mode(case, prv, synth, ctx):
    """C5 (100 views, 800x800): the splat z-buffer render of every view (RGBA bit-exact; depth identical here, the north-star
    tolerance is 1e-5 relative) and the voxel-driven cast (Perception_3D::precept rays) of every view against the oracle."""
    w = synth.build_workload(prv, case["name"])
    assert sha(w["keys"]) == case["keys_sha"] and sha(w["pose_world"]) == case["pose_world_sha"] and sha(w["cloud"]) == case["cloud_sha"]
    ctx.set_variant(prv.VARIANT_AXIS)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    bits, counts, hit, _ = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_VOXEL, want_hit_rank=True)
    assert counts.tolist() == case["voxel_counts"]
    for v in range(case["n_views"]):
        assert sha(hit[v]) == case["voxel_hit_sha"][v], "voxel-mode hit ranks of view %d" % v
    ctx.set_cloud(w["cloud"], w["cloud_rgb"])
    for v0 in range(0, case["n_views"], 25):  # 25 views per call: 64 MB RGBA + 64 MB depth
        rgba, sdepth = ctx.render_views(w["pose_world"][v0:v0 + 25], case["point_size"])
        for k in range(rgba.shape[0]):
            assert sha(rgba[k]) == case["splat_rgba_sha"][v0 + k], "RGBA of view %d" % (v0 + k)
            assert sha(sdepth[k]) == case["splat_depth_sha"][v0 + k], "depth image of view %d" % (v0 + k)
            assert int((rgba[k][..., 3] > 0).sum()) == case["visible_px"][v0 + k]
    ctx.set_camera(w["intr"], 1.0)
