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
