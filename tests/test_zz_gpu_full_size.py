"""Full-size golden vectors on the GPU (tests/golden/golden_full.json, frozen from the CPU oracle by
tests/golden/make_golden_full.py): BASELINE C1 / C2 whole, sampled views of C3, renders and voxel-driven casts of C5.
The file name sorts last on purpose: these are the heaviest GPU tests."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_G = json.load(open(os.path.join(HERE, "golden", "golden_full.json")))
FULL, SAMPLES, RENDERS = _G["cases"], _G["samples"], _G["renders"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


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


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


_C3 = json.load(open(os.path.join(HERE, "golden", "golden_c3.json")))["cases"][0]
_C4 = json.load(open(os.path.join(HERE, "golden", "golden_c4.json")))["cases"]


@pytest.mark.gpu
def test_gpu_reproduces_the_whole_1024_view_workload(prv, synth, ctx):
    """ALL of C3 -- the north-star strong-scaling workload, 1024 views at 1280x960, 1.258 G rays -- against the oracle's frozen
    vectors (tests/golden/golden_c3.json, ~40 CPU-minutes to generate): every view's first-hit ranks, depths and coverage row,
    the counts, hits per view, the greedy sequence / gains / covered mask.  10 GB of per-pixel results are hashed in slices."""
    from concurrent.futures import ThreadPoolExecutor
    case = _C3
    w = synth.build_workload(prv, "C3")
    assert sha16(w["keys"]) == case["keys_sha"] and sha16(w["pose_world"]) == case["pose_world_sha"] and sha16(w["init_pos"]) == case["init_pos_sha"]
    ctx.set_variant(prv.VARIANT_AXIS)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    ctx.set_views(w["pose_world"], w["init_pos"])
    assert ctx.full_voxels == case["full_voxels"] and ctx.words == case["words"]
    ctx.cast_async(prv.MODE_DENSE, want_pixels=True)
    bits = ctx.get_bitsets()
    counts = ctx.get_coverage_counts()
    assert counts.tolist() == case["counts"]
    assert hashlib.sha256(bits.tobytes()).hexdigest() == case["rows_sha"]
    st = ctx.get_cast_stats()
    assert st["rays"] == case["rays"] and st["hits"] == case["hits"]
    V, step = case["n_views"], 32
    with ThreadPoolExecutor(8) as pool:
        for v0 in range(0, V, step):
            hit = ctx.get_hit_rank(prv.MODE_DENSE, v0, step)
            depth = ctx.get_depth(v0, step)
            hs = list(pool.map(sha16, [hit[k] for k in range(step)]))
            ds = list(pool.map(sha16, [depth[k] for k in range(step)]))
            for k in range(step):
                v = v0 + k
                assert hs[k] == case["hit_sha16"][v], "first-hit ranks of view %d" % v
                assert ds[k] == case["depth_sha16"][v], "depths of view %d" % v
                assert sha16(bits[v]) == case["row_sha16"][v], "coverage row of view %d" % v
                assert int((hit[k] != prv.NONE).sum()) == case["hits_per_view"][v]
    ctx.greedy_async(0, 64)
    seq, gain, cov = ctx.get_greedy()
    assert seq.tolist() == case["greedy_seq"] and gain.tolist() == case["greedy_gain"]
    assert hashlib.sha256(cov.tobytes()).hexdigest() == case["covered_sha"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", _C4, ids=["C4-object-%d" % c["obj_index"] for c in _C4])
def test_gpu_reproduces_full_size_c4_objects(case, prv, synth, ctx):
    """C4 (PRVNet dataset-generation batch, 64 superquadric objects x 100 views at 640x480): objects 0, 7 and 63 at full size."""
    w = synth.build_workload(prv, "C4", obj_index=case["obj_index"])
    assert sha16(w["keys"]) == case["keys_sha"] and sha16(w["pose_world"]) == case["pose_world_sha"]
    ctx.set_variant(prv.VARIANT_AXIS)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    bits, counts, hit, depth = ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE, want_hit_rank=True, want_depth=True)
    assert counts.tolist() == case["counts"] and hashlib.sha256(bits.tobytes()).hexdigest() == case["rows_sha"]
    for v in range(case["n_views"]):
        assert sha16(hit[v]) == case["hit_sha16"][v] and sha16(depth[v]) == case["depth_sha16"][v], "view %d" % v
    ctx.greedy_async(0, 64)
    seq, gain, cov = ctx.get_greedy()
    assert seq.tolist() == case["greedy_seq"] and gain.tolist() == case["greedy_gain"] and hashlib.sha256(cov.tobytes()).hexdigest() == case["covered_sha"]
