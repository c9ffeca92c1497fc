"""Freezes golden vectors of the FULL-SIZE BASELINE workloads C1 (32 views) and C2 (100 views, the bench default), 640x480,
computed by the CPU oracle, into golden_full.json: per view the SHA-256 of the first-hit ranks, of the depths and of the
coverage row, plus the coverage counts, the greedy sequence and the stage-independent counters (rays, hits).

    python tests/golden/make_golden_full.py        # ~1-2 minutes on 8 cores

The GPU suite checks the CUDA path against these at full size -- region cull included, which the small parity cases cannot
exercise -- without running the oracle on the GPU box; the CPU suite checks the host-compiled
per-ray code against a few of the same views.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import load_pkg  # noqa: E402
import oracle as orc  # noqa: E402

prv = load_pkg.load()
from nerf_prv_b200 import synth  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute(name, views=None):
    """views: None = every view (coverage counts, greedy); else only those views of the workload (hashes per listed view)."""
    w = synth.build_workload(prv, name)
    if views is not None:
        return compute_sample(name, w, views)
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                             list(w["intr"].coeffs))
    words = orc.bitset_words(m.n)
    rows, hit_sha, depth_sha, row_sha, counts = [], [], [], [], []
    st = orc.CastStats()
    for v in range(w["n_views"]):
        ok, r, d = m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v], stats=st)
        row = orc.bitset_from_ranks(r, words)
        rows.append(row)
        hit_sha.append(sha(r))
        depth_sha.append(sha(d))
        row_sha.append(sha(row))
        counts.append(int(np.unpackbits(row.view(np.uint8)).sum()))
    seq, gain, cov, scored = orc.greedy(np.stack(rows), 0, 64)
    s = st.as_dict()
    return {"name": name, "n_views": int(w["n_views"]), "size": [int(w["W"]), int(w["H"])], "full_voxels": int(m.n), "words": int(words),
            "keys_sha": sha(w["keys"]), "pose_world_sha": sha(w["pose_world"]), "init_pos_sha": sha(w["init_pos"]),
            "hit_sha": hit_sha, "depth_sha": depth_sha, "row_sha": row_sha, "counts": counts, "rays": int(s["rays"]), "hits": int(s["hits"]),
            "s_in": int(s["probes_in"]), "greedy_seq": seq.tolist(), "greedy_gain": gain.tolist(), "covered_sha": sha(cov)}


def compute_sample(name, w, views):
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                             list(w["intr"].coeffs))
    words = orc.bitset_words(m.n)
    hit_sha, depth_sha, row_sha, counts = [], [], [], []
    st = orc.CastStats()
    for v in views:
        ok, r, d = m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v], stats=st)
        row = orc.bitset_from_ranks(r, words)
        hit_sha.append(sha(r))
        depth_sha.append(sha(d))
        row_sha.append(sha(row))
        counts.append(int(np.unpackbits(row.view(np.uint8)).sum()))
    s = st.as_dict()
    return {"name": name, "n_views": int(w["n_views"]), "views": [int(v) for v in views], "size": [int(w["W"]), int(w["H"])], "full_voxels": int(m.n),
            "words": int(words), "keys_sha": sha(w["keys"]), "pose_world_sha": sha(w["pose_world"]), "init_pos_sha": sha(w["init_pos"]),
            "hit_sha": hit_sha, "depth_sha": depth_sha, "row_sha": row_sha, "counts": counts, "rays": int(s["rays"]), "hits": int(s["hits"])}


def compute_render(name):
    """C5: splat z-buffer render (RGBA + depth) and voxel-driven cast (Perception_3D::precept) of every view."""
    w = synth.build_workload(prv, name)
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                             list(w["intr"].coeffs))
    words = orc.bitset_words(m.n)
    rgba_sha, sdepth_sha, vox_hit_sha, vox_counts, visible_px = [], [], [], [], []
    for v in range(w["n_views"]):
        rgba, sdepth, _ = orc.splat(w["cloud"], w["cloud_rgb"], it, w["pose_world"][v], 5)
        rgba_sha.append(sha(rgba))
        sdepth_sha.append(sha(sdepth))
        visible_px.append(int((rgba[..., 3] > 0).sum()))
        ok, pts, ranks = m.precept(it, w["pose_world"][v], w["init_pos"][v])
        vox_hit_sha.append(sha(ranks))
        vox_counts.append(int(np.unpackbits(orc.bitset_from_ranks(ranks, words).view(np.uint8)).sum()))
    return {"name": name, "n_views": int(w["n_views"]), "size": [int(w["W"]), int(w["H"])], "full_voxels": int(m.n), "n_points": int(len(w["cloud"])),
            "keys_sha": sha(w["keys"]), "pose_world_sha": sha(w["pose_world"]), "cloud_sha": sha(w["cloud"]), "point_size": 5,
            "splat_rgba_sha": rgba_sha, "splat_depth_sha": sdepth_sha, "visible_px": visible_px, "voxel_hit_sha": vox_hit_sha, "voxel_counts": vox_counts}


if __name__ == "__main__":
    # C3: the 1024-view Fibonacci hemisphere at 1280x960 (the strong-scaling workload) -- a sample of its views
    out = {"generator": "tests/golden/make_golden_full.py (CPU oracle)", "cases": [compute("C1"), compute("C2")],
           "samples": [compute("C3", [0, 1, 100, 333, 512, 777, 1000, 1023])], "renders": [compute_render("C5")]}
    with open(os.path.join(HERE, "golden_full.json"), "w") as f:
        json.dump(out, f, indent=1)
    for c in out["cases"]:
        print(c["name"], c["n_views"], "views", c["rays"], "rays", c["hits"], "hits", "S_in", c["s_in"], "greedy", len(c["greedy_seq"]))
    for c in out["samples"]:
        print(c["name"], "views", c["views"], c["rays"], "rays", c["hits"], "hits")
    for c in out["renders"]:
        print(c["name"], c["n_views"], "views rendered at", c["size"], "visible px", sum(c["visible_px"]), "voxel-mode coverage", sum(c["voxel_counts"]))
