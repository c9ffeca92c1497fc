"""Freezes small golden vectors of the hot path, computed by the CPU oracle, into golden_small.json.

    python tests/golden/make_golden.py

The reference has no golden vectors of its own (parity unpinned, see oracle/prv_oracle.h); these pin the oracle against
accidental change and give the GPU path a fixture to hit that does not need the oracle at run time.
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

CASES = [("C1", 6, (160, 120)), ("C2", 4, (128, 96)), ("C4", 5, (96, 72))]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute(name, n_views, size):
    w = synth.build_workload(prv, name, n_views=n_views, size=size)
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                             list(w["intr"].coeffs))
    words = orc.bitset_words(m.n)
    dense_rows, voxel_rows, hit, depth, vox_hit = [], [], [], [], []
    st = orc.CastStats()
    for v in range(w["n_views"]):
        ok, r, d = m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v], stats=st)
        hit.append(r)
        depth.append(d)
        dense_rows.append(orc.bitset_from_ranks(r, words))
        ok, pts, ranks = m.precept(it, w["pose_world"][v], w["init_pos"][v])
        vox_hit.append(ranks)
        voxel_rows.append(orc.bitset_from_ranks(ranks, words))
    dense_rows = np.stack(dense_rows)
    voxel_rows = np.stack(voxel_rows)
    seq, gain, cov, scored = orc.greedy(dense_rows, 0, 64)
    rgba, sdepth, sidx = orc.splat(w["cloud"], w["cloud_rgb"], it, w["pose_world"][1], 5)
    return {
        "name": name, "n_views": n_views, "size": list(size), "n_points": int(len(w["cloud"])), "full_voxels": int(m.n), "words": int(words),
        "keys_sha": sha(w["keys"]), "pose_world_sha": sha(w["pose_world"]), "init_pos_sha": sha(w["init_pos"]),
        "pose_world_view0": [float.hex(float(x)) for x in w["pose_world"][0].reshape(-1)],
        "dense_counts": [int(np.unpackbits(r.view(np.uint8)).sum()) for r in dense_rows],
        "voxel_counts": [int(np.unpackbits(r.view(np.uint8)).sum()) for r in voxel_rows],
        "dense_rows_sha": sha(dense_rows), "voxel_rows_sha": sha(voxel_rows), "dense_hit_sha": sha(np.stack(hit)),
        "dense_depth_sha": sha(np.stack(depth)), "voxel_hit_sha": sha(np.stack(vox_hit)),
        "stats": st.as_dict(), "greedy_seq": seq.tolist(), "greedy_gain": gain.tolist(), "greedy_scored": int(scored),
        "splat_rgba_sha": sha(rgba), "splat_depth_sha": sha(sdepth), "splat_opaque": int((rgba[..., 3] == 255).sum()),
    }


if __name__ == "__main__":
    out = {"generator": "tests/golden/make_golden.py (CPU oracle)", "cases": [compute(*c) for c in CASES]}
    with open(os.path.join(HERE, "golden_small.json"), "w") as f:
        json.dump(out, f, indent=1)
    for c in out["cases"]:
        print(c["name"], c["full_voxels"], c["dense_counts"], c["greedy_seq"], c["stats"])
