"""Freezes the oracle's results for the WHOLE of BASELINE C3 -- the north-star strong-scaling workload: 1024 Fibonacci-
hemisphere views at 1280x960, 1.258 G rays -- into golden_c3.json, and for three full-size C4 objects (0, 7, 63; 100 views
at 640x480 each) into golden_c4.json.  Per view: SHA-256 (first 16 hex digits) of the first-hit ranks, the depths and the
coverage row, and the coverage count; per workload: the greedy sequence / gains / covered mask, rays, hits, S_in.

    python tests/golden/make_golden_c3.py [C3] [C4]      # C3: ~30-40 min on 8 cores (oracle ~0.6 M rays/s); resumable

The workload is synthesised through oracle.HostShim, i.e. without the product library.  The multi-GPU parity gate
(tests/test_zz_gpu_full_size.py, tests/test_gpu_multi.py, bench.py strong mode) compares every rank's gathered rows and
greedy sequence with this file.
"""
import hashlib
import json
import os
import pickle
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import load_pkg  # noqa: E402
import oracle as orc  # noqa: E402

load_pkg.load()
from nerf_prv_b200 import synth  # noqa: E402


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def compute(name, obj_index=0, ckpt=None):
    w = synth.build_workload(orc.HostShim, name, obj_index=obj_index)
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                             list(w["intr"].coeffs))
    words = orc.bitset_words(m.n)
    V = w["n_views"]
    state = {"rows": [], "hit": [], "depth": [], "row": [], "counts": [], "hits": [], "rays": 0, "s_in": 0}
    if ckpt and os.path.exists(ckpt):
        state = pickle.load(open(ckpt, "rb"))
    t0 = time.time()
    for v in range(len(state["rows"]), V):
        st = orc.CastStats()
        ok, r, d = m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v], stats=st)
        row = orc.bitset_from_ranks(r, words)
        state["rows"].append(row)
        state["hit"].append(sha16(r))
        state["depth"].append(sha16(d))
        state["row"].append(sha16(row))
        state["counts"].append(int(np.unpackbits(row.view(np.uint8)).sum()))
        s = st.as_dict()
        state["hits"].append(int(s["hits"]))
        state["rays"] += int(s["rays"])
        state["s_in"] += int(s["probes_in"])
        if ckpt and (v % 16 == 15 or v == V - 1):
            pickle.dump(state, open(ckpt + ".tmp", "wb"))
            os.replace(ckpt + ".tmp", ckpt)
            print("%s view %d/%d  %.0f s" % (name, v + 1, V, time.time() - t0), flush=True)
    rows = np.stack(state["rows"])
    seq, gain, cov, scored = orc.greedy(rows, 0, 64)
    return {"name": name, "obj_index": obj_index, "n_views": int(V), "size": [int(w["W"]), int(w["H"])], "full_voxels": int(m.n), "words": int(words),
            "keys_sha": sha16(w["keys"]), "pose_world_sha": sha16(w["pose_world"]), "init_pos_sha": sha16(w["init_pos"]),
            "hit_sha16": state["hit"], "depth_sha16": state["depth"], "row_sha16": state["row"], "counts": state["counts"], "hits_per_view": state["hits"],
            "rows_sha": hashlib.sha256(rows.tobytes()).hexdigest(), "rays": state["rays"], "hits": int(sum(state["hits"])), "s_in": state["s_in"],
            "greedy_seq": seq.tolist(), "greedy_gain": gain.tolist(), "covered_sha": hashlib.sha256(cov.tobytes()).hexdigest(),
            "views_scored": int(scored)}


if __name__ == "__main__":
    what = sys.argv[1:] or ["C3", "C4"]
    if "C4" in what:
        out = {"generator": "tests/golden/make_golden_c3.py (CPU oracle, workload through oracle.HostShim)",
               "cases": [compute("C4", k) for k in (0, 7, 63)]}
        json.dump(out, open(os.path.join(HERE, "golden_c4.json"), "w"), indent=0)
        for c in out["cases"]:
            print("C4 object", c["obj_index"], c["rays"], "rays", c["hits"], "hits", "greedy", len(c["greedy_seq"]))
    if "C3" in what:
        c = compute("C3", ckpt="/tmp/golden_c3.ckpt")
        out = {"generator": "tests/golden/make_golden_c3.py (CPU oracle, workload through oracle.HostShim)", "cases": [c]}
        json.dump(out, open(os.path.join(HERE, "golden_c3.json"), "w"), indent=0)
        print("C3", c["rays"], "rays", c["hits"], "hits", "S_in", c["s_in"], "greedy", c["greedy_seq"][:8], len(c["greedy_seq"]))
