#!/usr/bin/env python
"""Warp-level instruction model of coarse_kernel + march_kernel (see tools/march_model.cpp).  PLANNING TOOL.

    python tools/march_model.py [C1|C2|C3|C5] [view stride]

Prints, for the default pipeline and for candidate changes, the modelled warp-instructions of the coarse and march
kernels.  Calibration: the per-trip instruction counts below were read off the round-1 sm_100a SASS; the fixed per-warp
costs are set so that the default C2 totals equal ncu's 416.9 M warp-instructions for march_kernel and 94.5 M for
coarse_kernel (profiles/r1_ncu_source_lines_march.txt, profiles/r1_ncu_full_summary.md).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# SASS-derived costs (warp-instructions)
LOOP_ADD = 1.75      # counted DADD loops, unrolled by 4: 4 DADD + 3 per trip of 4
LOOP_FIX = 10.0      # per counted loop: trip-count set-up, remainder
CMP_ITER = 4.0       # while (t < thr) { t += d; n++; }
INNER4 = 81.0        # one trip of the 4-probes-in-flight in-AABB loop
INNER_EXIT = 25.0    # decode of the stopping cell
WALK_ITER = 14.0     # one float cell step of the brick walks
COARSE_FIX = None    # per coarse-kernel warp: ticket, view prefix, direction, slab of the grown box, cell set-up, compaction (calibrated)
MARCH_FIX = None     # per march-kernel warp: set-up, windows, epilogue (calibrated)


def load():
    out = "/tmp/libmarch_model.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-w", "-o", out,
                    os.path.join(ROOT, "tools", "march_model.cpp")], check=True)
    return C.CDLL(out)


REC = np.dtype([("pid", "<u4"), ("hit", "u1"), ("fine_keep", "u1", 3), ("walk8", "<u2"), ("walkf", "<u2", 3), ("p1", "<u2", 3), ("m2", "<u2", 3),
                ("e2", "<u2", 3), ("nprobe", "<u2"), ("kentry", "<u2", 3)], align=True)


def collect(lib, w, views):
    assert lib.model_rec_size() == REC.itemsize, (lib.model_rec_size(), REC.itemsize)
    keys = np.ascontiguousarray(w["keys"], dtype=np.uint16)
    cap = w["W"] * w["H"]
    per_view = []
    for v in views:
        recs = np.zeros(cap, dtype=REC)
        keep8 = np.zeros(cap, dtype=np.uint8)
        pw = np.ascontiguousarray(w["pose_world"][v], dtype=np.float64)
        ip = np.ascontiguousarray(w["init_pos"][v], dtype=np.float64)
        n = lib.model_view(keys.ctypes.data_as(C.c_void_p), C.c_uint32(len(keys)), C.c_double(w["resolution"]), C.byref(w["intr"]),
                           pw.ctypes.data_as(C.c_void_p), ip.ctypes.data_as(C.c_void_p), recs.ctypes.data_as(C.c_void_p),
                           keep8.ctypes.data_as(C.c_void_p), cap)
        assert n >= 0, "model_view failed: %d" % n
        per_view.append((recs[:n].copy(), keep8[:n].astype(bool)))
    return per_view


def warp_max(a, width=32):
    """max over consecutive groups of `width` entries (last group partial)."""
    n = len(a)
    if n == 0:
        return np.zeros(0)
    pad = (-n) % width
    b = np.concatenate([a, np.zeros(pad, dtype=a.dtype)]).reshape(-1, width)
    return b.max(axis=1)


def march_cost(recs, extra_approach=None, nprobe=None, fixed=0.0):
    """warp-instructions of march_kernel over `recs` in warp order."""
    if len(recs) == 0:
        return dict(total=0.0, approach=0.0, inner=0.0, fixed=0.0, warps=0, lane_eff=1.0)
    nprobe = recs["nprobe"].astype(np.int64) if nprobe is None else nprobe
    approach = 0.0
    for a in range(3):
        approach += (LOOP_ADD * warp_max(recs["p1"][:, a].astype(np.int64)) + LOOP_FIX).sum()
        approach += (LOOP_ADD * warp_max(recs["m2"][:, a].astype(np.int64)) + CMP_ITER * warp_max(recs["e2"][:, a].astype(np.int64)) + LOOP_FIX).sum()
    if extra_approach is not None:  # steps moved from the merged in-AABB loop to a per-axis approach (three counted loops again)
        approach += (LOOP_ADD * warp_max(extra_approach) + 3 * LOOP_FIX + 3 * CMP_ITER * 2).sum()
    trips = (nprobe + 3) // 4
    wt = warp_max(trips)
    inner = (INNER4 * wt + INNER_EXIT).sum()
    warps = len(wt)
    lane_eff = trips.sum() / max(1.0, (wt * 32).sum())
    return dict(total=approach + inner + fixed * warps, approach=approach, inner=inner, fixed=fixed * warps, warps=warps, lane_eff=lane_eff)


def coarse_cost(walk_iters, fixed=None):
    """warp-instructions of the coarse kernel: 256-ray chunks = 8 consecutive warps of queue 1."""
    wt = warp_max(walk_iters.astype(np.int64))
    return float((WALK_ITER * wt + (COARSE_FIX if fixed is None else fixed)).sum())


def main():
    import load_pkg
    prv = load_pkg.load()
    from nerf_prv_b200 import synth
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    stride = int(sys.argv[2]) if len(sys.argv) > 2 else (1 if name in ("C1", "C2") else 16)
    lib = load()
    global MARCH_FIX, COARSE_FIX
    # calibration on C2 (all 100 views): fixed per-warp costs such that the default totals = ncu's 416.9 M / 94.5 M
    cal_file = "/tmp/march_model_fix2.txt"
    if os.path.exists(cal_file):
        MARCH_FIX, COARSE_FIX = (float(x) for x in open(cal_file).read().split())
    else:
        wc = synth.build_workload(prv, "C2")
        pv = collect(lib, wc, range(wc["n_views"]))
        var = sum(march_cost(r[k])["total"] for r, k in pv)
        warps = sum(march_cost(r[k])["warps"] for r, k in pv)
        MARCH_FIX = float((416.9e6 - var) / warps)
        cvar = sum(coarse_cost(r["walk8"], 0.0) for r, _ in pv)
        cwarps = sum((len(r) + 31) // 32 for r, _ in pv)
        COARSE_FIX = float((94.47e6 - cvar) / cwarps)
        open(cal_file, "w").write("%r %r" % (MARCH_FIX, COARSE_FIX))
        print("calibration on C2: march loops %.1f M over %d warps -> fixed %.0f per warp; coarse walk %.1f M over %d warps -> fixed %.0f per warp" %
              (var / 1e6, warps, MARCH_FIX, cvar / 1e6, cwarps, COARSE_FIX))
    w = synth.build_workload(prv, name)
    views = list(range(0, w["n_views"], stride))
    pv = collect(lib, w, views)
    scale = w["n_views"] / len(views)

    def report(label, march_parts, coarse_total):
        t = sum(p["total"] for p in march_parts) * scale
        ap = sum(p["approach"] for p in march_parts) * scale
        inn = sum(p["inner"] for p in march_parts) * scale
        fx = sum(p["fixed"] for p in march_parts) * scale
        wr = sum(p["warps"] for p in march_parts) * scale
        trips_w = sum(p["lane_eff"] * p["inner"] for p in march_parts) / max(1.0, sum(p["inner"] for p in march_parts))
        print("%-44s march %7.1f M (approach %6.1f, in-AABB %6.1f @ %2.0f%% lanes, fixed %6.1f; %7.0f k warps)  coarse %6.1f M  sum %7.1f M" %
              (label, t / 1e6, ap / 1e6, inn / 1e6, 100 * trips_w, fx / 1e6, wr / 1e3, coarse_total * scale / 1e6, (t + coarse_total * scale) / 1e6))
        return t + coarse_total * scale

    print("%s: %d of %d views, %d slab survivors, %d marched by default, %d hits" % (name, len(views), w["n_views"], sum(len(r) for r, _ in pv),
                                                                                  sum(int(k.sum()) for _, k in pv), sum(int(r["hit"].sum()) for r, _ in pv)))
    base = report("default (8^3 bricks)", [march_cost(r[k], fixed=MARCH_FIX) for r, k in pv], sum(coarse_cost(r["walk8"]) for r, _ in pv))
    for ki, K in enumerate((4, 2, 1)):
        keep = [r["fine_keep"][:, ki].astype(bool) for r, _ in pv]
        cc = sum(coarse_cost(r["walkf"][:, ki]) for r, _ in pv)
        t = report("fine cull K=%d" % K, [march_cost(r[kp], fixed=MARCH_FIX) for (r, _), kp in zip(pv, keep)], cc)
        # + start the exact merged march at the box of the first set fine cell
        parts = []
        for (r, _), kp in zip(pv, keep):
            rr = r[kp]
            ke = np.minimum(rr["kentry"][:, ki].astype(np.int64), rr["nprobe"].astype(np.int64))
            parts.append(march_cost(rr, extra_approach=ke, nprobe=rr["nprobe"].astype(np.int64) - ke + 1, fixed=MARCH_FIX + 40))
        report("fine cull K=%d + entry at the fine cell" % K, parts, cc)
        # + hits and misses in different warps (perfect classification: upper bound)
        parts = []
        for (r, _), kp in zip(pv, keep):
            rr = r[kp]
            for cls in (rr[rr["hit"] == 1], rr[rr["hit"] == 0]):
                ke = np.minimum(cls["kentry"][:, ki].astype(np.int64), cls["nprobe"].astype(np.int64))
                parts.append(march_cost(cls, extra_approach=ke, nprobe=cls["nprobe"].astype(np.int64) - ke + 1, fixed=MARCH_FIX + 40))
        report("  ... + hits / misses in separate warps", parts, cc)
    parts = []
    for r, k in pv:
        rr = r[k]
        for cls in (rr[rr["hit"] == 1], rr[rr["hit"] == 0]):
            parts.append(march_cost(cls, fixed=MARCH_FIX))
    report("default + hits / misses in separate warps", parts, sum(coarse_cost(r["walk8"]) for r, _ in pv))


if __name__ == "__main__":
    main()
