"""Host-side time of each C-ABI call of the end-to-end step (C2): where the e2e/resident gap goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import load_pkg
prv = load_pkg.load()
from nerf_prv_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
w = synth.build_workload(prv, name)
ctx = prv.Context(0)
def t(label, fn, n=10):
    fn(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    ctx.sync()
    dt = (time.perf_counter() - t0) / n * 1e3
    print("%-28s %8.3f ms" % (label, dt))
    return r
t("set_map", lambda: ctx.set_map(w["keys"], w["map_rgb"], w["resolution"]))
t("set_camera", lambda: ctx.set_camera(w["intr"], 1.0))
t("set_views", lambda: ctx.set_views(w["pose_world"], w["init_pos"]))
t("cast_async+sync (no pixels)", lambda: (ctx.cast_async(prv.MODE_DENSE, False), ctx.sync()))
t("cast_async+sync (pixels)", lambda: (ctx.cast_async(prv.MODE_DENSE, True), ctx.sync()))
t("get_bitsets", lambda: ctx.get_bitsets())
t("get_coverage_counts", lambda: ctx.get_coverage_counts())
t("cast_views (one call)", lambda: ctx.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE))
t("greedy (one call)", lambda: ctx.greedy(0, 64))
t("voxel cast_async+sync", lambda: (ctx.cast_async(prv.MODE_VOXEL, False), ctx.sync()))
print("voxel stats", ctx.get_cast_stats())
ctx.set_cloud(w["cloud"], w["cloud_rgb"])
t("render_async+sync (V views)", lambda: (ctx.render_async(w["n_views"], 5), ctx.sync()), n=3)
print("timing", ctx.get_timing())
