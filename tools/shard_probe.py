"""One rank's share of the strong-scaling workload on ONE GPU: the views {r, r + N, r + 2N, ...} of C3 cast 20 times (L2 flushed in
between), kernel times from the C ABI's event spans.  For tuning the per-rank path without an N-GPU box.
usage: python tools/shard_probe.py [N=8] [rank=0]   (PRV_B200_LIB selects an A/B build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import load_pkg
prv = load_pkg.load()
import bench
synth = bench._synth()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
r = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w = synth.build_workload(prv, "C3")
ids = np.arange(r, w["n_views"], N)
ctx = prv.Context(0)
ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
ctx.set_camera(w["intr"], 1.0)
ctx.set_views(np.ascontiguousarray(w["pose_world"][ids]), np.ascontiguousarray(w["init_pos"][ids]))
for _ in range(3):
    ctx.cast_async(prv.MODE_DENSE, want_pixels=True)
    ctx.flush_l2()
ctx.sync()
ctx.timing_reset()
K = 20
for _ in range(K):
    ctx.cast_async(prv.MODE_DENSE, want_pixels=True)
    ctx.flush_l2()
ctx.sync()
t = ctx.get_timing()
print("%s views %d of C3/%d: cull+coarse %.4f ms  march %.4f ms  cast %.4f ms  (x%d = %.3f ms vs one GPU)" % (
    os.environ.get("PRV_B200_LIB", "default"), len(ids), N, t["cull_ms"] / K, t["march_ms"] / K, (t["cull_ms"] + t["march_ms"]) / K, N, N * (t["cull_ms"] + t["march_ms"]) / K))
