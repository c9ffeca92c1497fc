"""Times prv_greedy_async alone (CUDA events) for the cluster and the grid-barrier kernels on a bench workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import load_pkg
prv = load_pkg.load()
from nerf_prv_b200 import synth
for wl in sys.argv[1:] or ["C2", "C3"]:
    w = synth.build_workload(prv, wl)
    for mode in ("1", "0"):
        os.environ["PRV_GREEDY_CLUSTER"] = mode
        ctx = prv.Context(0)
        ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
        ctx.set_camera(w["intr"], 1.0)
        ctx.set_views(w["pose_world"], w["init_pos"])
        ctx.cast_async(prv.MODE_DENSE, want_pixels=False)
        ctx.sync()
        for it in (1, 17, 33):
            ms = []
            for i in range(8):
                ctx.flush_l2()
                ctx.event_record(0)
                ctx.greedy_async(0, it)
                ctx.event_record(1)
                ms.append(ctx.event_elapsed_ms(0, 1))
            print("   max_iter", it, "median ms %.4f" % sorted(ms)[len(ms) // 2])
        ms = []
        for i in range(12):
            ctx.flush_l2()
            ctx.event_record(0)
            ctx.greedy_async(0, 64)
            ctx.event_record(1)
            ms.append(ctx.event_elapsed_ms(0, 1))
        seq, gains, _ = ctx.get_greedy(64)
        print(wl, "cluster" if mode == "1" else "grid", "median ms %.4f min %.4f" % (sorted(ms)[len(ms) // 2], min(ms)), "len", len(seq), seq[:8].tolist())
        ctx.close()
