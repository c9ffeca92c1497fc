#!/usr/bin/env python
"""bench.py -- rays/s of the PRV ray-cast + coverage + greedy hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # own arm (sm_100a kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU oracle on the host cores

A "step" is one pass of the hot path over one object: dense per-pixel ray cast of every candidate view
(first-hit voxel rank + depth per pixel, coverage bitset per view), per-view coverage counts, and the
greedy set-cover selection (first view 0, num_of_max_iteration = 64).  Workload at N = 1: BASELINE config C2
(~200k-point synthetic cloud, 0.001 m voxels, the reference's 100-view hemisphere, 640x480).  N > 1: every rank
processes its own object of the same shape (objects sharded across GPUs as in config C4; no data-path
collective) => weak scaling; `--workload C3` runs the 1024-view strong-scaling case with the NCCL bitset
all-gather instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "rays_per_sec"
UNIT = "rays/s"
GREEDY_MAX_ITER = 64  # DefaultConfiguration.yaml:26 num_of_max_iteration


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _oracle_intr(orc, it):
    return orc.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, it.model, list(it.coeffs))


def cpu_sample(w, view_ids, threads):
    """Times the oracle (CPU restatement of the reference path) on a bounded sample of the workload: dense cast of
    `view_ids`, their coverage rows, and a greedy pass over those rows.  Returns (rays, seconds, stats)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.build()
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = _oracle_intr(orc, w["intr"])
    words = orc.bitset_words(m.n)
    st = orc.CastStats()
    t0 = time.perf_counter()
    rows = []
    for v in view_ids:
        ok, ranks, depth = m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v], max_range=1.0, want_depth=True, stats=st, num_threads=threads)
        rows.append(orc.bitset_from_ranks(ranks, words))
    orc.greedy(np.stack(rows), 0, GREEDY_MAX_ITER)
    dt = time.perf_counter() - t0
    return len(view_ids) * it.width * it.height, dt, st.as_dict()


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (oracle port; the reference itself cannot be compiled here,
    see DESIGN.md) on all host threads, each step a bounded sample of the same workload."""
    rank, world, local = _dist_env()
    if rank != 0:
        return 0
    import load_pkg
    prv = load_pkg.load()
    from nerf_prv_b200 import synth
    w = synth.build_workload(prv, args.workload)
    threads = os.cpu_count() or 1
    V = w["n_views"]
    per_step = max(1, args.ref_views)
    times, rays = [], 0
    for s in range(args.warmup + args.steps):
        ids = [(s * per_step + k) * 7 % V for k in range(per_step)]
        r, dt, _ = cpu_sample(w, ids, threads)
        if s >= args.warmup:
            times.append(dt)
            rays += r
    total = sum(times)
    value = rays / total
    sample = "%d of %d views of %s per step (dense cast + rows + greedy over the sampled rows), %d steps" % (per_step, V, args.workload, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / max(1, len(times)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": _config(w, args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def _config(w, args, world):
    return {"workload": "%s: synthetic %s cloud %d pts -> %d voxels @ %.3f m, %d hemisphere views, %dx%d, dense per-pixel cast + coverage + greedy(%d)"
                        % (args.workload, w["name"], len(w["cloud"]), len(w["keys"]), w["resolution"], w["n_views"], w["W"], w["H"], GREEDY_MAX_ITER),
            "objects_per_step": world if args.workload != "C3" else 1, "views": w["n_views"], "width": w["W"], "height": w["H"],
            "voxels": int(len(w["keys"])), "resolution_m": w["resolution"], "l2": "flushed between timed steps (256 MiB memset, untimed)",
            "variant": args.variant, "brick": getattr(args, "brick", 8), "brick_entry": bool(getattr(args, "brick_entry", 1))}


def run_own(args):
    rank, world, local = _dist_env()
    import load_pkg
    prv = load_pkg.load()
    from nerf_prv_b200 import synth
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    strong = args.workload == "C3" and world > 1
    from nerf_prv_b200 import sharding
    # every rank its own object (weak scaling) except in the C3 strong-scaling mode
    obj_index = rank if args.workload == "C4" else 0
    w = synth.build_workload(prv, args.workload, obj_index=obj_index)
    ctx = prv.Context(local)
    ctx.set_variant(args.variant)
    ctx.set_brick_cull(args.brick, bool(args.brick_entry))  # tuning only: results are identical for every setting (include/prv.h)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    V = w["n_views"]
    if strong:
        ids = sharding.pad_view_ids(sharding.shard_view_ids(V, rank, world), V, rank, world)  # interleaved view sharding
        real = ids < V
        shard_pose = np.where(real[:, None, None], w["pose_world"][np.minimum(ids, V - 1)], np.eye(4)[None])
        # padded slots: a position outside the key range -> "View out of map" -> empty coverage row
        shard_init = np.where(real[:, None], w["init_pos"][np.minimum(ids, V - 1)], 1.0e6)
        if rank == 0:
            uid = prv.comm_unique_id()
        else:
            uid = bytes(128)
        import torch
        t = torch.tensor(list(uid), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        ctx.comm_init(bytes(t.cpu().tolist()), rank, world)
        ctx.set_views(shard_pose, shard_init, view_ids=ids)
        local_views = int(real.sum())
    else:
        ctx.set_views(w["pose_world"], w["init_pos"])
        local_views = V
    rays_per_step_local = local_views * w["W"] * w["H"]

    def step():
        ctx.cast_async(prv.MODE_DENSE, want_pixels=True)
        if strong:
            ctx.allgather_bitsets_async()
        ctx.greedy_async(0, GREEDY_MAX_ITER)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        step()
        ctx.flush_l2()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.timing_reset()
    ctx.reset_counters()
    step_ms = []
    barrier()
    for _ in range(args.steps):
        ctx.event_record(0)
        step()
        ctx.event_record(1)
        step_ms.append(ctx.event_elapsed_ms(0, 1))  # CUDA events on the launching stream
        ctx.flush_l2()                              # untimed: next step starts with a cold L2
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    timing = ctx.get_timing()
    counters = ctx.get_counters()
    stats = ctx.get_cast_stats()
    total_ms = float(sum(step_ms))
    if dist is not None:
        import torch
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    rays_total = (rays_per_step_local * world if not strong else V * w["W"] * w["H"]) * args.steps
    value = rays_total / (total_ms * 1e-3)
    seq, gains, _ = ctx.get_greedy(GREEDY_MAX_ITER)
    views_scored = sum(max(V - 1 - k, 0) for k in range(min(len(seq), GREEDY_MAX_ITER)))

    # ---- end-to-end through the host-buffer C ABI: H2D of the map + poses, D2H of bitsets, counts, greedy sequence
    e2e_steps = max(1, min(args.steps, 5))
    pw = np.ascontiguousarray(w["pose_world"] if not strong else shard_pose)
    ip = np.ascontiguousarray(w["init_pos"] if not strong else shard_init)
    try:
        import torch
        def pinned(a):
            t = torch.from_numpy(a).pin_memory()
            return t.numpy(), t
        keys_h, _k = pinned(np.ascontiguousarray(w["keys"]))
        rgb_h, _r = pinned(np.ascontiguousarray(w["map_rgb"]))
        pw_h, _p = pinned(pw)
        ip_h, _i = pinned(ip)
        # result buffers of the step (coverage rows + counts) are caller-owned pinned memory too
        bits_h, _b = pinned(np.zeros((pw.shape[0], ctx.words), dtype=np.uint64))
        cnt_h, _c = pinned(np.zeros(pw.shape[0], dtype=np.uint32))
        pinned_note = "pinned"
    except Exception:
        keys_h, rgb_h, pw_h, ip_h = w["keys"], w["map_rgb"], pw, ip
        bits_h = cnt_h = None
        pinned_note = "pageable"
    def e2e_step():
        ctx.set_map(keys_h, rgb_h, w["resolution"])
        ctx.set_camera(w["intr"], 1.0)
        if not strong:
            ctx.cast_views(pw_h, ip_h, mode=prv.MODE_DENSE, want_bitsets=True, want_counts=True, out_bitsets=bits_h, out_counts=cnt_h)
        else:
            ctx.set_views(pw_h, ip_h, view_ids=ids)
            ctx.cast_async(prv.MODE_DENSE, False)
            ctx.allgather_bitsets_async()
            ctx.get_bitsets()
            ctx.get_coverage_counts()
        return ctx.greedy(0, GREEDY_MAX_ITER)
    e2e_step()
    barrier()
    ctx.reset_counters()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e_seq, _g = e2e_step()
    ctx.sync()
    e2e_dt = time.perf_counter() - t0
    c2 = ctx.get_counters()
    if dist is not None:
        import torch
        t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = (rays_per_step_local * world if not strong else V * w["W"] * w["H"]) * e2e_steps / e2e_dt
    assert e_seq.tolist() == seq.tolist(), "e2e and resident paths disagree"

    per_rank = None
    if dist is not None:
        mine = {"rank": rank, "step_ms": float(sum(step_ms)) / args.steps, "cast_ms": timing["cast_ms"] / args.steps,
                "greedy_ms": timing["greedy_ms"] / args.steps, "allgather_ms": timing["other_ms"] / args.steps, "marched": stats["marched"]}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (march_kernel): algorithmic bytes per launch / CUDA-event time per launch.
    # S_in (in-AABB probes of the literal algorithm) is a property of the input: it is counted once, untimed, with the
    # FAST variant, which executes every probe of the reference algorithm (the AXIS pipeline proves some misses cheaper).
    peak, peak_src = _peaks()
    bitmap_bytes = ctx.map_bytes()
    ctx.set_variant(prv.VARIANT_FAST)
    ctx.cast_async(prv.MODE_DENSE, want_pixels=False)
    s_in = ctx.get_cast_stats()["probes_in"]
    ctx.set_variant(args.variant)
    per_launch_rays = stats["rays"]
    alg_total = 4 * s_in + 8 * per_launch_rays + local_views * (bitmap_bytes + ctx.words * 8)
    launches = max(1, timing["march_launches"] if args.variant == 2 else timing["cast_launches"])
    march_ms = (timing["march_ms"] if args.variant == 2 else timing["cast_ms"]) / launches
    # cull + coarse + march of one step (the K_CULL span holds two launches per step: cull_kernel and coarse_kernel)
    pipeline_ms = (timing["cull_ms"] + timing["march_ms"]) / args.steps
    # the cull kernel writes the 8 B/ray "no hit" records of the rays it proves to miss; everything else is the march kernel's
    alg_march = alg_total - 8 * (per_launch_rays - stats["marched"]) if args.variant == 2 else alg_total
    achieved = alg_march / (march_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "march_kernel" if args.variant == 2 else "raycast_kernel", "algorithmic_bytes_per_launch": int(alg_march),
                "ms_per_launch": march_ms, "peak_source": peak_src, "share_of_step": march_ms * launches / total_ms,
                "s_in_probes": int(s_in),
                "cast_pipeline": {"kernels": "cull_kernel+coarse_kernel+march_kernel", "algorithmic_bytes": int(alg_total), "ms": pipeline_ms,
                                  "achieved": alg_total / (pipeline_ms * 1e-3) / 1e9, "frac": alg_total / (pipeline_ms * 1e-3) / 1e9 / peak},
                "note": "per ray 4*S_in + 8 B, per view bitmap + bitset row (SURVEY 8(d)); the march is FP64-add/issue bound, not bandwidth bound (DESIGN.md)"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(args.workload)
        except Exception:
            pass

    # ---- CPU baseline beside it: the oracle on a bounded sample of the same workload
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample_ids = [int(i) for i in np.linspace(0, V - 1, args.cpu_views).astype(int)]
        r, dt, cst = cpu_sample(w, sample_ids, threads)
        cpu = {"value": r / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d of %d views of %s (dense cast + rows + greedy over the sampled rows), %.1f s" % (len(sample_ids), V, args.workload, dt)}

    # the reference's own execution structure (one std::thread per voxel in batches of
    # num_of_thread = 20, pose inverse per voxel, sparse lookup; main.cpp:124-130, 238-284) on one view of C1
    ref_structure = None
    if not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle as orc
            w1 = synth.build_workload(prv, "C1", n_views=1)
            m1 = orc.Map.from_keys(w1["keys"], w1["map_rgb"], w1["resolution"])
            t0 = time.perf_counter()
            m1.precept_threads(_oracle_intr(orc, w1["intr"]), w1["pose_world"][0], w1["init_pos"][0], 1.0, 20)
            dt = time.perf_counter() - t0
            ref_structure = {"value": m1.n / dt, "unit": "voxel rays/s", "sample": "Perception_3D::precept of 1 view of C1 (%d voxels), %.2f s" % (m1.n, dt),
                             "structure": "std::thread per voxel, batches of 20, joined per batch"}
        except Exception as exc:  # never let the side baseline break the bench line
            ref_structure = {"error": str(exc)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": _config(w, args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": c2["h2d_bytes"] // e2e_steps, "d2h_bytes_per_step": c2["d2h_bytes"] // e2e_steps,
                    "steps": e2e_steps, "host_memory": pinned_note},
            "gpu_launches": counters["kernel_launches"], "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "cpu_reference_structure": ref_structure,
            "views_scored_per_sec": views_scored * world / max(1e-9, timing["greedy_ms"] / args.steps * 1e-3) if not strong else
                                    views_scored / max(1e-9, timing["greedy_ms"] / args.steps * 1e-3),
            "kernel_ms_per_step": {k: timing[k] / args.steps for k in ("cast_ms", "cull_ms", "march_ms", "count_ms", "greedy_ms", "other_ms")},
            "per_rank": per_rank, "cast_stats": stats, "greedy_len": int(len(seq)), "greedy_seq": [int(x) for x in seq], "coverage_rate": float(gains.sum()) / max(1, ctx.full_voxels)}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    # stdout carries exactly ONE JSON line: anything libraries print there (e.g. "NCCL version ...") goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--variant", type=int, default=2)
    ap.add_argument("--brick", type=int, default=8, choices=[4, 8, 16], help="prv_set_brick_cull: brick edge of the conservative cull in voxels")
    ap.add_argument("--brick-entry", type=int, default=1, choices=[0, 1], help="prv_set_brick_cull: start the exact march at the first set brick")
    ap.add_argument("--cpu-views", type=int, default=16, help="views in the cpu_baseline sample (~10-30 s of CPU work across the host cores)")
    ap.add_argument("--ref-views", type=int, default=2, help="views per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
