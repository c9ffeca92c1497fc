#!/usr/bin/env python
"""bench.py -- rays/s of the PRV ray-cast + coverage + greedy hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # own arm (sm_100a kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU oracle on the host cores

A "step" is one pass of the hot path over one object: dense per-pixel ray cast of every candidate view
(first-hit voxel rank + depth per pixel, coverage bitset per view), per-view coverage counts, and the
greedy set-cover selection (first view 0, num_of_max_iteration = 64).

Default workload at EVERY N: BASELINE config C3, the north-star's "1024-view workload" (1024 Fibonacci-hemisphere
views at 1280x960, 1.258 G rays per step).  N = 1 casts all views on one GPU; N > 1 shards the views interleaved
across the ranks, all-gathers the coverage rows over NCCL inside the timed step and runs the selection replicated
=> STRONG scaling, same total work at every N.  The result of every timed run -- gathered rows, counts, greedy
sequence, on every rank -- is compared with the CPU oracle's frozen vectors (tests/golden/golden_c3.json).

At N = 1 the line also carries config C2 (BASELINE configs[1]: 100 views, 640x480, 0.001 m voxels; `c2`), the splat
z-buffer render of C5 (`c5_splat`) and the rooflines of the march, greedy and splat kernels.
`--workload C1|C2|C4|C5` selects another config as the headline (N > 1: one object per rank, weak scaling).
"""
import argparse
import hashlib
import json
import math
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
L2_BYTES = 126 << 20


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


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.build()
    return orc


def _oracle_intr(orc, it):
    return orc.make_intrinsics(it.width, it.height, it.fx, it.fy, it.ppx, it.ppy, it.model, list(it.coeffs))


def _synth():
    import load_pkg
    load_pkg.load()  # (imports the package; libprv_b200.so itself is only dlopen'ed by the first C-ABI call)
    from nerf_prv_b200 import synth
    return synth


def cpu_sample(orc, w, view_ids, threads):
    """Times the oracle (CPU restatement of the reference path) on a bounded sample of the workload: dense cast of
    `view_ids`, their coverage rows, and a greedy pass over those rows.  Returns (rays, seconds, stats)."""
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


def _workload_text(name, w):
    return ("%s: synthetic %s cloud %d pts -> %d voxels @ %.3f m, %d hemisphere views, %dx%d, dense per-pixel cast + coverage + greedy(%d)"
            % (name, w["name"], len(w["cloud"]), len(w["keys"]), w["resolution"], w["n_views"], w["W"], w["H"], GREEDY_MAX_ITER))


def _config(name, w, args, world, strong, l2_note):
    return {"workload": _workload_text(name, w), "objects_per_step": 1 if (strong or world == 1) else world, "views": w["n_views"], "width": w["W"],
            "height": w["H"], "voxels": int(len(w["keys"])), "resolution_m": w["resolution"], "l2": l2_note,
            "sharding": ("views interleaved over %d ranks + all-gather of the coverage rows in the timed step (%s)" %
                         (world, "peer-memory stores over NVLink fused into the count kernel" if getattr(args, "gather_used", args.gather) == "p2p" else "ncclAllGather")) if strong else
                        ("one object per rank, no collective" if world > 1 else "single GPU"),
            "variant": args.variant, "brick": args.brick, "brick_entry": bool(args.brick_entry), "stage_smem": bool(args.stage_smem), "stage_l2": bool(args.stage_l2)}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm on all host threads -- the oracle port (the reference itself cannot
    be compiled here), nothing of the product: the workload is synthesised through oracle.HostShim and
    libprv_b200.so is never loaded.  Each step casts a bounded sample of the workload's views (every view in turn over the
    steps), sized by a calibration view so that one step takes about --ref-seconds."""
    rank, world, local = _dist_env()
    if rank != 0:
        return 0
    orc = _oracle()
    synth = _synth()
    w = synth.build_workload(orc.HostShim, args.workload)
    threads = os.cpu_count() or 1
    V = w["n_views"]
    _, dt1, _ = cpu_sample(orc, w, [V // 2], threads)  # calibration (untimed)
    per_step = int(max(1, min(V, round(args.ref_seconds / max(dt1, 1e-3)))))
    times, rays = [], 0
    stride = max(1, V // per_step)
    for s in range(args.warmup + args.steps):
        ids = [(s + k * stride) % V for k in range(per_step)]  # spread over the hemisphere, shifted every step
        r, dt, _ = cpu_sample(orc, w, ids, threads)
        if s >= args.warmup:
            times.append(dt)
            rays += r
    total = sum(times)
    value = rays / total
    sample = ("%d of %d views of %s per step, spread over the hemisphere and shifted every step (dense cast + rows + greedy over the sampled rows), %d steps"
              % (per_step, V, args.workload, args.steps))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / max(1, len(times)), "higher_is_better": True, "scaling": "strong" if args.workload == "C3" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _config(args.workload, w, args, 1, False, "n/a (host)"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "product_library_loaded": any("libprv_b200" in l for l in open("/proc/self/maps"))}
    emit(line)
    return 0


class Runner:
    """One resident workload on one ctx: the timed step, kernel timing, e2e, rooflines."""

    def __init__(self, prv, name, w, args, local, dist, rank, world, strong):
        from nerf_prv_b200 import sharding
        self.prv, self.name, self.w, self.args, self.dist, self.rank, self.world, self.strong = prv, name, w, args, dist, rank, world, strong
        ctx = self.ctx = prv.Context(local)
        ctx.set_variant(args.variant)
        ctx.set_brick_cull(args.brick, bool(args.brick_entry))  # tuning only: results are identical for every setting (include/prv.h)
        ctx.set_staging(bool(args.stage_smem), bool(args.stage_l2))
        ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
        ctx.set_camera(w["intr"], 1.0)
        V = self.V = w["n_views"]
        if strong:
            ids = self.ids = sharding.pad_view_ids(sharding.shard_view_ids(V, rank, world), V, rank, world)  # interleaved view sharding
            real = ids < V
            self.pose = np.ascontiguousarray(np.where(real[:, None, None], w["pose_world"][np.minimum(ids, V - 1)], np.eye(4)[None]))
            # padded slots: a position outside the key range -> "View out of map" -> empty coverage row
            self.init = np.ascontiguousarray(np.where(real[:, None], w["init_pos"][np.minimum(ids, V - 1)], 1.0e6))
            import torch
            uid = prv.comm_unique_id() if rank == 0 else bytes(128)
            t = torch.tensor(list(uid), dtype=torch.uint8, device="cuda")
            dist.broadcast(t, 0)
            ctx.comm_init(bytes(t.cpu().tolist()), rank, world)
            self.p2p = args.gather == "p2p"
            if self.p2p:
                # peer-memory exchange: every rank's arena handle to every rank, then rows travel as NVLink stores from the count kernel.
                # If any rank cannot map a peer (no P2P / IPC between the devices) every rank uses ncclAllGather instead; the
                # transport actually used is named in config.sharding.
                ok = 1
                try:
                    handles = [None] * world
                    dist.all_gather_object(handles, ctx.p2p_export())
                    ctx.p2p_import(handles, rank, world)
                except prv.PrvError as exc:
                    sys.stderr.write("bench.py: rank %d: peer-memory exchange unavailable (%s); falling back to ncclAllGather\n" % (rank, exc))
                    ok = 0
                flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if int(flag.item()) == 0:
                    ctx.comm_destroy_p2p()
                    self.p2p = False
            args.gather_used = "p2p" if self.p2p else "nccl"
            ctx.set_views(self.pose, self.init, view_ids=ids)
            dist.barrier()
            self.local_views = int(real.sum())
        else:
            self.ids = None
            self.p2p = False
            self.pose, self.init = np.ascontiguousarray(w["pose_world"]), np.ascontiguousarray(w["init_pos"])
            ctx.set_views(self.pose, self.init)
            self.local_views = V
        self.px = w["W"] * w["H"]
        self.rays_local = self.local_views * self.px
        self.rays_job = V * self.px if (strong or world == 1) else self.rays_local * world
        # Small workloads (per-pixel tables up to ~2x the L2): every step timed on its own between two events, L2 flushed in
        # between, no overlap between steps (round 1's method).  Large ones: the K steps are enqueued back to back inside one
        # event bracket so that the scoring of step k overlaps the cast of step k+1; the L2 is still flushed between steps (a
        # memset on the cast stream) and the memsets' own device time is taken out of the bracket.
        self.per_step_timing = self.rays_local * 8 <= 2 * L2_BYTES

    def step(self, flush=False):
        self.ctx.cast_async(self.prv.MODE_DENSE, want_pixels=True, publish=self.p2p)
        if flush:
            self.ctx.flush_l2()  # between the cast and its scoring: both streams continue behind it
        if self.strong:
            self.ctx.allgather_bitsets_async()
        self.ctx.greedy_async(0, GREEDY_MAX_ITER)

    def barrier(self):
        self.ctx.sync()
        if self.dist is not None:
            self.dist.barrier()

    def _max_over_ranks(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, steps, warmup):
        """EXACTLY `steps` steps between two device events, enqueued back to back (the scoring of step k overlaps the cast of
        step k+1 on the ctx's second stream); barrier + synchronise on both sides; max over ranks."""
        ctx = self.ctx
        for _ in range(warmup):
            self.step()
            ctx.flush_l2()
        self.barrier()
        ctx.timing_reset()
        ctx.reset_counters()
        self.barrier()
        if self.per_step_timing:
            total = 0.0
            for _ in range(steps):
                ctx.event_record(0)
                self.step()
                ctx.event_record(1)
                total += ctx.event_elapsed_ms(0, 1)
                ctx.flush_l2()  # untimed
            self.barrier()
            self.timing = ctx.get_timing()
        else:
            ctx.event_record(0)
            for k in range(steps):
                self.step(flush=True)  # L2 flushed after every cast; the memsets are measured on their own and subtracted below
            ctx.event_record(1)
            total = ctx.event_elapsed_ms(0, 1)
            self.barrier()
            self.timing = ctx.get_timing()
            self.flush_ms = self.timing["flush_ms"]
            total -= self.flush_ms
        self.counters = ctx.get_counters()
        self.stats = ctx.get_cast_stats()
        self.local_ms = total
        return self._max_over_ranks(total)

    def l2_note(self):
        if self.per_step_timing:
            return "flushed between timed steps (256 MiB memset, untimed; steps timed one by one)"
        return ("flushed in every timed step (256 MiB memset after each cast, before its scoring and the next cast; the K steps run back to back inside "
                "one event bracket, nothing of the path runs underneath a memset, and the memsets' device time, %.3f ms in total, is subtracted); each "
                "step also rewrites %.2f GB of per-pixel tables per GPU"
                % (getattr(self, "flush_ms", 0.0), self.rays_local * 8 / 1e9))

    def e2e(self, steps):
        """The same step through the host-buffer C ABI: H2D of keys / colours / poses from pinned memory, D2H of the coverage
        rows, counts and the greedy sequence, every step, inside the timed region (wall clock around synchronous calls)."""
        prv, ctx, w = self.prv, self.ctx, self.w
        import torch

        def pinned(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t.numpy(), t
        keep = []
        keys_h, t = pinned(w["keys"]); keep.append(t)
        rgb_h, t = pinned(w["map_rgb"]); keep.append(t)
        pw_h, t = pinned(self.pose); keep.append(t)
        ip_h, t = pinned(self.init); keep.append(t)
        bits_h, t = pinned(np.zeros((self.pose.shape[0], ctx.words), dtype=np.uint64)); keep.append(t)
        cnt_h, t = pinned(np.zeros(self.pose.shape[0], dtype=np.uint32)); keep.append(t)

        def one():
            ctx.set_map(keys_h, rgb_h, w["resolution"])
            ctx.set_camera(w["intr"], 1.0)
            if not self.strong:
                ctx.cast_views(pw_h, ip_h, mode=prv.MODE_DENSE, want_bitsets=True, want_counts=True, out_bitsets=bits_h, out_counts=cnt_h)
            else:
                ctx.set_views(pw_h, ip_h, view_ids=self.ids)
                ctx.cast_async(prv.MODE_DENSE, False, publish=self.p2p)
                ctx.allgather_bitsets_async()
                ctx.get_bitsets()
                ctx.get_coverage_counts()
            return ctx.greedy(0, GREEDY_MAX_ITER)
        one()
        self.barrier()
        ctx.reset_counters()
        t0 = time.perf_counter()
        for _ in range(steps):
            seq, _g = one()
        ctx.sync()
        dt = self._max_over_ranks(time.perf_counter() - t0)
        c = ctx.get_counters()
        return {"value": self.rays_job * steps / dt, "unit": UNIT, "h2d_bytes_per_step": c["h2d_bytes"] // steps, "d2h_bytes_per_step": c["d2h_bytes"] // steps,
                "steps": steps, "host_memory": "pinned"}, seq

    def rooflines(self, steps, total_ms):
        """Dominant kernel (march_kernel): algorithmic bytes per launch / CUDA-event time per launch, SURVEY 8(d): per ray
        4*S_in + 8 B, per view bitmap + bitset row.  S_in (in-AABB probes of the literal algorithm) is a property of the input:
        counted once, untimed, with the FAST variant, which executes every probe of the reference algorithm."""
        prv, ctx, timing, stats = self.prv, self.ctx, self.timing, self.stats
        peak, peak_src = _peaks()
        ctx.set_variant(prv.VARIANT_FAST)
        ctx.cast_async(prv.MODE_DENSE, want_pixels=False)
        s_in = ctx.get_cast_stats()["probes_in"]
        ctx.set_variant(self.args.variant)
        axis = self.args.variant == 2
        rays = stats["rays"]
        words = ctx.words
        alg_total = 4 * s_in + 8 * rays + self.local_views * (ctx.map_bytes() + words * 8)
        launches = max(1, timing["march_launches"] if axis else timing["cast_launches"])
        march_ms = (timing["march_ms"] if axis else timing["cast_ms"]) / launches
        pipeline_ms = (timing["cull_ms"] + timing["march_ms"]) / steps
        # the cull / coarse kernels write the 8 B/ray "no hit" records of the rays they prove to miss; the rest is the march kernel's
        alg_march = alg_total - 8 * (rays - stats["marched"]) if axis else alg_total
        achieved = alg_march / (march_ms * 1e-3) / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get(self.name)
            except Exception:
                pass
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "march_kernel" if axis else "raycast_kernel", "algorithmic_bytes_per_launch": int(alg_march), "ms_per_launch": march_ms,
                "peak_source": peak_src, "share_of_step": march_ms * launches / max(1e-9, self.local_ms), "s_in_probes": int(s_in),
                "cast_pipeline": {"kernels": "cull_kernel+coarse_kernel+march_kernel", "algorithmic_bytes": int(alg_total), "ms": pipeline_ms,
                                  "achieved": alg_total / (pipeline_ms * 1e-3) / 1e9, "frac": alg_total / (pipeline_ms * 1e-3) / 1e9 / peak},
                "note": "per ray 4*S_in + 8 B, per view bitmap + bitset row (SURVEY 8(d)); the march is FP64-add / issue bound, its DRAM traffic is ~2 % of "
                        "these bytes (DESIGN.md section 7): a work rate expressed in bytes, not a bandwidth"}
        # scoring path: per view scored per iteration its row (8*words B) + the covered mask amortised over the table
        nrows = self.V if (self.strong or self.world == 1) else self.V
        n_sel = self.greedy_len
        scored = sum(max(nrows - 1 - k, 0) for k in range(min(n_sel, GREEDY_MAX_ITER)))
        iters = max(1, min(n_sel, GREEDY_MAX_ITER))
        g_ms = timing["greedy_ms"] / max(1, timing["greedy_launches"]) * (timing["greedy_launches"] / steps)
        g_bytes = scored * words * 8 + iters * words * 8
        g_ach = g_bytes / (g_ms * 1e-3) / 1e9
        roof_g = {"bound": "hbm", "kernel": ["greedy_cluster_kernel", "greedy_persistent_kernel", "greedy_iter_kernel"][max(0, ctx.greedy_path)],
                  "algorithmic_bytes_per_step": int(g_bytes), "ms_per_step": g_ms, "achieved": g_ach, "peak": peak, "unit": "GB/s", "frac": g_ach / peak,
                  "views_scored_per_step": int(scored), "iterations": int(iters), "us_per_iteration": 1e3 * g_ms / iters,
                  "note": "the table lives in the cluster's shared memory: no HBM traffic after the first pass; the selection is a chain of dependent "
                          "arg-max exchanges (one per pick), so its floor is iterations x exchange latency, not bytes / bandwidth (DESIGN.md 4.4)"}
        return roof, roof_g, scored


def _sha_rows(rows):
    return hashlib.sha256(np.ascontiguousarray(rows).tobytes()).hexdigest()


def parity_c3(runner, golden):
    """Every rank: the table its selection ran over (its own rows at N = 1, the all-gathered table at N > 1) re-ordered by view
    id, the coverage counts and the greedy result against the oracle's frozen full-size C3 vectors."""
    ctx, V = runner.ctx, runner.V
    seq, gains, cov = ctx.get_greedy()
    if runner.strong:
        rows, ids = ctx.get_gathered()
        table = np.zeros((V, ctx.words), dtype=np.uint64)
        real = ids < V
        table[ids[real]] = rows[real]
        pad_empty = bool(np.all(rows[~real] == 0)) if (~real).any() else True
    else:
        table = ctx.get_bitsets()
        pad_empty = True
    counts = np.unpackbits(table.view(np.uint8), axis=1).sum(axis=1)
    res = {"rows_sha": _sha_rows(table) == golden["rows_sha"], "counts": counts.tolist() == golden["counts"], "padding_rows_empty": pad_empty,
           "greedy_seq": seq.tolist() == golden["greedy_seq"], "greedy_gain": gains.tolist() == golden["greedy_gain"],
           "covered_sha": hashlib.sha256(cov.tobytes()).hexdigest() == golden["covered_sha"]}
    res["ok"] = all(res.values())
    return res, seq, gains


def measure(prv, synth, name, args, local, dist, rank, world, want_cpu, emit_golden=True):
    strong = name == "C3" and world > 1
    obj_index = rank if name == "C4" else 0
    w = synth.build_workload(prv, name, obj_index=obj_index)
    R = Runner(prv, name, w, args, local, dist, rank, world, strong)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms = R.timed(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    value = R.rays_job * args.steps / (total_ms * 1e-3)
    timing, stats = R.timing, R.stats

    # ---- parity gate on the timed configuration
    parity = None
    seq, gains, _ = R.ctx.get_greedy()
    R.greedy_len = len(seq)
    gpath = os.path.join(ROOT, "tests", "golden", "golden_c3.json")
    if name == "C3" and os.path.exists(gpath) and not args.no_parity:
        golden = json.load(open(gpath))["cases"][0]
        mine, seq, gains = parity_c3(R, golden)
        if dist is not None:
            allr = [None] * world
            dist.all_gather_object(allr, mine)
        else:
            allr = [mine]
        parity = {"golden": "tests/golden/golden_c3.json (CPU oracle, all 1024 views, 1.258 G rays)", "ranks_ok": [bool(r["ok"]) for r in allr],
                  "checks": {k: all(r[k] for r in allr) for k in mine if k != "ok"}, "ok": all(r["ok"] for r in allr)}

    # ---- a sustained region of >= 1 s when the K contract steps are shorter than that
    sustained = None
    step_s = total_ms * 1e-3 / args.steps
    if total_ms < 1000.0 and not args.no_sustained:
        n = int(min(20000, max(args.steps, math.ceil(1.05 / step_s))))
        keep_t, keep_s, keep_c, keep_l = R.timing, R.stats, R.counters, R.local_ms
        ms = R.timed(n, 0)
        sustained = {"steps": n, "seconds": ms * 1e-3, "ms_per_step": ms / n, "value": R.rays_job * n / (ms * 1e-3), "unit": UNIT}
        R.timing, R.stats, R.counters, R.local_ms = keep_t, keep_s, keep_c, keep_l

    # ---- end to end through the host-buffer C ABI
    e2e, e_seq = R.e2e(max(20, min(args.steps, 50)))
    assert e_seq.tolist() == seq.tolist(), "e2e and resident paths disagree"

    per_rank = None
    if dist is not None:
        mine = {"rank": rank, "step_ms": R.local_ms / args.steps, "cast_ms": timing["cast_ms"] / args.steps, "greedy_ms": timing["greedy_ms"] / args.steps,
                "allgather_ms": timing["gather_ms"] / args.steps, "marched": stats["marched"], "local_views": R.local_views}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    if rank != 0:
        R.ctx.close()
        return None

    roof, roof_g, scored = R.rooflines(args.steps, total_ms)
    cpu = None
    if want_cpu:
        orc = _oracle()
        threads = os.cpu_count() or 1
        V = w["n_views"]
        _, dt1, _ = cpu_sample(orc, w, [V // 2], threads)
        nv = int(max(2, min(V, round(15.0 / max(dt1, 1e-3)))))
        sample_ids = sorted(set(int(i) for i in np.linspace(0, V - 1, nv).astype(int)))
        r, dt, _ = cpu_sample(orc, w, sample_ids, threads)
        cpu = {"value": r / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d of %d views of %s (dense cast + rows + greedy over the sampled rows), %.1f s" % (len(sample_ids), V, name, dt)}
    greedy_ms = timing["greedy_ms"] / args.steps
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if (strong or (name == "C3" and world == 1)) else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": _config(name, w, args, world, strong, R.l2_note()),
            "e2e": e2e, "gpu_launches": R.counters["kernel_launches"], "clocks": clocks, "roofline": roof, "roofline_greedy": roof_g,
            "cpu_baseline": cpu, "parity": parity, "sustained": sustained,
            "views_scored_per_sec": scored * (1 if (strong or world == 1) else world) / max(1e-9, greedy_ms * 1e-3),
            "kernel_ms_per_step": {k: timing[k] / args.steps for k in ("cast_ms", "cull_ms", "march_ms", "count_ms", "greedy_ms", "gather_ms", "other_ms")},
            "per_rank": per_rank, "cast_stats": stats, "greedy_len": int(len(seq)), "greedy_seq": [int(x) for x in seq],
            "coverage_rate": float(gains.sum()) / max(1, R.ctx.full_voxels)}
    R.ctx.close()
    return line


def measure_c5_splat(prv, synth, args, local):
    """Config C5's render leg inside a timed step: the splat z-buffer of 100 views at 800x800 (points -> 64-bit atomicMin on
    packed depth|index -> resolve to RGBA + depth), poses resident.  Algorithmic bytes per view:
    16 P + W H (8 clear + 8 resolve-read + 4 rgba + 4 depth)."""
    w = synth.build_workload(prv, "C5")
    ctx = prv.Context(local)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    ctx.set_views(w["pose_world"], w["init_pos"])
    ctx.set_cloud(w["cloud"], w["cloud_rgb"])
    V, P, px = w["n_views"], len(w["cloud"]), w["W"] * w["H"]
    for _ in range(3):
        ctx.render_async(V, 5)
    ctx.sync()
    ctx.timing_reset()
    steps = max(20, args.steps)
    ctx.event_record(0)
    for _ in range(steps):
        ctx.render_async(V, 5)  # (each step writes 0.5 GB of corner cells + 0.5 GB of images: >> L2)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1) / steps
    t = ctx.get_timing()
    peak, _ = _peaks()
    alg = V * (16 * P + px * 24)
    ctx.close()
    return {"workload": "C5 render leg: %d views %dx%d, %d points, point_size 5" % (V, w["W"], w["H"], P), "steps": steps, "ms_per_step": ms,
            "views_per_sec": V / (ms * 1e-3), "pixels_per_sec": V * px / (ms * 1e-3),
            "kernel_ms_per_step": {"splat_points_ms": t["splat_ms"] / steps, "splat_resolve_ms": t["resolve_ms"] / steps},
            "roofline_splat": {"bound": "hbm", "kernels": "memset(corner)+splat_points_kernel+splat_resolve_kernel", "algorithmic_bytes_per_step": int(alg),
                               "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak}}


def measure_ensemble(prv, local):
    """SURVEY 8(f) #2, the reference's live "candidate views scored" loop (nbv_loop cases 2 / 3, main.cpp:2039-2161): per-pixel
    variance over the ensemble renders at W/16 x H/16, summed per view in the reference's order, arg-max.  Host-buffer API:
    the images travel H2D inside the timed call (as they would from instant-ngp's output), the scores come back."""
    rng = np.random.default_rng(5)
    V, E, W, H = 100, 2, 80, 45  # 100 candidate views, ensemble_num 2 (Share_Data.hpp:505-507), 1280x720 / 16 (main.cpp:1796-1806)
    images = rng.integers(0, 256, size=(V, E, H, W, 4), dtype=np.uint8)
    ctx = prv.Context(local)
    out = {"views": V, "ensemble_num": E, "width": W, "height": H}
    for method in (2, 3):
        ctx.score_ensemble(images, method)
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            best, scores = ctx.score_ensemble(images, method)
        dt = (time.perf_counter() - t0) / n
        out["method_%d" % method] = {"ms_per_call": dt * 1e3, "views_scored_per_sec": V / dt, "h2d_bytes_per_call": int(images.nbytes), "best_view": int(best)}
    ctx.close()
    return out


def run_own(args):
    rank, world, local = _dist_env()
    import load_pkg
    prv = load_pkg.load()
    synth = _synth()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure(prv, synth, args.workload, args, local, dist, rank, world, want_cpu=not args.no_cpu_baseline)
    if rank == 0 and world == 1 and not args.no_extras:
        # the other single-GPU configurations of BASELINE.json, in the same line
        if args.workload != "C2":
            sub = argparse.Namespace(**vars(args))
            sub.steps, sub.no_sustained = max(args.steps, 50), True
            c2 = measure(prv, synth, "C2", sub, local, None, 0, 1, want_cpu=False)
            line["c2"] = {k: c2[k] for k in ("value", "unit", "steps", "ms_per_step", "config", "e2e", "roofline", "roofline_greedy", "kernel_ms_per_step",
                                             "cast_stats", "views_scored_per_sec", "greedy_len", "coverage_rate")}
        line["c5_splat"] = measure_c5_splat(prv, synth, args, local)
        line["ensemble_scoring"] = measure_ensemble(prv, local)
        if not args.no_cpu_baseline:
            # the reference's own execution structure (one std::thread per voxel in batches of
            # num_of_thread = 20, pose inverse per voxel, sparse lookup; main.cpp:124-130, 238-284) on one view of C1
            try:
                orc = _oracle()
                w1 = synth.build_workload(prv, "C1", n_views=1)
                m1 = orc.Map.from_keys(w1["keys"], w1["map_rgb"], w1["resolution"])
                t0 = time.perf_counter()
                m1.precept_threads(_oracle_intr(orc, w1["intr"]), w1["pose_world"][0], w1["init_pos"][0], 1.0, 20)
                dt = time.perf_counter() - t0
                line["cpu_reference_structure"] = {"value": m1.n / dt, "unit": "voxel rays/s", "structure": "std::thread per voxel, batches of 20, joined per batch",
                                                   "sample": "Perception_3D::precept of 1 view of C1 (%d voxels), %.2f s" % (m1.n, dt)}
            except Exception as exc:  # never let the side baseline break the bench line
                line["cpu_reference_structure"] = {"error": str(exc)}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and line.get("parity") and not line["parity"]["ok"]:
        sys.stderr.write("bench.py: PARITY FAILED against tests/golden/golden_c3.json: %s\n" % json.dumps(line["parity"]))
        return 3
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--variant", type=int, default=2)
    ap.add_argument("--brick", type=int, default=0, choices=[0, 4, 8, 16], help="prv_set_brick_cull: brick edge of the conservative cull in voxels")
    ap.add_argument("--brick-entry", type=int, default=1, choices=[0, 1], help="prv_set_brick_cull: start the exact march at the first set brick")
    ap.add_argument("--stage-smem", type=int, default=1, choices=[0, 1], help="A/B: padded bitmap staged in the march blocks' shared memory")
    ap.add_argument("--stage-l2", type=int, default=0, choices=[0, 1], help="A/B: L2 persisting window over the padded bitmap")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="N > 1: transport of the coverage-row all-gather")
    ap.add_argument("--ref-seconds", type=float, default=6.0, help="reference arm: CPU seconds per step (sets the views sampled per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the C2 / C5 sub-measurements")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
