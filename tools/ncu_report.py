"""Summarises ncu reports into profiles/: per-kernel metric table (--set full), DRAM traffic per launch, launch list shares.
    python tools/ncu_report.py full gpurun_out/fin/prof_C3.ncu-rep profiles/r2_ncu_full_C3.md [traffic-json-key]
    python tools/ncu_report.py launches gpurun_out/fin/launches_C3.csv profiles/r2_launch_list_C3.md
"""
import collections
import csv
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def short(name):
    return name.split("(")[0].replace("void ", "")


def full(rep, out, key=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(k) for k in KEYS if k in hdr}
    ik = hdr.index("Kernel Name")
    kernels = rows[2:]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none (%s)\n\nOne captured launch per column.\n\n" % os.path.basename(rep))
        f.write("| metric | " + " | ".join(short(k[ik]) for k in kernels) + " |\n|---|" + "---|" * len(kernels) + "\n")
        for k in KEYS:
            if k in idx:
                f.write("| %s (%s) | " % (k, units[idx[k]]) + " | ".join(r[idx[k]] for r in kernels) + " |\n")
    traffic = {}
    for r in kernels:
        def val(k):
            v = float(r[idx[k]].replace(",", ""))
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[idx[k]].lower(), 1)
        traffic[short(r[ik])] = int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    print(json.dumps(traffic))
    if key:
        tf = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
        d = json.load(open(tf)) if os.path.exists(tf) else {}
        march = [v for k, v in traffic.items() if k.startswith("march_kernel")]
        if march:
            d[key] = march[0]
        d.setdefault("_per_kernel", {})[key] = traffic
        json.dump(d, open(tf, "w"), indent=1)


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(short(r[ik]), []).append(float(r[iv].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(out, "w") as f:
        f.write("# ncu launch list (%s)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and serialised: compare "
                "SHARES.\n\n| kernel | launches | mean us | share of GPU time |\n|---|---|---|---|\n" % os.path.basename(path))
        for k, v in d.items():
            f.write("| %s | %d | %.1f | %.1f %% |\n" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        launches(sys.argv[2], sys.argv[3])
