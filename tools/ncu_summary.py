"""Summarises an ncu report (--set full) and a launch list (--metrics gpu__time_duration.sum) into profiles/*.md|json.

    python tools/ncu_summary.py gpurun_out/r1_prof_cast.ncu-rep gpurun_out/r1_launches.csv profiles/r1
"""
import collections
import csv
import json
import subprocess
import sys

rep, launches, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
idx = {k: hdr.index(k) for k in KEYS if k in hdr}
kernels = []
for r in rows[2:]:
    kernels.append({k: r[i] for k, i in idx.items()})
    kernels[-1]["_units"] = {k: units[i] for k, i in idx.items()}
with open(out + "_ncu_full_summary.md", "w") as f:
    f.write("# ncu --set full summary (%s)\n\nOne captured launch per kernel, `--clock-control none`, command: `python bench.py --steps 1 --warmup 3` (C2).\n\n" % rep)
    f.write("| metric | " + " | ".join(k["Kernel Name"].split("(")[0].replace("void ", "") for k in kernels) + " |\n|---|" + "---|" * len(kernels) + "\n")
    for key in KEYS[1:]:
        if key in idx:
            f.write("| %s (%s) | " % (key, units[idx[key]]) + " | ".join(k[key] for k in kernels) + " |\n")
traffic = {}
for k in kernels:
    def val(key):
        v = float(k[key].replace(",", ""))
        u = k["_units"][key].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    traffic[k["Kernel Name"].split("(")[0].replace("void ", "")] = int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
json.dump(traffic, open(out + "_dram_traffic_per_launch.json", "w"), indent=1)
# launch list
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
h = rows[0]
ik, iv = h.index("Kernel Name"), h.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ik].split("(")[0].replace("void ", ""), []).append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in d.values())
with open(out + "_launch_list_summary.md", "w") as f:
    f.write("# ncu launch list summary (%s)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 --warmup 3` (C2); "
            "per-launch times are cold-cache and serialised: compare SHARES.\n\n| kernel | launches | mean ns | share of GPU time |\n|---|---|---|---|\n" % launches)
    for k, v in d.items():
        f.write("| %s | %d | %.0f | %.1f %% |\n" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
print(open(out + "_launch_list_summary.md").read())
print(json.dumps(traffic))
