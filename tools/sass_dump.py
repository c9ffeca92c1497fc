"""SASS evidence under profiles/r2_sass/: cuobjdump -sass of the hot kernels of nerf-prv_b200/libprv_b200.so and instruction histograms.
usage: python tools/sass_dump.py"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_sass")
os.makedirs(OUT, exist_ok=True)
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "nerf-prv_b200", "libprv_b200.so")], capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur is not None:
        funcs[cur].append(line)
KEEP = {"march_kernelILi256ELi4ELb0": "march_kernelILi256ELi4ELb0.sass", "march_kernelILi256ELi4ELb1": "march_kernelILi256ELi4ELb1.sass", "cull_kernelILb0": "cull_kernelILb0.sass",
        "coarse_kernelILi8ELb0": "coarse_kernelILi8ELb0.sass", "greedy_cluster_kernel": "greedy_cluster_kernel.sass"}
HOT = ["march_kernel", "coarse_kernel", "cull_kernel", "count_nonwhite", "splat_resolve", "splat_points", "greedy_cluster", "popcount_rows", "map_check", "count_publish", "deproj_table"]
hist = ["# SASS instruction histograms of the hot kernels (cuobjdump -sass nerf-prv_b200/libprv_b200.so, sm_100a only, final round-2 kernels)", "",
        "No contraction on this path, so no `UTC*MMA` / `LDTM` is expected; what to look for instead: `DADD` / `DFMA` / `DSETP` (the exact FP64 DDA), `LDG.E.CONSTANT` / `LDS` probes, `REDUX` (warp arg-max / counters), `ATOMG` / `RED` (coverage rows, z-buffer, tickets), and in the cluster greedy the distributed-shared-memory traffic (`ST.ASYNC`-class `STAS`, `SYNCS` mbarrier ops, `UCGABAR`).", ""]
for name, lines in funcs.items():
    for key, fn in KEEP.items():
        if key in name:
            open(os.path.join(OUT, fn), "w").write("Function : %s\n" % name + "\n".join(lines) + "\n")
    if any(h in name for h in HOT):
        ops = collections.Counter()
        n = 0
        for l in lines:
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
            if m:
                ops[m.group(1)] += 1
                n += 1
        hist += ["## `%s` -- %d instructions" % (name, n), "", ", ".join("%s %d" % kv for kv in ops.most_common(28)), ""]
open(os.path.join(OUT, "instruction_histograms.md"), "w").write("\n".join(hist))
print("wrote", OUT)
