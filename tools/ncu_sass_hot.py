"""Hottest SASS instructions of a kernel in an ncu report, with their stall reasons.
usage: python tools/ncu_sass_hot.py report.ncu-rep kernel-substring [top N]"""
import csv, io, subprocess, sys
rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kernel = None; hdr = None; data = {}
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name": kernel = r[1]; continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) >= len(hdr) - 2 and kernel and ksub in kernel:
        data.setdefault(kernel, []).append(r)
for k, L in data.items():
    ci = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = sum(int(r[ci]) for r in L); tot_i = sum(int(r[ii]) for r in L)
    print("==", k, "instr", tot_i, "samples", tot_s)
    agg = {}
    for r in L:
        for i in stalls:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print("stall mix:", ", ".join("%s %.1f%%" % (h[6:], 100 * v / max(tot_s, 1)) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:10]))
    for n, r in sorted(enumerate(L), key=lambda x: -int(x[1][ci]))[:top]:
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stalls), reverse=True)[:3]
        print("%5d %5.2f%% smp %5.2f%% inst thr %4.1f  %-60s %s" % (n, 100 * int(r[ci]) / max(tot_s, 1), 100 * int(r[ii]) / max(tot_i, 1), int(r[ti]) / max(int(r[ii]), 1), r[1].strip()[:60], " ".join("%s:%d" % (b, a) for a, b in st if a)))
