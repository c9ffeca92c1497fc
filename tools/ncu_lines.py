"""Per-source-line instruction / sample shares from an ncu report (needs --import-source on and -lineinfo).
usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
ksub = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; kernel = None; data = {}
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name": kernel = r[1]; continue
    if len(r) >= 2 and r[0] == "Function Name": kernel = r[1]; continue
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) < 9 or r[0] in ("Line No", ""): continue
    try:
        ln = int(r[0]); inst = int(r[7]); samp = int(r[6]) if r[6] != "-" else 0; thr = int(r[8]) if r[8] not in ("-", "") else 0
    except ValueError:
        continue
    data.setdefault(kernel, []).append((inst, samp, thr, cur_file, ln, r[1].strip()[:100]))
for k, lines in data.items():
    if ksub not in (k or ""): continue
    ti = sum(x[0] for x in lines); ts = sum(x[1] for x in lines)
    print("==", k, "warp-instr", ti, "samples", ts)
    for inst, samp, thr, f, ln, src in sorted(lines, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% smp thr/inst %4.1f  %s:%d  %s" % (100 * inst / max(ti, 1), 100 * samp / max(ts, 1), thr / max(inst, 1), f, ln, src))
