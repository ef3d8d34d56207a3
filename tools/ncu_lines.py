#!/usr/bin/env python
"""Per-source-line totals from an ncu report (needs -lineinfo): tools/ncu_lines.py X.ncu-rep [KERNEL_INDEX] [N]
Prints, for the chosen launch, the N source lines with most executed warp instructions, with their stall samples."""
import csv, io, subprocess, sys
rep = sys.argv[1]
kid = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# sections start with a "Function Name" row; pick kernel section #kid among distinct launches
starts = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
# group by launches: a launch = consecutive files until the next section whose file path repeats the first file
files = [(rows[i - 2][1] if i >= 2 else "?") for i in starts]
first = files[0]
launch_starts = [j for j, f in enumerate(files) if f == first]
lo = launch_starts[kid]
hi = launch_starts[kid + 1] if kid + 1 < len(launch_starts) else len(starts)
recs = []
tot_i = tot_s = 0
for j in range(lo, hi):
    i = starts[j]
    hdr = rows[i]
    col = {h: c for c, h in reversed(list(enumerate(hdr)))}
    end = starts[j + 1] - 2 if j + 1 < len(starts) else len(rows)
    for r in rows[i + 1:end]:
        if len(r) < 10 or not r[0].strip().isdigit(): continue
        num = lambda v: int(v) if v.strip().isdigit() else 0
        ins = num(r[col["Instructions Executed"]]); smp = num(r[col["# Samples"]])
        tot_i += ins; tot_s += smp
        recs.append((ins, smp, files[j].split("/")[-1], r[0], r[1].strip()[:110]))
print("launch %d: warp instructions %d, samples %d" % (kid, tot_i, tot_s))
for ins, smp, f, ln, src in sorted(recs, reverse=True)[:n]:
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100.0 * ins / max(tot_i, 1), 100.0 * smp / max(tot_s, 1), f, ln, src))

# optional: totals per line range, e.g.  RANGES="stage:218-245,phase1:246-302" tools/ncu_lines.py ...
import os
if os.environ.get("RANGES"):
    print("--- per range (pmg_apply_sweep.h lines)")
    for spec in os.environ["RANGES"].split(","):
        name, rng = spec.split(":"); a, b = map(int, rng.split("-"))
        ti = sum(r[0] for r in recs if r[2] == "pmg_apply_sweep.h" and a <= int(r[3]) <= b)
        ts = sum(r[1] for r in recs if r[2] == "pmg_apply_sweep.h" and a <= int(r[3]) <= b)
        print("%-10s %5.1f%% inst %5.1f%% samples" % (name, 100.0 * ti / tot_i, 100.0 * ts / tot_s))
    ti = sum(r[0] for r in recs if r[2] != "pmg_apply_sweep.h"); ts = sum(r[1] for r in recs if r[2] != "pmg_apply_sweep.h")
    print("%-10s %5.1f%% inst %5.1f%% samples" % ("other files", 100.0 * ti / tot_i, 100.0 * ts / tot_s))
