#!/usr/bin/env python
"""Shares of a V-cycle's kernels from an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv python bench.py ...
   tools/launch_summary.py launches.csv LAUNCHES_PER_CYCLE [SKIP_FROM_END]
Takes the last LAUNCHES_PER_CYCLE launches before the last SKIP_FROM_END ones (default 0) -- one V-cycle of the bench --
and prints total / average time and share per (kernel, grid).  ncu times are cold-cache and serialised: compare SHARES."""
import collections
import csv
import re
import sys


def main():
    path, per_cycle = sys.argv[1], int(sys.argv[2])
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "us")
        us = val / 1000.0 if unit in ("ns", "nsecond") else val * 1000.0 if unit in ("ms", "msecond") else val
        name = re.sub(r"\(.*$", "", d["Kernel Name"]).replace("(int)", "").replace("void ", "").replace("<unnamed>::", "")
        rows.append((name.strip(), d.get("Grid Size", ""), us))
    end = len(rows) - skip
    sel = rows[max(0, end - per_cycle):end]
    tot = sum(r[2] for r in sel) or 1.0
    agg = collections.OrderedDict()
    for name, grid, us in sel:
        k = (name, grid)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    print("%d launches in the list; the %d before the last %d: sum %.1f us" % (len(rows), len(sel), skip, tot))
    for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s grid %-18s n=%3d  total %8.1f us  avg %7.1f us  %5.1f%%" % (name[:62], grid, n, us, us / n, 100.0 * us / tot))


if __name__ == "__main__":
    main()
