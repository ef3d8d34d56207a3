#!/usr/bin/env python
"""Summarise the source page of an ncu report: tools/ncu_hot.py SRC.csv [N]
(SRC.csv from `ncu -i X.ncu-rep --page source --csv`).  Prints the N hottest SASS instructions by stall samples,
the stall-reason totals, and the executed-instruction mix."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(r for r in rows if r and r[0] == "Address")
body = rows[rows.index(hdr) + 1:]
col = {h: i for i, h in enumerate(hdr)}
S, X = col["# Samples"], col["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter(); mix = Counter(); samples = 0
recs = []
for k, r in enumerate(body):
    if len(r) < len(hdr) or not r[S].strip().isdigit(): continue
    s = int(r[S] or 0); samples += s
    op = r[col["Source"]].split()[0] if r[col["Source"]] else "?"
    if op.startswith("@"): op = r[col["Source"]].split()[1]
    mix[op.split(".")[0]] += int(r[X] or 0)
    for h in stalls: tot[h] += int(r[col[h]] or 0)
    recs.append((s, k, r))
print("samples", samples)
print("stalls:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(samples, 1)) for h, v in tot.most_common(8)))
ti = sum(mix.values())
print("warp instructions %d: " % ti + ", ".join("%s %.1f%%" % (o, 100.0 * v / ti) for o, v in mix.most_common(14)))
for s, k, r in sorted(recs, reverse=True)[:n]:
    top = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print("%6d  #%5d  %-70s %s" % (s, k, r[col["Source"]][:70], " ".join("%s=%d" % (h, v) for v, h in top)))
