#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: instruction mix, stall reasons, hottest SASS lines."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
data = [r for r in rows[1:] if r[hdr.index("Instructions Executed")].isdigit()]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, samp, st = collections.Counter(), collections.Counter(), collections.Counter()
n_inst = n_samp = 0
for r in data:
    toks = r[ia].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    e, s = int(r[ie]), int(r[isamp])
    tot[op] += e; samp[op] += s; n_inst += e; n_samp += s
    for h in stalls:
        st[h] += int(r[hdr.index(h)])
print("total warp instrs", n_inst, "samples", n_samp, "static instrs", len(data))
for op, c in tot.most_common(22):
    print("%-10s %11d %5.1f%%   samples %5.1f%%" % (op, c, 100 * c / n_inst, 100 * samp[op] / max(n_samp, 1)))
print({k: v for k, v in st.most_common(10)})
if len(sys.argv) > 2:
    top = sorted(data, key=lambda r: -int(r[isamp]))[: int(sys.argv[2])]
    for r in top:
        print(r[isamp], r[ie], r[ia][:90])
