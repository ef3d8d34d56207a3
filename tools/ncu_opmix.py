#!/usr/bin/env python
"""Executed warp instructions per SASS opcode of one kernel of an ncu report captured with --import-source on:
tools/ncu_opmix.py X.ncu-rep [kernel-regex] [dofs]   (dofs: print warp instructions per DoF)"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else "."; dofs = float(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first kernel only
hdr = None; mix = collections.Counter(); samples = collections.Counter(); n = 0; on = False
for r in rows:
    if r and r[0] == "Kernel Name":
        on = False
        if re.search(kre, r[1]):
            n += 1
            if n > 1: break
            on = True; print(r[1])
        continue
    if not on: continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", d["Source"])
    if not m: continue
    op = m.group(1)
    mix[op] += int(d["Instructions Executed"]); samples[op] += int(d["# Samples"])
tot = sum(mix.values()); ts = sum(samples.values()) or 1
print("total executed warp instructions %d%s" % (tot, (" = %.2f per DoF" % (tot / dofs)) if dofs else ""))
for op, c in mix.most_common(28):
    print("%-10s %12d %5.1f%%  %s samples %4.1f%%" % (op, c, 100.0 * c / tot, ("%.3f/DoF" % (c / dofs)) if dofs else "", 100.0 * samples[op] / ts))
