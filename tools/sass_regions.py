#!/usr/bin/env python3
"""Static code-footprint report of a kernel: instructions and bytes between block barriers, extent of the main loop.

    python tools/sass_regions.py <object-or-executable> <kernel-name-regex>

Reads `cuobjdump -sass`, finds the first kernel whose mangled name matches, and prints
  * the total instruction count, the opcode histogram (top 12) and the DFMA share,
  * the regions separated by BAR.SYNC (for the line-marching apply kernel: staging + y sweep | x sweep | z sweep),
  * the largest backward branch (= the steady-state loop) and its size in bytes.
Development tool (CPU only, no GPU needed): the numbers it prints are static, not executed counts.
"""
import re
import subprocess
import sys
from collections import Counter

INSTR = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?\s*(.*?);")


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], check=True, capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name:
            body.append(line)
    if name:
        yield name, body


def main():
    path, pat = sys.argv[1], re.compile(sys.argv[2])
    for name, body in kernels(path):
        if not pat.search(name):
            continue
        ins = []
        for line in body:
            m = INSTR.match(line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2), m.group(4)))
        print(name)
        hist = Counter(op for _, op, _ in ins)
        print("  instructions %d (%.1f KB), DFMA %d (%.0f %%)" % (len(ins), len(ins) * 16 / 1024.0, hist["DFMA"],
                                                                  100.0 * hist["DFMA"] / max(1, len(ins))))
        print("  top opcodes: " + ", ".join("%s %d" % kv for kv in hist.most_common(12)))
        bars = [a for a, op, _ in ins if op == "BAR"]
        prev = 0
        for i, a in enumerate(bars + [ins[-1][0]]):
            seg = [x for x in ins if prev <= x[0] < a]
            print("  region %d: 0x%05x..0x%05x  %5d instructions, %5d DFMA" % (i, prev, a, len(seg), sum(1 for x in seg if x[1] == "DFMA")))
            prev = a
        best = None
        for a, op, args in ins:
            if op == "BRA":
                m = re.search(r"0x([0-9a-f]+)", args)
                if m and int(m.group(1), 16) < a and (best is None or a - int(m.group(1), 16) > best[1] - best[0]):
                    best = (int(m.group(1), 16), a)
        if best:
            print("  largest loop: 0x%05x..0x%05x = %.1f KB (%d instructions)" % (best[0], best[1], (best[1] - best[0]) / 1024.0,
                                                                                   (best[1] - best[0]) // 16))
        return
    sys.exit("no kernel matches")


if __name__ == "__main__":
    main()
