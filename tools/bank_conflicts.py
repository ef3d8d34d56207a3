#!/usr/bin/env python
"""Offline shared-memory bank-conflict model of the apply tile program (64-bit accesses: a warp is served in
two 16-lane halves; a half is conflict-free iff its 16 addresses fall into 16 different 8-byte banks).
Prints the wavefronts per warp instruction for candidate paddings of T1 (STRIDE) and O (OX)."""
import itertools
import sys


def wavefronts(addrs):
    """addrs: list of 32 double-indices (or None for inactive lanes)"""
    w = 0
    for h in range(2):
        lanes = [a for a in addrs[16 * h:16 * h + 16] if a is not None]
        if not lanes:
            continue
        banks = {}
        for a in lanes:
            banks.setdefault(a % 16, set()).add(a)
        w += max(len(v) for v in banks.values())
    return w


def model(P, BX, BY, STRIDE, OXPAD):
    N1 = P + 1
    CXC, CYC = BX + 1, BY + 1
    NITEM = CXC * CYC * N1
    NT = (NITEM + 31) // 32 * 32
    OX = CXC * N1 + OXPAD
    def item(t):
        return t % CXC, (t // CXC) % N1, t // (CXC * N1)
    def ioff(tcx, j, tcy):
        return (tcx + CXC * (j + N1 * tcy)) * STRIDE
    tot = {"F_store": [0, 0], "Y_load": [0, 0], "B_load": [0, 0], "O_store": [0, 0], "E_load": [0, 0]}
    for w0 in range(0, NT, 32):
        lanes = [t if t < NITEM else None for t in range(w0, w0 + 32)]
        for m, a in itertools.product(range(N1), range(N1)):
            ad = [None if t is None else ioff(*item(t)) + m * N1 + a for t in lanes]
            tot["F_store"][0] += wavefronts(ad); tot["F_store"][1] += 1
            tot["B_load"][0] += wavefronts(ad); tot["B_load"][1] += 1
        for m, j in itertools.product(range(N1), range(N1)):
            ad = []
            for t in lanes:
                if t is None: ad.append(None); continue
                tcx, a, tcy = item(t)
                ad.append(ioff(tcx, j, tcy) + m * N1 + a)
            tot["Y_load"][0] += wavefronts(ad); tot["Y_load"][1] += 1
        for k, i in itertools.product(range(1), range(N1)):
            ad = []
            for t in lanes:
                if t is None: ad.append(None); continue
                tcx, j, tcy = item(t)
                ad.append((tcy * N1 + j) * OX + tcx * N1 + i)
            tot["O_store"][0] += wavefronts(ad); tot["O_store"][1] += 1
    EW, EH = BX * P + 1, BY * P + 1
    for w0 in range(0, EW * EH, 32):
        ad = []
        for col in range(w0, w0 + 32):
            if col >= EW * EH: ad.append(None); continue
            ix, iy = col % EW, col // EW
            tx1, il, ty1, jl = ix // P + 1, ix % P, iy // P + 1, iy % P
            ad.append((ty1 * N1 + jl) * OX + tx1 * N1 + il if tx1 < CXC and ty1 < CYC else None)
        tot["E_load"][0] += wavefronts(ad); tot["E_load"][1] += 1
    return {k: v[0] / max(v[1], 1) for k, v in tot.items()}


if __name__ == "__main__":
    P, BX, BY = (int(x) for x in sys.argv[1:4])
    N1 = P + 1
    base = N1 * N1
    for stride in range(base, base + 10):
        for pad in range(0, 6):
            r = model(P, BX, BY, stride, pad)
            score = r["F_store"] + r["Y_load"] * 2 + r["B_load"] + r["O_store"] * P / N1 + r["E_load"] * 0.3
            print("STRIDE %3d OXPAD %d  " % (stride, pad) + "  ".join("%s %.2f" % kv for kv in r.items()) + "  score %.2f" % score)
