#!/usr/bin/env python
"""Compare the oracle with vectors dumped from the REAL reference (dump_reference_vectors.cc):
   python tests/golden/dealii_dump/compare_with_reference_dump.py <prefix> <degree> <refinements>
Prints the relative l2 differences (apply 1e-12, transfers 1e-13, V-cycle / CG history 1e-10 are the bars of
BASELINE.json's north_star) and exits 1 if one is exceeded.  Needs only numpy + the built oracle (make -C oracle)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyoracle as O
from helpers import rel_l2, splitmix_src


def main():
    prefix, p, ref = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    n = 2 ** ref
    levels = [(p, n >> k) for k in range(ref, -1, -1)]
    mfs = [O.MatrixFree(3, d, c) for d, c in levels]
    trs = [O.Transfer(mfs[l - 1], mfs[l], "h") for l in range(1, len(levels))]
    vc = O.VCycle(mfs, trs)
    load = lambda name: np.fromfile("%s_%s.f64" % (prefix, name))
    src = splitmix_src(mfs[-1].n_dofs)
    ok = True

    def check(name, got, ref_, tol):
        nonlocal ok
        e = rel_l2(got, ref_)
        print("%-28s %.3e %s" % (name, e, "ok" if e <= tol else "FAIL"))
        ok = ok and e <= tol

    check("synthetic source", src, load("src"), 0.0)
    check("LaplaceOperator::vmult", mfs[-1].vmult(src), load("vmult"), 1e-12)
    srcc = splitmix_src(mfs[-2].n_dofs, salt=5)
    check("prolongate_and_add", trs[-1].prolongate_and_add(np.zeros(mfs[-1].n_dofs), srcc), load("prolongated"), 1e-13)
    check("restrict_and_add", trs[-1].restrict_and_add(np.zeros(mfs[-2].n_dofs), src), load("restricted"), 1e-13)
    res = src.copy()
    res[mfs[-1].constrained()] = 0.0
    check("VCycleMultigrid::vmult", vc.vmult(res), load("vcycle"), 1e-10)
    lines = open(prefix + "_cg.txt").read().split()
    it_ref, hist_ref = int(lines[0]), np.array([float(x) for x in lines[1:]])
    x, it, hist, rc = O.cg_solve(mfs[-1], mfs[-1].assemble_rhs(), vc)
    print("CG iterations: oracle %d, reference %d" % (it, it_ref))
    ok = ok and it == it_ref
    m = min(len(hist), len(hist_ref))
    e = float(np.max(np.abs(hist[:m] - hist_ref[:m])) / hist_ref[0])
    print("%-28s %.3e %s" % ("CG residual history", e, "ok" if e <= 1e-10 else "FAIL"))
    sys.exit(0 if ok and e <= 1e-10 else 1)


if __name__ == "__main__":
    main()
