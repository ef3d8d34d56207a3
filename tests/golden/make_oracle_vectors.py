#!/usr/bin/env python
"""Regenerates tests/golden/oracle_vectors.npz from the oracle (run from the repo root).

The reference ships no golden vectors and cannot be run here, so these vectors do NOT pin the oracle to the
reference; they pin the oracle (and, through the parity tests, the CUDA path) against regressions.  The
independent pins are the analytic anchors in anchors.json."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyoracle as O
from helpers import hierarchy_levels, splitmix_src

out = {}
for p, n in [(1, (3, 2, 2)), (2, (3, 2, 2)), (4, (2, 2, 1))]:
    mf = O.MatrixFree(3, p, n)
    src = splitmix_src(mf.n_dofs)
    out["vmult_p%d" % p] = mf.vmult(src)
    out["dinv_p%d" % p] = mf.compute_diagonal()
for kind, p, n in [("h", 2, 8), ("hp", 4, 4)]:
    levels = hierarchy_levels(kind, p, n)
    mfs = [O.MatrixFree(3, q, c) for (q, c) in levels]
    trs = [O.Transfer(mfs[l - 1], mfs[l], "h" if levels[l][0] == levels[l - 1][0] else "p") for l in range(1, len(levels))]
    vc = O.VCycle(mfs, trs)
    b = mfs[-1].assemble_rhs()
    x, it, hist, rc = O.cg_solve(mfs[-1], b, vc)
    est = np.array([[e[0], e[1], e[2], e[3]] for e in vc.estimate()])
    out["cg_%s_p%d_n%d_hist" % (kind, p, n)] = hist
    out["cg_%s_p%d_n%d_est" % (kind, p, n)] = est
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz"), **out)
print({k: v.shape for k, v in out.items()})
