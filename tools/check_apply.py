#!/usr/bin/env python3
"""Quick GPU parity check of the line-marching apply kernel against the oracle without pytest / torch (seconds):
    python tools/check_apply.py [degree ...]        (default: 4)
Forces the line-marching kernel (PMG_TILE_VARIANT=1 unless set: 4 = its pipelined variant), compares vmult with the oracle on two meshes and two Dirichlet
masks per degree, tolerance 1e-12 relative l2 (north_star)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("portable-multigrid_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
os.environ.setdefault("PMG_TILE_VARIANT", "1")  # 4 = the pipelined variant
import numpy as np
import pmg_b200 as G
import pyoracle as O
from helpers import rel_l2, splitmix_src

ctx = G.Context(0)
bad = 0
for p in [int(a) for a in sys.argv[1:]] or [4]:
    for n, faces in (((9, 8, 5), 0x3F), ((5, 9, 11), 0x15), ((17, 16, 9), 0x3F)):
        mf = O.MatrixFree(3, p, n, faces=faces)
        src = splitmix_src(mf.n_dofs, salt=p)
        op = G.LaplaceOperator(ctx, p, n, dirichlet_faces=faces)
        s, d = op.vector_from(src), op.initialize_dof_vector()
        op.vmult(d, s)
        err = rel_l2(d.export_host(), mf.vmult(src))
        print("Q%d %s faces=%#x: rel l2 error %.2e %s" % (p, n, faces, err, "ok" if err <= 1e-12 else "FAIL"), flush=True)
        bad += err > 1e-12
sys.exit(1 if bad else 0)
