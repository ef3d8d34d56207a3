#!/usr/bin/env python
"""Time the fused Chebyshev step on small Q1/Q2 meshes (coarse multigrid levels): tools/small_levels.py
Run once per kernel choice: PMG_TILE_VARIANT=0 (line-marching kernel) or 2 (cell-tile kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "portable-multigrid_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pmg_b200 as G
from helpers import splitmix_src

ctx = G.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream())
for p, n in [(1, 1), (1, 2), (1, 4), (1, 8), (1, 16), (1, 32), (1, 64), (2, 16), (2, 32), (2, 64)]:
    op = G.LaplaceOperator(ctx, p, n)
    N = op.m()
    u, b = op.vector_from(splitmix_src(N, salt=1)), op.vector_from(splitmix_src(N, salt=2))
    xo = op.initialize_dof_vector()
    for _ in range(20):
        op.chebyshev_step(xo, u, xo, b, 0.3, 0.1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync()
    k = 500
    e0.record(stream)
    for _ in range(k):
        op.chebyshev_step(xo, u, xo, b, 0.3, 0.1)
    e1.record(stream)
    ctx.sync()
    print("variant %s p=%d n=%d N=%d: %.2f us per fused step" % (os.environ.get("PMG_TILE_VARIANT", "0"), p, n, N, e0.elapsed_time(e1) / k * 1e3), flush=True)
