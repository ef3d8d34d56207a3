#!/usr/bin/env python
"""Launch the hot kernels a few times (for ncu / timing sweeps): tools/run_kernels.py DEGREE CELLS [REPS] [MODE]
MODE: apply | cheb | both | sweep (sweep = CUDA-event timing of apply and fused Chebyshev step for degrees 1..8).
Environment: COEF=1 times the variable-coefficient operator, SWEEP_DOFS the size of the sweep (default 100e6)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "portable-multigrid_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import pmg_b200 as G
from helpers import splitmix_src


def time_ms(ctx, stream, fn, k=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync()
    e0.record(stream)
    for _ in range(k):
        fn()
    e1.record(stream)
    ctx.sync()
    return e0.elapsed_time(e1) / k


def main():
    p = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    mode = sys.argv[4] if len(sys.argv) > 4 else "both"
    ctx = G.Context(0)
    coef = int(os.environ.get("COEF", "0"))
    stream = torch.cuda.ExternalStream(ctx.stream())
    if mode == "sweep":
        target = float(os.environ.get("SWEEP_DOFS", "100e6"))
        for p in range(1, 9):
            n = max(2, round((target ** (1.0 / 3.0) - 1) / p))
            op = G.LaplaceOperator(ctx, p, n, coefficient=coef)
            N = op.m()
            u, b = op.vector_from(splitmix_src(N, salt=1)), op.vector_from(splitmix_src(N, salt=2))
            z, xo = op.initialize_dof_vector(), op.initialize_dof_vector()
            ta = time_ms(ctx, stream, lambda: op.vmult(z, u), 10)
            tc = time_ms(ctx, stream, lambda: op.chebyshev_step(xo, u, xo, b, 0.3, 0.1), 10)
            print("p=%d n=%d N=%d apply %.3f ms %.1f GDoF/s (%.3f of 6538.9 GB/s @16B) | cheb-step %.3f ms %.1f GDoF/s (%.3f @32B)"
                  % (p, n, N, ta, N / ta / 1e6, 16 * N / ta / 1e6 / 6538.9, tc, N / tc / 1e6, 32 * N / tc / 1e6 / 6538.9), flush=True)
            del u, b, z, xo, op
        return
    op = G.LaplaceOperator(ctx, p, n, coefficient=coef)
    N = op.m()
    u, b = op.vector_from(splitmix_src(N, salt=1)), op.vector_from(splitmix_src(N, salt=2))
    z, xo = op.initialize_dof_vector(), op.initialize_dof_vector()
    for _ in range(reps):
        if mode in ("apply", "both"):
            op.vmult(z, u)
        if mode in ("cheb", "both"):
            op.chebyshev_step(xo, u, xo, b, 0.3, 0.1)
    ctx.sync()
    print("done", p, n, N, float(np.abs(z.export_host()).sum()))


if __name__ == "__main__":
    main()
