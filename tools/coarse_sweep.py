#!/usr/bin/env python
"""V-cycle time against the size bound of the single-CTA coarse kernel (PMG_COARSE_MAX_WORK; 0 = per-level kernels only):
tools/coarse_sweep.py [c1|c2]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "portable-multigrid_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pmg_b200 as G
from helpers import hierarchy_levels, splitmix_src

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
levels = hierarchy_levels("hp", 4, 64) if cfg == "c2" else hierarchy_levels("h", 2, 64)
cheb = 5 if cfg == "c2" else 3
ctx = G.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream())
for work in ("0", "1e3", "1e4", "5e4", "1e5", "4e5", "1.2e6"):
    os.environ["PMG_COARSE_MAX_WORK"] = work
    ops, transfers, smoothers, mg = G.build_hierarchy(ctx, levels, degree=cheb)
    top = ops[-1]
    r, z = top.vector_from(splitmix_src(top.m(), salt=1)), top.initialize_dof_vector()
    counts = []
    for _ in range(4):
        b = ctx.launch_count(); mg.vmult(z, r); counts.append(ctx.launch_count() - b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync(); e0.record(stream)
    for _ in range(20):
        mg.vmult(z, r)
    e1.record(stream); ctx.sync()
    print("%s max_work=%s launches/cycle=%d V-cycle %.3f ms" % (cfg, work, counts[0], e0.elapsed_time(e1) / 20), flush=True)
    del ops, transfers, smoothers, mg, r, z
