#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_check.py
Every rank builds the slab-distributed hierarchy, rank 0 additionally runs the CPU oracle; operator apply,
transfers, V-cycle and CG must agree with the oracle to the single-GPU tolerances and the iteration counts
must be identical (partition independence)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "portable-multigrid_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import pmg_b200 as G
from helpers import rel_l2, splitmix_src


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    obj = [G.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    ctx = G.Context(local, rank, world, obj[0])
    ctx.set_coarse_threshold(int(os.environ.get("PMG_COARSE_THRESHOLD", "2000")))
    ok = True

    def report(name, err, tol):
        nonlocal ok
        if rank == 0 or not (err <= tol):
            print("[rank %d] %-44s %.3e %s" % (rank, name, err, "ok" if err <= tol else "FAIL"), flush=True)
        ok = ok and (err <= tol)

    import pyoracle as O

    # 1. operator apply + vector ops on a distributed level, degrees 1..4
    for p, n in [(1, (9, 7, 8 * world)), (2, (6, 7, 4 * world)), (3, (5, 4, 2 * world)), (4, (4, 5, 2 * world))]:
        mf = O.MatrixFree(3, p, n)
        src = splitmix_src(mf.n_dofs, salt=p)
        op = G.LaplaceOperator(ctx, p, n)
        s, d = op.vector_from(src), op.initialize_dof_vector()
        op.vmult(d, s)
        report("vmult p=%d n=%s" % (p, n), rel_l2(d.export_host(), mf.vmult(src)), 1e-12)
        report("dot p=%d" % p, abs(s.dot(s) - float(src @ src)) / float(src @ src), 1e-13)
    # 1b. the rank's part of a vector and the ghost operations (LinearAlgebra::distributed::Vector update_ghost_values /
    #     zero_out_ghost_values / compress(add)) on slabs
    for p, n in [(2, (8, 6, 4 * world)), (3, (4, 4, 2 * world))]:  # > coarse threshold: distributed levels
        mf = O.MatrixFree(3, p, n)
        op = G.LaplaceOperator(ctx, p, n)
        g = splitmix_src(mf.n_dofs, salt=40 + p)
        v = op.initialize_dof_vector()
        plane, z0, nzl, lo, hi = v.local_range()
        tot = torch.tensor([hi - lo], device="cuda")
        dist.all_reduce(tot)
        report("owned planes tile the mesh p=%d" % p, float(abs(int(tot.item()) - mf.nd[2])), 0.0)
        v.import_owned(g[lo * plane:hi * plane].copy())  # ghosts still zero
        loc = v.export_local()
        own = slice((lo - z0) * plane, (hi - z0) * plane)
        ghost = np.ones(nzl * plane, bool)
        ghost[own] = False
        report("import_owned leaves ghosts p=%d" % p, float(np.abs(loc[ghost]).max() if ghost.any() else 0.0), 0.0)
        report("export_owned round trip p=%d" % p, float(np.abs(v.export_owned() - g[lo * plane:hi * plane]).max()), 0.0)
        v.update_ghost_values()
        report("update_ghost_values p=%d" % p, float(np.abs(v.export_local() - g[z0 * plane:(z0 + nzl) * plane]).max()), 0.0)
        report("export_host after update p=%d" % p, float(np.abs(v.export_host() - g).max()), 0.0)
        v.zero_out_ghost_values()
        loc = v.export_local()
        report("zero_out_ghost_values p=%d" % p, float(np.abs(loc[ghost]).max() if ghost.any() else 0.0) + float(np.abs(loc[own] - g[lo * plane:hi * plane]).max()), 0.0)
        # compress(add): every stored copy of a dof is summed into its owner.  All stored planes = 1 -> an owned plane ends up
        # with 1 + the number of neighbours that store it as a ghost
        v.import_local(np.ones(nzl * plane))
        v.compress_add()
        got = v.export_owned().reshape(hi - lo, plane)[:, 0]
        cnt = np.ones(mf.nd[2])
        ranges = [None] * world
        dist.all_gather_object(ranges, (z0, nzl, lo, hi))
        for (rz0, rn, rlo, rhi) in ranges:
            for z in list(range(rz0, rlo)) + list(range(rhi, rz0 + rn)):
                cnt[z] += 1
        report("compress_add p=%d" % p, float(np.abs(got - cnt[lo:hi]).max()), 0.0)
    # 2. transfers between two distributed levels, and between a distributed and a gathered level
    for kind, pc, pf, nc in [("h", 2, 2, (3, 2, 2 * world)), ("p", 1, 3, (3, 4, 2 * world)), ("h", 1, 1, (4, 4, world))]:
        nf = tuple(2 * c for c in nc) if kind == "h" else nc
        mc, mfine = O.MatrixFree(3, pc, nc), O.MatrixFree(3, pf, nf)
        tr = O.Transfer(mc, mfine, kind)
        oc, of = G.LaplaceOperator(ctx, pc, nc), G.LaplaceOperator(ctx, pf, nf)
        t = G.GeometricTransfer(oc, of) if kind == "h" else G.PolynomialTransfer(oc, of)
        xc, xf = splitmix_src(mc.n_dofs, salt=5), splitmix_src(mfine.n_dofs, salt=6)
        d0 = splitmix_src(mfine.n_dofs, salt=7)
        dc, df = oc.vector_from(xc), of.vector_from(d0)
        t.prolongate_and_add(df, dc)
        report("prolongate %s %d->%d nc=%s" % (kind, pc, pf, nc), rel_l2(df.export_host(), tr.prolongate_and_add(d0.copy(), xc)), 1e-13)
        c0 = splitmix_src(mc.n_dofs, salt=8)
        dc2, df2 = oc.vector_from(c0), of.vector_from(xf)
        t.restrict_and_add(dc2, df2)
        got = dc2.export_host()
        report("restrict   %s %d->%d nc=%s" % (kind, pc, pf, nc), rel_l2(got, tr.restrict_and_add(c0.copy(), xf)), 1e-13)
    # 3. V-cycle + CG: h-MG Q2 and hp-MG Q4, coarse levels gathered on rank 0
    for levels in ([(2, (m, m, m)) for m in (1, 2, 4, 8, 16)], [(1, (1, 1, 1)), (1, (2, 2, 2)), (1, (4, 4, 4)), (1, (8, 8, 8)), (2, (8, 8, 8)), (4, (8, 8, 8))]):
        mfs = [O.MatrixFree(3, p, n) for p, n in levels]
        trs = [O.Transfer(mfs[l - 1], mfs[l], "h" if levels[l][0] == levels[l - 1][0] else "p") for l in range(1, len(levels))]
        vc = O.VCycle(mfs, trs)
        ops, transfers, smoothers, mg = G.build_hierarchy(ctx, levels)
        top = ops[-1]
        r = splitmix_src(mfs[-1].n_dofs, mfs[-1].constrained(), salt=14)
        dr, dz = top.vector_from(r), top.initialize_dof_vector()
        zref = vc.vmult(r)
        for rep in range(3):
            mg.vmult(dz, dr)
            report("V-cycle rep %d top=%s" % (rep, levels[-1]), rel_l2(dz.export_host(), zref), 1e-10)
        b = top.initialize_dof_vector()
        top.assemble_rhs(b)
        x = top.initialize_dof_vector()
        it, hist, rc = G.cg_solve(top, x, b, mg)
        xr, itr, histr, rcr = O.cg_solve(mfs[-1], mfs[-1].assemble_rhs(), vc)
        report("CG iterations %d vs oracle %d" % (it, itr), float(abs(it - itr)), 0.0)
        report("CG residual history", float(np.max(np.abs(hist - histr[: len(hist)]) / histr[0])) if len(hist) == len(histr) else 1.0, 1e-10)
        report("solution", rel_l2(x.export_host(), xr), 1e-9)
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("applies with fused ghost push (no exchange of their own): %d, kernel launches: %d" % (ctx.fused_halo_count(), ctx.launch_count()), flush=True)
        print("DIST CHECK", "PASSED" if t.item() == 1.0 else "FAILED", "on", world, "GPUs", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
