"""N>1 host logic on the CPU: world_size-2 gloo processes reproduce the product's slab layout, halo
exchange pattern (pmg_core.c: p ghost planes from below, 1 from above, no compress) and owned-range dot
products, with the CUDA tile program run under the host emulator on every rank.  The assembled result must
equal the serial oracle, and a Jacobi-CG run on the distributed operator must follow the serial residuals."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import rel_l2, splitmix_src
from test_host_and_emulator import emu_apply, slab_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _halo_update(rank, world, p, plane, sl, v):
    """mirror of pmg_halo_update: top p owned planes -> upper's lower ghosts; first owned plane -> lower's upper ghost"""
    z0, nzl, _, _, zol, zoh = sl
    reqs = []
    if rank + 1 < world:
        send = torch.from_numpy(v[(zoh - p - z0) * plane:(zoh - z0) * plane].copy())
        recv_up = torch.empty(plane, dtype=torch.float64)
        reqs += [dist.isend(send, rank + 1), dist.irecv(recv_up, rank + 1)]
    if rank > 0:
        send2 = torch.from_numpy(v[(zol - z0) * plane:(zol - z0 + 1) * plane].copy())
        recv_lo = torch.empty(plane * p, dtype=torch.float64)
        reqs += [dist.isend(send2, rank - 1), dist.irecv(recv_lo, rank - 1)]
    for r in reqs:
        r.wait()
    if rank + 1 < world:
        v[(zoh - z0) * plane:(zoh - z0 + 1) * plane] = recv_up.numpy()
    if rank > 0:
        v[:p * plane] = recv_lo.numpy()


def _worker(rank, world, port, p, n, out_path):
    import sys

    sys.path.insert(0, os.path.join(ROOT, "portable-multigrid_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pmg_b200 as G
    import pyoracle as O

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libemu.so"))
    lo, hi, ok = G.host_partition(n[2], world, rank)
    assert ok
    sl = slab_of(p, n, lo, hi)
    z0, nzl, _, _, zol, zoh = sl
    Nx, Ny = n[0] * p + 1, n[1] * p + 1
    plane = Nx * Ny
    N = plane * (n[2] * p + 1)
    u_glob = splitmix_src(N, salt=11)
    own = slice((zol - z0) * plane, (zoh - z0) * plane)

    def to_local(g):
        v = np.zeros(nzl * plane)
        v[own] = g[zol * plane:zoh * plane]  # ghosts deliberately stale (zero) until the exchange
        return v

    def dot(a, b):
        t = torch.tensor([float(a[own] @ b[own])], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def apply(v):
        _halo_update(rank, world, p, plane, sl, v)
        o = emu_apply(emu, p, n, v, slab=sl)
        o[np.isnan(o)] = 0.0
        return o

    u = to_local(u_glob)
    Au = apply(u)
    # Jacobi-preconditioned CG on A x = b, distributed
    mf = O.MatrixFree(3, p, n)
    dinv_g = mf.compute_diagonal()
    b_g = mf.assemble_rhs()
    dinv, b = to_local(dinv_g), to_local(b_g)
    x, r = np.zeros_like(b), b.copy()
    z = dinv * r
    pv = z.copy()
    rz = dot(r, z)
    hist = [np.sqrt(dot(r, r))]
    for it in range(8):
        Ap = apply(pv)
        alpha = rz / dot(pv, Ap)
        x[own] += alpha * pv[own]
        r[own] -= alpha * Ap[own]
        hist.append(np.sqrt(dot(r, r)))
        z = dinv * r
        rz_new = dot(r, z)
        pv[own] = z[own] + (rz_new / rz) * pv[own]
        rz = rz_new
    # gather owned slabs on rank 0
    parts = [None] * world
    dist.all_gather_object(parts, (zol, zoh, Au[own].copy(), np.array(hist)))
    if rank == 0:
        got = np.zeros(N)
        for a, bnd, vals, _ in parts:
            got[a * plane:bnd * plane] = vals
        np.savez(out_path, got=got, hist=parts[0][3], u=u_glob)
    dist.destroy_process_group()


@pytest.mark.parametrize("p,n", [(2, (3, 3, 4)), (3, (2, 3, 2))])
def test_two_rank_slab_apply_and_cg(p, n, emu, oracle, tmp_path):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), p, n, out), nprocs=2, join=True)
    res = np.load(out)
    mf = oracle.MatrixFree(3, p, n)
    assert rel_l2(res["got"], mf.vmult(res["u"])) < 1e-13
    # serial Jacobi-CG residual history
    dinv, b = mf.compute_diagonal(), mf.assemble_rhs()
    x, r = np.zeros_like(b), b.copy()
    z = dinv * r
    pv, rz = z.copy(), r @ z
    hist = [np.linalg.norm(r)]
    for it in range(8):
        Ap = mf.vmult(pv)
        alpha = rz / (pv @ Ap)
        x += alpha * pv
        r -= alpha * Ap
        hist.append(np.linalg.norm(r))
        z = dinv * r
        rz_new = r @ z
        pv = z + (rz_new / rz) * pv
        rz = rz_new
    assert np.allclose(res["hist"], hist, rtol=1e-10, atol=1e-14)
