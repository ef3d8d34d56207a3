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


# ---- the smoother's chain of fused applies (host/pmg_smoother.c, csrc/pmg_apply_plane_launch.h) on two ranks ---------------------
def _chain_worker(rank, world, port, p, n, degree, out_path):
    """Mirror of pmg_chebyshev_smooth_chain on z-slabs: ONE stand-alone exchange before the first apply, then every apply of the
    plane-per-step kernel (host emulator) copies its boundary planes into the neighbour's ghost planes -- here through shared
    memory segments that stand for the CUDA-IPC mapping, with a barrier standing for the flag words."""
    import sys
    from multiprocessing import shared_memory

    sys.path.insert(0, os.path.join(ROOT, "portable-multigrid_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pmg_b200 as G
    import pyoracle as O

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libemu.so"))
    emu.emu_plane_set_push.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    emu.emu_plane_set_push.restype = None
    lo, hi, ok = G.host_partition(n[2], world, rank)
    assert ok
    slabs = [slab_of(p, n, *G.host_partition(n[2], world, r)[:2]) for r in range(world)]
    sl = slabs[rank]
    z0, nzl, _, _, zol, zoh = sl
    plane = (n[0] * p + 1) * (n[1] * p + 1)
    N = plane * (n[2] * p + 1)
    # the two ping-pong vectors of every rank live in named shared memory: the neighbours map them
    name = lambda r, v: "pmgt%d_%d_%d" % (port, r, v)
    mine = [shared_memory.SharedMemory(name=name(rank, v), create=True, size=nzl * plane * 8) for v in range(2)]
    dist.barrier()
    peers = {r: [shared_memory.SharedMemory(name=name(r, v)) for v in range(2)] for r in (rank - 1, rank + 1) if 0 <= r < world}
    arr = lambda shm, nz: np.ndarray((nz * plane,), dtype=np.float64, buffer=shm.buf)
    vec = [arr(m, nzl) for m in mine]
    peer_vec = {r: [arr(s, slabs[r][1]) for s in shms] for r, shms in peers.items()}
    try:
        mf = O.MatrixFree(3, p, n)
        u_g, b_g = splitmix_src(N, mf.constrained(), salt=21), splitmix_src(N, mf.constrained(), salt=22)
        own = slice((zol - z0) * plane, (zoh - z0) * plane)
        cut = lambda g: g[z0 * plane:(z0 + nzl) * plane].copy()
        vec[0][:] = 0.0
        vec[0][own] = u_g[zol * plane:zoh * plane]  # ghosts stale until the one exchange of the chain
        vec[1][:] = np.nan
        b = cut(b_g)
        theta, delta = 1.7, 1.1
        cur, other = 0, 1
        _halo_update(rank, world, p, plane, sl, vec[cur])  # the chain's only stand-alone exchange
        rhok, sigma = delta / theta, theta / delta
        for i in range(degree):
            dist.barrier()  # the flag words: the neighbours have finished the launch before this one
            lo_ptr = peer_vec[rank - 1][other].ctypes.data if rank > 0 else None
            hi_ptr = peer_vec[rank + 1][other].ctypes.data if rank + 1 < world else None
            emu.emu_plane_set_push(lo_ptr, slabs[rank - 1][0] if rank > 0 else 0, hi_ptr, slabs[rank + 1][0] if rank + 1 < world else 0)
            if i == 0:
                emu_apply(emu, p, n, vec[cur], mode=2, b=b, f2=1.0 / theta, slab=sl, chunks=2, out=vec[other], kernel="plane")
            else:
                rhokp = 1.0 / (2.0 * sigma - rhok)
                f1, f2 = rhokp * rhok, 2.0 * rhokp / delta
                rhok = rhokp
                emu_apply(emu, p, n, vec[cur], mode=3, b=b, xold=vec[other], f1=f1, f2=f2, slab=sl, chunks=2, out=vec[other], kernel="plane")
            cur, other = other, cur
        emu.emu_plane_set_push(None, 0, None, 0)
        dist.barrier()
        parts = [None] * world
        dist.all_gather_object(parts, (z0, nzl, vec[cur].copy()))
        if rank == 0:
            np.savez(out_path, u=u_g, b=b_g, **{"local%d" % r: parts[r][2] for r in range(world)},
                     ranges=np.array([[parts[r][0], parts[r][1]] for r in range(world)]))
        dist.barrier()
    finally:
        del vec, peer_vec
        for shms in peers.values():
            for s in shms:
                s.close()
        dist.barrier()
        for m in mine:
            m.close()
            m.unlink()
        dist.destroy_process_group()


@pytest.mark.parametrize("p,n,degree", [(2, (3, 2, 4), 5), (4, (2, 2, 4), 3)])
def test_two_rank_fused_smoother_chain(p, n, degree, emu, oracle, tmp_path):
    out = str(tmp_path / "chain.npz")
    mp.spawn(_chain_worker, args=(2, _free_port(), p, n, degree, out), nprocs=2, join=True)
    res = np.load(out)
    mf = oracle.MatrixFree(3, p, n)
    dinv, u, b = mf.compute_diagonal(), res["u"], res["b"]
    theta, delta = 1.7, 1.1
    x_old, x = u, u + (1.0 / theta) * dinv * (b - mf.vmult(u))
    rhok, sigma = delta / theta, theta / delta
    for _ in range(degree - 1):
        rhokp = 1.0 / (2.0 * sigma - rhok)
        f1, f2 = rhokp * rhok, 2.0 * rhokp / delta
        rhok = rhokp
        x_old, x = x, x + f1 * (x - x_old) + f2 * dinv * (b - mf.vmult(x))
    plane = mf.nd[0] * mf.nd[1]
    for r, (z0, nzl) in enumerate(res["ranges"]):
        # every stored plane of every rank -- owned AND ghost -- holds the serial iterate: the last apply pushed its planes too
        assert rel_l2(res["local%d" % r], x[z0 * plane:(z0 + nzl) * plane]) < 1e-12
