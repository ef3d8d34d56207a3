"""CPU tests of the product's host logic and of the CUDA tile program run under the host emulator.

No compute entry of libpmg.so is called here (there is no GPU): the library must load, export every
symbol include/*.h declares, fail loudly without a device, and its host-only helpers must agree with the
oracle.  The kernels' index logic / ownership rules / fused epilogues are checked by executing the same
source (csrc/pmg_apply_tile.h) thread by thread on the CPU (tests/emu)."""
import ctypes as C
import functools
import os
import re

import numpy as np
import pytest

from helpers import hierarchy_levels, rel_l2, splitmix_src

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)


def P(a):
    return None if a is None else a.ctypes.data_as(dp)


# ---- C-ABI surface ---------------------------------------------------------------------------------
def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pmgk?_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pmg):
    lib = pmg.lib()
    names = _declared("pmg.h") + _declared("pmg_kernels.h")
    assert len(names) > 80
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(pmg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pmg.PmgError) as e:
        pmg.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "portable-multigrid_b200")
    for d, _, files in os.walk(pkg):
        if os.sep + "build" in d:
            continue
        for f in files:
            if f.endswith((".c", ".h", ".cu", ".py", "Makefile")):
                txt = open(os.path.join(d, f), errors="replace").read()
                assert "pyoracle" not in txt and "liborc" not in txt and "orc.h" not in txt, os.path.join(d, f)


# ---- host helpers vs oracle ------------------------------------------------------------------------
@pytest.mark.parametrize("p", range(1, 10))
def test_fastdiag_tables(p, pmg, oracle):
    S, lam = pmg.host_fastdiag_tables(p)
    M, K = pmg.host_pencil(p)
    Sv, Dco, w = oracle.shape_tables(p)
    Mo, Ko = Sv.T @ np.diag(w) @ Sv, (Dco @ Sv).T @ np.diag(w) @ (Dco @ Sv)  # the reference's quadrature
    assert np.abs(M - Mo).max() < 1e-14 and np.abs(K - Ko).max() < 1e-11
    assert np.abs(S.T @ S - M).max() < 1e-14
    assert np.abs(S.T @ np.diag(lam) @ S - K).max() < 1e-11
    assert lam[0] == 0.0 and np.all(np.diff(lam) > 0)
    assert np.linalg.cond(S) < 40  # well conditioned change of basis: no accuracy lost


@pytest.mark.parametrize("p", range(1, 10))
def test_prolongation_matrices(p, pmg, oracle):
    assert np.abs(pmg.host_prolongation_1d(0, p) - oracle.h_prolongation_1d(p)).max() < 1e-14
    for pf in range(p + 1, 10):
        assert np.abs(pmg.host_prolongation_1d(1, p, pf) - oracle.p_prolongation_1d(p, pf)).max() < 1e-14


def test_chebyshev_parameters_and_tridiag(pmg, oracle):
    theta, delta, k = pmg.host_chebyshev_parameters(0.3, 1.9, 15.0, 5)
    lmax = 1.2 * 1.9
    assert k == 5 and abs(theta - 0.5 * (lmax + lmax / 15)) < 1e-15 and abs(delta - 0.5 * (lmax - lmax / 15)) < 1e-15
    theta, delta, k = pmg.host_chebyshev_parameters(1.0, 1.0, 1e-3, pmg.PMG_INVALID_DEGREE)
    assert k == 3 and abs(theta - 1.1) < 1e-15  # one-dof coarse level of the geometric driver
    rng = np.random.default_rng(0)
    for n in (1, 2, 9, 40):
        d, e = rng.standard_normal(n) + 3, rng.standard_normal(max(n - 1, 0))
        ev = oracle.tridiag_eigenvalues(d, e)
        lo, hi = pmg.host_tridiag_extreme_eigenvalues(d, e)
        assert abs(lo - ev[0]) < 1e-11 and abs(hi - ev[-1]) < 1e-11


def test_partition(pmg):
    for nz, R in [(320, 8), (64, 2), (40, 4), (8, 8)]:
        slabs = [pmg.host_partition(nz, R, r) for r in range(R)]
        assert all(ok for _, _, ok in slabs)
        assert slabs[0][0] == 0 and slabs[-1][1] == nz
        assert all(slabs[r][1] == slabs[r + 1][0] for r in range(R - 1))
        fine = [pmg.host_partition(2 * nz, R, r) for r in range(R)]
        assert all(f[0] == 2 * c[0] and f[1] == 2 * c[1] for f, c in zip(fine, slabs))  # refinement keeps slabs nested
    lo, hi, ok = pmg.host_partition(5, 2, 0)
    assert not ok and (lo, hi) == (0, 5)  # not divisible: the level lives on rank 0
    assert pmg.host_partition(5, 2, 1)[:2] == (0, 0)
    assert pmg.host_partition(7, 1, 0) == (0, 7, True)


KERNELS = ["sweep", "celltile", "plane"]


# ---- the CUDA tile programs under the host emulator --------------------------------------------------
def _tables(emu, p, h):
    n = p + 1
    S, lam, tab = np.zeros(n * n), np.zeros(n), np.zeros((p + 2) ** 3)
    emu.pmg_fe_fastdiag(C.c_int(p), P(S), P(lam))
    emu.pmg_fe_dinv_table(C.c_int(p), P(np.asarray(h, dtype=np.float64)), C.c_int(3), P(tab))
    return S, lam, tab


def emu_apply(emu, p, n, u, mode=0, b=None, xold=None, f1=0.0, f2=0.0, small=1, chunks=1, faces=0x3F, slab=None, dinv_vec=None, out=None,
              kernel="sweep"):
    """Run one launch of the apply kernel under the host emulator.  kernel = "sweep" (line-marching kernel, the
    default launch) or "celltile" (cell-tile kernel, PMG_TILE_VARIANT >= 2)."""
    nx, ny, nz = n
    h = np.array([1.0 / nx, 1.0 / ny, 1.0 / nz])
    S, lam, tab = _tables(emu, p, h)
    Nz = nz * p + 1
    z0, nzl, czlo, czhi, zol, zoh = (0, Nz, 0, nz, 0, Nz) if slab is None else slab
    if out is None:
        out = np.full(u.shape, np.nan)
    if kernel in ("sweep", "plane"):
        M, K = np.zeros((p + 1) ** 2), np.zeros((p + 1) ** 2)
        emu.pmg_fe_pencil(C.c_int(p), P(M), P(K))
        rc = (emu.emu_sweep if kernel == "sweep" else emu.emu_plane)(p, small, nx, ny, nz, C.c_uint(faces), z0, nzl, czlo, czhi, zol, zoh, chunks, P(M), P(K), P(h), mode,
                           P(u), P(b), P(xold), P(out), C.c_double(f1), C.c_double(f2), P(dinv_vec), P(tab))
    else:
        rc = emu.emu_apply(p, small, nx, ny, nz, C.c_uint(faces), z0, nzl, czlo, czhi, zol, zoh, chunks, P(S), P(lam), P(h), mode,
                           P(u), P(b), P(xold), P(out), C.c_double(f1), C.c_double(f2), P(dinv_vec), P(tab))
    assert rc == 0
    return out


def slab_of(p, n, cz_lo, cz_hi):
    """(z0, nzl, cz_lo, cz_hi, z_own_lo, z_own_hi) of the product's ghost layout (pmg_core.c)."""
    Nz = n[2] * p + 1
    z0 = 0 if cz_lo == 0 else cz_lo * p - p
    z_end = Nz if cz_hi == n[2] else cz_hi * p + 1
    return z0, z_end - z0, cz_lo, cz_hi, cz_lo * p, (Nz if cz_hi == n[2] else cz_hi * p)


@pytest.mark.parametrize("p", range(1, 10))
@pytest.mark.parametrize("small,chunks", [(1, 1), (1, 3), (0, 2)])
@pytest.mark.parametrize("kernel", KERNELS)
def test_emulated_apply_matches_oracle(p, small, chunks, kernel, emu, oracle):
    if kernel == "plane" and small == 0 and p > 6:
        pytest.skip("the plane-per-step kernel ships tiles for degrees 1..6 (csrc/pmg_apply_plane_tiles.inc)")
    n = (5, 4, 3) if p < 5 else (3, 2, 3)
    mf = oracle.MatrixFree(3, p, n)
    u = splitmix_src(mf.n_dofs, salt=p)
    out = emu_apply(emu, p, n, u, small=small, chunks=chunks, kernel=kernel)
    assert not np.isnan(out).any()  # every dof is written by exactly one owner
    assert rel_l2(out, mf.vmult(u)) < 1e-13


@pytest.mark.parametrize("p", [1, 2, 3, 4])
@pytest.mark.parametrize("kernel", KERNELS)
def test_emulated_epilogues_and_faces(p, kernel, emu, oracle):
    n = (4, 3, 4)
    emu_apply = functools.partial(globals()["emu_apply"], kernel=kernel)
    mf = oracle.MatrixFree(3, p, n)
    u, b, xo = (splitmix_src(mf.n_dofs, salt=s) for s in (1, 2, 3))
    Au, dinv = mf.vmult(u), mf.compute_diagonal()
    f1, f2 = 0.37, 0.81
    assert rel_l2(emu_apply(emu, p, n, u, mode=1, b=b), b - Au) < 1e-13
    assert rel_l2(emu_apply(emu, p, n, u, mode=2, b=b, f2=f2), u + f2 * dinv * (b - Au)) < 1e-13
    ref = u + f1 * (u - xo) + f2 * dinv * (b - Au)
    assert rel_l2(emu_apply(emu, p, n, u, mode=3, b=b, xold=xo, f1=f1, f2=f2), ref) < 1e-13
    x2 = xo.copy()
    emu_apply(emu, p, n, u, mode=3, b=b, xold=x2, f1=f1, f2=f2, out=x2)  # x_old overwritten in place
    assert rel_l2(x2, ref) < 1e-13
    assert rel_l2(emu_apply(emu, p, n, u, mode=3, b=b, f1=f1, f2=f2), u + f1 * u + f2 * dinv * (b - Au)) < 1e-13
    assert rel_l2(emu_apply(emu, p, n, u, mode=3, b=b, xold=xo, f1=f1, f2=f2, dinv_vec=dinv), ref) < 1e-13
    for faces in (0x00, 0x15, 0x2A, 0x03):
        m2 = oracle.MatrixFree(3, p, n, faces=faces)
        A2, d2 = m2.vmult(u), m2.compute_diagonal()
        assert rel_l2(emu_apply(emu, p, n, u, faces=faces), A2) < 1e-13
        assert rel_l2(emu_apply(emu, p, n, u, mode=2, b=b, f2=f2, faces=faces), u + f2 * d2 * (b - A2)) < 1e-13


@pytest.mark.parametrize("p,splits", [(1, [(0, 2), (2, 4)]), (3, [(0, 1), (1, 3), (3, 4)]), (4, [(0, 2), (2, 4)])])
@pytest.mark.parametrize("kernel", KERNELS)
def test_emulated_slabs_cover_the_serial_result(p, splits, kernel, emu, oracle):
    """Each rank applies the operator to its z-slab (ghost layer below, ghost plane above): the owned parts
    tile the serial result and nothing outside the owned planes is written."""
    n = (3, 4, 4)
    mf = oracle.MatrixFree(3, p, n)
    u = splitmix_src(mf.n_dofs, salt=7)
    ref = mf.vmult(u)
    plane = mf.nd[0] * mf.nd[1]
    got = np.full(mf.n_dofs, np.nan)
    for lo, hi in splits:
        z0, nzl, _, _, zol, zoh = sl = slab_of(p, n, lo, hi)
        ul = u[z0 * plane:(z0 + nzl) * plane].copy()
        for chunks in (1, 2):  # a slab launched whole takes whatever chunking fills the SMs, one chunk included
            ol = emu_apply(emu, p, n, ul, slab=sl, chunks=chunks, kernel=kernel)
            owned = np.zeros(nzl * plane, bool)
            owned[(zol - z0) * plane:(zoh - z0) * plane] = True
            assert np.isnan(ol[~owned]).all() and not np.isnan(ol[owned]).any()
            assert rel_l2(ol[owned], ref[zol * plane:zoh * plane]) < 1e-13
        got[zol * plane:zoh * plane] = ol[owned]
    assert rel_l2(got, ref) < 1e-13


@pytest.mark.parametrize("p,small", [(1, 1), (2, 0), (3, 1), (4, 0), (4, 3), (6, 1)])
@pytest.mark.parametrize("mode,faces", [(0, 0x3F), (3, 0x3F), (3, 0x09), (0, 0x16)])
def test_emulated_fused_ghost_push(p, small, mode, faces, emu, oracle):
    """Fused ghost push of the plane-per-step kernel (csrc/pmg_apply_plane.h epilogue): every rank's launch stores its
    boundary planes of the result into the neighbours' arrays, so that afterwards each rank's ghost planes of `out` hold what
    an exchange would have delivered -- and nothing else is written there."""
    n = (3, 4, 6)
    splits = [(0, 2), (2, 4), (4, 6)]
    # faces: Dirichlet faces of the mesh (bits x-, x+, y-, y+, z-, z+); without a Dirichlet high face in x / y the tiles also hold
    # the mesh's last vertex lines, which the push has to carry too
    mf = oracle.MatrixFree(3, p, n, faces=faces)
    u, b, xo = (splitmix_src(mf.n_dofs, salt=s) for s in (7, 8, 9))
    Au = mf.vmult(u)
    f1, f2 = 0.3, 0.6
    ref = Au if mode == 0 else u + f1 * (u - xo) + f2 * mf.compute_diagonal() * (b - Au)
    plane = mf.nd[0] * mf.nd[1]
    slabs = [slab_of(p, n, lo, hi) for lo, hi in splits]
    outs = [np.full(sl[1] * plane, np.nan) for sl in slabs]
    emu.emu_plane_set_push.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    emu.emu_plane_set_push.restype = None
    try:
        for r, sl in enumerate(slabs):
            z0, nzl = sl[0], sl[1]
            lo = outs[r - 1].ctypes.data if r > 0 else None
            hi = outs[r + 1].ctypes.data if r + 1 < len(slabs) else None
            emu.emu_plane_set_push(lo, slabs[r - 1][0] if r > 0 else 0, hi, slabs[r + 1][0] if r + 1 < len(slabs) else 0)
            cut = lambda v: v[z0 * plane:(z0 + nzl) * plane].copy()
            emu_apply(emu, p, n, cut(u), mode=mode, b=cut(b), xold=cut(xo), f1=f1, f2=f2, small=small, chunks=2, slab=sl, out=outs[r], kernel="plane",
                      faces=faces, dinv_vec=cut(mf.compute_diagonal()) if mode else None)
    finally:
        emu.emu_plane_set_push(None, 0, None, 0)
    for r, sl in enumerate(slabs):  # every stored plane of every rank now equals the serial result: owned and ghost planes
        z0, nzl = sl[0], sl[1]
        assert not np.isnan(outs[r]).any()
        assert rel_l2(outs[r], ref[z0 * plane:(z0 + nzl) * plane]) < 1e-13


@pytest.mark.parametrize("p", range(1, 9))
@pytest.mark.parametrize("small,chunks", [(2, 2), (3, 1), (3, 3), (4, 1), (4, 2)])
def test_emulated_plane_kernel_variants(p, small, chunks, emu, oracle):
    """Plane-per-step kernel (csrc/pmg_apply_plane.h): threads of a phase run in descending order (small = 2: a hazard between
    threads of one phase would show), the steps of a cell layer rolled (small = 3); mixed Dirichlet faces, fused step."""
    if small == 4 and p == 1:
        pytest.skip("two x items per cell row need two nodes")
    n = (5, 3, 4) if p < 5 else (3, 2, 3)
    for faces in (0x3F, 0x19):
        mf = oracle.MatrixFree(3, p, n, faces=faces)
        u, b, xo = (splitmix_src(mf.n_dofs, salt=s + p) for s in (1, 2, 3))
        Au, dinv = mf.vmult(u), mf.compute_diagonal()
        out = emu_apply(emu, p, n, u, small=small, chunks=chunks, faces=faces, kernel="plane")
        assert not np.isnan(out).any()
        assert rel_l2(out, Au) < 1e-13
        ref = u + 0.3 * (u - xo) + 0.9 * dinv * (b - Au)
        assert rel_l2(emu_apply(emu, p, n, u, mode=3, b=b, xold=xo, f1=0.3, f2=0.9, small=small, chunks=chunks, faces=faces, kernel="plane"), ref) < 1e-13


@pytest.mark.parametrize("p", range(1, 9))
@pytest.mark.parametrize("small", [2, 3])
def test_emulated_sweep_with_two_segments_per_line(p, small, emu, oracle):
    """SG = 2 (csrc/pmg_apply_sweep.h): the y lines of phase 1 and the x lines of phase 2 are marched by two threads each; phase 2
    works in place, so a segment reads what its neighbour overwrites before either starts.  small = 3 runs the threads of every
    phase in descending order, which turns a missed hand-over into a wrong result.  (Measured slower on B200, kept as a knob.)"""
    n = (5, 4, 3) if p < 5 else (4, 3, 2)
    for faces in (0x3F, 0x15):
        mf = oracle.MatrixFree(3, p, n, faces=faces)
        u, b, xo = (splitmix_src(mf.n_dofs, salt=s) for s in (71, 72, 73))
        Au, dinv = mf.vmult(u), mf.compute_diagonal()
        assert rel_l2(emu_apply(emu, p, n, u, small=small, chunks=2, faces=faces), Au) < 1e-13
        ref = u + 0.3 * (u - xo) + 0.7 * dinv * (b - Au)
        assert rel_l2(emu_apply(emu, p, n, u, mode=3, b=b, xold=xo, f1=0.3, f2=0.7, small=small, faces=faces), ref) < 1e-13


@pytest.mark.parametrize("p", [2, 3, 4])
def test_emulated_shipped_tiles_with_rolled_cell_loops(p, emu, oracle):
    """The launch tiles of csrc/pmg_apply_sweep_tiles.inc with RL = 1 -- the configuration the Q4 plain apply is launched with
    (PmgSweepModeTune, csrc/pmg_apply_sweep_launch.h) -- on a mesh wider than one tile in x and y."""
    n = (9, 10, 3) if p < 4 else (6, 5, 3)
    mf = oracle.MatrixFree(3, p, n)
    u = splitmix_src(mf.n_dofs, salt=p)
    out = emu_apply(emu, p, n, u, small=8, chunks=2)
    assert not np.isnan(out).any() and rel_l2(out, mf.vmult(u)) < 1e-13


@pytest.mark.parametrize("p", range(1, 9))
@pytest.mark.parametrize("small", [4, 5])
def test_emulated_sweep_with_rolled_cell_loops(p, small, emu, oracle):
    """RL = 1 (csrc/pmg_apply_sweep.h): the cell loops of the y and x sweeps stay rolled (smaller code); for odd degrees the
    parity of a cell's first staged row is then a run-time value.  Odd and even cell counts per tile, all epilogues.
    small = 5: A2 = 1 as well, the x sweep's work items move to the other half of the CTA's threads in every second step (several
    steps per chunk, so both assignments run)."""
    n = (5, 5, 5) if p < 5 else (3, 4, 3)
    for faces in (0x3F, 0x2A):
        mf = oracle.MatrixFree(3, p, n, faces=faces)
        u, b, xo = (splitmix_src(mf.n_dofs, salt=s) for s in (81, 82, 83))
        Au, dinv = mf.vmult(u), mf.compute_diagonal()
        assert rel_l2(emu_apply(emu, p, n, u, small=small, chunks=2, faces=faces), Au) < 1e-13
        ref = u + 0.3 * (u - xo) + 0.7 * dinv * (b - Au)
        assert rel_l2(emu_apply(emu, p, n, u, mode=3, b=b, xold=xo, f1=0.3, f2=0.7, small=small, faces=faces), ref) < 1e-13




# ---- variable-coefficient tile program (csrc/pmg_apply_var.h) under the host emulator ---------------------
def emu_var(emu, p, n, u, mode=0, b=None, xold=None, f1=0.0, f2=0.0, small=1, chunks=1, faces=0x3F, slab=None, dinv_vec=None,
            want_dinv=False):
    nx, ny, nz = n
    h = np.array([1.0 / nx, 1.0 / ny, 1.0 / nz])
    Nz = nz * p + 1
    z0, nzl, czlo, czhi, zol, zoh = (0, Nz, 0, nz, 0, Nz) if slab is None else slab
    out = np.full(u.shape, np.nan)
    dv = np.empty(u.shape) if want_dinv else None
    rc = emu.emu_var(p, small, nx, ny, nz, C.c_uint(faces), z0, nzl, czlo, czhi, zol, zoh, chunks, P(h), mode, P(u), P(b), P(xold),
                     P(out), C.c_double(f1), C.c_double(f2), P(dinv_vec), P(dv))
    assert rc == 0
    return (out, dv) if want_dinv else out


@pytest.mark.parametrize("p", range(1, 10))
@pytest.mark.parametrize("small,chunks", [(1, 1), (1, 3), (0, 2)])
def test_emulated_variable_coefficient_apply_matches_oracle(p, small, chunks, emu, oracle):
    n = (5, 4, 3) if p < 5 else (3, 2, 3)
    mf = oracle.MatrixFree(3, p, n, coef="c5")
    u = splitmix_src(mf.n_dofs, salt=p)
    out = emu_var(emu, p, n, u, small=small, chunks=chunks)
    assert not np.isnan(out).any()
    assert rel_l2(out, mf.vmult(u)) < 1e-13
    const = oracle.MatrixFree(3, p, n).vmult(u)
    assert rel_l2(out, const) > 1e-2  # the coefficient matters: not the constant-coefficient operator


@pytest.mark.parametrize("p", [1, 2, 3, 5])
def test_emulated_variable_coefficient_diagonal_epilogues_faces(p, emu, oracle):
    n = (3, 4, 3) if p < 5 else (2, 3, 2)
    mf = oracle.MatrixFree(3, p, n, coef="c5")
    u, b, xo = (splitmix_src(mf.n_dofs, salt=s) for s in (21, 22, 23))
    Au, dinv = mf.vmult(u), mf.compute_diagonal()
    f1, f2 = 0.3, 0.8
    out, dv = emu_var(emu, p, n, u, mode=2, b=b, f2=f2, want_dinv=True)
    assert rel_l2(dv, dinv) < 1e-13  # diagonal by the squared-shape-function formula == e_i^T A e_i of the oracle
    assert rel_l2(out, u + f2 * dinv * (b - Au)) < 1e-13
    assert rel_l2(emu_var(emu, p, n, u, mode=1, b=b), b - Au) < 1e-13
    assert rel_l2(emu_var(emu, p, n, u, mode=3, b=b, xold=xo, f1=f1, f2=f2), u + f1 * (u - xo) + f2 * dinv * (b - Au)) < 1e-13
    for faces in (0x00, 0x2A):
        m2 = oracle.MatrixFree(3, p, n, faces=faces, coef="c5")
        assert rel_l2(emu_var(emu, p, n, u, faces=faces), m2.vmult(u)) < 1e-13


@pytest.mark.parametrize("p,splits", [(1, [(0, 2), (2, 4)]), (3, [(0, 1), (1, 3), (3, 4)])])
def test_emulated_variable_coefficient_slabs(p, splits, emu, oracle):
    """z-slabs: each rank stores the coefficient of its own cell layers plus the ghost layer below."""
    n = (3, 4, 4)
    mf = oracle.MatrixFree(3, p, n, coef="c5")
    u, b = splitmix_src(mf.n_dofs, salt=7), splitmix_src(mf.n_dofs, salt=8)
    ref = u + 0.6 * mf.compute_diagonal() * (b - mf.vmult(u))
    plane = mf.nd[0] * mf.nd[1]
    got = np.full(mf.n_dofs, np.nan)
    for lo, hi in splits:
        z0, nzl, _, _, zol, zoh = sl = slab_of(p, n, lo, hi)
        ul, bl = u[z0 * plane:(z0 + nzl) * plane].copy(), b[z0 * plane:(z0 + nzl) * plane].copy()
        ol = emu_var(emu, p, n, ul, mode=2, b=bl, f2=0.6, slab=sl, chunks=2)
        owned = np.zeros(nzl * plane, bool)
        owned[(zol - z0) * plane:(zoh - z0) * plane] = True
        assert np.isnan(ol[~owned]).all() and not np.isnan(ol[owned]).any()
        got[zol * plane:zoh * plane] = ol[owned]
    assert rel_l2(got, ref) < 1e-13


# ---- the 2-D kernels' per-DoF functions (csrc/pmg_dim2.h) under the host emulator ----------------------------------
@pytest.mark.parametrize("p", range(1, 10))
@pytest.mark.parametrize("faces", [0xF, 0x5, 0x0])
def test_emulated_2d_apply_and_epilogues(p, faces, emu, oracle):
    n = (5, 3) if p < 5 else (3, 2)
    mf = oracle.MatrixFree(2, p, n, faces=faces)
    u, b, xo = (splitmix_src(mf.n_dofs, salt=s) for s in (41, 42, 43))
    Au, dinv = mf.vmult(u), mf.compute_diagonal()

    def run(mode, f1=0.0, f2=0.0, xold=None, dinv_vec=None, out=None):
        out = np.empty(mf.n_dofs) if out is None else out
        assert emu.emu2_apply(p, n[0], n[1], C.c_uint(faces), mode, P(u), P(b), P(xold), P(out), C.c_double(f1), C.c_double(f2), P(dinv_vec)) == 0
        return out

    assert rel_l2(run(0), Au) < 1e-13
    assert np.array_equal(run(0)[mf.constrained()], u[mf.constrained()])
    assert rel_l2(run(1), b - Au) < 1e-13
    assert rel_l2(run(2, f2=0.7), u + 0.7 * dinv * (b - Au)) < 1e-13
    ref = u + 0.3 * (u - xo) + 0.7 * dinv * (b - Au)
    assert rel_l2(run(3, 0.3, 0.7, xold=xo), ref) < 1e-13
    assert rel_l2(run(3, 0.3, 0.7, xold=xo, dinv_vec=dinv), ref) < 1e-13
    x2 = xo.copy()
    run(3, 0.3, 0.7, xold=x2, out=x2)  # x_old overwritten in place
    assert rel_l2(x2, ref) < 1e-13


@pytest.mark.parametrize("kind,pc,pf,n", [("h", 1, 1, (3, 2)), ("h", 2, 2, (3, 2)), ("h", 3, 3, (2, 3)), ("h", 7, 7, (1, 2)),
                                          ("p", 1, 2, (3, 2)), ("p", 3, 7, (2, 3)), ("p", 6, 7, (2, 2)), ("p", 2, 4, (3, 3))])
def test_emulated_2d_transfers(kind, pc, pf, n, emu, oracle):
    nf = tuple(2 * c for c in n) if kind == "h" else n
    mc, mf = oracle.MatrixFree(2, pc, n), oracle.MatrixFree(2, pf, nf)
    t = oracle.Transfer(mc, mf, kind)
    xc, rf = splitmix_src(mc.n_dofs, mc.constrained(), salt=44), splitmix_src(mf.n_dofs, salt=45)
    d0, c0 = splitmix_src(mf.n_dofs, salt=46), splitmix_src(mc.n_dofs, salt=47)  # "_and_add": the targets are not zero
    ref_p, ref_r = t.prolongate_and_add(d0.copy(), xc), t.restrict_and_add(c0.copy(), rf)
    dp, dr = d0.copy(), c0.copy()
    k = 0 if kind == "h" else 1
    assert emu.emu2_prolongate_and_add(k, pc, pf, n[0], n[1], C.c_uint(0xF), P(dp), P(xc)) == 0
    assert emu.emu2_restrict_and_add(k, pc, pf, n[0], n[1], C.c_uint(0xF), P(dr), P(rf)) == 0
    assert rel_l2(dp, ref_p) < 1e-13 and rel_l2(dr, ref_r) < 1e-13


# ---- the single-CTA coarse V-cycle program (csrc/pmg_coarse_cycle.h) under the host emulator ------------------------
@pytest.mark.parametrize("p,n,faces", [(1, 8, 0x3F), (2, 4, 0x3F), (1, 4, 0x15), (3, 2, 0x3F), (4, 2, 0x3F)])
def test_emulated_coarse_cycle_matches_oracle_vcycle(p, n, faces, emu, oracle):
    """All levels of a small h-hierarchy in one program: same result as the oracle's VCycleMultigrid::vmult (the Chebyshev
    parameters come from the oracle's eigenvalue estimates, as the host passes its own to the kernel)."""
    levels = hierarchy_levels("h", p, n)
    mfs = [oracle.MatrixFree(3, q, m, faces=faces) for (q, m) in levels]
    trs = [oracle.Transfer(mfs[l - 1], mfs[l], "h") for l in range(1, len(levels))]
    vc = oracle.VCycle(mfs, trs)
    est = vc.estimate()
    L = len(levels)
    deg = (C.c_int * L)(*[e[2] for e in est])
    theta, delta = (C.c_double * L)(*[e[4] for e in est]), (C.c_double * L)(*[e[5] for e in est])
    r = splitmix_src(mfs[-1].n_dofs, mfs[-1].constrained(), salt=61)
    out = np.full(mfs[-1].n_dofs, np.nan)
    assert emu.emu_coarse_cycle(p, L, (C.c_int * 3)(1, 1, 1), C.c_uint(faces), 2, 2, deg, theta, delta, P(out), P(r)) == 0
    assert rel_l2(out, vc.vmult(r)) < 1e-12


def test_roofline_traffic_file_matches_the_committed_ncu_capture():
    """bench.py reports roofline.traffic from profiles/r02_cheb_step_q4_c2.json; that number must be the dram bytes of the fused
    Chebyshev step in the committed raw page of the ncu capture (profiles/r02_plane_q4_c2_ncu_raw.csv), not a free constant."""
    import csv
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "profiles", "r02_cheb_step_q4_c2.json")) as f:
        j = json.load(f)
    with open(os.path.join(root, "profiles", "r02_plane_q4_c2_ncu_raw.csv"), newline="") as f:
        rows = list(csv.reader(f))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    seen = []
    for r in rows[2:]:
        if "4, 4, 4, 96, 5, 0, 3" not in r[col["Kernel Name"]]:  # the fused step: FM = 3, rolled layer loop
            continue
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale[units[col["dram__bytes_write.sum"]]]
        seen.append(rd + wr)
    assert seen, "the capture holds launches of the fused Chebyshev step"
    assert min(abs(s - j["dram_bytes_per_launch"]) for s in seen) <= 1e-3 * j["dram_bytes_per_launch"]
    assert abs(j["algorithmic_bytes_per_launch"] - 32 * 16974593) < 1
    assert 0.9 < j["dram_bytes_per_launch"] / j["algorithmic_bytes_per_launch"] < 1.1  # no wasted re-reads
