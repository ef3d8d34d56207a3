"""CPU tests of the oracle (oracle/): known answers, committed vectors and the invariants the reference's
algorithm must satisfy.  The reference has no tests of its own, so these are the pins."""
import json
import os

import numpy as np
import pytest

from helpers import hierarchy_levels, rel_l2, splitmix_src

HERE = os.path.dirname(os.path.abspath(__file__))
ANCHORS = json.load(open(os.path.join(HERE, "golden", "anchors.json")))
VEC = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))


def test_gll_nodes_and_quadrature(oracle):
    for p, nodes in ANCHORS["gll_nodes"].items():
        assert np.allclose(oracle.gauss_lobatto(int(p) + 1), nodes, atol=5e-8)
    for n in range(1, 11):
        x, w = oracle.gauss_legendre(n)
        assert abs(w.sum() - 1.0) < 1e-14
        for k in range(2 * n):  # exact for degree 2n-1 on [0,1]
            assert abs((w * x ** k).sum() - 1.0 / (k + 1)) < 1e-13


@pytest.mark.parametrize("key", sorted(ANCHORS["discrete_3d"]))
def test_known_answers_3d(key, oracle):
    p, n = map(int, key.split(","))
    N, bnorm, unorm = ANCHORS["discrete_3d"][key]
    mf = oracle.MatrixFree(3, p, n)
    assert mf.n_dofs == N
    b = mf.assemble_rhs()
    assert abs(np.linalg.norm(b) - bnorm) < 2e-10
    x, it, hist, rc = oracle.cg_solve(mf, b, None, rel_tol=1e-13)
    assert rc == 0
    assert abs(mf.l2_norm_solution(x) - unorm) < 2e-10
    assert abs(mf.l2_norm_solution(x) - ANCHORS["analytic"]["3d"]) < 5e-4


@pytest.mark.parametrize("key", sorted(ANCHORS["discrete_2d"]))
def test_known_answers_2d(key, oracle):
    p, n = map(int, key.split(","))
    N, bnorm, unorm = ANCHORS["discrete_2d"][key]
    mf = oracle.MatrixFree(2, p, n)
    assert mf.n_dofs == N
    b = mf.assemble_rhs()
    assert abs(np.linalg.norm(b) - bnorm) < 2e-10
    x, it, hist, rc = oracle.cg_solve(mf, b, None, rel_tol=1e-13)
    assert rc == 0 and abs(mf.l2_norm_solution(x) - unorm) < 2e-10


def test_committed_vectors(oracle):
    for p, n in [(1, (3, 2, 2)), (2, (3, 2, 2)), (4, (2, 2, 1))]:
        mf = oracle.MatrixFree(3, p, n)
        assert rel_l2(mf.vmult(splitmix_src(mf.n_dofs)), VEC["vmult_p%d" % p]) < 1e-14
        assert rel_l2(mf.compute_diagonal(), VEC["dinv_p%d" % p]) < 1e-14


@pytest.mark.parametrize("dim,p,n", [(3, 1, (3, 4, 2)), (3, 2, (3, 2, 2)), (3, 3, (2, 2, 3)), (3, 5, (2, 1, 2)), (2, 4, (3, 5)), (2, 7, (2, 3))])
def test_operator_invariants(dim, p, n, oracle):
    mf = oracle.MatrixFree(dim, p, n)
    rng = np.random.default_rng(p)
    c = mf.constrained()
    u, v = rng.standard_normal(mf.n_dofs), rng.standard_normal(mf.n_dofs)
    Au_full = mf.vmult(u)
    assert np.array_equal(Au_full[c], u[c])  # identity on constrained dofs
    u[c] = 0
    v[c] = 0
    Au, Av = mf.vmult(u), mf.vmult(v)
    assert abs(u @ Av - v @ Au) < 1e-11 * np.linalg.norm(u) * np.linalg.norm(Av)  # symmetry
    assert u @ Au > 0  # positive definite on the free dofs
    # diagonal == e_i^T A e_i on unconstrained dofs; 1 on constrained
    dinv = mf.compute_diagonal()
    for i in rng.choice(np.flatnonzero(~c), size=min(12, (~c).sum()), replace=False):
        e = np.zeros(mf.n_dofs)
        e[i] = 1
        assert abs(mf.vmult(e)[i] * dinv[i] - 1.0) < 1e-12
    assert np.all(dinv[c] == 1.0)


def test_neumann_operator_annihilates_constants(oracle):
    mf = oracle.MatrixFree(3, 3, (3, 2, 2), faces=0)
    assert np.abs(mf.vmult(np.ones(mf.n_dofs))).max() < 1e-12


def test_variable_coefficient_operator_is_spd(oracle):
    mf = oracle.MatrixFree(3, 2, (3, 3, 3), coef="c5")
    rng = np.random.default_rng(0)
    c = mf.constrained()
    u, v = rng.standard_normal(mf.n_dofs), rng.standard_normal(mf.n_dofs)
    u[c] = 0
    v[c] = 0
    assert abs(u @ mf.vmult(v) - v @ mf.vmult(u)) < 1e-11 * np.linalg.norm(u) * np.linalg.norm(v)
    assert u @ mf.vmult(u) > 0


@pytest.mark.parametrize("kind,dim,pc,pf,n", [("h", 3, 2, 2, (2, 3, 2)), ("h", 3, 4, 4, (1, 2, 2)), ("h", 2, 3, 3, (3, 2)),
                                              ("p", 3, 1, 2, (3, 2, 2)), ("p", 3, 2, 4, (2, 2, 3)), ("p", 2, 3, 7, (2, 3)), ("p", 3, 6, 7, (2, 1, 2))])
def test_transfer_invariants(kind, dim, pc, pf, n, oracle):
    nf = tuple(2 * c for c in n) if kind == "h" else n
    mc, mf = oracle.MatrixFree(dim, pc, n), oracle.MatrixFree(dim, pf, nf)
    t = oracle.Transfer(mc, mf, kind)
    rng = np.random.default_rng(1)
    xc, rf = rng.standard_normal(mc.n_dofs), rng.standard_normal(mf.n_dofs)
    xc[mc.constrained()] = 0
    Pc = t.prolongate_and_add(np.zeros(mf.n_dofs), xc)
    Rr = t.restrict_and_add(np.zeros(mc.n_dofs), rf)
    assert abs(rf @ Pc - Rr @ xc) < 1e-12 * np.linalg.norm(rf) * np.linalg.norm(Pc)  # restriction = prolongation^T
    assert np.all(Pc[mf.constrained()] == 0) and np.all(Rr[mc.constrained()] == 0)
    # a polynomial of degree <= pc that vanishes on the boundary is prolongated exactly
    def nodes(m):
        g = oracle.gauss_lobatto(m.p + 1)
        axes = [np.concatenate([(c + g[:-1]) / m.ncell[d] for c in range(m.ncell[d])] + [[1.0]]) for d in range(dim)]
        return np.meshgrid(*axes[::-1], indexing="ij")[::-1]
    def f(X):
        v = np.ones_like(X[0])
        for x in X:
            v = v * x * (1 - x)
        return v.ravel() if pc >= 2 else None
    fc = f(nodes(mc))
    if fc is not None:
        assert rel_l2(t.prolongate_and_add(np.zeros(mf.n_dofs), fc), f(nodes(mf))) < 1e-12


def test_tridiagonal_eigenvalues(oracle):
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 17, 60):
        d, e = rng.standard_normal(n) + 3, rng.standard_normal(max(n - 1, 0))
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        assert np.allclose(oracle.tridiag_eigenvalues(d, e), np.linalg.eigvalsh(T), atol=1e-12)


def test_chebyshev_is_the_chebyshev_polynomial(oracle):
    """With A = I on constrained dofs only (a 1-cell Q1 mesh: every dof constrained) the smoother must return
    p_k(1) * src where p_k is the degree-k Chebyshev iteration polynomial for [alpha, lambda_max]."""
    mf = oracle.MatrixFree(3, 1, 1)
    src = np.arange(1.0, 9.0)
    dst, info = oracle.chebyshev_vmult(mf, src, degree=5)
    lmax, theta, delta = info["lambda_max"], info["theta"], info["delta"]
    assert abs(lmax - 1.2) < 1e-14 and info["cg_iterations"] == 1
    x, xold = src / theta, np.zeros(8)
    rho, sigma = delta / theta, theta / delta
    for k in range(4):
        rhon = 1.0 / (2 * sigma - rho)
        x, xold = x + rhon * rho * (x - xold) + 2 * rhon / delta * (src - x), x
        rho = rhon
    assert rel_l2(dst, x) < 1e-14


@pytest.mark.parametrize("kind,p,n", [("h", 2, 8), ("hp", 4, 4)])
def test_vcycle_cg_regression_and_symmetry(kind, p, n, oracle):
    levels = hierarchy_levels(kind, p, n)
    mfs = [oracle.MatrixFree(3, q, c) for (q, c) in levels]
    trs = [oracle.Transfer(mfs[l - 1], mfs[l], "h" if levels[l][0] == levels[l - 1][0] else "p") for l in range(1, len(levels))]
    vc = oracle.VCycle(mfs, trs)
    b = mfs[-1].assemble_rhs()
    x, it, hist, rc = oracle.cg_solve(mfs[-1], b, vc)
    gold = VEC["cg_%s_p%d_n%d_hist" % (kind, p, n)]
    assert rc == 0 and it == len(gold) - 1
    assert np.all(np.abs(hist - gold) <= 1e-10 * gold[0])
    est = np.array([[e[0], e[1], e[2], e[3]] for e in vc.estimate()])
    assert np.allclose(est, VEC["cg_%s_p%d_n%d_est" % (kind, p, n)], rtol=1e-9)
    assert hist[-1] <= 1e-12 * hist[0] and it <= 6
    assert abs(mfs[-1].l2_norm_solution(x) - ANCHORS["analytic"]["3d"]) < 2e-4
    # V(2,2) with a symmetric smoother is a symmetric preconditioner
    rng = np.random.default_rng(5)
    c = mfs[-1].constrained()
    u, v = rng.standard_normal(mfs[-1].n_dofs), rng.standard_normal(mfs[-1].n_dofs)
    u[c] = 0
    v[c] = 0
    assert abs(u @ vc.vmult(v) - v @ vc.vmult(u)) < 1e-10 * np.linalg.norm(u) * np.linalg.norm(v)
