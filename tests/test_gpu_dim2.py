"""GPU parity of the 2-D path (the reference's polynomial_multigrid driver is dim = 2, fe_degree = 7, seven p-levels,
source/polynomial_multigrid/program.cc:439-441) through the C-ABI against the CPU oracle.  Tolerances as in test_gpu_parity.py;
the analytic anchors are SURVEY.md 8c's 2-D values."""
import json
import os

import numpy as np
import pytest

from helpers import rel_l2, splitmix_src

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_2d_vmult_diagonal_and_fused_steps(p, pmg, ctx, oracle):
    n = (37, 21) if p < 4 else (9, 14)
    mf = oracle.MatrixFree(2, p, n)
    u, b, xo = (splitmix_src(mf.n_dofs, salt=s) for s in (51, 52, 53))
    Au, dinv = mf.vmult(u), mf.compute_diagonal()
    op = pmg.LaplaceOperator(ctx, p, n, dim=2)
    assert op.m() == mf.n_dofs
    s, d = op.vector_from(u), op.initialize_dof_vector()
    op.vmult(d, s)
    out = d.export_host()
    assert rel_l2(out, Au) <= 1e-12
    c = mf.constrained()
    assert np.array_equal(out[c], u[c])
    op.compute_diagonal()
    assert rel_l2(op.get_matrix_diagonal_inverse().export_host(), dinv) <= 1e-12
    vb, vx = op.vector_from(b), op.vector_from(xo)
    op.residual(d, vb, s)
    assert rel_l2(d.export_host(), b - Au) <= 1e-12
    op.chebyshev_step(vx, s, vx, vb, 0.3, 0.7)
    assert rel_l2(vx.export_host(), u + 0.3 * (u - xo) + 0.7 * dinv * (b - Au)) <= 1e-12


@pytest.mark.parametrize("kind,pc,pf,n", [("h", 2, 2, (9, 6)), ("h", 7, 7, (3, 4)), ("p", 6, 7, (8, 5)), ("p", 1, 2, (16, 9)), ("p", 2, 4, (7, 7))])
def test_2d_transfers(kind, pc, pf, n, pmg, ctx, oracle):
    nf = tuple(2 * c for c in n) if kind == "h" else n
    mc, mf = oracle.MatrixFree(2, pc, n), oracle.MatrixFree(2, pf, nf)
    t_ref = oracle.Transfer(mc, mf, kind)
    oc, of = pmg.LaplaceOperator(ctx, pc, n, dim=2), pmg.LaplaceOperator(ctx, pf, nf, dim=2)
    t = pmg.GeometricTransfer(oc, of) if kind == "h" else pmg.PolynomialTransfer(oc, of)
    xc, rf = splitmix_src(mc.n_dofs, mc.constrained(), salt=54), splitmix_src(mf.n_dofs, salt=55)
    d0, c0 = splitmix_src(mf.n_dofs, salt=56), splitmix_src(mc.n_dofs, salt=57)
    vd, vc = of.vector_from(d0), oc.vector_from(c0)
    t.prolongate_and_add(vd, oc.vector_from(xc))
    t.restrict_and_add(vc, of.vector_from(rf))
    assert rel_l2(vd.export_host(), t_ref.prolongate_and_add(d0.copy(), xc)) <= 1e-13
    assert rel_l2(vc.export_host(), t_ref.restrict_and_add(c0.copy(), rf)) <= 1e-13


@pytest.mark.parametrize("fe_degree,n", [(7, 4), (3, 8), (4, 16)])
def test_2d_polynomial_multigrid_driver_hierarchy(fe_degree, n, pmg, ctx, oracle):
    """The reference driver's hierarchy: p = 1 .. fe_degree on one mesh, p -> p+1 transfers (program.cc:144-158, 241-242),
    V(2,2), Chebyshev(5), CG to 1e-12 ||b||: same smoother set-up, iteration count and residual history as the oracle."""
    levels = [(q, (n, n)) for q in range(1, fe_degree + 1)]
    mfs = [oracle.MatrixFree(2, q, m) for (q, m) in levels]
    trs = [oracle.Transfer(mfs[l - 1], mfs[l], "p") for l in range(1, len(levels))]
    vc_ref = oracle.VCycle(mfs, trs)
    ops, transfers, smoothers, mg = pmg.build_hierarchy(ctx, levels, dim=2)
    top = ops[-1]
    est = vc_ref.estimate()
    for l, sm in enumerate(smoothers):
        info = sm.info()
        assert info["degree"] == est[l][2], (l, info, est[l])
        assert info["cg_iterations"] == est[l][3]
        assert abs(info["lambda_max"] - est[l][1]) <= 1e-8 * est[l][1]
    r = splitmix_src(mfs[-1].n_dofs, mfs[-1].constrained(), salt=58)
    z_ref = vc_ref.vmult(r)
    dr, dz = top.vector_from(r), top.initialize_dof_vector()
    for rep in range(3):
        mg.vmult(dz, dr)
        assert rel_l2(dz.export_host(), z_ref) <= 1e-10, rep
    b_ref = mfs[-1].assemble_rhs()
    b = top.initialize_dof_vector()
    top.assemble_rhs(b)
    assert rel_l2(b.export_host(), b_ref) <= 1e-14
    x_ref, it_ref, hist_ref, rc_ref = oracle.cg_solve(mfs[-1], b_ref, vc_ref)
    x = top.initialize_dof_vector()
    it, hist, rc = pmg.cg_solve(top, x, b, mg)
    assert rc == 0 and rc_ref == 0 and it == it_ref
    assert np.all(np.abs(hist - hist_ref) <= 1e-10 * hist_ref[0])
    norm = top.solution_norm(x)
    assert abs(norm - mfs[-1].l2_norm_solution(x_ref)) <= 1e-10
    assert abs(norm - 0.0412614896) < 2e-4  # analytic ||u||_L2 of -Laplace u = 1 on the unit square


def test_2d_known_answer_anchors(pmg, ctx):
    """||b||_2 and ||u_h||_L2 of the exact FE_Q / Gauss discretisation (tests/golden/anchors.json)."""
    anchors = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "anchors.json")))
    assert anchors["discrete_2d"]
    for key, (n_dofs, b_norm, u_norm) in anchors["discrete_2d"].items():
        p, n = (int(v) for v in key.split(","))
        levels = [(p, (m, m)) for m in (n // 2, n)]
        ops, transfers, smoothers, mg = pmg.build_hierarchy(ctx, levels, dim=2)
        top = ops[-1]
        assert top.m() == n_dofs
        b, x = top.initialize_dof_vector(), top.initialize_dof_vector()
        top.assemble_rhs(b)
        assert abs(b.l2_norm() - b_norm) <= 1e-9
        it, hist, rc = pmg.cg_solve(top, x, b, mg)
        assert rc == 0
        assert abs(top.solution_norm(x) - u_norm) <= 2e-9
