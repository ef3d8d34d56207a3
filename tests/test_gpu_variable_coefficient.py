"""GPU parity of the variable-coefficient operator -div(a grad u), a = 1/(0.05 + 2|x|^2) (BASELINE config 5), through the
C-ABI against the CPU oracle (the reference has no coefficient: the oracle multiplies a(x_q) into JxW).
Tolerances as in test_gpu_parity.py."""
import numpy as np
import pytest

from helpers import hierarchy_levels, rel_l2, splitmix_src

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_variable_coefficient_vmult_diagonal_and_fused_steps(p, pmg, ctx, oracle):
    n = {1: (17, 16, 18), 2: (13, 14, 9), 3: (11, 10, 7), 4: (9, 8, 5)}.get(p, (5, 6, 4))
    mf = oracle.MatrixFree(3, p, n, coef="c5")
    u, b, xo = (splitmix_src(mf.n_dofs, salt=s) for s in (31, 32, 33))
    Au = mf.vmult(u)
    op = pmg.LaplaceOperator(ctx, p, n, coefficient=1)
    s, d = op.vector_from(u), op.initialize_dof_vector()
    op.vmult(d, s)
    out = d.export_host()
    assert rel_l2(out, Au) <= 1e-12
    c = mf.constrained()
    assert np.array_equal(out[c], u[c])
    dinv = mf.compute_diagonal()
    op.compute_diagonal()
    assert rel_l2(op.get_matrix_diagonal_inverse().export_host(), dinv) <= 1e-12
    vb, vx = op.vector_from(b), op.vector_from(xo)
    op.residual(d, vb, s)
    assert rel_l2(d.export_host(), b - Au) <= 1e-12
    op.chebyshev_step(d, s, None, vb, 0.0, 0.7)
    assert rel_l2(d.export_host(), u + 0.7 * dinv * (b - Au)) <= 1e-12
    op.chebyshev_step(vx, s, vx, vb, 0.3, 0.7)  # x_old overwritten in place
    assert rel_l2(vx.export_host(), u + 0.3 * (u - xo) + 0.7 * dinv * (b - Au)) <= 1e-12


@pytest.mark.parametrize("faces", [0x00, 0x15, 0x2A])
def test_variable_coefficient_mixed_boundary(faces, pmg, ctx, oracle):
    p, n = 3, (6, 5, 7)
    mf = oracle.MatrixFree(3, p, n, faces=faces, coef="c5")
    u = splitmix_src(mf.n_dofs, salt=34)
    op = pmg.LaplaceOperator(ctx, p, n, dirichlet_faces=faces, coefficient=1)
    s, d = op.vector_from(u), op.initialize_dof_vector()
    op.vmult(d, s)
    assert rel_l2(d.export_host(), mf.vmult(u)) <= 1e-12


def test_variable_coefficient_large_is_symmetric(pmg, ctx):
    """Size-independent property at a size the oracle would not finish: <A u, v> = <u, A v> (Q5, 24^3 cells, 1.8 M DoFs)."""
    op = pmg.LaplaceOperator(ctx, 5, 24, coefficient=1)
    m = op.m()
    u, v = op.vector_from(splitmix_src(m, salt=35)), op.vector_from(splitmix_src(m, salt=36))
    Au, Av = op.initialize_dof_vector(), op.initialize_dof_vector()
    op.vmult(Au, u)
    op.vmult(Av, v)
    a, b = Au.dot(v), Av.dot(u)
    assert abs(a - b) <= 1e-12 * max(abs(a), abs(b))


@pytest.mark.parametrize("kind,p,n", [("h", 2, 8), ("hp", 4, 8), ("hp", 5, 4)])
def test_variable_coefficient_vcycle_and_cg(kind, p, n, pmg, ctx, oracle):
    levels = hierarchy_levels(kind, p, n)
    mfs = [oracle.MatrixFree(3, q, m, coef="c5") for (q, m) in levels]
    trs = [oracle.Transfer(mfs[l - 1], mfs[l], "h" if levels[l][0] == levels[l - 1][0] else "p") for l in range(1, len(levels))]
    vc_ref = oracle.VCycle(mfs, trs)
    ops, transfers, smoothers, mg = pmg.build_hierarchy(ctx, levels, coefficient=1)
    top = ops[-1]
    est = vc_ref.estimate()
    for l, sm in enumerate(smoothers):
        info = sm.info()
        assert info["degree"] == est[l][2], (l, info, est[l])
        assert info["cg_iterations"] == est[l][3]
        assert abs(info["lambda_max"] - est[l][1]) <= 1e-8 * est[l][1]
    r = splitmix_src(mfs[-1].n_dofs, mfs[-1].constrained(), salt=37)
    z_ref = vc_ref.vmult(r)
    dr, dz = top.vector_from(r), top.initialize_dof_vector()
    for rep in range(3):  # eager warm-up, graph capture, graph replay
        mg.vmult(dz, dr)
        assert rel_l2(dz.export_host(), z_ref) <= 1e-10, rep
    # load vector of f = 1: the oracle's JxW carries a(x), so the right-hand side comes from the constant-coefficient level
    b_ref = oracle.MatrixFree(3, *levels[-1]).assemble_rhs()
    b = top.initialize_dof_vector()
    top.assemble_rhs(b)
    assert rel_l2(b.export_host(), b_ref) <= 1e-14
    x_ref, it_ref, hist_ref, rc_ref = oracle.cg_solve(mfs[-1], b_ref, vc_ref)
    x = top.initialize_dof_vector()
    it, hist, rc = pmg.cg_solve(top, x, b, mg)
    assert rc == 0 and rc_ref == 0
    assert it == it_ref
    assert np.all(np.abs(hist - hist_ref) <= 1e-10 * hist_ref[0])
    assert rel_l2(x.export_host(), x_ref) <= 1e-9
    assert abs(top.solution_norm(x) - mfs[-1].l2_norm_solution(x_ref)) <= 1e-10
