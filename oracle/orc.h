/*
 * oracle/orc.h -- CPU restatement of dealii-X/portable-multigrid's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product library (libpmg.so,
 * portable-multigrid_b200/) links, imports or calls this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker or the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or recorded
 * outputs, and cannot be built here (needs deal.II >= 9.8.0, Kokkos, MPI,
 * p4est -- source/geometric_multigrid/CMakeLists.txt:32).  The oracle is
 * therefore anchored on (i) the reference's own algorithm, restated function
 * by function with file:line citations below, in the reference's data layout
 * (stored local_to_global, Dirichlet mask, per-q-point inv_jacobian / JxW,
 * the same 12 one-dimensional sweeps, scatter-add), (ii) the published
 * deal.II algorithms for the un-vendored pieces (FE_Q on Gauss-Lobatto nodes,
 * QGauss, PreconditionChebyshev, SolverCG; deal.II >= 9.8.0 is the only pin),
 * and (iii) the analytic known answers in tests/golden/anchors.json.
 *
 * All arrays use the reference's Kokkos LayoutLeft convention (first index
 * fastest): local_to_global(i,cell) -> [i + n_loc*cell], JxW(q,cell) ->
 * [q + n_q*cell], inv_jacobian(q,cell,d,e) -> [q + n_q*(cell + n_cells*(d + dim*e))].
 */
#ifndef ORC_H
#define ORC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INVALID 0xFFFFFFFFu /* numbers::invalid_unsigned_int */
#define ORC_MAX_DEGREE 9        /* OperatorDispatchFactory::max_degree, portable_laplace_operator_base.h:65 */

/* ---- 1-D finite element tables (fe1d.c) -------------------------------- */
void orc_gauss_legendre(int n, double *x, double *w); /* QGauss<1>(n) on [0,1] */
void orc_gauss_lobatto(int n, double *x);             /* FE_Q(n-1) support points on [0,1] */
void orc_lagrange(int n, const double *nodes, double x, double *val, double *der);
/* shape_values[q*(p+1)+i] = phi_i(g_q); co_shape_gradients[q*(p+1)+r] = L_r'(g_q), L = Lagrange basis on the Gauss points */
void orc_shape_tables(int p, double *shape_values, double *co_shape_gradients, double *gauss_w);
/* h-transfer 1-D matrix, stored [i*(2p+1) + j + c*p] (portable_geometric_transfer.h:1307-1314) */
void orc_h_prolongation_1d(int p, double *P);
/* p-transfer 1-D matrix, stored [i_coarse*(pf+1) + j_fine] (portable_polynomial_tranfer.h:957-976) */
void orc_p_prolongation_1d(int pc, int pf, double *P);

/* ---- matrix-free data of one level (mesh.c) ---------------------------- */
typedef struct orc_mf {
  int dim, p;
  int n[3];            /* cells per direction (n[2]=1 in 2-D) */
  int nd[3];           /* dofs per direction n*p+1 */
  double h[3];         /* cell size */
  unsigned faces;      /* Dirichlet faces bitmask: bit 2d = low face of dir d, 2d+1 = high face */
  int64_t n_cells, n_dofs;
  int n_loc, n_q;      /* (p+1)^dim */
  uint32_t *local_to_global; /* (i,cell) */
  uint32_t *mask;            /* (i,cell): ORC_INVALID on constrained dofs, else global index */
  double *inv_jacobian;      /* (q,cell,d,e) */
  double *JxW;               /* (q,cell) */
  uint8_t *constrained;      /* per global dof */
  double shape_values[(ORC_MAX_DEGREE + 1) * (ORC_MAX_DEGREE + 1)];
  double co_shape_gradients[(ORC_MAX_DEGREE + 1) * (ORC_MAX_DEGREE + 1)];
  double gauss_w[ORC_MAX_DEGREE + 1];
  /* cell colouring (parity colouring of the structured grid; plays the role of
     MatrixFree's colored graph with use_coloring=true) */
  int n_colors;
  int64_t *color_start; /* n_colors+1 */
  int64_t *color_cells; /* cell ids grouped by colour */
  double *inv_diag;     /* filled by orc_compute_diagonal */
} orc_mf;

/* coef: optional callback a(x,y,z) evaluated at quadrature points and multiplied into JxW
   (variable-coefficient extension; NULL = reference behaviour) */
typedef double (*orc_coef_fn)(double x, double y, double z);
orc_mf *orc_mf_create(int dim, int p, int nx, int ny, int nz, unsigned dirichlet_faces, orc_coef_fn coef);
void orc_mf_destroy(orc_mf *mf);
double orc_coef_c5(double x, double y, double z); /* a = 1/(0.05 + 2|x|^2) */

/* ---- Laplace operator (laplace.c) -------------------------------------- */
/* LaplaceOperator::vmult, portable_laplace_operator.h:557-719 (body :227-381) */
void orc_vmult(const orc_mf *mf, double *dst, const double *src);
/* LaplaceOperator::compute_diagonal, :752-917 (body :71-210); stores mf->inv_diag */
void orc_compute_diagonal(orc_mf *mf);
/* one cell of LocalLaplaceOperator::operator(): values[n_loc] in place; scratch >= 4*n_q */
void orc_cell_apply(const orc_mf *mf, int64_t cell, double *values, double *scratch);

/* ---- transfers (transfer.c) -------------------------------------------- */
typedef struct orc_transfer {
  int kind; /* 0 = geometric (h), 1 = polynomial (p) */
  int dim, pc, pf;
  int nc1, nf1;  /* 1-D sizes: h: p+1, 2p+1 ; p: pc+1, pf+1 */
  int64_t n_cells; /* coarse cells (h) or cells (p) */
  int n_loc_c, n_loc_f;
  uint32_t *idx_coarse; /* (i,cell); h: ORC_INVALID on constrained coarse dofs (:1453) */
  uint32_t *idx_fine;   /* (i,cell) */
  uint32_t *mask_coarse, *mask_fine; /* p-transfer only (:1175-1267) */
  double *weights;      /* (i,cell) */
  double *P;            /* nc1 x nf1, row = coarse node */
  int64_t n_dofs_c, n_dofs_f;
  const orc_mf *coarse, *fine;
} orc_transfer;

orc_transfer *orc_transfer_create_h(const orc_mf *coarse, const orc_mf *fine);
orc_transfer *orc_transfer_create_p(const orc_mf *coarse, const orc_mf *fine);
void orc_transfer_destroy(orc_transfer *t);
void orc_prolongate_and_add(const orc_transfer *t, double *dst_fine, const double *src_coarse);
void orc_restrict_and_add(const orc_transfer *t, double *dst_coarse, const double *src_fine);

/* ---- smoother / V-cycle / CG (solver.c) -------------------------------- */
typedef struct orc_chebyshev {
  const orc_mf *mf;
  double smoothing_range;
  int degree;             /* -1 = numbers::invalid_unsigned_int (auto) */
  int eig_cg_n_iterations;
  int initialized;
  double lambda_min, lambda_max, theta, delta;
  int cg_iterations;
} orc_chebyshev;

void orc_chebyshev_init(orc_chebyshev *c, const orc_mf *mf, double smoothing_range, int degree, int eig_cg_n_iterations);
void orc_chebyshev_estimate(orc_chebyshev *c);
void orc_chebyshev_vmult(orc_chebyshev *c, double *dst, const double *src);

typedef struct orc_vcycle {
  int n_levels;
  orc_mf **mf;              /* [n_levels] level 0 = coarsest */
  orc_transfer **transfer;  /* [n_levels]; transfer[l] maps l-1 <-> l, transfer[0]=NULL */
  orc_chebyshev *smoother;  /* [n_levels] */
  int pre, post;
} orc_vcycle;

void orc_vcycle_vmult(orc_vcycle *v, double *dst, const double *src);

/* SolverCG with SolverControl(max_it, tol): returns 0 on convergence; history has last_step+1 entries (res norms) */
int orc_cg_solve(const orc_mf *A, double *x, const double *b, orc_vcycle *precond, int max_it, double tol,
                 int *last_step, double *history, int history_cap);

/* symmetric tridiagonal eigenvalues (ascending), used by the Lanczos estimate */
void orc_tridiag_eigenvalues(int n, const double *diag, const double *offdiag, double *eig);

/* assemble_rhs (f = 1), source/geometric_multigrid/program.cc:289-334 */
void orc_assemble_rhs(const orc_mf *mf, double *rhs);
/* "solution norm" with QGauss(p+2), program.cc:382-395 */
double orc_l2_norm_solution(const orc_mf *mf, const double *u);

/* BLAS-1 helpers (a11) */
double orc_dot(int64_t n, const double *a, const double *b);
double orc_l2(int64_t n, const double *a);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
