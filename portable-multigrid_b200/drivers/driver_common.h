/* driver_common.h -- shared pieces of the C drivers (re-creations of the reference's program.cc files). */
#ifndef DRIVER_COMMON_H
#define DRIVER_COMMON_H
#include <pmg.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CK(call)                                                                      \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != PMG_OK) {                                                              \
      fprintf(stderr, "\n----------------------------------------------------\n"      \
                      "Exception on processing:\n%s -> %d: %s\nAborting!\n"           \
                      "----------------------------------------------------\n",       \
              #call, rc_, pmg_last_error());                                          \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

#define MAXL 32

typedef struct { int degree, n; } level_t;

static double now_s(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

static int arg_int(int argc, char **argv, const char *name, int def)
{
  for (int i = 1; i + 1 < argc; ++i)
    if (!strcmp(argv[i], name)) return atoi(argv[i + 1]);
  return def;
}

/* builds operators, transfers, smoothers (program.cc:203-287) and solves (:336-364); prints the reference's lines */
static int solve_hierarchy(pmg_context *ctx, const level_t *lv, int L, int pre, int post, int cheb_degree)
{
  pmg_operator *ops[MAXL];
  pmg_transfer *tr[MAXL];
  pmg_chebyshev *sm[MAXL];
  memset(tr, 0, sizeof(tr));
  for (int l = 0; l < L; ++l)
    CK(pmg_laplace_operator_create(ctx, 3, lv[l].degree, lv[l].n, lv[l].n, lv[l].n, PMG_ALL_FACES, 0, &ops[l]));
  for (int l = 1; l < L; ++l) {
    if (lv[l].degree == lv[l - 1].degree) CK(pmg_transfer_create_geometric(ops[l - 1], ops[l], &tr[l]));
    else CK(pmg_transfer_create_polynomial(ops[l - 1], ops[l], &tr[l]));
  }
  for (int l = 0; l < L; ++l) {
    int64_t m = 0;
    CK(pmg_laplace_operator_m(ops[l], &m));
    CK(pmg_laplace_operator_compute_diagonal(ops[l]));
    if (l > 0) CK(pmg_chebyshev_create(ops[l], 15.0, cheb_degree, 10, &sm[l]));
    else CK(pmg_chebyshev_create(ops[l], 1e-3, PMG_INVALID_DEGREE, (int)(m > 2000000000 ? 2000000000 : m), &sm[l]));
  }
  pmg_vcycle *mg;
  CK(pmg_vcycle_create(ops, tr, sm, L, pre, post, &mg));
  pmg_operator *A = ops[L - 1];
  pmg_vector *rhs, *x;
  CK(pmg_laplace_operator_initialize_dof_vector(A, &rhs));
  CK(pmg_laplace_operator_initialize_dof_vector(A, &x));
  CK(pmg_laplace_operator_assemble_rhs(A, rhs));
  double bnorm = 0.0;
  CK(pmg_vector_l2_norm(rhs, &bnorm));
  int64_t n_dofs = 0;
  CK(pmg_laplace_operator_m(A, &n_dofs));
  /* set-up (eigenvalue estimates) outside the timed solve, like the lazy first vmult of the reference */
  for (int l = 0; l < L; ++l) CK(pmg_chebyshev_info(sm[l], NULL, NULL, NULL, NULL));
  int last_step = 0;
  CK(pmg_sync(ctx));
  const double t0 = now_s();
  CK(pmg_cg_solve(A, x, rhs, mg, (int)(n_dofs > 100000 ? 100000 : n_dofs), 1e-12 * bnorm, &last_step, NULL, 0));
  CK(pmg_sync(ctx));
  const double dt = now_s() - t0;
  printf("  Solver converged in %d iterations.\n", last_step);
  double norm = 0.0;
  CK(pmg_laplace_operator_solution_norm(A, x, &norm));
  printf("  solution norm: %.10g\n", norm);
  printf("  [b200] solve time %.3f ms, %.3f GDoF/s (DoFs x iterations / time)\n", dt * 1e3, (double)n_dofs * last_step / dt / 1e9);
  pmg_vector_destroy(rhs); pmg_vector_destroy(x);
  pmg_vcycle_destroy(mg);
  for (int l = 0; l < L; ++l) { pmg_chebyshev_destroy(sm[l]); if (tr[l]) pmg_transfer_destroy(tr[l]); }
  for (int l = 0; l < L; ++l) pmg_laplace_operator_destroy(ops[l]);
  return 0;
}
#endif
