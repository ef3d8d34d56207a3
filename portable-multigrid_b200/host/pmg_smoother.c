/*
 * pmg_smoother.c -- Chebyshev-Jacobi smoother of the C-ABI (host C).
 *
 * Stands in for deal.II's PreconditionChebyshev<LaplaceOperatorBase, Vector> with the
 * DiagonalMatrix inner preconditioner, as configured by the reference drivers
 * (source/geometric_multigrid/program.cc:267-285) and called from VCycleMultigrid::smooth
 * (include/multigrid/portable_v_cycle_multigrid.h:116-125).  Numerical spec: DESIGN.md.
 *
 * B200-first: every Chebyshev step is ONE kernel (operator apply + Jacobi scaling + three-term
 * update, csrc/pmg_apply_tile.h epilogue); the reference runs an apply kernel that writes A x to
 * HBM followed by an unfused vector update that re-reads it.  The inverse diagonal comes from a
 * (p+2)^3 table, not from a vector stream.  smooth(u, rhs) is evaluated as the Chebyshev iteration
 * with initial guess u (algebraically identical to "r = rhs - A u; d = Cheb(r); u += d", one apply
 * per step, no residual / correction vectors), so a smoothing step of degree k costs
 * k fused passes of 32 B/DoF instead of k applies + (3k+2) vector passes.
 */
#include "pmg_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

int pmg_chebyshev_create(pmg_operator *op, double smoothing_range, int degree, int eig_cg_n_iterations, pmg_chebyshev **out)
{
  if (!op || !out || eig_cg_n_iterations < 0 || (degree < 1 && degree != PMG_INVALID_DEGREE) || smoothing_range <= 0.0) {
    pmg_set_error("chebyshev_create: bad arguments");
    return PMG_ERR_ARG;
  }
  pmg_chebyshev *s = (pmg_chebyshev *)calloc(1, sizeof(*s));
  if (!s) return PMG_ERR_NOMEM;
  s->op = op; s->smoothing_range = smoothing_range; s->degree = degree; s->eig_cg_n_iterations = eig_cg_n_iterations;
  PMG_CHECK(pmg_vector_create_layout(op->ctx, &op->lay, &s->t0));
  PMG_CHECK(pmg_vector_create_layout(op->ctx, &op->lay, &s->t1));
  *out = s;
  return PMG_OK;
}

int pmg_chebyshev_destroy(pmg_chebyshev *s)
{
  if (!s) return PMG_OK;
  pmg_vector_destroy(s->t0);
  pmg_vector_destroy(s->t1);
  free(s);
  return PMG_OK;
}

/* lmax: the largest eigenvalue as the smoother uses it (the estimate times its safety factor, or the configured value) */
static int chebyshev_parameters(double lmin, double lmax, double smoothing_range, int degree_in, double *theta, double *delta,
                                int *degree_out);

int pmg_host_chebyshev_parameters(double lmin, double lmax_est, double smoothing_range, int degree_in,
                                  double *theta, double *delta, int *degree_out)
{
  if (!theta || !delta || !degree_out) return PMG_ERR_ARG;
  return chebyshev_parameters(lmin, 1.2 * lmax_est /* safety factor: CG is in general not converged */, smoothing_range, degree_in,
                              theta, delta, degree_out);
}

static int chebyshev_parameters(double lmin, double lmax, double smoothing_range, int degree_in, double *theta, double *delta,
                                int *degree_out)
{
  const double alpha = (smoothing_range > 1.0) ? lmax / smoothing_range : fmin(0.9 * lmax, lmin);
  int degree = degree_in;
  if (degree == PMG_INVALID_DEGREE) {
    /* Chebyshev error estimate (Varga, Matrix Iterative Analysis, sec. 5.1) */
    const double actual_range = lmax / alpha;
    const double sigma = (1.0 - sqrt(1.0 / actual_range)) / (1.0 + sqrt(1.0 / actual_range));
    const double eps = smoothing_range;
    degree = 1 + (int)(log(1.0 / eps + sqrt(1.0 / eps / eps - 1.0)) / log(1.0 / sigma));
  }
  *delta = (lmax - alpha) * 0.5;
  *theta = (lmax + alpha) * 0.5;
  *degree_out = degree;
  return PMG_OK;
}

/* extreme eigenvalues of a symmetric tridiagonal matrix by Sturm-sequence bisection */
static int sturm_count(int n, const double *d, const double *e, double x)
{
  int count = 0;
  double q = 1.0;
  for (int i = 0; i < n; ++i) {
    const double e2 = (i > 0) ? e[i - 1] * e[i - 1] : 0.0;
    q = d[i] - x - ((i > 0) ? e2 / q : 0.0);
    if (q == 0.0) q = 1e-300;
    if (q < 0.0) ++count;
  }
  return count; /* number of eigenvalues < x */
}

static double kth_eigenvalue(int n, const double *d, const double *e, int k, double lo, double hi)
{
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid == lo || mid == hi) break;
    if (sturm_count(n, d, e, mid) > k) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

int pmg_host_tridiag_extreme_eigenvalues(int n, const double *diag, const double *offdiag, double *lmin, double *lmax)
{
  if (n < 1 || !diag || !lmin || !lmax || (n > 1 && !offdiag)) return PMG_ERR_ARG;
  double lo = diag[0], hi = diag[0];
  for (int i = 0; i < n; ++i) { /* Gershgorin bounds */
    const double r = ((i > 0) ? fabs(offdiag[i - 1]) : 0.0) + ((i < n - 1) ? fabs(offdiag[i]) : 0.0);
    if (diag[i] - r < lo) lo = diag[i] - r;
    if (diag[i] + r > hi) hi = diag[i] + r;
  }
  const double pad = 1e-12 * fmax(fabs(lo), fabs(hi)) + 1e-300;
  *lmin = kth_eigenvalue(n, diag, offdiag, 0, lo - pad, hi + pad);
  *lmax = kth_eigenvalue(n, diag, offdiag, n - 1, lo - pad, hi + pad);
  return PMG_OK;
}

/* PreconditionChebyshev::estimate_eigenvalues: Jacobi-preconditioned CG on A x = v from x = 0 with
   v_i = (global_i mod 11) - mean, IterationNumberControl(eig_cg_n_iterations, 1e-10); eigenvalues of the
   Lanczos matrix of iterations 1..it-1. */
/* the Lanczos part: every resource is released on every path */
static int estimate_lanczos(pmg_chebyshev *s, double *lmin, double *lmax)
{
  pmg_operator *op = s->op;
  pmg_context *ctx = op->ctx;
  const pmg_layout *l = &op->lay;
  pmg_vector *r = NULL, *z = NULL, *pv = NULL, *Ap = NULL;
  double *diag = NULL, *off = NULL;
  int rc = PMG_OK, nt = 0, it = 0;
#define EST_CHECK(call) do { rc = (call); if (rc != PMG_OK) goto done; } while (0)
  const int max_it = s->eig_cg_n_iterations;
  /* one entry per iteration; the coarsest level asks for "as many as it takes" (up to 2^31): grow on demand */
  size_t cap = (size_t)(max_it < 4096 ? max_it : 4096) + 2;
  diag = (double *)calloc(cap, sizeof(double));
  off = (double *)calloc(cap, sizeof(double));
  if (!diag || !off) { pmg_set_error("chebyshev estimate: out of host memory"); rc = PMG_ERR_NOMEM; goto done; }
  EST_CHECK(pmg_vector_create_layout(ctx, l, &r));
  EST_CHECK(pmg_vector_create_layout(ctx, l, &z));
  EST_CHECK(pmg_vector_create_layout(ctx, l, &pv));
  EST_CHECK(pmg_vector_create_layout(ctx, l, &Ap));
  /* start vector on all stored planes (global index => identical on any rank count) */
  EST_CHECK(pmgk_set_mod11(r->d, l->plane * l->z0, l->n_local, ctx->stream));
  double mean = 0.0, res = 0.0, rz = 0.0;
  EST_CHECK(pmg_vector_mean_value(r, &mean));
  EST_CHECK(pmgk_set(z->d, -mean, l->n_local, ctx->stream));
  EST_CHECK(pmg_vector_add(r, 1.0, z)); /* r = v - mean */
  EST_CHECK(pmg_vector_l2_norm(r, &res));
  if (res > 1e-10) {
    EST_CHECK(pmgk_scale_dinv(&op->lv, 1.0, r->d, z->d, ctx->stream));
    EST_CHECK(pmg_vector_copy(pv, z));
    EST_CHECK(pmg_vector_dot(r, z, &rz));
    double alpha_prev = 0.0, beta_prev = 0.0, eigen_beta_alpha = 0.0;
    for (;;) {
      ++it;
      EST_CHECK(pmg_laplace_operator_vmult(op, Ap, pv));
      double pAp = 0.0;
      EST_CHECK(pmg_vector_dot(pv, Ap, &pAp));
      const double alpha = rz / pAp;
      EST_CHECK(pmg_vector_add(r, -alpha, Ap));
      EST_CHECK(pmg_vector_l2_norm(r, &res));
      if (it > 1) {
        if ((size_t)nt + 2 > cap) {
          const size_t ncap = 2 * cap;
          double *nd = (double *)realloc(diag, ncap * sizeof(double)), *no = nd ? (double *)realloc(off, ncap * sizeof(double)) : NULL;
          if (nd) diag = nd;
          if (no) off = no;
          if (!nd || !no) { pmg_set_error("chebyshev estimate: out of host memory"); rc = PMG_ERR_NOMEM; goto done; }
          cap = ncap;
        }
        diag[nt] = 1.0 / alpha_prev + eigen_beta_alpha;
        eigen_beta_alpha = beta_prev / alpha_prev;
        off[nt] = sqrt(beta_prev) / alpha_prev;
        ++nt;
      }
      if (res <= 1e-10 || it >= max_it) break;
      EST_CHECK(pmgk_scale_dinv(&op->lv, 1.0, r->d, z->d, ctx->stream));
      double rz_new = 0.0;
      EST_CHECK(pmg_vector_dot(r, z, &rz_new));
      const double beta = rz_new / rz;
      EST_CHECK(pmg_vector_sadd(pv, beta, 1.0, z));
      rz = rz_new;
      alpha_prev = alpha; beta_prev = beta;
    }
  }
  s->cg_iterations = it;
  if (nt > 0) EST_CHECK(pmg_host_tridiag_extreme_eigenvalues(nt, diag, off, lmin, lmax));
done:
#undef EST_CHECK
  free(diag); free(off);
  pmg_vector_destroy(r); pmg_vector_destroy(z); pmg_vector_destroy(pv); pmg_vector_destroy(Ap);
  return rc;
}

int pmg_chebyshev_estimate(pmg_chebyshev *s)
{
  if (s && s->op) PMG_CHECK(pmg_enter(s->op->ctx));
  double lmin = 1.0, lmax = 1.0;
  s->cg_iterations = 0;
  /* no estimate run (eig_cg_n_iterations == 0): the configured largest eigenvalue (1) is used as it is, like deal.II does;
     the 1.2 safety factor belongs to an unconverged CG estimate only */
  const double safety = (s->eig_cg_n_iterations > 0) ? 1.2 : 1.0;
  if (s->eig_cg_n_iterations > 0) PMG_CHECK(estimate_lanczos(s, &lmin, &lmax));
  int degree = s->degree;
  PMG_CHECK(chebyshev_parameters(lmin, safety * lmax, s->smoothing_range, s->degree, &s->theta, &s->delta, &degree));
  s->degree = degree;
  s->lambda_min = lmin;
  s->lambda_max = safety * lmax;
  s->initialized = 1;
  return PMG_OK;
}

int pmg_chebyshev_info(pmg_chebyshev *s, double *lambda_min, double *lambda_max, int *degree, int *cg_iterations)
{
  if (!s) return PMG_ERR_ARG;
  if (!s->initialized) PMG_CHECK(pmg_chebyshev_estimate(s));
  if (lambda_min) *lambda_min = s->lambda_min;
  if (lambda_max) *lambda_max = s->lambda_max;
  if (degree) *degree = s->degree;
  if (cg_iterations) *cg_iterations = s->cg_iterations;
  return PMG_OK;
}

/* One smooth(): k = degree fused passes.  Buffers ping-pong between u and tmp; *result tells the
   caller where the final iterate lives. */
int pmg_chebyshev_smooth(pmg_chebyshev *s, pmg_vector *u, const pmg_vector *rhs, pmg_vector *tmp, int zero_guess,
                         pmg_vector **result)
{
  return pmg_chebyshev_smooth_chain(s, u, rhs, tmp, zero_guess, result, 0, 0, NULL);
}

/* smooth() as a link of a longer chain of applies (the V-cycle: smooth, smooth, residual): chain_in = u's ghost planes were
   pushed by the caller's previous fused apply (the first apply here consumes them, no exchange); want_out = the caller's next
   operation is a fused apply that reads the result (the last apply here pushes too); *pushed_out = it did. */
int pmg_chebyshev_smooth_chain(pmg_chebyshev *s, pmg_vector *u, const pmg_vector *rhs, pmg_vector *tmp, int zero_guess,
                               pmg_vector **result, int chain_in, int want_out, int *pushed_out)
{
  if (pushed_out) *pushed_out = 0;
  if (s && s->op) PMG_CHECK(pmg_enter(s->op->ctx));
  if (!s->initialized) PMG_CHECK(pmg_chebyshev_estimate(s));
  pmg_operator *op = s->op;
  pmg_context *ctx = op->ctx;
  *result = u;
  if (!op->lay.active) return PMG_OK;
  const double theta = s->theta, delta = s->delta;
  pmg_vector *cur = u, *other = tmp;
  int other_is_xold = 0; /* does `other` hold the previous iterate? */
  /* Slabs with neighbours: the applies of one smooth() form a chain u -> tmp -> u -> ...; each pushes the boundary planes
     of its result into the neighbours' ghost planes from its own epilogue (one fused compute + exchange kernel), so only the
     first apply is preceded by an exchange.  n_applies: the fused passes of this call. */
  const int chained = pmg_apply_chain_ok(op, PMGK_CHEB_STEP, u->d, tmp->d);
  const int n_steps = (s->degree >= 2 && fabs(delta) >= 1e-40) ? s->degree - 1 : 0;
  const int n_applies = n_steps + (zero_guess ? 0 : 1);
  int i_apply = 0;
#define PMG_SMOOTH_APPLY(mode, xold_, f1_, f2_)                                                                                   \
  do {                                                                                                                            \
    if (chained) PMG_CHECK(pmg_apply_chained(op, mode, cur->d, rhs->d, xold_, other->d, f1_, f2_, i_apply > 0 || chain_in,        \
                                             i_apply + 1 < n_applies || want_out));                                               \
    else PMG_CHECK(pmg_apply_with_halo(op, mode, cur->d, rhs->d, xold_, other->d, f1_, f2_));                                    \
    ++i_apply;                                                                                                                    \
  } while (0)
  /* step 0: x1 = x0 + theta^-1 Dinv (rhs - A x0) */
  if (zero_guess) {
    PMG_CHECK(pmgk_scale_dinv(&op->lv, 1.0 / theta, rhs->d, cur->d, ctx->stream)); /* x1 in cur, x0 = 0 */
  } else {
    PMG_SMOOTH_APPLY(PMGK_CHEB_FIRST, NULL, 0.0, 1.0 / theta);
    pmg_vector *t = cur; cur = other; other = t; /* cur = x1, other = x0 */
    other_is_xold = 1;
  }
  if (s->degree >= 2 && fabs(delta) >= 1e-40) {
    double rhok = delta / theta;
    const double sigma = theta / delta;
    for (int k = 0; k < s->degree - 1; ++k) {
      const double rhokp = 1.0 / (2.0 * sigma - rhok);
      const double factor1 = rhokp * rhok, factor2 = 2.0 * rhokp / delta;
      rhok = rhokp;
      /* x_{k+2} = x_{k+1} + f1 (x_{k+1} - x_k) + f2 Dinv (rhs - A x_{k+1}), written over x_k */
      PMG_SMOOTH_APPLY(PMGK_CHEB_STEP, other_is_xold ? other->d : NULL, factor1, factor2);
      pmg_vector *t = cur; cur = other; other = t;
      other_is_xold = 1;
    }
  }
#undef PMG_SMOOTH_APPLY
  if (pushed_out) *pushed_out = chained && want_out && n_applies > 0;
  *result = cur;
  return PMG_OK;
}

/* PreconditionChebyshev::vmult: dst = Cheb_k(src) from a zero initial guess */
int pmg_chebyshev_vmult(pmg_chebyshev *s, pmg_vector *dst, const pmg_vector *src)
{
  if (!s || !dst || !src || dst == src) { pmg_set_error("chebyshev_vmult: bad arguments"); return PMG_ERR_ARG; }
  if (!pmg_layout_same(&dst->lay, &s->op->lay) || !pmg_layout_same(&src->lay, &s->op->lay)) {
    pmg_set_error("chebyshev_vmult: vectors are not initialised for the smoother's operator");
    return PMG_ERR_ARG;
  }
  pmg_vector *res = NULL;
  PMG_CHECK(pmg_chebyshev_smooth(s, dst, src, s->t0, 1, &res));
  if (res != dst) PMG_CHECK(pmg_vector_copy(dst, res));
  return PMG_OK;
}
