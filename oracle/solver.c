/*
 * oracle/solver.c -- Chebyshev-Jacobi smoother, V-cycle, CG, rhs and solution norm
 * (test infrastructure, see orc.h).
 *
 * V-cycle: restates include/multigrid/portable_v_cycle_multigrid.h
 *   vmult :79-94, smooth :96-126, v_cycle :128-190 (including its per-call temporaries).
 * Smoother / Krylov solver are un-vendored deal.II (>= 9.8.0, the reference's only pin:
 * source/geometric_multigrid/CMakeLists.txt:32); restated from the published
 * PreconditionChebyshev / SolverCG algorithms as written down in DESIGN.md "Numerical spec":
 *   - PreconditionChebyshev::vmult: x1 = theta^-1 D^-1 b; for k>=1
 *       x+ = x + rho_k rho_{k-1} (x - x-) + (2 rho_k / delta) D^-1 (b - A x),
 *       rho_k = 1/(2 sigma - rho_{k-1}), rho_0 = delta/theta, sigma = theta/delta.
 *   - eigenvalue estimate: v_i = (i mod 11) - mean, PCG(D^-1) on A x = v from x = 0 with
 *       IterationNumberControl(eig_cg_n_iterations, 1e-10); Lanczos tridiagonal from the CG
 *       coefficients of iterations 1..it-1 (deal.II >= 9.5 pushes iteration k's coefficients
 *       during iteration k+1); lambda_max *= 1.2; no tridiagonal => lambda_min = lambda_max = 1.
 *   - SolverCG: PCG, convergence check on ||r||_2 before the first and after every iteration.
 * Call sites: source/geometric_multigrid/program.cc:267-285 (smoother parameters), :342-355
 * (V(2,2), tol 1e-12 ||b||, "Solver converged in K iterations"), :289-334 (rhs), :382-395 (norm).
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

double orc_dot(int64_t n, const double *a, const double *b)
{
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

double orc_l2(int64_t n, const double *a) { return sqrt(orc_dot(n, a, a)); }

static double *vec_new(int64_t n) { return (double *)calloc((size_t)n, sizeof(double)); }

/* eigenvalues of a symmetric tridiagonal matrix by the implicit QL algorithm */
void orc_tridiag_eigenvalues(int n, const double *diag, const double *offdiag, double *eig)
{
  if (n <= 0) return;
  double *d = eig;
  double *e = (double *)calloc((size_t)n + 1, sizeof(double));
  for (int i = 0; i < n; ++i) d[i] = diag[i];
  for (int i = 0; i < n - 1; ++i) e[i] = offdiag[i];
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (++iter > 200) break;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i], b = c * e[i];
          r = hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
          s = f / r; c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p; e[l] = g; e[m] = 0.0;
      }
    } while (m != l);
  }
  /* sort ascending */
  for (int i = 1; i < n; ++i) {
    double v = d[i]; int j = i - 1;
    while (j >= 0 && d[j] > v) { d[j + 1] = d[j]; --j; }
    d[j + 1] = v;
  }
  free(e);
}

/* ---- Chebyshev ---------------------------------------------------------- */
void orc_chebyshev_init(orc_chebyshev *c, const orc_mf *mf, double smoothing_range, int degree, int eig_cg_n_iterations)
{
  memset(c, 0, sizeof(*c));
  c->mf = mf; c->smoothing_range = smoothing_range; c->degree = degree;
  c->eig_cg_n_iterations = eig_cg_n_iterations;
}

void orc_chebyshev_estimate(orc_chebyshev *c)
{
  const orc_mf *mf = c->mf;
  const int64_t n = mf->n_dofs;
  const double *dinv = mf->inv_diag;
  double lmin = 1.0, lmax = 1.0;
  c->cg_iterations = 0;
  if (c->eig_cg_n_iterations > 0) {
    double *b = vec_new(n), *x = vec_new(n), *r = vec_new(n), *z = vec_new(n), *pv = vec_new(n), *Ap = vec_new(n);
    /* set_initial_guess: (global index mod 11) minus the mean value */
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) { b[i] = (double)(i % 11); sum += b[i]; }
    const double mean = sum / (double)n;
    for (int64_t i = 0; i < n; ++i) b[i] -= mean;
    const int max_it = c->eig_cg_n_iterations;
    double *diag = (double *)calloc((size_t)max_it + 2, sizeof(double));
    double *off = (double *)calloc((size_t)max_it + 2, sizeof(double));
    int nt = 0;
    memcpy(r, b, sizeof(double) * n);
    double res = orc_l2(n, r);
    int it = 0;
    if (res > 1e-10) {
      for (int64_t i = 0; i < n; ++i) z[i] = dinv[i] * r[i];
      memcpy(pv, z, sizeof(double) * n);
      double rz = orc_dot(n, r, z);
      double alpha_prev = 0.0, beta_prev = 0.0, eigen_beta_alpha = 0.0;
      for (;;) {
        ++it;
        orc_vmult(mf, Ap, pv);
        const double alpha = rz / orc_dot(n, pv, Ap);
        for (int64_t i = 0; i < n; ++i) { x[i] += alpha * pv[i]; r[i] -= alpha * Ap[i]; }
        res = orc_l2(n, r);
        if (it > 1) {
          /* coefficients of iteration it-1 become available now */
          diag[nt] = 1.0 / alpha_prev + eigen_beta_alpha;
          eigen_beta_alpha = beta_prev / alpha_prev;
          off[nt] = sqrt(beta_prev) / alpha_prev;
          ++nt;
        }
        if (res <= 1e-10 || it >= max_it) break;
        for (int64_t i = 0; i < n; ++i) z[i] = dinv[i] * r[i];
        const double rz_new = orc_dot(n, r, z);
        const double beta = rz_new / rz;
        for (int64_t i = 0; i < n; ++i) pv[i] = z[i] + beta * pv[i];
        rz = rz_new;
        alpha_prev = alpha; beta_prev = beta;
      }
    }
    c->cg_iterations = it;
    if (nt > 0) {
      double *eig = (double *)calloc((size_t)nt, sizeof(double));
      orc_tridiag_eigenvalues(nt, diag, off, eig);
      lmin = eig[0]; lmax = eig[nt - 1];
      free(eig);
    }
    lmax *= 1.2; /* safety factor */
    free(diag); free(off);
    free(b); free(x); free(r); free(z); free(pv); free(Ap);
  }
  const double alpha = (c->smoothing_range > 1.0) ? lmax / c->smoothing_range : fmin(0.9 * lmax, lmin);
  if (c->degree < 0) {
    const double actual_range = lmax / alpha;
    const double sigma = (1.0 - sqrt(1.0 / actual_range)) / (1.0 + sqrt(1.0 / actual_range));
    const double eps = c->smoothing_range;
    c->degree = 1 + (int)(log(1.0 / eps + sqrt(1.0 / eps / eps - 1.0)) / log(1.0 / sigma));
  }
  c->lambda_min = lmin; c->lambda_max = lmax;
  c->delta = (lmax - alpha) * 0.5;
  c->theta = (lmax + alpha) * 0.5;
  c->initialized = 1;
}

void orc_chebyshev_vmult(orc_chebyshev *c, double *dst, const double *src)
{
  if (!c->initialized) orc_chebyshev_estimate(c); /* lazy, on the first vmult */
  const orc_mf *mf = c->mf;
  const int64_t n = mf->n_dofs;
  const double *dinv = mf->inv_diag;
  double *x = vec_new(n), *xold = vec_new(n), *t1 = vec_new(n);
  const double f0 = 1.0 / c->theta;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) x[i] = f0 * src[i] * dinv[i];
  if (c->degree >= 2 && fabs(c->delta) >= 1e-40) {
    double rhok = c->delta / c->theta;
    const double sigma = c->theta / c->delta;
    for (int k = 0; k < c->degree - 1; ++k) {
      orc_vmult(mf, t1, x);
      const double rhokp = 1.0 / (2.0 * sigma - rhok);
      const double factor1 = rhokp * rhok, factor2 = 2.0 * rhokp / c->delta;
      rhok = rhokp;
      if (k == 0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i)
          xold[i] = (1.0 + factor1) * x[i] + factor2 * dinv[i] * (src[i] - t1[i]);
      } else {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i)
          xold[i] = (1.0 + factor1) * x[i] - factor1 * xold[i] + factor2 * dinv[i] * (src[i] - t1[i]);
      }
      double *sw = x; x = xold; xold = sw;
    }
  }
  memcpy(dst, x, sizeof(double) * n);
  free(x); free(xold); free(t1);
}

/* ---- V-cycle ------------------------------------------------------------ */
static void smooth(orc_vcycle *v, double *u, const double *rhs, int level)
{
  const orc_mf *mf = v->mf[level];
  const int64_t n = mf->n_dofs;
  double *r = vec_new(n), *d = vec_new(n); /* :116-118 */
  orc_vmult(mf, r, u);                       /* :120 */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) r[i] = -r[i] + rhs[i]; /* r.sadd(-1, rhs) :121 */
  orc_chebyshev_vmult(&v->smoother[level], d, r);        /* :123 */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) u[i] += d[i];          /* :125 */
  free(r); free(d);
}

static void v_cycle(orc_vcycle *v, double *dst, const double *src, int level)
{
  if (level == 0) { smooth(v, dst, src, 0); return; } /* :148-154 */
  const orc_mf *mf = v->mf[level];
  const int64_t n = mf->n_dofs, nc = v->mf[level - 1]->n_dofs;
  for (int s = 0; s < v->pre; ++s) smooth(v, dst, src, level); /* :157-160 */
  double *residual = vec_new(n);
  orc_vmult(mf, residual, dst);                               /* :165 */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) residual[i] = -residual[i] + src[i]; /* :166 */
  double *coarse_residual = vec_new(nc);
  orc_restrict_and_add(v->transfer[level], coarse_residual, residual); /* :172 */
  double *coarse_correction = vec_new(nc);
  v_cycle(v, coarse_correction, coarse_residual, level - 1);           /* :179 */
  orc_prolongate_and_add(v->transfer[level], dst, coarse_correction);  /* :182 */
  for (int s = 0; s < v->post; ++s) smooth(v, dst, src, level);        /* :185-188 */
  free(residual); free(coarse_residual); free(coarse_correction);
}

void orc_vcycle_vmult(orc_vcycle *v, double *dst, const double *src)
{
  memset(dst, 0, sizeof(double) * (size_t)v->mf[v->n_levels - 1]->n_dofs); /* dst = 0 :92 */
  v_cycle(v, dst, src, v->n_levels - 1);
}

/* ---- CG ----------------------------------------------------------------- */
int orc_cg_solve(const orc_mf *A, double *x, const double *b, orc_vcycle *precond, int max_it, double tol,
                 int *last_step, double *history, int history_cap)
{
  const int64_t n = A->n_dofs;
  double *r = vec_new(n), *z = vec_new(n), *p = vec_new(n), *Ap = vec_new(n);
  int it = 0, converged = 0;
  /* r = b - A x */
  orc_vmult(A, Ap, x);
  for (int64_t i = 0; i < n; ++i) r[i] = b[i] - Ap[i];
  double res = orc_l2(n, r);
  if (history && history_cap > 0) history[0] = res;
  if (res <= tol) converged = 1;
  double rz = 0.0;
  if (!converged) {
    if (precond) orc_vcycle_vmult(precond, z, r); else memcpy(z, r, sizeof(double) * n);
    memcpy(p, z, sizeof(double) * n);
    rz = orc_dot(n, r, z);
  }
  while (!converged && it < max_it) {
    ++it;
    orc_vmult(A, Ap, p);
    const double alpha = rz / orc_dot(n, p, Ap);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
    res = orc_l2(n, r);
    if (history && it < history_cap) history[it] = res;
    if (res <= tol) { converged = 1; break; }
    if (precond) orc_vcycle_vmult(precond, z, r); else memcpy(z, r, sizeof(double) * n);
    const double rz_new = orc_dot(n, r, z);
    const double beta = rz_new / rz;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    rz = rz_new;
  }
  if (last_step) *last_step = it;
  free(r); free(z); free(p); free(Ap);
  return converged ? 0 : 1;
}

/* ---- rhs and solution norm ---------------------------------------------- */
void orc_assemble_rhs(const orc_mf *mf, double *rhs)
{
  /* cell_rhs(i) = sum_q phi_i(x_q) * 1 * JxW(q); constrained rows dropped
     (distribute_local_to_global with homogeneous constraints), program.cc:311-323 */
  const int dim = mf->dim, n1 = mf->p + 1, nl = mf->n_loc, nq = mf->n_q;
  const double *S = mf->shape_values;
  memset(rhs, 0, sizeof(double) * (size_t)mf->n_dofs);
  const int nz = (dim == 3) ? n1 : 1;
  for (int64_t cell = 0; cell < mf->n_cells; ++cell) {
    const uint32_t *l2g = mf->local_to_global + (int64_t)nl * cell;
    for (int iz = 0; iz < nz; ++iz)
      for (int iy = 0; iy < n1; ++iy)
        for (int ix = 0; ix < n1; ++ix) {
          double s = 0.0;
          for (int qz = 0; qz < nz; ++qz)
            for (int qy = 0; qy < n1; ++qy)
              for (int qx = 0; qx < n1; ++qx) {
                const int q = qx + n1 * (qy + n1 * qz);
                double phi = S[qx * n1 + ix] * S[qy * n1 + iy];
                if (dim == 3) phi *= S[qz * n1 + iz];
                s += phi * mf->JxW[q + (int64_t)nq * cell];
              }
          const uint32_t g = l2g[ix + n1 * (iy + n1 * iz)];
          if (!mf->constrained[g]) rhs[g] += s;
        }
  }
}

double orc_l2_norm_solution(const orc_mf *mf, const double *u)
{
  /* integrate_difference(u_h, 0, QGauss(p+2), L2_norm), program.cc:382-395 */
  const int dim = mf->dim, p = mf->p, n1 = p + 1, m = p + 2, nl = mf->n_loc;
  double gq[ORC_MAX_DEGREE + 2], gw[ORC_MAX_DEGREE + 2], gll[ORC_MAX_DEGREE + 1];
  double E[(ORC_MAX_DEGREE + 2) * (ORC_MAX_DEGREE + 1)];
  orc_gauss_legendre(m, gq, gw);
  orc_gauss_lobatto(n1, gll);
  for (int q = 0; q < m; ++q) orc_lagrange(n1, gll, gq[q], E + q * n1, NULL);
  const int nz = (dim == 3) ? n1 : 1, mz = (dim == 3) ? m : 1;
  double total = 0.0;
  for (int64_t cell = 0; cell < mf->n_cells; ++cell) {
    const uint32_t *l2g = mf->local_to_global + (int64_t)nl * cell;
    double vol = 1.0;
    for (int d = 0; d < dim; ++d) vol *= mf->h[d];
    for (int qz = 0; qz < mz; ++qz)
      for (int qy = 0; qy < m; ++qy)
        for (int qx = 0; qx < m; ++qx) {
          double val = 0.0;
          for (int iz = 0; iz < nz; ++iz)
            for (int iy = 0; iy < n1; ++iy)
              for (int ix = 0; ix < n1; ++ix) {
                double phi = E[qx * n1 + ix] * E[qy * n1 + iy];
                if (dim == 3) phi *= E[qz * n1 + iz];
                val += phi * u[l2g[ix + n1 * (iy + n1 * iz)]];
              }
          double w = gw[qx] * gw[qy] * ((dim == 3) ? gw[qz] : 1.0) * vol;
          total += val * val * w;
        }
  }
  return sqrt(total);
}
