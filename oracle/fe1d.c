/*
 * oracle/fe1d.c -- 1-D quadrature / Lagrange tables (test infrastructure, see orc.h).
 *
 * Restates the un-vendored deal.II pieces the reference consumes through
 * Portable::MatrixFree::PrecomputedData (shape_values, co_shape_gradients;
 * used at include/operators/portable_laplace_operator.h:99-101,274-276):
 * FE_Q<1>(p) = Lagrange basis on the p+1 Gauss-Lobatto points of [0,1],
 * QGauss<1>(p+1) = Gauss-Legendre.  Published algorithms, no reference text.
 */
#include "orc.h"
#include <math.h>
#include <string.h>

/* Legendre P_n(x) and P_n'(x) on [-1,1] by the three-term recurrence */
static void legendre(int n, double x, double *P, double *dP)
{
  double p0 = 1.0, p1 = x;
  if (n == 0) { *P = 1.0; *dP = 0.0; return; }
  for (int k = 2; k <= n; ++k) {
    double pk = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
    p0 = p1; p1 = pk;
  }
  *P = p1;
  *dP = n * (x * p1 - p0) / (x * x - 1.0);
}

void orc_gauss_legendre(int n, double *x, double *w)
{
  for (int i = 0; i < n; ++i) {
    double z = -cos(M_PI * (i + 0.75) / (n + 0.5));
    for (int it = 0; it < 100; ++it) {
      double P, dP;
      legendre(n, z, &P, &dP);
      double dz = P / dP;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    double P, dP;
    legendre(n, z, &P, &dP);
    x[i] = 0.5 * (z + 1.0);
    w[i] = 1.0 / ((1.0 - z * z) * dP * dP); /* = (2/((1-z^2) dP^2)) / 2 */
  }
  /* symmetrise */
  for (int i = 0; i < n / 2; ++i) {
    double a = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
    x[i] = a; x[n - 1 - i] = 1.0 - a;
    double b = 0.5 * (w[i] + w[n - 1 - i]);
    w[i] = b; w[n - 1 - i] = b;
  }
  if (n % 2) x[n / 2] = 0.5;
}

void orc_gauss_lobatto(int n, double *x)
{
  /* n points: endpoints and the roots of P'_{n-1} */
  const int m = n - 1;
  x[0] = 0.0; x[n - 1] = 1.0;
  for (int i = 1; i < n - 1; ++i) {
    double z = -cos(M_PI * i / m);
    for (int it = 0; it < 100; ++it) {
      /* Newton on q(z) = P'_m(z): q' = (2 z P'_m - m(m+1) P_m)/(1-z^2) */
      double P, dP;
      legendre(m, z, &P, &dP);
      double d2P = (2.0 * z * dP - m * (m + 1.0) * P) / (1.0 - z * z);
      double dz = dP / d2P;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    x[i] = 0.5 * (z + 1.0);
  }
  for (int i = 0; i < n / 2; ++i) {
    double a = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
    x[i] = a; x[n - 1 - i] = 1.0 - a;
  }
  if (n % 2) x[n / 2] = 0.5;
}

void orc_lagrange(int n, const double *nodes, double x, double *val, double *der)
{
  for (int i = 0; i < n; ++i) {
    double denom = 1.0;
    for (int j = 0; j < n; ++j)
      if (j != i) denom *= (nodes[i] - nodes[j]);
    double v = 1.0;
    for (int j = 0; j < n; ++j)
      if (j != i) v *= (x - nodes[j]);
    double d = 0.0;
    for (int k = 0; k < n; ++k) {
      if (k == i) continue;
      double t = 1.0;
      for (int j = 0; j < n; ++j)
        if (j != i && j != k) t *= (x - nodes[j]);
      d += t;
    }
    if (val) val[i] = v / denom;
    if (der) der[i] = d / denom;
  }
}

void orc_shape_tables(int p, double *shape_values, double *co_shape_gradients, double *gauss_w)
{
  const int n = p + 1;
  double gll[ORC_MAX_DEGREE + 1], g[ORC_MAX_DEGREE + 1], w[ORC_MAX_DEGREE + 1];
  double v[ORC_MAX_DEGREE + 1], d[ORC_MAX_DEGREE + 1];
  orc_gauss_lobatto(n, gll);
  orc_gauss_legendre(n, g, w);
  for (int q = 0; q < n; ++q) {
    orc_lagrange(n, gll, g[q], v, NULL);
    for (int i = 0; i < n; ++i) shape_values[q * n + i] = v[i];
    orc_lagrange(n, g, g[q], NULL, d);
    for (int r = 0; r < n; ++r) co_shape_gradients[q * n + r] = d[r];
    if (gauss_w) gauss_w[q] = w[q];
  }
}

static double snap(double v)
{
  /* FE_Q prolongation/embedding entries at coinciding nodes are exactly 0 or 1 */
  if (fabs(v) < 1e-13) return 0.0;
  if (fabs(v - 1.0) < 1e-13) return 1.0;
  return v;
}

void orc_h_prolongation_1d(int p, double *P)
{
  /* P[i*(2p+1) + j + c*p] = phi_i^{parent}((c + xi_j)/2): parent basis at the child's
     support points; rows = coarse node, columns = the 2p+1 nodes of the two children
     (include/multigrid/portable_geometric_transfer.h:1287-1314) */
  const int n = p + 1, nf = 2 * p + 1;
  double gll[ORC_MAX_DEGREE + 1], v[ORC_MAX_DEGREE + 1];
  orc_gauss_lobatto(n, gll);
  for (int c = 0; c < 2; ++c)
    for (int j = 0; j < n; ++j) {
      orc_lagrange(n, gll, 0.5 * (c + gll[j]), v, NULL);
      for (int i = 0; i < n; ++i) P[i * nf + j + c * p] = snap(v[i]);
    }
}

void orc_p_prolongation_1d(int pc, int pf, double *P)
{
  /* P[i*(pf+1)+j] = phi_i^{coarse}(xi_j^{fine}) : embedding FE_Q(pc) -> FE_Q(pf)
     (FETools::get_projection_matrix for nested spaces;
      include/multigrid/portable_polynomial_tranfer.h:957-976) */
  const int nc = pc + 1, nf = pf + 1;
  double gc[ORC_MAX_DEGREE + 1], gf[ORC_MAX_DEGREE + 1], v[ORC_MAX_DEGREE + 1];
  orc_gauss_lobatto(nc, gc);
  orc_gauss_lobatto(nf, gf);
  for (int j = 0; j < nf; ++j) {
    orc_lagrange(nc, gc, gf[j], v, NULL);
    for (int i = 0; i < nc; ++i) P[i * nf + j] = snap(v[i]);
  }
}
