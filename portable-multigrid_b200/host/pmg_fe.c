/*
 * pmg_fe.c -- 1-D finite-element tables of the product library (host C, no CUDA).
 *
 * Provides what the reference obtains from deal.II's FE_Q<1>, QGauss<1> and ShapeInfo
 * (reference include/operators/portable_laplace_operator.h:469-482; 1-D transfer matrices
 * include/multigrid/portable_geometric_transfer.h:1287-1314 and
 * include/multigrid/portable_polynomial_tranfer.h:957-976), in the form the B200 kernels want:
 *   - the 1-D mass/stiffness pencil (M, K) of FE_Q(p) on [0,1] integrated with QGauss(p+1),
 *   - its simultaneous diagonalisation  M = S^T S,  K = S^T diag(lam) S,
 *   - 1-D prolongation matrices for the h- and p-transfer,
 *   - 1-D diagonals for the tabulated inverse diagonal.
 * Independent of oracle/ (barycentric Lagrange evaluation, Jacobi eigen-solver).
 */
#include "pmg_internal.h"
#include <math.h>
#include <string.h>

#define NMAX (PMG_MAX_DEGREE + 2)

static void legendre_pd(int n, double x, double *pn, double *dpn)
{
  /* P_n and P_n' on [-1,1] via Bonnet's recursion */
  double a = 1.0, b = x, da = 0.0, db = 1.0;
  if (n == 0) { *pn = 1.0; *dpn = 0.0; return; }
  for (int k = 1; k < n; ++k) {
    const double c = ((2 * k + 1) * x * b - k * a) / (k + 1);
    const double dc = da + (2 * k + 1) * b;
    a = b; b = c; da = db; db = dc;
  }
  *pn = b; *dpn = db;
}

void pmg_fe_gauss(int n, double *x, double *w)
{
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5));
    double pn, dpn;
    for (int it = 0; it < 60; ++it) {
      legendre_pd(n, z, &pn, &dpn);
      const double dz = pn / dpn;
      z -= dz;
      if (fabs(dz) < 4e-16) break;
    }
    legendre_pd(n, z, &pn, &dpn);
    const double wi = 1.0 / ((1.0 - z * z) * dpn * dpn);
    x[n - 1 - i] = 0.5 * (1.0 + z); w[n - 1 - i] = wi;
    x[i] = 0.5 * (1.0 - z);         w[i] = wi;
  }
}

void pmg_fe_gll(int n, double *x)
{
  /* zeros of (1-z^2) P'_{n-1}(z) */
  const int m = n - 1;
  x[0] = 0.0; x[m] = 1.0;
  for (int i = 1; i <= m / 2; ++i) {
    double z = cos(M_PI * i / m);
    for (int it = 0; it < 60; ++it) {
      double pm, dpm;
      legendre_pd(m, z, &pm, &dpm);
      /* (1-z^2) P_m'' = 2 z P_m' - m(m+1) P_m */
      const double d2 = (2.0 * z * dpm - m * (m + 1.0) * pm) / (1.0 - z * z);
      const double dz = dpm / d2;
      z -= dz;
      if (fabs(dz) < 4e-16) break;
    }
    x[m - i] = 0.5 * (1.0 + z);
    x[i] = 0.5 * (1.0 - z);
  }
  if (m % 2 == 0) x[m / 2] = 0.5;
}

/* values and derivatives of the Lagrange basis on `nodes` at x (barycentric form) */
void pmg_fe_lagrange(int n, const double *nodes, double x, double *val, double *der)
{
  double wb[NMAX];
  for (int i = 0; i < n; ++i) {
    double w = 1.0;
    for (int j = 0; j < n; ++j) if (j != i) w *= (nodes[i] - nodes[j]);
    wb[i] = 1.0 / w;
  }
  int hit = -1;
  for (int i = 0; i < n; ++i) if (x == nodes[i]) hit = i;
  if (hit < 0) {
    double ell = 1.0, s1 = 0.0;
    for (int j = 0; j < n; ++j) { ell *= (x - nodes[j]); s1 += 1.0 / (x - nodes[j]); }
    for (int i = 0; i < n; ++i) {
      const double li = ell * wb[i] / (x - nodes[i]);
      if (val) val[i] = li;
      if (der) der[i] = li * (s1 - 1.0 / (x - nodes[i]));
    }
  } else {
    for (int i = 0; i < n; ++i) {
      if (val) val[i] = (i == hit) ? 1.0 : 0.0;
      if (der && i != hit) der[i] = (wb[i] / wb[hit]) / (nodes[hit] - nodes[i]);
    }
    if (der) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) if (j != hit) s += 1.0 / (nodes[hit] - nodes[j]);
      der[hit] = s;
    }
  }
}

void pmg_fe_pencil(int p, double *M, double *K)
{
  const int n = p + 1;
  double gll[NMAX], g[NMAX], w[NMAX], v[NMAX], d[NMAX];
  pmg_fe_gll(n, gll);
  pmg_fe_gauss(n, g, w);
  memset(M, 0, sizeof(double) * n * n);
  memset(K, 0, sizeof(double) * n * n);
  for (int q = 0; q < n; ++q) {
    pmg_fe_lagrange(n, gll, g[q], v, d);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        M[i * n + j] += w[q] * v[i] * v[j];
        K[i * n + j] += w[q] * d[i] * d[j];
      }
  }
}

/* cyclic Jacobi for a symmetric n x n matrix: A = Q diag(ev) Q^T, Q columns = eigenvectors */
static void jacobi_eig(int n, double *A, double *Q, double *ev)
{
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Q[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-60) break;
    for (int pi = 0; pi < n; ++pi)
      for (int qi = pi + 1; qi < n; ++qi) {
        const double apq = A[pi * n + qi];
        if (apq == 0.0) continue;
        const double th = (A[qi * n + qi] - A[pi * n + pi]) / (2.0 * apq);
        const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + pi], akq = A[k * n + qi];
          A[k * n + pi] = c * akp - s * akq;
          A[k * n + qi] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[pi * n + k], aqk = A[qi * n + k];
          A[pi * n + k] = c * apk - s * aqk;
          A[qi * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double qkp = Q[k * n + pi], qkq = Q[k * n + qi];
          Q[k * n + pi] = c * qkp - s * qkq;
          Q[k * n + qi] = s * qkp + c * qkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) ev[i] = A[i * n + i];
}

void pmg_fe_fastdiag(int p, double *S, double *lam)
{
  /* M = L L^T; C = L^-1 K L^-T = Q diag(lam) Q^T; S = Q^T L^T  =>  M = S^T S, K = S^T diag(lam) S */
  const int n = p + 1;
  double M[NMAX * NMAX], K[NMAX * NMAX], L[NMAX * NMAX], Cm[NMAX * NMAX], Q[NMAX * NMAX], T[NMAX * NMAX];
  pmg_fe_pencil(p, M, K);
  memset(L, 0, sizeof(L));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = M[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = (i == j) ? sqrt(s) : s / L[j * n + j];
    }
  /* T = L^-1 K  (forward substitution on columns), C = T L^-T */
  for (int c = 0; c < n; ++c)
    for (int i = 0; i < n; ++i) {
      double s = K[i * n + c];
      for (int k = 0; k < i; ++k) s -= L[i * n + k] * T[k * n + c];
      T[i * n + c] = s / L[i * n + i];
    }
  for (int r = 0; r < n; ++r)
    for (int i = 0; i < n; ++i) {
      double s = T[r * n + i];
      for (int k = 0; k < i; ++k) s -= L[i * n + k] * Cm[r * n + k];
      Cm[r * n + i] = s / L[i * n + i];
    }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) { const double a = 0.5 * (Cm[i * n + j] + Cm[j * n + i]); Cm[i * n + j] = Cm[j * n + i] = a; }
  double ev[NMAX];
  jacobi_eig(n, Cm, Q, ev);
  /* sort eigenpairs ascending (selection sort on columns of Q) */
  int order[NMAX];
  for (int i = 0; i < n; ++i) order[i] = i;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (ev[order[j]] < ev[order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
  for (int a = 0; a < n; ++a) {
    const int col = order[a];
    lam[a] = (a == 0) ? 0.0 : ev[col]; /* the constant mode has eigenvalue exactly 0 */
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int k = 0; k <= i; ++k) s += Q[k * n + col] * L[i * n + k]; /* (Q^T L^T)(a,i) = sum_k Q(k,a) L(i,k) */
      S[a * n + i] = s;
    }
  }
}

static double snap01(double v)
{
  if (fabs(v) < 1e-13) return 0.0;
  if (fabs(v - 1.0) < 1e-13) return 1.0;
  return v;
}

void pmg_fe_prolongation_h(int p, double *Pm)
{
  /* rows: p+1 coarse nodes; columns: the 2p+1 nodes of the two children */
  const int n = p + 1, nf = 2 * p + 1;
  double gll[NMAX], v[NMAX];
  pmg_fe_gll(n, gll);
  for (int j = 0; j < nf; ++j) {
    const int child = (j <= p) ? 0 : 1;
    const int jj = j - child * p;
    pmg_fe_lagrange(n, gll, 0.5 * (child + gll[jj]), v, NULL);
    for (int i = 0; i < n; ++i) Pm[i * nf + j] = snap01(v[i]);
  }
}

void pmg_fe_prolongation_p(int pc, int pf, double *Pm)
{
  const int nc = pc + 1, nf = pf + 1;
  double gc[NMAX], gf[NMAX], v[NMAX];
  pmg_fe_gll(nc, gc);
  pmg_fe_gll(nf, gf);
  for (int j = 0; j < nf; ++j) {
    pmg_fe_lagrange(nc, gc, gf[j], v, NULL);
    for (int i = 0; i < nc; ++i) Pm[i * nf + j] = snap01(v[i]);
  }
}

void pmg_fe_diag_1d(int p, double *Md, double *Kd)
{
  double M[NMAX * NMAX], K[NMAX * NMAX];
  const int n = p + 1;
  pmg_fe_pencil(p, M, K);
  for (int i = 0; i < n; ++i) { Md[i] = M[i * n + i]; Kd[i] = K[i * n + i]; }
}

void pmg_fe_dinv_table(int p, const double h[3], int dim, double *tab)
{
  /* position types per direction: 0 = interior vertex, 1..p-1 = cell-interior node,
     p = vertex on the low boundary, p+1 = vertex on the high boundary */
  const int T = p + 2;
  double Md[NMAX], Kd[NMAX], m1[3][NMAX], k1[3][NMAX];
  pmg_fe_diag_1d(p, Md, Kd);
  for (int d = 0; d < 3; ++d)
    for (int t = 0; t < T; ++t) {
      double mm, kk;
      if (t == 0) { mm = Md[0] + Md[p]; kk = Kd[0] + Kd[p]; }
      else if (t < p) { mm = Md[t]; kk = Kd[t]; }
      else if (t == p) { mm = Md[0]; kk = Kd[0]; }
      else { mm = Md[p]; kk = Kd[p]; }
      m1[d][t] = mm * h[d];
      k1[d][t] = kk / h[d];
    }
  for (int tz = 0; tz < T; ++tz)
    for (int ty = 0; ty < T; ++ty)
      for (int tx = 0; tx < T; ++tx) {
        double dg;
        if (dim == 3)
          dg = k1[0][tx] * m1[1][ty] * m1[2][tz] + m1[0][tx] * k1[1][ty] * m1[2][tz] + m1[0][tx] * m1[1][ty] * k1[2][tz];
        else
          dg = k1[0][tx] * m1[1][ty] + m1[0][tx] * k1[1][ty];
        tab[tx + T * (ty + T * tz)] = 1.0 / dg;
      }
}

/* Tables of the quadrature-point formulation (variable-coefficient operator, csrc/pmg_apply_var.h), what the reference
   reads from MatrixFree's shape_values / co_shape_gradients (include/operators/portable_laplace_operator.h:267-357):
   Sq[q*n+i] = phi_i(x_q) (nodal GLL basis at Gauss point q), Dco[q*n+r] = derivative of the Lagrange basis ON the Gauss
   points at Gauss point q (so grad at the quadrature points = Dco (Sq u)), G[q*n+i] = phi_i'(x_q), gq / gw = Gauss points
   and weights on [0,1].  Any output may be NULL. */
void pmg_fe_shape_tables(int p, double *Sq, double *Dco, double *G, double *gq, double *gw)
{
  const int n = p + 1;
  double gll[NMAX], g[NMAX], w[NMAX], v[NMAX], d[NMAX];
  pmg_fe_gll(n, gll);
  pmg_fe_gauss(n, g, w);
  for (int q = 0; q < n; ++q) {
    pmg_fe_lagrange(n, gll, g[q], v, d);
    for (int i = 0; i < n; ++i) {
      if (Sq) Sq[q * n + i] = v[i];
      if (G) G[q * n + i] = d[i];
    }
    if (Dco) {
      pmg_fe_lagrange(n, g, g[q], NULL, d);
      for (int r = 0; r < n; ++r) Dco[q * n + r] = d[r];
    }
    if (gq) gq[q] = g[q];
    if (gw) gw[q] = w[q];
  }
}
