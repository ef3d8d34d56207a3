/*
 * oracle/laplace.c -- restatement of the reference's matrix-free Laplace operator
 * (test infrastructure, see orc.h).
 *
 * Follows include/operators/portable_laplace_operator.h:
 *   LocalLaplaceOperator::operator()   :227-381  -> orc_cell_apply + gather/scatter in orc_vmult
 *   LaplaceOperator::vmult             :557-719  -> orc_vmult
 *   LaplaceDiagonalOperator::operator():71-210   -> cell_diagonal
 *   LaplaceOperator::compute_diagonal  :752-917  -> orc_compute_diagonal
 * The tensor-product primitive EvaluatorTensorProduct<evaluate_general,...>::values /
 * co_gradients is un-vendored deal.II; `sweep` restates its documented semantics
 * (SURVEY.md 8c': dof_to_quad applies the 1-D matrix, otherwise its transpose; add
 * accumulates).
 */
#include "orc.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* out(.., a, ..) (+)= sum_b M(a,b) in(.., b, ..) along `dir`;
   M(a,b) = mat[a*n+b] if dof_to_quad, else mat[b*n+a] (transpose). */
static void sweep(int dim, int n, int dir, const double *mat, int dof_to_quad, int add,
                  const double *in, double *out)
{
  const int stride = (dir == 0) ? 1 : (dir == 1) ? n : n * n;
  /* the n^(dim-1) lines along `dir`: their starts are o1 * s1 + o2 * s2 over the two other directions */
  const int s1 = (dir == 0) ? n : 1, s2 = (dir == 2) ? n : n * n;
  const int n2 = (dim == 3) ? n : 1;
  double line[ORC_MAX_DEGREE + 1];
  for (int o2 = 0; o2 < n2; ++o2)
    for (int o1 = 0; o1 < n; ++o1) {
      const int base = o1 * s1 + o2 * s2;
      for (int b = 0; b < n; ++b) line[b] = in[base + b * stride];
      for (int a = 0; a < n; ++a) {
        double s = 0.0;
        if (dof_to_quad)
          for (int b = 0; b < n; ++b) s += mat[a * n + b] * line[b];
        else
          for (int b = 0; b < n; ++b) s += mat[b * n + a] * line[b];
        if (add) out[base + a * stride] += s;
        else out[base + a * stride] = s;
      }
    }
}

void orc_cell_apply(const orc_mf *mf, int64_t cell, double *values, double *scratch)
{
  const int dim = mf->dim, n = mf->p + 1, nq = mf->n_q;
  const int64_t nc = mf->n_cells;
  double *grad = scratch; /* gradients(q,d) -> grad[d*nq + q] */
  const double *S = mf->shape_values, *D = mf->co_shape_gradients;

  /* 1. transform to the collocation space (:282-286) */
  for (int d = 0; d < dim; ++d) sweep(dim, n, d, S, 1, 0, values, values);
  /* 2. gradients in the collocation space (:289-296) */
  for (int d = 0; d < dim; ++d) sweep(dim, n, d, D, 1, 0, values, grad + d * nq);
  /* 3. q-point operation g <- JxW * J^{-1} J^{-T} g (:301-325) */
  for (int q = 0; q < nq; ++q) {
    double g[3], t[3];
    for (int d1 = 0; d1 < dim; ++d1) {
      double tmp = 0.0;
      for (int d2 = 0; d2 < dim; ++d2)
        tmp += mf->inv_jacobian[q + (int64_t)nq * (cell + nc * (d2 + dim * d1))] * grad[d2 * nq + q];
      g[d1] = tmp;
    }
    const double jxw = mf->JxW[q + (int64_t)nq * cell];
    for (int d1 = 0; d1 < dim; ++d1) {
      double tmp = 0.0;
      for (int d2 = 0; d2 < dim; ++d2)
        tmp += mf->inv_jacobian[q + (int64_t)nq * (cell + nc * (d1 + dim * d2))] * g[d2];
      t[d1] = tmp * jxw;
    }
    for (int d = 0; d < dim; ++d) grad[d * nq + q] = t[d];
  }
  /* 4. derivatives transposed, highest direction first, accumulate (:332-350) */
  sweep(dim, n, dim - 1, D, 0, 0, grad + (dim - 1) * nq, values);
  for (int d = dim - 2; d >= 0; --d) sweep(dim, n, d, D, 0, 1, grad + d * nq, values);
  /* 5. back to the nodal space, z,y,x (:353-357) */
  for (int d = dim - 1; d >= 0; --d) sweep(dim, n, d, S, 0, 0, values, values);
}

void orc_vmult(const orc_mf *mf, double *dst, const double *src)
{
  const int nl = mf->n_loc;
  /* dst = 0 (:570) */
  memset(dst, 0, sizeof(double) * (size_t)mf->n_dofs);
  for (int col = 0; col < mf->n_colors; ++col) {
    const int64_t c0 = mf->color_start[col], c1 = mf->color_start[col + 1];
#pragma omp parallel
    {
      double *values = (double *)malloc(sizeof(double) * nl);
      double *scratch = (double *)malloc(sizeof(double) * 4 * nl);
#pragma omp for schedule(static)
      for (int64_t k = c0; k < c1; ++k) {
        const int64_t cell = mf->color_cells[k];
        const uint32_t *l2g = mf->local_to_global + (int64_t)nl * cell;
        const uint32_t *msk = mf->mask + (int64_t)nl * cell;
        /* read dof values, zero at masked dofs (:245-258) */
        for (int i = 0; i < nl; ++i) values[i] = (msk[i] == ORC_INVALID) ? 0.0 : src[l2g[i]];
        orc_cell_apply(mf, cell, values, scratch);
        /* distribute, skipping masked dofs; colouring => plain += (:362-370) */
        for (int i = 0; i < nl; ++i)
          if (msk[i] != ORC_INVALID) dst[l2g[i]] += values[i];
      }
      free(values); free(scratch);
    }
  }
  /* matrix_free.copy_constrained_values(src, dst) (:718) */
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < mf->n_dofs; ++g)
    if (mf->constrained[g]) dst[g] = src[g];
}

/* diagonal of one cell: n_loc unit-vector applies (:104-209) */
static void cell_diagonal(const orc_mf *mf, int64_t cell, double *diag, double *values, double *scratch)
{
  const int nl = mf->n_loc;
  for (int i = 0; i < nl; ++i) {
    for (int j = 0; j < nl; ++j) values[j] = (i == j) ? 1.0 : 0.0;
    orc_cell_apply(mf, cell, values, scratch);
    diag[i] = values[i];
  }
}

void orc_compute_diagonal(orc_mf *mf)
{
  const int nl = mf->n_loc, nq = mf->n_q, dim = mf->dim;
  const int64_t nc = mf->n_cells;
  if (!mf->inv_diag) mf->inv_diag = (double *)malloc(sizeof(double) * (size_t)mf->n_dofs);
  double *diag = mf->inv_diag;
  memset(diag, 0, sizeof(double) * (size_t)mf->n_dofs);
  double *values = (double *)malloc(sizeof(double) * nl);
  double *scratch = (double *)malloc(sizeof(double) * 4 * nl);
  double *cd = (double *)malloc(sizeof(double) * nl);
  /* The reference recomputes every cell; cells whose geometry block is bitwise equal to
     the previous cell's reuse its result (same numbers, fewer flops). */
  double *geo_prev = (double *)malloc(sizeof(double) * nq * (dim * dim + 1));
  double *geo = (double *)malloc(sizeof(double) * nq * (dim * dim + 1));
  int have_prev = 0;
  for (int64_t cell = 0; cell < nc; ++cell) {
    for (int q = 0; q < nq; ++q) {
      geo[q] = mf->JxW[q + (int64_t)nq * cell];
      for (int de = 0; de < dim * dim; ++de)
        geo[nq * (1 + de) + q] = mf->inv_jacobian[q + (int64_t)nq * (cell + nc * de)];
    }
    if (!have_prev || memcmp(geo, geo_prev, sizeof(double) * nq * (dim * dim + 1)) != 0) {
      cell_diagonal(mf, cell, cd, values, scratch);
      memcpy(geo_prev, geo, sizeof(double) * nq * (dim * dim + 1));
      have_prev = 1;
    }
    const uint32_t *l2g = mf->local_to_global + (int64_t)nl * cell;
    for (int i = 0; i < nl; ++i) diag[l2g[i]] += cd[i];
  }
  /* set_constrained_values(1.0) (:906), then invert (:910-916) */
  for (int64_t g = 0; g < mf->n_dofs; ++g) {
    if (mf->constrained[g]) diag[g] = 1.0;
    diag[g] = 1.0 / diag[g];
  }
  free(values); free(scratch); free(cd); free(geo); free(geo_prev);
}
