/*
 * oracle/mesh.c -- reference-layout matrix-free data for a structured box mesh
 * (test infrastructure, see orc.h).
 *
 * Emits the arrays Portable::MatrixFree::PrecomputedData holds and the
 * reference kernels consume: local_to_global(i,cell), inv_jacobian(q,cell,d,e),
 * JxW(q,cell) (include/operators/portable_laplace_operator.h:144,158,254) and the
 * Dirichlet mask of setup_dirichlet_boundary_dofs_masks (:487-555:
 * mask(i,cell) = constrained ? invalid_unsigned_int : global index), with local
 * dofs in lexicographic order (x fastest).  The mesh is the drivers' mesh: the
 * unit hyper-cube, uniformly refined, boundary id 0 = homogeneous Dirichlet
 * (source/geometric_multigrid/program.cc:130,163-166,404-417).
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

double orc_coef_c5(double x, double y, double z)
{
  return 1.0 / (0.05 + 2.0 * (x * x + y * y + z * z));
}

orc_mf *orc_mf_create(int dim, int p, int nx, int ny, int nz, unsigned faces, orc_coef_fn coef)
{
  orc_mf *mf = (orc_mf *)calloc(1, sizeof(orc_mf));
  mf->dim = dim; mf->p = p;
  mf->n[0] = nx; mf->n[1] = ny; mf->n[2] = (dim == 3) ? nz : 1;
  for (int d = 0; d < 3; ++d) {
    mf->nd[d] = (d < dim) ? mf->n[d] * p + 1 : 1;
    mf->h[d] = 1.0 / mf->n[d];
  }
  mf->faces = faces;
  const int n1 = p + 1;
  mf->n_loc = mf->n_q = (dim == 3) ? n1 * n1 * n1 : n1 * n1;
  mf->n_cells = (int64_t)mf->n[0] * mf->n[1] * mf->n[2];
  mf->n_dofs = (int64_t)mf->nd[0] * mf->nd[1] * mf->nd[2];
  orc_shape_tables(p, mf->shape_values, mf->co_shape_gradients, mf->gauss_w);

  double gq[ORC_MAX_DEGREE + 1], gw[ORC_MAX_DEGREE + 1];
  orc_gauss_legendre(n1, gq, gw);

  /* constrained dofs: all dofs on a Dirichlet face */
  mf->constrained = (uint8_t *)calloc((size_t)mf->n_dofs, 1);
  for (int64_t g = 0; g < mf->n_dofs; ++g) {
    int c[3];
    int64_t r = g;
    c[0] = (int)(r % mf->nd[0]); r /= mf->nd[0];
    c[1] = (int)(r % mf->nd[1]); r /= mf->nd[1];
    c[2] = (int)r;
    uint8_t con = 0;
    for (int d = 0; d < dim; ++d) {
      if (c[d] == 0 && (faces >> (2 * d) & 1u)) con = 1;
      if (c[d] == mf->nd[d] - 1 && (faces >> (2 * d + 1) & 1u)) con = 1;
    }
    mf->constrained[g] = con;
  }

  const int nl = mf->n_loc, nq = mf->n_q;
  const int64_t nc = mf->n_cells;
  mf->local_to_global = (uint32_t *)malloc(sizeof(uint32_t) * nl * nc);
  mf->mask = (uint32_t *)malloc(sizeof(uint32_t) * nl * nc);
  mf->inv_jacobian = (double *)calloc((size_t)nq * nc * dim * dim, sizeof(double));
  mf->JxW = (double *)malloc(sizeof(double) * nq * nc);

  const int nzl = (dim == 3) ? n1 : 1;
#pragma omp parallel for schedule(static)
  for (int64_t cell = 0; cell < nc; ++cell) {
    int cc[3];
    int64_t r = cell;
    cc[0] = (int)(r % mf->n[0]); r /= mf->n[0];
    cc[1] = (int)(r % mf->n[1]); r /= mf->n[1];
    cc[2] = (int)r;
    for (int kz = 0; kz < nzl; ++kz)
      for (int ky = 0; ky < n1; ++ky)
        for (int kx = 0; kx < n1; ++kx) {
          const int i = kx + n1 * (ky + n1 * kz);
          const int64_t gx = (int64_t)cc[0] * p + kx, gy = (int64_t)cc[1] * p + ky;
          const int64_t gz = (dim == 3) ? (int64_t)cc[2] * p + kz : 0;
          const int64_t g = gx + mf->nd[0] * (gy + mf->nd[1] * gz);
          mf->local_to_global[i + (int64_t)nl * cell] = (uint32_t)g;
          mf->mask[i + (int64_t)nl * cell] = mf->constrained[g] ? ORC_INVALID : (uint32_t)g;
          /* geometry at quadrature point q == i (same tensor index) */
          const int q = i;
          double jxw = gw[kx] * gw[ky] * ((dim == 3) ? gw[kz] : 1.0);
          for (int d = 0; d < dim; ++d) jxw *= mf->h[d];
          if (coef) {
            const double x = (cc[0] + gq[kx]) * mf->h[0], y = (cc[1] + gq[ky]) * mf->h[1];
            const double z = (dim == 3) ? (cc[2] + gq[kz]) * mf->h[2] : 0.0;
            jxw *= coef(x, y, z);
          }
          mf->JxW[q + (int64_t)nq * cell] = jxw;
          for (int d = 0; d < dim; ++d)
            mf->inv_jacobian[q + (int64_t)nq * (cell + nc * (d + dim * d))] = 1.0 / mf->h[d];
        }
  }

  /* parity colouring: cells of one colour share no dof */
  mf->n_colors = 1 << dim;
  mf->color_start = (int64_t *)calloc(mf->n_colors + 1, sizeof(int64_t));
  mf->color_cells = (int64_t *)malloc(sizeof(int64_t) * nc);
  int64_t pos = 0;
  for (int col = 0; col < mf->n_colors; ++col) {
    mf->color_start[col] = pos;
    for (int64_t cell = 0; cell < nc; ++cell) {
      int64_t r = cell;
      int cx = (int)(r % mf->n[0]); r /= mf->n[0];
      int cy = (int)(r % mf->n[1]); r /= mf->n[1];
      int cz = (int)r;
      int c = (cx & 1) | ((cy & 1) << 1) | ((cz & 1) << 2);
      if (c == col) mf->color_cells[pos++] = cell;
    }
  }
  mf->color_start[mf->n_colors] = pos;
  return mf;
}

void orc_mf_destroy(orc_mf *mf)
{
  if (!mf) return;
  free(mf->local_to_global); free(mf->mask); free(mf->inv_jacobian); free(mf->JxW);
  free(mf->constrained); free(mf->color_start); free(mf->color_cells); free(mf->inv_diag);
  free(mf);
}
