/*
 * pmg_operator.c -- LaplaceOperator of the C-ABI (host C).
 *
 * Mirrors Portable::LaplaceOperator<dim,fe_degree,number>
 * (reference include/operators/portable_laplace_operator.h:383-461) method by method; the
 * cell loop itself is csrc/pmg_apply_sweep.h.  The constructor replaces MatrixFree::reinit +
 * setup_dirichlet_boundary_dofs_masks (:463-555): on the structured box nothing needs to be
 * stored per cell, only the 1-D tables.
 */
#include "pmg_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

int pmg_laplace_operator_create(pmg_context *ctx, int dim, int degree, int nx, int ny, int nz,
                                unsigned faces, int coefficient, pmg_operator **out)
{
  if (!ctx || !out || nx < 1 || ny < 1 || nz < 1) { pmg_set_error("operator_create: bad arguments"); return PMG_ERR_ARG; }
  if (dim != 3) { pmg_set_error("only dim = 3 is compiled (the reference's 2-D driver is out of this round's scope)"); return PMG_ERR_UNSUPPORTED; }
  if (degree < 1 || degree > PMG_MAX_DEGREE) { pmg_set_error("degree %d outside 1..%d", degree, PMG_MAX_DEGREE); return PMG_ERR_UNSUPPORTED; }
  if (coefficient != 0) { pmg_set_error("variable coefficient operator not available yet"); return PMG_ERR_UNSUPPORTED; }
  if ((int64_t)(nx * (int64_t)degree + 1) * (ny * (int64_t)degree + 1) >= (int64_t)1 << 31) return PMG_ERR_ARG;
  pmg_operator *op = (pmg_operator *)calloc(1, sizeof(*op));
  if (!op) return PMG_ERR_NOMEM;
  op->ctx = ctx; op->dim = dim; op->degree = degree; op->coefficient = coefficient; op->faces = faces & PMG_ALL_FACES;
  PMG_CHECK(pmg_layout_make(ctx, degree, nx, ny, nz, &op->lay));
  pmgk_level *lv = &op->lv;
  lv->degree = degree;
  lv->nx = nx; lv->ny = ny; lv->nz = nz;
  lv->Nx = op->lay.Nx; lv->Ny = op->lay.Ny; lv->Nz = op->lay.Nz;
  lv->faces = op->faces;
  lv->z0 = op->lay.z0; lv->nzl = op->lay.nzl;
  lv->cz_lo = op->lay.cz_lo; lv->cz_hi = op->lay.cz_hi;
  lv->z_own_lo = op->lay.z_own_lo; lv->z_own_hi = op->lay.z_own_hi;
  lv->h[0] = 1.0 / nx; lv->h[1] = 1.0 / ny; lv->h[2] = 1.0 / nz;
  pmg_fe_fastdiag(degree, lv->S, lv->lam);
  pmg_fe_pencil(degree, lv->Mref, lv->Kref);
  const int T = degree + 2;
  double *tab = (double *)malloc(sizeof(double) * T * T * T);
  if (!tab) { free(op); return PMG_ERR_NOMEM; }
  pmg_fe_dinv_table(degree, lv->h, dim, tab);
  PMG_CUDA(cudaSetDevice(ctx->device));
  PMG_CUDA(cudaMalloc((void **)&op->d_dinv_tab, sizeof(double) * T * T * T));
  PMG_CUDA(cudaMemcpyAsync(op->d_dinv_tab, tab, sizeof(double) * T * T * T, cudaMemcpyHostToDevice, ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  free(tab);
  lv->dinv_tab = op->d_dinv_tab;
  lv->dinv_vec = NULL;
  const char *tv = getenv("PMG_TILE_VARIANT");
  lv->tile_variant = tv ? atoi(tv) : 0;
  *out = op;
  return PMG_OK;
}

int pmg_laplace_operator_destroy(pmg_operator *op)
{
  if (!op) return PMG_OK;
  cudaStreamSynchronize(op->ctx->stream);
  if (op->dinv) pmg_vector_destroy(op->dinv);
  for (int i = 0; i < 4; ++i) if (op->cg_ws[i]) pmg_vector_destroy(op->cg_ws[i]);
  cudaFree(op->d_dinv_tab);
  free(op);
  return PMG_OK;
}

static int check_vec(const pmg_operator *op, const pmg_vector *v, const char *what)
{
  if (!v || v->ctx != op->ctx || !pmg_layout_same(&v->lay, &op->lay)) {
    pmg_set_error("%s: vector is not initialised for this operator (initialize_dof_vector)", what);
    return PMG_ERR_ARG;
  }
  return PMG_OK;
}

static int apply_mode(const pmg_operator *op, int mode, pmg_vector *dst, const pmg_vector *src, const pmg_vector *b,
                      const pmg_vector *xold, double f1, double f2)
{
  pmg_context *ctx = op->ctx;
  if (!op->lay.active) return PMG_OK;
  /* src.update_ghost_values() (:661); there is no compress(add): the kernel owns complete rows */
  PMG_CHECK(pmg_halo_update(ctx, &src->lay, src->d));
  PMG_CHECK(pmgk_apply(&op->lv, mode, src->d, b ? b->d : NULL, xold ? xold->d : NULL, dst->d, f1, f2, ctx->stream));
  return PMG_OK;
}

int pmg_laplace_operator_vmult(const pmg_operator *op, pmg_vector *dst, const pmg_vector *src)
{
  if (!op) return PMG_ERR_ARG;
  PMG_CHECK(check_vec(op, dst, "vmult(dst)"));
  PMG_CHECK(check_vec(op, src, "vmult(src)"));
  if (dst == src) { pmg_set_error("vmult: dst and src must differ"); return PMG_ERR_ARG; }
  return apply_mode(op, PMGK_APPLY, dst, src, NULL, NULL, 0.0, 0.0);
}

int pmg_laplace_operator_Tvmult(const pmg_operator *op, pmg_vector *dst, const pmg_vector *src)
{
  return pmg_laplace_operator_vmult(op, dst, src); /* symmetric (:721-735) */
}

int pmg_laplace_operator_residual(const pmg_operator *op, pmg_vector *dst, const pmg_vector *b, const pmg_vector *src)
{
  if (!op) return PMG_ERR_ARG;
  PMG_CHECK(check_vec(op, dst, "residual(dst)"));
  PMG_CHECK(check_vec(op, src, "residual(src)"));
  PMG_CHECK(check_vec(op, b, "residual(b)"));
  if (dst == src) { pmg_set_error("residual: dst and src must differ"); return PMG_ERR_ARG; }
  return apply_mode(op, PMGK_RESIDUAL, dst, src, b, NULL, 0.0, 0.0);
}

int pmg_laplace_operator_chebyshev_step(const pmg_operator *op, pmg_vector *dst, const pmg_vector *src,
                                        const pmg_vector *xold, const pmg_vector *b, double f1, double f2)
{
  if (!op) return PMG_ERR_ARG;
  PMG_CHECK(check_vec(op, dst, "chebyshev_step(dst)"));
  PMG_CHECK(check_vec(op, src, "chebyshev_step(src)"));
  PMG_CHECK(check_vec(op, b, "chebyshev_step(b)"));
  if (xold) PMG_CHECK(check_vec(op, xold, "chebyshev_step(xold)"));
  if (dst == src) { pmg_set_error("chebyshev_step: dst and src must differ"); return PMG_ERR_ARG; }
  return apply_mode(op, PMGK_CHEB_STEP, dst, src, b, xold, f1, f2);
}

int pmg_laplace_operator_initialize_dof_vector(const pmg_operator *op, pmg_vector **vec)
{
  if (!op || !vec) return PMG_ERR_ARG;
  return pmg_vector_create_layout(op->ctx, &op->lay, vec);
}

int pmg_laplace_operator_compute_diagonal(pmg_operator *op)
{
  if (!op) return PMG_ERR_ARG;
  if (!op->dinv) PMG_CHECK(pmg_vector_create_layout(op->ctx, &op->lay, &op->dinv));
  if (!op->lay.active) return PMG_OK;
  /* diagonal of the tensor-product cell matrices summed at shared dofs, 1 on constrained dofs,
     inverted (:752-917).  The fused smoother keeps using the (p+2)^3 table: same numbers,
     no 8 B/DoF stream. */
  return pmgk_fill_dinv(&op->lv, op->dinv->d, op->ctx->stream);
}

int pmg_laplace_operator_get_matrix_diagonal_inverse(const pmg_operator *op, const pmg_vector **dinv)
{
  if (!op || !dinv) return PMG_ERR_ARG;
  if (!op->dinv) { pmg_set_error("compute_diagonal() has not been called"); return PMG_ERR_STATE; }
  *dinv = op->dinv;
  return PMG_OK;
}

int pmg_laplace_operator_m(const pmg_operator *op, int64_t *m) { if (!op || !m) return PMG_ERR_ARG; *m = op->lay.n_global; return PMG_OK; }
int pmg_laplace_operator_n(const pmg_operator *op, int64_t *n) { return pmg_laplace_operator_m(op, n); }

int pmg_laplace_operator_el(const pmg_operator *op, int64_t row, int64_t col, double *value)
{
  if (!op || !value || row < 0 || row >= op->lay.n_global) return PMG_ERR_ARG;
  if (row != col) { pmg_set_error("el(): only diagonal entries are available (reference: ExcNotImplemented)"); return PMG_ERR_UNSUPPORTED; }
  if (!op->dinv) { pmg_set_error("el(): compute_diagonal() has not been called"); return PMG_ERR_STATE; }
  /* 1 / inverse_diagonal(row,row) (:953); the owner of the row reads it, the others get it by allreduce */
  pmg_context *ctx = op->ctx;
  const pmg_layout *l = &op->lay;
  const int gz = (int)(row / l->plane);
  double v = 0.0;
  PMG_CUDA(cudaMemsetAsync(ctx->scalars + 1, 0, sizeof(double), ctx->stream));
  if (l->active && gz >= l->z_own_lo && gz < l->z_own_hi)
    PMG_CUDA(cudaMemcpyAsync(ctx->scalars + 1, op->dinv->d + (row - l->plane * l->z0), sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  PMG_CHECK(pmg_allreduce_sum(ctx, ctx->scalars + 1, 1));
  PMG_CUDA(cudaMemcpyAsync(ctx->h_scalars + 1, ctx->scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  v = ctx->h_scalars[1];
  *value = 1.0 / v;
  return PMG_OK;
}

int pmg_laplace_operator_degree(const pmg_operator *op, int *degree) { if (!op || !degree) return PMG_ERR_ARG; *degree = op->degree; return PMG_OK; }

int pmg_laplace_operator_cells(const pmg_operator *op, int *nx, int *ny, int *nz)
{
  if (!op) return PMG_ERR_ARG;
  if (nx) *nx = op->lay.nx;
  if (ny) *ny = op->lay.ny;
  if (nz) *nz = op->lay.nz;
  return PMG_OK;
}

int pmg_laplace_operator_vmult_host(const pmg_operator *op, double *dst_host, const double *src_host)
{
  if (!op || !dst_host || !src_host) return PMG_ERR_ARG;
  pmg_vector *s = NULL, *d = NULL;
  PMG_CHECK(pmg_laplace_operator_initialize_dof_vector(op, &s));
  PMG_CHECK(pmg_laplace_operator_initialize_dof_vector(op, &d));
  int rc = pmg_vector_import_host(s, src_host);
  if (!rc) rc = pmg_laplace_operator_vmult(op, d, s);
  if (!rc) rc = pmg_vector_export_host(d, dst_host);
  pmg_vector_destroy(s);
  pmg_vector_destroy(d);
  return rc;
}

/* ---- driver helpers: right-hand side and solution norm ---------------------------------- */
/* Load vector of f = 1 (reference source/geometric_multigrid/program.cc:289-334).  On the box mesh
   the cell vector is the tensor product of the 1-D load b1[i] = sum_q w_q phi_i(x_q) h, so the global
   vector is separable too: rhs(x,y,z) = bx(x) by(y) bz(z), constrained rows dropped.  Assembled on
   the host once (setup path). */
int pmg_laplace_operator_assemble_rhs(const pmg_operator *op, pmg_vector *rhs)
{
  if (!op) return PMG_ERR_ARG;
  PMG_CHECK(check_vec(op, rhs, "assemble_rhs"));
  const pmg_layout *l = &op->lay;
  if (!l->active) return PMG_OK;
  const int p = op->degree, n1 = p + 1;
  double gll[PMG_MAX_DEGREE + 2], g[PMG_MAX_DEGREE + 2], w[PMG_MAX_DEGREE + 2], v[PMG_MAX_DEGREE + 2], b1[PMG_MAX_DEGREE + 2];
  pmg_fe_gll(n1, gll);
  pmg_fe_gauss(n1, g, w);
  memset(b1, 0, sizeof(b1));
  for (int q = 0; q < n1; ++q) {
    pmg_fe_lagrange(n1, gll, g[q], v, NULL);
    for (int i = 0; i < n1; ++i) b1[i] += w[q] * v[i];
  }
  const int N[3] = {l->Nx, l->Ny, l->Nz}, nc[3] = {l->nx, l->ny, l->nz};
  double *line[3];
  for (int d = 0; d < 3; ++d) {
    line[d] = (double *)calloc((size_t)N[d], sizeof(double));
    if (!line[d]) return PMG_ERR_NOMEM;
    for (int c = 0; c < nc[d]; ++c)
      for (int i = 0; i < n1; ++i) line[d][c * p + i] += b1[i] * op->lv.h[d];
    if (op->faces >> (2 * d) & 1u) line[d][0] = 0.0;
    if (op->faces >> (2 * d + 1) & 1u) line[d][N[d] - 1] = 0.0;
  }
  double *host = (double *)malloc(sizeof(double) * (size_t)l->n_local);
  if (!host) return PMG_ERR_NOMEM;
  for (int lz = 0; lz < l->nzl; ++lz)
    for (int y = 0; y < l->Ny; ++y) {
      const double f = line[2][l->z0 + lz] * line[1][y];
      double *row = host + ((int64_t)lz * l->Ny + y) * l->Nx;
      for (int x = 0; x < l->Nx; ++x) row[x] = f * line[0][x];
    }
  PMG_CUDA(cudaMemcpyAsync(rhs->d, host, sizeof(double) * (size_t)l->n_local, cudaMemcpyHostToDevice, op->ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(op->ctx->stream));
  free(host);
  for (int d = 0; d < 3; ++d) free(line[d]);
  return PMG_OK;
}

/* ||u_h||_L2 with QGauss(p+2) (program.cc:382-395), evaluated on the host from the exported vector
   by sum factorisation cell by cell. */
int pmg_laplace_operator_solution_norm(const pmg_operator *op, const pmg_vector *u, double *norm)
{
  if (!op || !norm) return PMG_ERR_ARG;
  PMG_CHECK(check_vec(op, u, "solution_norm"));
  const pmg_layout *l = &op->lay;
  const int p = op->degree, n1 = p + 1, m = p + 2;
  double *host = (double *)malloc(sizeof(double) * (size_t)l->n_global);
  if (!host) return PMG_ERR_NOMEM;
  PMG_CHECK(pmg_vector_export_host(u, host));
  double gll[PMG_MAX_DEGREE + 2], g[PMG_MAX_DEGREE + 3], w[PMG_MAX_DEGREE + 3];
  double E[(PMG_MAX_DEGREE + 3) * (PMG_MAX_DEGREE + 2)];
  pmg_fe_gll(n1, gll);
  pmg_fe_gauss(m, g, w);
  for (int q = 0; q < m; ++q) pmg_fe_lagrange(n1, gll, g[q], E + q * n1, NULL);
  const double vol = op->lv.h[0] * op->lv.h[1] * op->lv.h[2];
  double total = 0.0;
  double *t1 = (double *)malloc(sizeof(double) * m * n1 * n1), *t2 = (double *)malloc(sizeof(double) * m * m * n1);
  for (int cz = 0; cz < l->nz; ++cz)
    for (int cy = 0; cy < l->ny; ++cy)
      for (int cx = 0; cx < l->nx; ++cx) {
        /* x */
        for (int k = 0; k < n1; ++k)
          for (int j = 0; j < n1; ++j) {
            const double *row = host + ((int64_t)(cz * p + k) * l->Ny + (cy * p + j)) * l->Nx + cx * p;
            for (int q = 0; q < m; ++q) {
              double s = 0.0;
              for (int i = 0; i < n1; ++i) s += E[q * n1 + i] * row[i];
              t1[(k * n1 + j) * m + q] = s;
            }
          }
        /* y */
        for (int k = 0; k < n1; ++k)
          for (int qy = 0; qy < m; ++qy)
            for (int qx = 0; qx < m; ++qx) {
              double s = 0.0;
              for (int j = 0; j < n1; ++j) s += E[qy * n1 + j] * t1[(k * n1 + j) * m + qx];
              t2[(k * m + qy) * m + qx] = s;
            }
        /* z + accumulate */
        for (int qz = 0; qz < m; ++qz)
          for (int qy = 0; qy < m; ++qy)
            for (int qx = 0; qx < m; ++qx) {
              double s = 0.0;
              for (int k = 0; k < n1; ++k) s += E[qz * n1 + k] * t2[(k * m + qy) * m + qx];
              total += s * s * w[qx] * w[qy] * w[qz] * vol;
            }
      }
  free(t1); free(t2); free(host);
  *norm = sqrt(total);
  return PMG_OK;
}
