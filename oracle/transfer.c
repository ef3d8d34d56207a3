/*
 * oracle/transfer.c -- restatement of the reference's h- and p-multigrid transfers
 * (test infrastructure, see orc.h).
 *
 * h-transfer: include/multigrid/portable_geometric_transfer.h
 *   CellProlongationKernel::operator() :150-387, CellRestrictionKernel::operator() :450-684,
 *   prolongate_and_add :760-823, restrict_and_add :825-888, 1-D matrix :1287-1314,
 *   setup_weights :1329-1387, setup_dof_indices :1389-1487 (constrained coarse -> invalid :1453).
 * p-transfer: include/multigrid/portable_polynomial_tranfer.h
 *   CellProlongationKernel :103-326, CellRestrictionKernel :390-615, prolongate_and_add :674-786,
 *   restrict_and_add :788-901, 1-D matrix :957-976, weights + masks :1033-1268.
 * Both kernels contract the three directions in the reference's order with the
 * reference's index algebra (x,y,z for prolongation; z,y,x for restriction).
 */
#include "orc.h"
#include <stdlib.h>
#include <string.h>

static int ipow(int a, int b) { int r = 1; while (b--) r *= a; return r; }

static void setup_weights(orc_transfer *t, const uint8_t *constrained_fine)
{
  /* weight = 1 / (number of patches (h) or cells (p) that contain the fine dof),
     zero on constrained fine dofs (:1336-1349 / p :1119-1132) */
  double *w = (double *)calloc((size_t)t->n_dofs_f, sizeof(double));
  const int64_t tot = (int64_t)t->n_loc_f * t->n_cells;
  for (int64_t k = 0; k < tot; ++k) w[t->idx_fine[k]] += 1.0;
  for (int64_t g = 0; g < t->n_dofs_f; ++g)
    if (w[g] > 0) w[g] = 1.0 / w[g];
  for (int64_t g = 0; g < t->n_dofs_f; ++g)
    if (constrained_fine[g]) w[g] = 0.0;
  t->weights = (double *)malloc(sizeof(double) * tot);
  for (int64_t k = 0; k < tot; ++k) t->weights[k] = w[t->idx_fine[k]];
  free(w);
}

orc_transfer *orc_transfer_create_h(const orc_mf *coarse, const orc_mf *fine)
{
  const int dim = coarse->dim, p = coarse->p;
  if (fine->p != p || fine->dim != dim) return NULL;
  for (int d = 0; d < dim; ++d)
    if (fine->n[d] != 2 * coarse->n[d]) return NULL; /* AssertThrow non-isotropic (:1055) */
  orc_transfer *t = (orc_transfer *)calloc(1, sizeof(orc_transfer));
  t->kind = 0; t->dim = dim; t->pc = p; t->pf = p;
  t->nc1 = p + 1; t->nf1 = 2 * p + 1;
  t->n_cells = coarse->n_cells;
  t->n_loc_c = ipow(t->nc1, dim); t->n_loc_f = ipow(t->nf1, dim);
  t->n_dofs_c = coarse->n_dofs; t->n_dofs_f = fine->n_dofs;
  t->coarse = coarse; t->fine = fine;
  t->P = (double *)malloc(sizeof(double) * t->nc1 * t->nf1);
  orc_h_prolongation_1d(p, t->P);
  t->idx_coarse = (uint32_t *)malloc(sizeof(uint32_t) * t->n_loc_c * t->n_cells);
  t->idx_fine = (uint32_t *)malloc(sizeof(uint32_t) * t->n_loc_f * t->n_cells);
  const int nzc = (dim == 3) ? t->nc1 : 1, nzf = (dim == 3) ? t->nf1 : 1;
  for (int64_t cell = 0; cell < t->n_cells; ++cell) {
    int64_t r = cell;
    const int cx = (int)(r % coarse->n[0]); r /= coarse->n[0];
    const int cy = (int)(r % coarse->n[1]); r /= coarse->n[1];
    const int cz = (int)r;
    /* coarse indices are the coarse cell's lexicographic dofs, invalid when constrained */
    for (int i = 0; i < t->n_loc_c; ++i) {
      const uint32_t g = coarse->local_to_global[i + (int64_t)t->n_loc_c * cell];
      t->idx_coarse[i + (int64_t)t->n_loc_c * cell] = coarse->constrained[g] ? ORC_INVALID : g;
    }
    /* fine patch: the (2p+1)^dim dofs of the 2^dim children, lexicographic */
    for (int kz = 0; kz < nzf; ++kz)
      for (int ky = 0; ky < t->nf1; ++ky)
        for (int kx = 0; kx < t->nf1; ++kx) {
          const int i = kx + t->nf1 * (ky + t->nf1 * kz);
          const int64_t gx = (int64_t)cx * 2 * p + kx, gy = (int64_t)cy * 2 * p + ky;
          const int64_t gz = (dim == 3) ? (int64_t)cz * 2 * p + kz : 0;
          t->idx_fine[i + (int64_t)t->n_loc_f * cell] = (uint32_t)(gx + fine->nd[0] * (gy + fine->nd[1] * gz));
        }
    (void)nzc;
  }
  setup_weights(t, fine->constrained);
  return t;
}

orc_transfer *orc_transfer_create_p(const orc_mf *coarse, const orc_mf *fine)
{
  const int dim = coarse->dim;
  if (fine->dim != dim || coarse->p >= fine->p) return NULL;
  for (int d = 0; d < dim; ++d)
    if (fine->n[d] != coarse->n[d]) return NULL;
  orc_transfer *t = (orc_transfer *)calloc(1, sizeof(orc_transfer));
  t->kind = 1; t->dim = dim; t->pc = coarse->p; t->pf = fine->p;
  t->nc1 = t->pc + 1; t->nf1 = t->pf + 1;
  t->n_cells = fine->n_cells;
  t->n_loc_c = coarse->n_loc; t->n_loc_f = fine->n_loc;
  t->n_dofs_c = coarse->n_dofs; t->n_dofs_f = fine->n_dofs;
  t->coarse = coarse; t->fine = fine;
  t->P = (double *)malloc(sizeof(double) * t->nc1 * t->nf1);
  orc_p_prolongation_1d(t->pc, t->pf, t->P);
  /* same cell on both levels (cell_lists_fine_to_coarse is the identity here) */
  t->idx_coarse = (uint32_t *)malloc(sizeof(uint32_t) * t->n_loc_c * t->n_cells);
  t->idx_fine = (uint32_t *)malloc(sizeof(uint32_t) * t->n_loc_f * t->n_cells);
  memcpy(t->idx_coarse, coarse->local_to_global, sizeof(uint32_t) * t->n_loc_c * t->n_cells);
  memcpy(t->idx_fine, fine->local_to_global, sizeof(uint32_t) * t->n_loc_f * t->n_cells);
  t->mask_coarse = (uint32_t *)malloc(sizeof(uint32_t) * t->n_loc_c * t->n_cells);
  t->mask_fine = (uint32_t *)malloc(sizeof(uint32_t) * t->n_loc_f * t->n_cells);
  memcpy(t->mask_coarse, coarse->mask, sizeof(uint32_t) * t->n_loc_c * t->n_cells);
  memcpy(t->mask_fine, fine->mask, sizeof(uint32_t) * t->n_loc_f * t->n_cells);
  setup_weights(t, fine->constrained);
  return t;
}

void orc_transfer_destroy(orc_transfer *t)
{
  if (!t) return;
  free(t->idx_coarse); free(t->idx_fine); free(t->mask_coarse); free(t->mask_fine);
  free(t->weights); free(t->P); free(t);
}

/* fine = (P x P x P) coarse, contraction order x, y, z with the reference's
   intermediate layouts (h :263-366, p :199-293). P(row=coarse m.., col=fine) = P[k*Nf + m]. */
static void cell_prolongate(int dim, int Nc, int Nf, const double *P, const double *vc, double *vf,
                            double *tmp1, double *tmp2)
{
  if (dim == 2) {
    /* tmp[i*Nf + j] = sum_k P[j + k*Nf] vc[i*Nc + k]   (i: coarse y, j: fine x) */
    for (int i = 0; i < Nc; ++i)
      for (int j = 0; j < Nf; ++j) {
        double sum = P[j] * vc[i * Nc];
        for (int k = 1; k < Nc; ++k) sum += P[j + k * Nf] * vc[i * Nc + k];
        tmp1[i * Nf + j] = sum;
      }
    /* vf[i + j*Nf] = sum_k P[j + k*Nf] tmp[i + k*Nf]   (i: fine x, j: fine y) */
    for (int i = 0; i < Nf; ++i)
      for (int j = 0; j < Nf; ++j) {
        double sum = P[j] * tmp1[i];
        for (int k = 1; k < Nc; ++k) sum += P[j + k * Nf] * tmp1[i + k * Nf];
        vf[i + j * Nf] = sum;
      }
    return;
  }
  /* x: tmp1[(z*Nc + y)*Nf + xf] */
  for (int i = 0; i < Nc; ++i)
    for (int j = 0; j < Nc; ++j)
      for (int m = 0; m < Nf; ++m) {
        const int base = (i * Nc + j) * Nc;
        double sum = P[m] * vc[base];
        for (int k = 1; k < Nc; ++k) sum += P[m + k * Nf] * vc[base + k];
        tmp1[(i * Nc + j) * Nf + m] = sum;
      }
  /* y: tmp2[xf + (z*Nf + yf)*Nf] */
  for (int i = 0; i < Nf; ++i)
    for (int j = 0; j < Nc; ++j)
      for (int m = 0; m < Nf; ++m) {
        const int base = i + j * Nf * Nc;
        double sum = P[m] * tmp1[base];
        for (int k = 1; k < Nc; ++k) sum += P[m + k * Nf] * tmp1[base + k * Nf];
        tmp2[i + (j * Nf + m) * Nf] = sum;
      }
  /* z: vf[(yf + zf*Nf)*Nf + xf] */
  for (int i = 0; i < Nf; ++i)
    for (int j = 0; j < Nf; ++j)
      for (int m = 0; m < Nf; ++m) {
        const int base = i * Nf + j;
        double sum = P[m] * tmp2[base];
        for (int k = 1; k < Nc; ++k) sum += P[m + k * Nf] * tmp2[base + k * Nf * Nf];
        vf[(i + m * Nf) * Nf + j] = sum;
      }
}

/* coarse = (P^T x P^T x P^T) fine, contraction order z, y, x (h :567-668, p :496-588) */
static void cell_restrict(int dim, int Nc, int Nf, const double *P, const double *vf, double *vc,
                          double *tmp1, double *tmp2)
{
  if (dim == 2) {
    /* tmp[i + j*Nf] = sum_k P[j*Nf + k] vf[i + k*Nf]  (i: fine x, j: coarse y) */
    for (int i = 0; i < Nf; ++i)
      for (int j = 0; j < Nc; ++j) {
        double sum = P[j * Nf] * vf[i];
        for (int k = 1; k < Nf; ++k) sum += P[j * Nf + k] * vf[i + k * Nf];
        tmp1[i + j * Nf] = sum;
      }
    /* vc[i*Nc + j] = sum_k P[j*Nf + k] tmp[i*Nf + k]  (i: coarse y, j: coarse x) */
    for (int i = 0; i < Nc; ++i)
      for (int j = 0; j < Nc; ++j) {
        double sum = P[j * Nf] * tmp1[i * Nf];
        for (int k = 1; k < Nf; ++k) sum += P[j * Nf + k] * tmp1[i * Nf + k];
        vc[i * Nc + j] = sum;
      }
    return;
  }
  /* z: tmp1[(yf + zc*Nf)*Nf + xf] */
  for (int i = 0; i < Nf; ++i)
    for (int j = 0; j < Nf; ++j)
      for (int m = 0; m < Nc; ++m) {
        const int base = i * Nf + j;
        double sum = P[m * Nf] * vf[base];
        for (int k = 1; k < Nf; ++k) sum += P[m * Nf + k] * vf[base + k * Nf * Nf];
        tmp1[(i + m * Nf) * Nf + j] = sum;
      }
  /* y: tmp2[xf + (zc*Nc + yc)*Nf] */
  for (int i = 0; i < Nf; ++i)
    for (int j = 0; j < Nc; ++j)
      for (int m = 0; m < Nc; ++m) {
        const int base = i + j * Nf * Nf;
        double sum = P[m * Nf] * tmp1[base];
        for (int k = 1; k < Nf; ++k) sum += P[m * Nf + k] * tmp1[base + k * Nf];
        tmp2[i + (j * Nc + m) * Nf] = sum;
      }
  /* x: vc[(zc*Nc + yc)*Nc + xc] */
  for (int i = 0; i < Nc; ++i)
    for (int j = 0; j < Nc; ++j)
      for (int m = 0; m < Nc; ++m) {
        const int base = (i * Nc + j) * Nf;
        double sum = P[m * Nf] * tmp2[base];
        for (int k = 1; k < Nf; ++k) sum += P[m * Nf + k] * tmp2[base + k];
        vc[(i * Nc + j) * Nc + m] = sum;
      }
}

/* Cells are visited colour by colour (cells of one colour share no dof, so `+=` replaces
   the reference's atomic_add; the result is independent of the thread count). */
static void colour_lists(const orc_transfer *t, int *n_colors, const int64_t **start, const int64_t **cells)
{
  const orc_mf *m = (t->kind == 0) ? t->coarse : t->fine;
  *n_colors = m->n_colors; *start = m->color_start; *cells = m->color_cells;
}

void orc_prolongate_and_add(const orc_transfer *t, double *dst, const double *src)
{
  const int nlc = t->n_loc_c, nlf = t->n_loc_f;
  int n_colors; const int64_t *cstart, *ccells;
  colour_lists(t, &n_colors, &cstart, &ccells);
  for (int col = 0; col < n_colors; ++col) {
#pragma omp parallel
    {
      double *vc = (double *)malloc(sizeof(double) * nlc);
      double *vf = (double *)malloc(sizeof(double) * nlf);
      double *tmp1 = (double *)malloc(sizeof(double) * nlf);
      double *tmp2 = (double *)malloc(sizeof(double) * nlf);
#pragma omp for schedule(static)
      for (int64_t k = cstart[col]; k < cstart[col + 1]; ++k) {
        const int64_t cell = ccells[k];
        const uint32_t *ic = t->idx_coarse + (int64_t)nlc * cell;
        const uint32_t *jf = t->idx_fine + (int64_t)nlf * cell;
        const double *w = t->weights + (int64_t)nlf * cell;
        /* read coarse values; h: invalid index -> 0 (:164-174); p: read unmasked (:115-121) */
        for (int i = 0; i < nlc; ++i) vc[i] = (ic[i] == ORC_INVALID) ? 0.0 : src[ic[i]];
        cell_prolongate(t->dim, t->nc1, t->nf1, t->P, vc, vf, tmp1, tmp2);
        for (int i = 0; i < nlf; ++i) vf[i] *= w[i]; /* apply weights (:369-374) */
        if (t->kind == 0) {
          for (int i = 0; i < nlf; ++i) dst[jf[i]] += vf[i]; /* :378-385 */
        } else {
          const uint32_t *mf_ = t->mask_fine + (int64_t)nlf * cell;
          for (int i = 0; i < nlf; ++i)
            if (mf_[i] != ORC_INVALID) dst[jf[i]] += vf[i]; /* p :306-324 */
        }
      }
      free(vc); free(vf); free(tmp1); free(tmp2);
    }
  }
}

void orc_restrict_and_add(const orc_transfer *t, double *dst, const double *src)
{
  const int nlc = t->n_loc_c, nlf = t->n_loc_f;
  int n_colors; const int64_t *cstart, *ccells;
  colour_lists(t, &n_colors, &cstart, &ccells);
  for (int col = 0; col < n_colors; ++col) {
#pragma omp parallel
    {
      double *vc = (double *)malloc(sizeof(double) * nlc);
      double *vf = (double *)malloc(sizeof(double) * nlf);
      double *tmp1 = (double *)malloc(sizeof(double) * nlf);
      double *tmp2 = (double *)malloc(sizeof(double) * nlf);
#pragma omp for schedule(static)
      for (int64_t k = cstart[col]; k < cstart[col + 1]; ++k) {
        const int64_t cell = ccells[k];
        const uint32_t *ic = t->idx_coarse + (int64_t)nlc * cell;
        const uint32_t *jf = t->idx_fine + (int64_t)nlf * cell;
        const double *w = t->weights + (int64_t)nlf * cell;
        for (int i = 0; i < nlf; ++i) vf[i] = src[jf[i]] * w[i]; /* read + weights (:462-478) */
        cell_restrict(t->dim, t->nc1, t->nf1, t->P, vf, vc, tmp1, tmp2);
        if (t->kind == 0) {
          for (int i = 0; i < nlc; ++i)
            if (ic[i] != ORC_INVALID) dst[ic[i]] += vc[i]; /* :674-682 */
        } else {
          const uint32_t *mc = t->mask_coarse + (int64_t)nlc * cell;
          for (int i = 0; i < nlc; ++i)
            if (mc[i] != ORC_INVALID) dst[ic[i]] += vc[i]; /* p :592-612 */
        }
      }
      free(vc); free(vf); free(tmp1); free(tmp2);
    }
  }
}
