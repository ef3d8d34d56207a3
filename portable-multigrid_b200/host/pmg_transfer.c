/*
 * pmg_transfer.c -- MGTransferBase of the C-ABI: geometric (h) and polynomial (p) transfer (host C).
 *
 * Mirrors Portable::GeometricTransfer (reference include/multigrid/portable_geometric_transfer.h:
 * reinit :892-1327, prolongate_and_add :760-823, restrict_and_add :825-888) and
 * Portable::PolynomialTransfer (include/multigrid/portable_polynomial_tranfer.h: reinit :903-1031,
 * prolongate_and_add :674-786, restrict_and_add :788-901).  reinit shrinks to the 1-D matrix:
 * indices, weights and masks are computed inside the kernels (csrc/pmg_transfer.cu).
 * The reference's per-call temporaries (:782-789, :846-852) are gone: kernels write the caller's
 * vectors directly; the only workspace is the cell-local restriction scratch, allocated once.
 *
 * Multi-GPU: when both levels are slab-distributed a rank's fine slab is the refinement of its
 * coarse slab (pmg_core.c partition), so prolongation needs only the coarse ghost update and
 * restriction one compress(add) of the coarse upper ghost plane.  When the coarse level lives on
 * rank 0 only (below the DoF threshold), the FINE vector is gathered to / scattered from rank 0
 * and the transfer runs there.
 */
#include "pmg_internal.h"
#include <stdlib.h>
#include <string.h>

static int transfer_create(int kind, const pmg_operator *coarse, const pmg_operator *fine, pmg_transfer **out)
{
  if (!coarse || !fine || !out || coarse->ctx != fine->ctx) { pmg_set_error("transfer_create: bad arguments"); return PMG_ERR_ARG; }
  const pmg_layout *c = &coarse->lay, *f = &fine->lay;
  if (coarse->faces != fine->faces || coarse->dim != fine->dim) { pmg_set_error("transfer: levels differ in boundary description"); return PMG_ERR_ARG; }
  if (kind == 0) {
    if (coarse->degree != fine->degree || f->nx != 2 * c->nx || f->ny != 2 * c->ny || (coarse->dim == 3 && f->nz != 2 * c->nz)) {
      pmg_set_error("geometric transfer: fine mesh is not the coarse mesh refined once (reference AssertThrow, portable_geometric_transfer.h:1055)");
      return PMG_ERR_ARG;
    }
  } else {
    if (coarse->degree >= fine->degree || f->nx != c->nx || f->ny != c->ny || f->nz != c->nz) {
      pmg_set_error("polynomial transfer: need the same mesh and p_coarse < p_fine");
      return PMG_ERR_ARG;
    }
  }
  pmg_context *ctx = coarse->ctx;
  pmg_transfer *t = (pmg_transfer *)calloc(1, sizeof(*t));
  if (!t) return PMG_ERR_NOMEM;
  t->ctx = ctx; t->kind = kind; t->coarse = coarse; t->fine = fine;
  const int NC = coarse->degree + 1, NF = (kind == 0) ? 2 * coarse->degree + 1 : fine->degree + 1;
  double P[(PMG_MAX_DEGREE + 1) * (2 * PMG_MAX_DEGREE + 1)];
  if (kind == 0) pmg_fe_prolongation_h(coarse->degree, P);
  else pmg_fe_prolongation_p(coarse->degree, fine->degree, P);
  PMG_CUDA(cudaMalloc((void **)&t->d_P, sizeof(double) * NC * NF));
  PMG_CUDA(cudaMemcpyAsync(t->d_P, P, sizeof(double) * NC * NF, cudaMemcpyHostToDevice, ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  /* layouts: distributed/distributed, gathered/gathered, or fine distributed over a gathered coarse level */
  const int mixed = (!f->gathered && c->gathered && ctx->n_ranks > 1);
  if (mixed) {
    /* fine-level vector in rank-0 layout */
    pmg_layout full = *f;
    full.gathered = 1; full.active = (ctx->rank == 0);
    full.cz_lo = 0; full.cz_hi = full.active ? f->nz : 0;
    full.z0 = 0; full.nzl = full.active ? f->Nz : 0;
    full.z_own_lo = 0; full.z_own_hi = full.active ? f->Nz : 0;
    full.lower = full.upper = -1;
    full.n_local = full.plane * full.nzl;
    PMG_CHECK(pmg_vector_create_layout(ctx, &full, &t->gather_buf));
  }
  if (c->active) {
    pmgk_level cl = coarse->lv, fl = fine->lv;
    if (mixed) { fl.z0 = 0; fl.nzl = f->Nz; fl.cz_lo = 0; fl.cz_hi = f->nz; fl.z_own_lo = 0; fl.z_own_hi = f->Nz; }
    const int64_t ns = pmgk_restrict_scratch_doubles(kind, &cl, &fl);
    if (ns > 0) PMG_CUDA(cudaMalloc((void **)&t->d_scratch, sizeof(double) * (size_t)ns));
  }
  *out = t;
  return PMG_OK;
}

int pmg_transfer_create_geometric(const pmg_operator *coarse, const pmg_operator *fine, pmg_transfer **t)
{
  return transfer_create(0, coarse, fine, t);
}

int pmg_transfer_create_polynomial(const pmg_operator *coarse, const pmg_operator *fine, pmg_transfer **t)
{
  return transfer_create(1, coarse, fine, t);
}

int pmg_transfer_destroy(pmg_transfer *t)
{
  if (!t) return PMG_OK;
  cudaStreamSynchronize(t->ctx->stream);
  cudaFree(t->d_P); cudaFree(t->d_scratch);
  if (t->gather_buf) pmg_vector_destroy(t->gather_buf);
  free(t);
  return PMG_OK;
}

static int check(const pmg_transfer *t, const pmg_vector *fine, const pmg_vector *coarse)
{
  if (!t || !fine || !coarse || !pmg_layout_same(&fine->lay, &t->fine->lay) || !pmg_layout_same(&coarse->lay, &t->coarse->lay)) {
    pmg_set_error("transfer: vectors do not belong to the transfer's levels");
    return PMG_ERR_ARG;
  }
  return PMG_OK;
}

/* owned slabs of a distributed vector <-> full vector on rank 0 */
static int gather_to_root(pmg_context *ctx, const pmg_vector *dist, pmg_vector *full)
{
  const pmg_layout *l = &dist->lay;
  PMG_NCCL(ncclGroupStart());
  if (ctx->rank == 0) {
    for (int r = 1; r < ctx->n_ranks; ++r) {
      int lo, hi;
      pmg_host_partition(l->nz, ctx->n_ranks, r, &lo, &hi);
      const int zlo = lo * l->degree, zhi = (r == ctx->n_ranks - 1) ? l->Nz : hi * l->degree;
      PMG_NCCL_IN_GROUP(ncclRecv(full->d + l->plane * zlo, (size_t)(l->plane * (zhi - zlo)), ncclDouble, r, ctx->comm, ctx->stream));
    }
  } else {
    PMG_NCCL_IN_GROUP(ncclSend(dist->d + l->plane * (l->z_own_lo - l->z0), (size_t)(l->plane * (l->z_own_hi - l->z_own_lo)), ncclDouble, 0, ctx->comm, ctx->stream));
  }
  PMG_NCCL(ncclGroupEnd());
  pmg_count_launch(1);
  if (ctx->rank == 0)
    PMG_CHECK(pmgk_copy(full->d + l->plane * l->z_own_lo, dist->d + l->plane * (l->z_own_lo - l->z0),
                        l->plane * (l->z_own_hi - l->z_own_lo), ctx->stream));
  return PMG_OK;
}

/* dist(owned) += full slab */
static int scatter_add_from_root(pmg_context *ctx, const pmg_vector *full, pmg_vector *dist)
{
  const pmg_layout *l = &dist->lay;
  double *tmp = NULL;
  const int64_t n_own = l->plane * (l->z_own_hi - l->z_own_lo);
  if (ctx->rank != 0) PMG_CUDA(cudaMallocAsync((void **)&tmp, sizeof(double) * (size_t)n_own, ctx->stream));
  PMG_NCCL(ncclGroupStart());
  if (ctx->rank == 0) {
    for (int r = 1; r < ctx->n_ranks; ++r) {
      int lo, hi;
      pmg_host_partition(l->nz, ctx->n_ranks, r, &lo, &hi);
      const int zlo = lo * l->degree, zhi = (r == ctx->n_ranks - 1) ? l->Nz : hi * l->degree;
      PMG_NCCL_IN_GROUP(ncclSend(full->d + l->plane * zlo, (size_t)(l->plane * (zhi - zlo)), ncclDouble, r, ctx->comm, ctx->stream));
    }
  } else {
    PMG_NCCL_IN_GROUP(ncclRecv(tmp, (size_t)n_own, ncclDouble, 0, ctx->comm, ctx->stream));
  }
  PMG_NCCL(ncclGroupEnd());
  pmg_count_launch(1);
  double *own = dist->d + l->plane * (l->z_own_lo - l->z0);
  if (ctx->rank == 0) PMG_CHECK(pmgk_axpby(own, 1.0, own, 1.0, full->d + l->plane * l->z_own_lo, n_own, ctx->stream));
  else {
    PMG_CHECK(pmgk_axpby(own, 1.0, own, 1.0, tmp, n_own, ctx->stream));
    PMG_CUDA(cudaFreeAsync(tmp, ctx->stream));
  }
  return PMG_OK;
}

int pmg_transfer_prolongate_and_add(const pmg_transfer *t, pmg_vector *dst_fine, const pmg_vector *src_coarse)
{
  if (dst_fine) PMG_CHECK(pmg_enter(dst_fine->ctx));
  PMG_CHECK(check(t, dst_fine, src_coarse));
  pmg_context *ctx = t->ctx;
  if (t->gather_buf) {
    pmg_vector *full = t->gather_buf;
    if (ctx->rank == 0) {
      pmgk_level fl = t->fine->lv;
      fl.z0 = 0; fl.nzl = t->fine->lay.Nz; fl.cz_lo = 0; fl.cz_hi = t->fine->lay.nz; fl.z_own_lo = 0; fl.z_own_hi = t->fine->lay.Nz;
      PMG_CHECK(pmgk_set(full->d, 0.0, full->lay.n_local, ctx->stream));
      PMG_CHECK(pmgk_prolongate_and_add(t->kind, &t->coarse->lv, &fl, t->d_P, full->d, src_coarse->d, ctx->stream));
    }
    return scatter_add_from_root(ctx, full, dst_fine);
  }
  if (!t->coarse->lay.active) return PMG_OK;
  PMG_CHECK(pmg_halo_update(ctx, &src_coarse->lay, src_coarse->d)); /* src.update_ghost_values() (:779) */
  return pmgk_prolongate_and_add(t->kind, &t->coarse->lv, &t->fine->lv, t->d_P, dst_fine->d, src_coarse->d, ctx->stream);
}

int pmg_transfer_restrict_and_add(const pmg_transfer *t, pmg_vector *dst_coarse, const pmg_vector *src_fine)
{
  if (dst_coarse) PMG_CHECK(pmg_enter(dst_coarse->ctx));
  PMG_CHECK(check(t, src_fine, dst_coarse));
  pmg_context *ctx = t->ctx;
  if (t->gather_buf) {
    pmg_vector *full = t->gather_buf;
    PMG_CHECK(gather_to_root(ctx, src_fine, full));
    if (ctx->rank == 0) {
      pmgk_level fl = t->fine->lv;
      fl.z0 = 0; fl.nzl = t->fine->lay.Nz; fl.cz_lo = 0; fl.cz_hi = t->fine->lay.nz; fl.z_own_lo = 0; fl.z_own_hi = t->fine->lay.Nz;
      PMG_CHECK(pmgk_restrict_and_add(t->kind, &t->coarse->lv, &fl, t->d_P, dst_coarse->d, full->d, t->d_scratch, ctx->stream));
    }
    return PMG_OK;
  }
  if (!t->coarse->lay.active) return PMG_OK;
  const int distributed = (ctx->has_comm && !t->coarse->lay.gathered);
  if (distributed) PMG_CHECK(pmg_vector_zero_out_ghost_values(dst_coarse));
  PMG_CHECK(pmgk_restrict_and_add(t->kind, &t->coarse->lv, &t->fine->lv, t->d_P, dst_coarse->d, src_fine->d, t->d_scratch, ctx->stream));
  if (distributed) PMG_CHECK(pmg_vector_compress_add(dst_coarse)); /* vec_coarse.compress(add) (:875) */
  return PMG_OK;
}
