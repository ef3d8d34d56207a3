/*
 * pmg_core.c -- context, z-slab layouts, distributed vector and halo exchange (host C).
 *
 * Rebuilds, for one 8xB200 box, what the reference takes from deal.II:
 *   LinearAlgebra::distributed::Vector<double, MemorySpace::Default> + Utilities::MPI::Partitioner
 *   (used at reference include/base/portable_laplace_operator_base.h:23-25,58-59 and in
 *   include/operators/portable_laplace_operator.h:635-661,713-716).
 * B200-first differences (DESIGN.md "Distributed vector"):
 *   - the structured box is cut into z-slabs of whole cell layers, so every halo is a contiguous
 *     range of x-y planes of the lexicographic vector: no pack/unpack kernels, no index lists;
 *   - ghost region = one cell layer (p planes) below + one plane above the owned planes; the
 *     apply kernel recomputes the neighbour's boundary cell layer instead of doing the
 *     reference's compress(add) after every cell loop, which is what lets the smoother update be
 *     fused into the apply (one exchange per operator application instead of two);
 *   - exchange = ncclSend/ncclRecv pairs over NVLink inside one group on the compute stream,
 *     scalar reductions = ncclAllReduce on a device scalar.
 */
#include "pmg_internal.h"
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char tls_error[512];
static int64_t g_launches = 0;

void pmg_set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
  va_end(ap);
}

void pmg_count_launch(int n) { __atomic_fetch_add(&g_launches, (int64_t)n, __ATOMIC_RELAXED); }

int pmg_enter(const pmg_context *ctx)
{
  if (!ctx) return PMG_ERR_ARG;
  int d = -1;
  if (cudaGetDevice(&d) == cudaSuccess && d == ctx->device) return PMG_OK;
  PMG_CUDA(cudaSetDevice(ctx->device));
  return PMG_OK;
}

const char *pmg_last_error(void) { return tls_error; }
const char *pmg_version(void) { return "portable-multigrid_b200 0.1 (sm_100a)"; }

/* ---- context -------------------------------------------------------------- */
static int context_init(pmg_context *ctx, const void *nccl_id);

static int context_common(pmg_context **out, int device, int rank, int n_ranks, const void *nccl_id)
{
  if (!out || n_ranks < 1 || rank < 0 || rank >= n_ranks) { pmg_set_error("pmg_context_create: bad arguments"); return PMG_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    pmg_set_error("no CUDA device: this library has no CPU fallback");
    return PMG_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { pmg_set_error("device %d out of range (%d devices)", device, ndev); return PMG_ERR_ARG; }
  if (n_ranks > 1 && !nccl_id) { pmg_set_error("distributed context needs an ncclUniqueId"); return PMG_ERR_ARG; }
  PMG_CUDA(cudaSetDevice(device));
  pmg_context *ctx = (pmg_context *)calloc(1, sizeof(*ctx));
  if (!ctx) return PMG_ERR_NOMEM;
  ctx->device = device; ctx->rank = rank; ctx->n_ranks = n_ranks;
  ctx->coarse_threshold = 262144;
  const int rc = context_init(ctx, nccl_id);
  if (rc != PMG_OK) { pmg_context_destroy(ctx); return rc; } /* destroy releases whatever was created */
  *out = ctx;
  return PMG_OK;
}

static int context_init(pmg_context *ctx, const void *nccl_id)
{
  const int rank = ctx->rank, n_ranks = ctx->n_ranks;
  PMG_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  PMG_CUDA(cudaMalloc((void **)&ctx->work, sizeof(double) * (size_t)pmgk_dot_work_doubles()));
  PMG_CUDA(cudaMalloc((void **)&ctx->scalars, sizeof(double) * 64));
  PMG_CUDA(cudaMemset(ctx->scalars, 0, sizeof(double) * 64));
  PMG_CUDA(cudaMallocHost((void **)&ctx->h_scalars, sizeof(double) * 64));
  ctx->sm_count = pmgk_device_sm_count();
  if (n_ranks > 1) {
    ncclUniqueId id;
    memcpy(&id, nccl_id, sizeof(id));
    PMG_NCCL(ncclCommInitRank(&ctx->comm, n_ranks, id, rank));
    ctx->has_comm = 1;
    PMG_CHECK(pmg_p2p_init(ctx)); /* ghost planes over NVLink peer memory when CUDA IPC works on every rank (pmg_p2p.c) */
    /* PMG_HALO_OVERLAP=1: halo exchange on its own stream / communicator, overlapped with the interior z-chunks of the apply
       (pmg_operator.c).  OFF by default: measured on 2 B200s (profiles/r01_halo_overlap_2gpu.txt) the split launch is slower
       -- fused step 0.321 against 0.271 ms, V-cycle 9.52 against 7.93 ms: the apply kernel fills every SM, so the NCCL kernel
       only gets in when the interior launch drains, and the second launch + two event hops are added on top. */
    const char *ov = getenv("PMG_HALO_OVERLAP");
    if (ov && atoi(ov) != 0) {
      PMG_NCCL(ncclCommSplit(ctx->comm, 0, rank, &ctx->halo_comm, NULL));
      /* highest priority: the exchange's kernels take the first CTA slots the running apply frees */
      int prio_lo = 0, prio_hi = 0;
      PMG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      PMG_CUDA(cudaStreamCreateWithPriority(&ctx->halo_stream, cudaStreamNonBlocking, prio_hi));
      PMG_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
      PMG_CUDA(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
      ctx->overlap = 1;
    }
  }
  return PMG_OK;
}

int pmg_context_create(pmg_context **ctx, int device) { return context_common(ctx, device, 0, 1, NULL); }

int pmg_context_create_distributed(pmg_context **ctx, int device, int rank, int n_ranks, const void *nccl_id)
{
  return context_common(ctx, device, rank, n_ranks, nccl_id);
}

int pmg_nccl_unique_id(void *out128)
{
  if (!out128) return PMG_ERR_ARG;
  ncclUniqueId id;
  PMG_NCCL(ncclGetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return PMG_OK;
}

int pmg_context_destroy(pmg_context *ctx)
{
  if (!ctx) return PMG_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->overlap) {
    cudaStreamSynchronize(ctx->halo_stream);
    ncclCommDestroy(ctx->halo_comm);
    cudaEventDestroy(ctx->ev_ready); cudaEventDestroy(ctx->ev_halo);
    cudaStreamDestroy(ctx->halo_stream);
  }
  if (ctx->has_comm) pmg_p2p_shutdown(ctx);
  if (ctx->has_comm) ncclCommDestroy(ctx->comm);
  cudaFree(ctx->work); cudaFree(ctx->scalars); cudaFreeHost(ctx->h_scalars);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  free(ctx);
  return PMG_OK;
}

int pmg_sync(pmg_context *ctx)
{
  if (!ctx) return PMG_ERR_ARG;
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->p2p.enabled && ctx->p2p.fused && ctx->p2p.mailbox) {
    /* the fused ghost exchange bounds its waits for the neighbour GPUs (csrc/pmg_apply_plane_launch.h: ~4 s); a wait that gave
       up left the error word set: the results since then are not to be trusted, and the caller hears about it here */
    uint64_t err = 0;
    PMG_CUDA(cudaMemcpy(&err, ctx->p2p.mailbox + 14 /* PMG_FUSED_ERROR */, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) {
      pmg_set_error("fused ghost exchange: a neighbour rank did not arrive within the time limit (rank %d)", ctx->rank);
      return PMG_ERR_CUDA;
    }
  }
  return PMG_OK;
}

int pmg_context_rank(const pmg_context *ctx, int *rank, int *n_ranks)
{
  if (!ctx) return PMG_ERR_ARG;
  if (rank) *rank = ctx->rank;
  if (n_ranks) *n_ranks = ctx->n_ranks;
  return PMG_OK;
}

int pmg_context_set_coarse_threshold(pmg_context *ctx, int64_t n_dofs)
{
  if (!ctx || n_dofs < 0) return PMG_ERR_ARG;
  ctx->coarse_threshold = n_dofs;
  return PMG_OK;
}

void *pmg_context_stream(pmg_context *ctx) { return ctx ? (void *)ctx->stream : NULL; }
/* kernels and collectives this PROCESS has enqueued (all contexts, all threads; the counter is atomic) */
int64_t pmg_context_fused_halo_count(const pmg_context *ctx) { return ctx ? ctx->p2p.n_fused : 0; }
int64_t pmg_context_launch_count(const pmg_context *ctx) { (void)ctx; return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

/* ---- partition -------------------------------------------------------------- */
/* A level is cut into n_ranks z-slabs of nz / n_ranks cell layers iff n_ranks divides nz.
   Refinement doubles nz, so a rank's fine slab is exactly the refinement of its coarse slab and
   the h-transfer between two distributed levels never crosses ranks; p-levels share the mesh and
   therefore the slabs.  Levels whose layer count is not divisible (the coarse end of every
   hierarchy) live on rank 0 only. */
int pmg_host_partition(int nz, int n_ranks, int rank, int *cz_lo, int *cz_hi)
{
  if (nz < 1 || n_ranks < 1 || rank < 0 || rank >= n_ranks || !cz_lo || !cz_hi) return PMG_ERR_ARG;
  if (n_ranks == 1) { *cz_lo = 0; *cz_hi = nz; return PMG_OK; }
  if (nz % n_ranks != 0) { /* not distributable: everything on rank 0 */
    *cz_lo = 0; *cz_hi = (rank == 0) ? nz : 0;
    return 1;
  }
  const int per = nz / n_ranks;
  *cz_lo = per * rank;
  *cz_hi = per * (rank + 1);
  return PMG_OK;
}

int pmg_layout_make(pmg_context *ctx, int dim, int degree, int nx, int ny, int nz, pmg_layout *lay)
{
  memset(lay, 0, sizeof(*lay));
  if (dim == 2) nz = 1; /* 2-D: a single dof plane; nz = 1 is a placeholder, the level is never cut into slabs */
  lay->nx = nx; lay->ny = ny; lay->nz = nz; lay->degree = degree;
  lay->Nx = nx * degree + 1; lay->Ny = ny * degree + 1; lay->Nz = (dim == 2) ? 1 : nz * degree + 1;
  lay->plane = (int64_t)lay->Nx * lay->Ny;
  lay->n_global = lay->plane * lay->Nz;
  lay->lower = lay->upper = -1;
  const int R = ctx->n_ranks, r = ctx->rank;
  int cz_lo = 0, cz_hi = nz;
  int distributed = 0;
  if (R > 1 && dim == 3 && lay->n_global >= ctx->coarse_threshold) {
    const int rc = pmg_host_partition(nz, R, r, &cz_lo, &cz_hi);
    if (rc < 0) return rc;
    distributed = (rc == 0);
  }
  if (!distributed) {
    lay->gathered = (R > 1);
    lay->active = (r == 0);
    lay->cz_lo = 0; lay->cz_hi = lay->active ? nz : 0;
    lay->z0 = 0; lay->nzl = lay->active ? lay->Nz : 0;
    lay->z_own_lo = 0; lay->z_own_hi = lay->active ? lay->Nz : 0;
  } else {
    lay->active = 1;
    lay->cz_lo = cz_lo; lay->cz_hi = cz_hi;
    lay->z_own_lo = cz_lo * degree;
    lay->z_own_hi = (r == R - 1) ? lay->Nz : cz_hi * degree;
    lay->z0 = (r == 0) ? 0 : cz_lo * degree - degree;
    const int z_end = (r == R - 1) ? lay->Nz : cz_hi * degree + 1;
    lay->nzl = z_end - lay->z0;
    lay->lower = (r > 0) ? r - 1 : -1;
    lay->upper = (r < R - 1) ? r + 1 : -1;
  }
  lay->n_local = lay->plane * lay->nzl;
  return PMG_OK;
}

int pmg_layout_same(const pmg_layout *a, const pmg_layout *b)
{
  return a->nx == b->nx && a->ny == b->ny && a->nz == b->nz && a->degree == b->degree && a->z0 == b->z0 &&
         a->nzl == b->nzl && a->gathered == b->gathered;
}

/* ---- vectors ---------------------------------------------------------------- */
int pmg_vector_create_layout(pmg_context *ctx, const pmg_layout *lay, pmg_vector **out)
{
  if (ctx) PMG_CHECK(pmg_enter(ctx));
  pmg_vector *v = (pmg_vector *)calloc(1, sizeof(*v));
  if (!v) return PMG_ERR_NOMEM;
  v->ctx = ctx; v->lay = *lay;
  if (lay->n_local > 0) {
    v->d = pmg_p2p_acquire(ctx, lay); /* a released array of this level that the neighbours already have mapped */
    const int fresh = (v->d == NULL);
    /* + 16 bytes: the apply kernel's bulk row copies fetch whole 16-byte granules (csrc/pmg_apply_sweep.h) */
    if (fresh && cudaMalloc((void **)&v->d, sizeof(double) * ((size_t)lay->n_local + 2)) != cudaSuccess) {
      pmg_set_error("cudaMalloc of %lld doubles failed", (long long)lay->n_local);
      free(v);
      return PMG_ERR_NOMEM;
    }
    int rc = pmgk_set(v->d, 0.0, lay->n_local, ctx->stream);
    if (!rc && fresh) rc = pmg_p2p_register(ctx, lay, v->d); /* collective: handle exchange with the slab neighbours */
    if (rc) { if (!pmg_p2p_release(ctx, v->d)) cudaFree(v->d); free(v); return rc; }
  }
  *out = v;
  return PMG_OK;
}

int pmg_vector_destroy(pmg_vector *v)
{
  if (!v) return PMG_OK;
  cudaStreamSynchronize(v->ctx->stream);
  if (!v->borrowed && !pmg_p2p_release(v->ctx, v->d)) cudaFree(v->d);
  free(v);
  return PMG_OK;
}

/* A vector over device memory the caller owns (the reference adapter hands in the storage of a
   LinearAlgebra::distributed::Vector<double, MemorySpace::Default>: no copy per vmult).  The array holds the rank's
   stored planes -- pmg_vector_local_range() -- in lexicographic order; it must stay valid while the handle lives. */
int pmg_vector_wrap(const pmg_vector *like, double *device_values, pmg_vector **out)
{
  if (!like || !device_values || !out) return PMG_ERR_ARG;
  if (((uintptr_t)device_values & 15) != 0) { pmg_set_error("pmg_vector_wrap: the array is not 16-byte aligned"); return PMG_ERR_ARG; }
  pmg_vector *v = (pmg_vector *)calloc(1, sizeof(*v));
  if (!v) return PMG_ERR_NOMEM;
  v->ctx = like->ctx; v->lay = like->lay; v->d = device_values; v->borrowed = 1;
  *out = v;
  return PMG_OK;
}

/* The rank's part of the vector (get_vector_partitioner analogue, include/base/portable_laplace_operator_base.h:58-59):
   stored dof planes [z0, z0 + n_planes) of plane_size dofs each, of which [z_own_lo, z_own_hi) are owned, the others ghosts. */
int pmg_vector_local_range(const pmg_vector *v, int64_t *plane_size, int *z0, int *n_planes, int *z_own_lo, int *z_own_hi)
{
  if (!v) return PMG_ERR_ARG;
  const pmg_layout *l = &v->lay;
  if (plane_size) *plane_size = l->plane;
  if (z0) *z0 = l->active ? l->z0 : 0;
  if (n_planes) *n_planes = l->active ? l->nzl : 0;
  if (z_own_lo) *z_own_lo = l->active ? l->z_own_lo : 0;
  if (z_own_hi) *z_own_hi = l->active ? l->z_own_hi : 0;
  return PMG_OK;
}

int pmg_vector_size(const pmg_vector *v, int64_t *n) { if (!v || !n) return PMG_ERR_ARG; *n = v->lay.n_global; return PMG_OK; }

int pmg_vector_locally_owned_size(const pmg_vector *v, int64_t *n)
{
  if (!v || !n) return PMG_ERR_ARG;
  *n = v->lay.plane * (v->lay.z_own_hi - v->lay.z_own_lo);
  return PMG_OK;
}

double *pmg_vector_device_ptr(pmg_vector *v) { return v ? v->d : NULL; }

static double *owned_ptr(const pmg_vector *v) { return v->d ? v->d + v->lay.plane * (v->lay.z_own_lo - v->lay.z0) : NULL; }
static int64_t owned_n(const pmg_vector *v) { return v->lay.plane * (v->lay.z_own_hi - v->lay.z_own_lo); }

static int check_pair(const pmg_vector *a, const pmg_vector *b)
{
  if (!a || !b || a->ctx != b->ctx || !pmg_layout_same(&a->lay, &b->lay)) {
    pmg_set_error("vectors are not compatible (different level or partitioner)");
    return PMG_ERR_ARG;
  }
  return PMG_OK;
}

int pmg_vector_set(pmg_vector *v, double value)
{
  if (v) PMG_CHECK(pmg_enter(v->ctx));
  if (!v) return PMG_ERR_ARG;
  return pmgk_set(v->d, value, v->lay.n_local, v->ctx->stream);
}

int pmg_vector_copy(pmg_vector *dst, const pmg_vector *src)
{
  if (dst) PMG_CHECK(pmg_enter(dst->ctx));
  PMG_CHECK(check_pair(dst, src));
  return pmgk_copy(dst->d, src->d, dst->lay.n_local, dst->ctx->stream);
}

int pmg_vector_scale(pmg_vector *v, double a)
{
  if (v) PMG_CHECK(pmg_enter(v->ctx));
  if (!v) return PMG_ERR_ARG;
  return pmgk_scale(v->d, a, v->lay.n_local, v->ctx->stream);
}

int pmg_vector_add(pmg_vector *v, double a, const pmg_vector *x)
{
  if (v) PMG_CHECK(pmg_enter(v->ctx));
  PMG_CHECK(check_pair(v, x));
  return pmgk_axpby(v->d, 1.0, v->d, a, x->d, v->lay.n_local, v->ctx->stream);
}

int pmg_vector_sadd(pmg_vector *v, double s, double a, const pmg_vector *x)
{
  if (v) PMG_CHECK(pmg_enter(v->ctx));
  PMG_CHECK(check_pair(v, x));
  return pmgk_axpby(v->d, s, v->d, a, x->d, v->lay.n_local, v->ctx->stream);
}

int pmg_allreduce_sum(pmg_context *ctx, double *dev_scalar, int count)
{
  if (ctx->has_comm) {
    PMG_NCCL(ncclAllReduce(dev_scalar, dev_scalar, (size_t)count, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    pmg_count_launch(1);
  }
  return PMG_OK;
}

static int fetch_scalar(pmg_context *ctx, int slot, double *result)
{
  PMG_CUDA(cudaMemcpyAsync(ctx->h_scalars + slot, ctx->scalars + slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  *result = ctx->h_scalars[slot];
  return PMG_OK;
}

/* x . y into device scalar `slot` of the context (summed over the ranks), no host round trip: CG keeps its scalars on the
   device and reads back one number per iteration (the residual norm it has to compare with the tolerance) */
int pmg_vector_dot_device(const pmg_vector *x, const pmg_vector *y, int slot)
{
  if (!x || !y || slot < 0 || slot >= 56) return PMG_ERR_ARG;
  pmg_context *ctx = x->ctx;
  if (x->lay.active) PMG_CHECK(pmgk_dot(owned_ptr(x), owned_ptr(y), owned_n(x), ctx->scalars + slot, ctx->work, ctx->stream));
  else PMG_CUDA(cudaMemsetAsync(ctx->scalars + slot, 0, sizeof(double), ctx->stream));
  return pmg_allreduce_sum(ctx, ctx->scalars + slot, 1);
}

int pmg_vector_dot(const pmg_vector *x, const pmg_vector *y, double *result)
{
  if (x) PMG_CHECK(pmg_enter(x->ctx));
  PMG_CHECK(check_pair(x, y));
  if (!result) return PMG_ERR_ARG;
  pmg_context *ctx = x->ctx;
  /* every rank takes part (a level gathered on rank 0 contributes 0 elsewhere), so all ranks see the same
     value and stay in lock-step in the iterative solvers */
  PMG_CHECK(pmgk_dot(owned_ptr(x), owned_ptr(y), owned_n(x), ctx->scalars, ctx->work, ctx->stream));
  PMG_CHECK(pmg_allreduce_sum(ctx, ctx->scalars, 1));
  return fetch_scalar(ctx, 0, result);
}

int pmg_vector_l2_norm(const pmg_vector *x, double *result)
{
  double d = 0.0;
  PMG_CHECK(pmg_vector_dot(x, x, &d));
  *result = (d > 0.0) ? __builtin_sqrt(d) : 0.0;
  return PMG_OK;
}

int pmg_vector_mean_value(const pmg_vector *x, double *result)
{
  if (!x || !result) return PMG_ERR_ARG;
  pmg_context *ctx = x->ctx;
  PMG_CHECK(pmgk_sum(owned_ptr(x), owned_n(x), ctx->scalars, ctx->work, ctx->stream));
  PMG_CHECK(pmg_allreduce_sum(ctx, ctx->scalars, 1));
  double s = 0.0;
  PMG_CHECK(fetch_scalar(ctx, 0, &s));
  *result = s / (double)x->lay.n_global;
  return PMG_OK;
}

/* ---- halo exchange ------------------------------------------------------------ */
int pmg_halo_update(pmg_context *ctx, const pmg_layout *lay, double *d)
{
  if (ctx->halo_hook && ctx->has_comm && !lay->gathered && lay->active) {
    ctx->halo_hook(ctx->halo_hook_user, 1);
    const int rc = pmg_halo_update_on(ctx, lay, d, ctx->comm, ctx->stream);
    ctx->halo_hook(ctx->halo_hook_user, 0);
    return rc;
  }
  return pmg_halo_update_on(ctx, lay, d, ctx->comm, ctx->stream);
}

int pmg_halo_update_on(pmg_context *ctx, const pmg_layout *lay, double *d, ncclComm_t comm, cudaStream_t stream)
{
  if (!ctx->has_comm || lay->gathered || !lay->active) return PMG_OK;
  {
    /* registered vectors: one push kernel over NVLink peer memory (csrc/pmg_halo.cu); anything else (wrapped arrays, boxes
       without CUDA IPC): the NCCL send / receive group below */
    int done = 0;
    PMG_CHECK(pmg_p2p_halo(ctx, lay, d, stream, &done));
    if (done) return PMG_OK;
  }
  const int p = lay->degree;
  const int64_t plane = lay->plane;
  PMG_NCCL(ncclGroupStart());
  if (lay->upper >= 0) {
    /* my top p owned planes -> upper neighbour's lower ghost layer; its first owned plane -> my upper ghost */
    PMG_NCCL_IN_GROUP(ncclSend(d + plane * (lay->z_own_hi - p - lay->z0), (size_t)(plane * p), ncclDouble, lay->upper, comm, stream));
    PMG_NCCL_IN_GROUP(ncclRecv(d + plane * (lay->z_own_hi - lay->z0), (size_t)plane, ncclDouble, lay->upper, comm, stream));
  }
  if (lay->lower >= 0) {
    PMG_NCCL_IN_GROUP(ncclSend(d + plane * (lay->z_own_lo - lay->z0), (size_t)plane, ncclDouble, lay->lower, comm, stream));
    PMG_NCCL_IN_GROUP(ncclRecv(d, (size_t)(plane * p), ncclDouble, lay->lower, comm, stream));
  }
  PMG_NCCL(ncclGroupEnd());
  pmg_count_launch(1);
  return PMG_OK;
}

int pmg_vector_update_ghost_values(pmg_vector *v)
{
  if (!v) return PMG_ERR_ARG;
  return pmg_halo_update(v->ctx, &v->lay, v->d);
}

int pmg_vector_zero_out_ghost_values(pmg_vector *v)
{
  if (!v) return PMG_ERR_ARG;
  const pmg_layout *l = &v->lay;
  if (!l->active) return PMG_OK;
  const int nlow = l->z_own_lo - l->z0, nup = (l->z0 + l->nzl) - l->z_own_hi;
  if (nlow > 0) PMG_CHECK(pmgk_set(v->d, 0.0, l->plane * nlow, v->ctx->stream));
  if (nup > 0) PMG_CHECK(pmgk_set(v->d + l->plane * (l->z_own_hi - l->z0), 0.0, l->plane * nup, v->ctx->stream));
  return PMG_OK;
}

/* ghost -> owner, summed (the reference's compress(VectorOperation::add)) */
int pmg_vector_compress_add(pmg_vector *v)
{
  if (!v) return PMG_ERR_ARG;
  pmg_context *ctx = v->ctx;
  const pmg_layout *l = &v->lay;
  if (!ctx->has_comm || l->gathered || !l->active) return PMG_OK;
  const int p = l->degree;
  const int64_t plane = l->plane;
  double *tmp = NULL;
  PMG_CUDA(cudaMallocAsync((void **)&tmp, sizeof(double) * (size_t)(plane * (p + 1)), ctx->stream));
  double *from_upper = tmp, *from_lower = tmp + plane * p;
  PMG_NCCL(ncclGroupStart());
  if (l->upper >= 0) {
    PMG_NCCL_IN_GROUP(ncclSend(v->d + plane * (l->z_own_hi - l->z0), (size_t)plane, ncclDouble, l->upper, ctx->comm, ctx->stream));
    PMG_NCCL_IN_GROUP(ncclRecv(from_upper, (size_t)(plane * p), ncclDouble, l->upper, ctx->comm, ctx->stream));
  }
  if (l->lower >= 0) {
    PMG_NCCL_IN_GROUP(ncclSend(v->d, (size_t)(plane * p), ncclDouble, l->lower, ctx->comm, ctx->stream));
    PMG_NCCL_IN_GROUP(ncclRecv(from_lower, (size_t)plane, ncclDouble, l->lower, ctx->comm, ctx->stream));
  }
  PMG_NCCL(ncclGroupEnd());
  pmg_count_launch(1);
  if (l->upper >= 0) {
    double *dst = v->d + plane * (l->z_own_hi - p - l->z0);
    PMG_CHECK(pmgk_axpby(dst, 1.0, dst, 1.0, from_upper, plane * p, ctx->stream));
  }
  if (l->lower >= 0) {
    double *dst = v->d + plane * (l->z_own_lo - l->z0);
    PMG_CHECK(pmgk_axpby(dst, 1.0, dst, 1.0, from_lower, plane, ctx->stream));
  }
  PMG_CUDA(cudaFreeAsync(tmp, ctx->stream));
  return PMG_OK;
}

/* ---- host import / export of the rank's own part: the owned planes only (ghost planes are not touched: an operator
   application refreshes them); host_owned holds plane_size * (z_own_hi - z_own_lo) doubles.  No collective, no staging of the
   global vector: this is what a distributed caller uses (pmg_vector_export_host assembles the GLOBAL vector on every rank). */
int pmg_vector_import_owned(pmg_vector *v, const double *host_owned)
{
  if (!v || !host_owned) return PMG_ERR_ARG;
  const pmg_layout *l = &v->lay;
  if (!l->active) return PMG_OK;
  const int64_t n = l->plane * (l->z_own_hi - l->z_own_lo);
  PMG_CUDA(cudaMemcpyAsync(v->d + l->plane * (l->z_own_lo - l->z0), host_owned, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, v->ctx->stream));
  return PMG_OK; /* stream-ordered: host_owned must stay unchanged until pmg_context_sync() or the next blocking call */
}

int pmg_vector_export_owned(const pmg_vector *v, double *host_owned)
{
  if (!v || !host_owned) return PMG_ERR_ARG;
  const pmg_layout *l = &v->lay;
  if (!l->active) return PMG_OK;
  const int64_t n = l->plane * (l->z_own_hi - l->z_own_lo);
  PMG_CUDA(cudaMemcpyAsync(host_owned, v->d + l->plane * (l->z_own_lo - l->z0), sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, v->ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return PMG_OK;
}

/* every stored plane, ghosts included (tests of the ghost operations) */
int pmg_vector_export_local(const pmg_vector *v, double *host_local)
{
  if (!v || !host_local) return PMG_ERR_ARG;
  const pmg_layout *l = &v->lay;
  if (!l->active) return PMG_OK;
  PMG_CUDA(cudaMemcpyAsync(host_local, v->d, sizeof(double) * (size_t)l->n_local, cudaMemcpyDeviceToHost, v->ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return PMG_OK;
}

int pmg_vector_import_local(pmg_vector *v, const double *host_local)
{
  if (!v || !host_local) return PMG_ERR_ARG;
  const pmg_layout *l = &v->lay;
  if (!l->active) return PMG_OK;
  PMG_CUDA(cudaMemcpyAsync(v->d, host_local, sizeof(double) * (size_t)l->n_local, cudaMemcpyHostToDevice, v->ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return PMG_OK;
}

/* ---- host import / export (global lexicographic order) -------------------------- */
int pmg_vector_import_host(pmg_vector *v, const double *host)
{
  if (!v || !host) return PMG_ERR_ARG;
  const pmg_layout *l = &v->lay;
  if (!l->active) return PMG_OK;
  /* every stored plane (owned and ghost) is filled from the global array */
  PMG_CUDA(cudaMemcpyAsync(v->d, host + l->plane * l->z0, sizeof(double) * (size_t)l->n_local, cudaMemcpyHostToDevice, v->ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(v->ctx->stream)); /* the caller may reuse `host` (pinned or not) as soon as this returns */
  return PMG_OK;
}

int pmg_vector_export_host(const pmg_vector *v, double *host)
{
  if (!v || !host) return PMG_ERR_ARG;
  pmg_context *ctx = v->ctx;
  const pmg_layout *l = &v->lay;
  if (!ctx->has_comm) {
    PMG_CUDA(cudaMemcpyAsync(host, v->d, sizeof(double) * (size_t)l->n_global, cudaMemcpyDeviceToHost, ctx->stream));
    PMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return PMG_OK;
  }
  /* all ranks assemble the full vector: allgather of owned slabs through a device staging buffer */
  double *full = NULL;
  PMG_CUDA(cudaMallocAsync((void **)&full, sizeof(double) * (size_t)l->n_global, ctx->stream));
  PMG_CUDA(cudaMemsetAsync(full, 0, sizeof(double) * (size_t)l->n_global, ctx->stream));
  if (l->active) {
    const int64_t n = l->plane * (l->z_own_hi - l->z_own_lo);
    PMG_CUDA(cudaMemcpyAsync(full + l->plane * l->z_own_lo, v->d + l->plane * (l->z_own_lo - l->z0), sizeof(double) * (size_t)n,
                             cudaMemcpyDeviceToDevice, ctx->stream));
  }
  /* owned slabs are disjoint and the rest is zero: a sum-allreduce assembles the vector */
  PMG_NCCL(ncclAllReduce(full, full, (size_t)l->n_global, ncclDouble, ncclSum, ctx->comm, ctx->stream));
  pmg_count_launch(1);
  PMG_CUDA(cudaMemcpyAsync(host, full, sizeof(double) * (size_t)l->n_global, cudaMemcpyDeviceToHost, ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  PMG_CUDA(cudaFreeAsync(full, ctx->stream));
  return PMG_OK;
}
