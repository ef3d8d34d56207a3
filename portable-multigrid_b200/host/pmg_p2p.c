/*
 * pmg_p2p.c -- peer mapping of the neighbours' vectors for the ghost-plane exchange over NVLink (csrc/pmg_halo.cu).
 *
 * One process per GPU: a rank reaches its slab neighbours' memory through CUDA IPC handles.  At context creation every rank
 * exports a mailbox (six flag words); for every slab-distributed vector created afterwards the ranks exchange the IPC handle of
 * its storage with rank - 1 and rank + 1 (two 64-byte messages over the context's NCCL communicator: set-up, not hot path) and
 * map the neighbours' copies.  pmg_halo_update_on() then finds the mapping of the array it is given and launches the push
 * kernel instead of the ncclSend / ncclRecv group.
 *
 * Everything here is collective over the ranks of the context and happens in the same order on all of them (the library is
 * used SPMD: every rank creates the same vectors in the same order).  If CUDA IPC is not usable on the box (containers without
 * a shared IPC namespace, PMG_P2P_HALO=0), every rank notices at context creation, the feature stays off and the NCCL path is
 * used: the result is the same, only slower.
 */
#include "pmg_internal.h"
#include <stdlib.h>
#include <string.h>

/* stand-alone exchanges switch from the NCCL group to the push kernel at this size (0 = never).  Measured on 8 B200s, config 5
   (profiles/r02_driver_c5_q5_160cells_8gpu*.txt): 25 MB per exchange 0.43 ms with the NCCL group, 0.18 ms with the push kernel;
   at 2.5 MB the two are level (profiles/r02_halo_p2p_2gpu.txt), and at 6-10 MB between fused launches (config 2 on 8 GPUs) the
   push kernel's two flag rounds cost more than they save (7.89 against 7.67 ms per cycle, and a 2.9 ms stall in the eager
   per-level profile: profiles/r02_bench_8gpu_push_kernel_from_4MB.json) -- hence 16 MB */
#define PMG_P2P_MIN_BYTES_DEFAULT 16000000
#define P2P_MSG_BYTES 128 /* cudaIpcMemHandle_t (64 bytes) + the first stored plane of the sender's slab + a validity word */

typedef struct p2p_msg {
  cudaIpcMemHandle_t handle;
  int z0;
  int ok;
  char pad[P2P_MSG_BYTES - sizeof(cudaIpcMemHandle_t) - 2 * sizeof(int)];
} p2p_msg;

/* my message to both neighbours, theirs to me (rank - 1 = lower, rank + 1 = upper; a missing neighbour leaves its slot zero) */
static int exchange_with_neighbours(pmg_context *ctx, const p2p_msg *mine, p2p_msg *from_lower, p2p_msg *from_upper)
{
  const int lower = ctx->rank - 1, upper = (ctx->rank + 1 < ctx->n_ranks) ? ctx->rank + 1 : -1;
  char *dev = (char *)ctx->p2p.msg_dev; /* 3 messages: mine, from lower, from upper */
  memset(from_lower, 0, sizeof(*from_lower));
  memset(from_upper, 0, sizeof(*from_upper));
  PMG_CUDA(cudaMemcpyAsync(dev, mine, P2P_MSG_BYTES, cudaMemcpyHostToDevice, ctx->stream));
  PMG_CUDA(cudaMemsetAsync(dev + P2P_MSG_BYTES, 0, 2 * P2P_MSG_BYTES, ctx->stream));
  PMG_NCCL(ncclGroupStart());
  if (lower >= 0) {
    PMG_NCCL_IN_GROUP(ncclSend(dev, P2P_MSG_BYTES, ncclChar, lower, ctx->comm, ctx->stream));
    PMG_NCCL_IN_GROUP(ncclRecv(dev + P2P_MSG_BYTES, P2P_MSG_BYTES, ncclChar, lower, ctx->comm, ctx->stream));
  }
  if (upper >= 0) {
    PMG_NCCL_IN_GROUP(ncclSend(dev, P2P_MSG_BYTES, ncclChar, upper, ctx->comm, ctx->stream));
    PMG_NCCL_IN_GROUP(ncclRecv(dev + 2 * P2P_MSG_BYTES, P2P_MSG_BYTES, ncclChar, upper, ctx->comm, ctx->stream));
  }
  PMG_NCCL(ncclGroupEnd());
  PMG_CUDA(cudaMemcpyAsync(from_lower, dev + P2P_MSG_BYTES, P2P_MSG_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
  PMG_CUDA(cudaMemcpyAsync(from_upper, dev + 2 * P2P_MSG_BYTES, P2P_MSG_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  return PMG_OK;
}

/* 1 on every rank iff `flag` is non-zero on every rank */
static int all_ranks_agree(pmg_context *ctx, int flag, int *all)
{
  double v = flag ? 1.0 : 0.0;
  PMG_CUDA(cudaMemcpyAsync(ctx->scalars + 60, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  PMG_NCCL(ncclAllReduce(ctx->scalars + 60, ctx->scalars + 60, 1, ncclDouble, ncclMin, ctx->comm, ctx->stream));
  PMG_CUDA(cudaMemcpyAsync(&v, ctx->scalars + 60, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PMG_CUDA(cudaStreamSynchronize(ctx->stream));
  *all = (v > 0.5);
  return PMG_OK;
}

/* context creation (distributed): the mailbox and its mapping on the neighbours.  Leaves ctx->p2p.enabled = 0 when IPC does
   not work on every rank. */
int pmg_p2p_init(pmg_context *ctx)
{
  pmg_p2p *pp = &ctx->p2p;
  memset(pp, 0, sizeof(*pp));
  if (!ctx->has_comm || ctx->n_ranks < 2) return PMG_OK;
  /* opt-in (PMG_P2P_HALO=1).  Measured on 2 B200s (profiles/r02_halo_p2p_2gpu.txt, C2 per GPU): V-cycle 7.05 ms with the push
     kernel against 6.76 ms with the NCCL group -- two flag round trips + the launch cost as much as NCCL's protocol, and neither
     overlaps with the apply.  Kept as the base of a push fused into the apply kernel's epilogue. */
  const char *env = getenv("PMG_P2P_HALO"), *envf = getenv("PMG_FUSED_HALO");
  const int explicit_push = (env && atoi(env) != 0);
  /* The fused push (the smoother's applies store their boundary planes into the neighbours' ghost planes themselves,
     csrc/pmg_apply_plane_launch.h) is on by default: it removes the exchange between two Chebyshev steps altogether. */
  const int fused = !(envf && atoi(envf) == 0);
  const int wanted = explicit_push || fused;
  PMG_CUDA(cudaMalloc(&pp->msg_dev, 3 * P2P_MSG_BYTES));
  PMG_CUDA(cudaMalloc((void **)&pp->mailbox, 16 * sizeof(uint64_t)));
  PMG_CUDA(cudaMemset(pp->mailbox, 0, 16 * sizeof(uint64_t)));
  p2p_msg mine, lo, up;
  memset(&mine, 0, sizeof(mine));
  mine.ok = wanted && cudaIpcGetMemHandle(&mine.handle, pp->mailbox) == cudaSuccess;
  if (!mine.ok) cudaGetLastError();
  PMG_CHECK(exchange_with_neighbours(ctx, &mine, &lo, &up));
  int ok = mine.ok;
  if (ok && ctx->rank > 0) {
    if (!lo.ok || cudaIpcOpenMemHandle((void **)&pp->mb_lower, lo.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; pp->mb_lower = NULL; cudaGetLastError(); }
  }
  if (ok && ctx->rank + 1 < ctx->n_ranks) {
    if (!up.ok || cudaIpcOpenMemHandle((void **)&pp->mb_upper, up.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; pp->mb_upper = NULL; cudaGetLastError(); }
  }
  int all = 0;
  PMG_CHECK(all_ranks_agree(ctx, ok, &all));
  pp->enabled = all;
  pp->explicit_push = all && explicit_push;
  {
    /* large exchanges are bandwidth bound and the push kernel moves them at NVLink speed with 148 CTAs, small ones are latency
       bound and level with NCCL (profiles/r02_halo_p2p_2gpu.txt): PMG_P2P_MIN_BYTES sets the switch-over (bytes one rank sends
       per exchange; default: see below) */
    const char *em = getenv("PMG_P2P_MIN_BYTES");
    pp->push_min_bytes = em ? (int64_t)atof(em) : PMG_P2P_MIN_BYTES_DEFAULT;
  }
  pp->fused = all && fused;
  return PMG_OK;
}

/* context destruction (collective): unmap the neighbours' arrays, wait until every rank has done so, then free the exported
   arrays -- cudaFree of memory another process still has mapped is undefined */
void pmg_p2p_shutdown(pmg_context *ctx)
{
  pmg_p2p *pp = &ctx->p2p;
  for (int i = 0; i < pp->n_reg; ++i) {
    if (pp->reg[i].peer_lower) cudaIpcCloseMemHandle((void *)pp->reg[i].peer_lower);
    if (pp->reg[i].peer_upper) cudaIpcCloseMemHandle((void *)pp->reg[i].peer_upper);
  }
  if (pp->mb_lower) cudaIpcCloseMemHandle(pp->mb_lower);
  if (pp->mb_upper) cudaIpcCloseMemHandle(pp->mb_upper);
  if (pp->enabled) { int all = 0; all_ranks_agree(ctx, 1, &all); }
  for (int i = 0; i < pp->n_reg; ++i) cudaFree(pp->reg[i].base);
  cudaFree(pp->mailbox);
  cudaFree(pp->msg_dev);
  memset(pp, 0, sizeof(*pp));
}

/* a slab-distributed vector was created: exchange handles with the neighbours and map their copies (collective) */
int pmg_p2p_register(pmg_context *ctx, const pmg_layout *lay, double *d)
{
  pmg_p2p *pp = &ctx->p2p;
  if (!pp->enabled || lay->gathered || !lay->active || !d) return PMG_OK;
  p2p_msg mine, lo, up;
  memset(&mine, 0, sizeof(mine));
  mine.z0 = lay->z0;
  mine.ok = pp->n_reg < PMG_P2P_MAX_REG && cudaIpcGetMemHandle(&mine.handle, d) == cudaSuccess;
  if (!mine.ok) cudaGetLastError();
  PMG_CHECK(exchange_with_neighbours(ctx, &mine, &lo, &up));
  pmg_p2p_reg r;
  memset(&r, 0, sizeof(r));
  r.base = d; r.in_use = 1;
  r.key[0] = lay->Nx; r.key[1] = lay->Ny; r.key[2] = lay->Nz; r.key[3] = lay->degree;
  int ok = mine.ok;
  if (ok && lay->lower >= 0) {
    if (!lo.ok || cudaIpcOpenMemHandle((void **)&r.peer_lower, lo.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; r.peer_lower = NULL; cudaGetLastError(); }
    r.lower_z0 = lo.z0;
  }
  if (ok && lay->upper >= 0) {
    if (!up.ok || cudaIpcOpenMemHandle((void **)&r.peer_upper, up.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; r.peer_upper = NULL; cudaGetLastError(); }
    r.upper_z0 = up.z0;
  }
  int all = 0;
  PMG_CHECK(all_ranks_agree(ctx, ok, &all));
  if (!all) { /* one rank could not map: nobody uses this vector's mapping (its exchanges go through NCCL) */
    if (r.peer_lower) cudaIpcCloseMemHandle((void *)r.peer_lower);
    if (r.peer_upper) cudaIpcCloseMemHandle((void *)r.peer_upper);
    return PMG_OK;
  }
  pp->reg[pp->n_reg++] = r;
  return PMG_OK;
}

/* An exported array is never freed while the context lives (its neighbours have it mapped): a destroyed vector leaves it here,
   and the next vector of the same level takes it over, mapping included -- the ranks create and destroy vectors in the same
   order, so they all pick the same entry.  Returns the array (zero it before use) or NULL. */
double *pmg_p2p_acquire(pmg_context *ctx, const pmg_layout *lay)
{
  pmg_p2p *pp = &ctx->p2p;
  if (!pp->enabled || lay->gathered || !lay->active) return NULL;
  for (int i = 0; i < pp->n_reg; ++i) {
    pmg_p2p_reg *r = &pp->reg[i];
    if (!r->in_use && r->key[0] == lay->Nx && r->key[1] == lay->Ny && r->key[2] == lay->Nz && r->key[3] == lay->degree) {
      r->in_use = 1;
      return r->base;
    }
  }
  return NULL;
}

/* 1: the array is an exported one and stays allocated (the caller must not free it); 0: not ours */
int pmg_p2p_release(pmg_context *ctx, double *d)
{
  pmg_p2p *pp = &ctx->p2p;
  if (!pp->enabled || !d) return 0;
  for (int i = 0; i < pp->n_reg; ++i)
    if (pp->reg[i].base == d) { pp->reg[i].in_use = 0; return 1; }
  return 0;
}

/* Every exchange runs both flag rounds.  Skipping the "ready" round when the previous exchange was on another vector was
   tried and is WRONG here (dist_check: V-cycle off by 5e-6): BLAS-1 kernels between two exchanges write all stored planes of
   their result, ghost planes included, and so overwrite what a neighbour that is one exchange ahead has already pushed.  The
   bookkeeping stays (last_base) for the day those kernels are restricted to owned planes. */
void pmg_p2p_forget(pmg_context *ctx) { ctx->p2p.last_base = NULL; }

/* ghost planes of array d (a registered vector's storage): 1 = done by the push kernel, 0 = not registered (use NCCL) */
int pmg_p2p_halo(pmg_context *ctx, const pmg_layout *lay, double *d, cudaStream_t stream, int *done)
{
  pmg_p2p *pp = &ctx->p2p;
  *done = 0;
  if (!pp->enabled) return PMG_OK;
  if (!pp->explicit_push) { /* by size: the same decision on every rank (equal slabs) */
    const int64_t bytes = 8 * lay->plane * (lay->degree + 1);
    if (pp->push_min_bytes <= 0 || bytes < pp->push_min_bytes) return PMG_OK;
  }
  for (int i = 0; i < pp->n_reg; ++i)
    if (pp->reg[i].base == d) {
      const pmg_p2p_reg *r = &pp->reg[i];
      const int p = lay->degree;
      const int64_t plane = lay->plane;
      /* my first owned plane z_own_lo is the lower neighbour's upper ghost plane; my top p owned planes
         [z_own_hi - p, z_own_hi) are the upper neighbour's lower ghost planes: same global plane indices, offset by the
         neighbour's first stored plane */
      const int64_t n_to_lower = (lay->lower >= 0) ? plane : 0, n_to_upper = (lay->upper >= 0) ? plane * p : 0;
      const int64_t dst_in_lower = (lay->lower >= 0) ? plane * (lay->z_own_lo - r->lower_z0) : 0;
      const int64_t dst_in_upper = (lay->upper >= 0) ? plane * (lay->z_own_hi - p - r->upper_z0) : 0;
      PMG_CHECK(pmgk_halo_push(d, lay->lower >= 0 ? (double *)r->peer_lower : NULL, lay->upper >= 0 ? (double *)r->peer_upper : NULL,
                               n_to_lower, plane * (lay->z_own_lo - lay->z0), dst_in_lower, n_to_upper,
                               plane * (lay->z_own_hi - p - lay->z0), dst_in_upper, pp->mailbox, lay->lower >= 0 ? pp->mb_lower : NULL,
                               lay->upper >= 0 ? pp->mb_upper : NULL, 1 /* see pmg_p2p_forget() */, stream));
      pp->last_base = d;
      *done = 1;
      return PMG_OK;
    }
  return PMG_OK;
}

int pmg_p2p_push_desc(pmg_context *ctx, const pmg_layout *lay, double *out, pmgk_push *desc)
{
  pmg_p2p *pp = &ctx->p2p;
  memset(desc, 0, sizeof(*desc));
  if (!pp->enabled || !pp->fused || lay->gathered || !lay->active) return 0;
  for (int i = 0; i < pp->n_reg; ++i)
    if (pp->reg[i].base == out) {
      const pmg_p2p_reg *r = &pp->reg[i];
      desc->out_lower = lay->lower >= 0 ? (double *)r->peer_lower : NULL;
      desc->out_upper = lay->upper >= 0 ? (double *)r->peer_upper : NULL;
      desc->lower_z0 = r->lower_z0; desc->upper_z0 = r->upper_z0;
      desc->mailbox = pp->mailbox;
      desc->mailbox_lower = lay->lower >= 0 ? pp->mb_lower : NULL;
      desc->mailbox_upper = lay->upper >= 0 ? pp->mb_upper : NULL;
      return 1;
    }
  return 0;
}
