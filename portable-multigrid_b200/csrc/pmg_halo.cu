// pmg_halo.cu -- ghost-plane exchange of a z-slab over NVLink peer memory: ONE kernel per exchange, no NCCL on the hot path.
//
// Replaces src.update_ghost_values() of the reference's vmult (include/operators/portable_laplace_operator.h:635-661) for the
// slab decomposition of host/pmg_core.c.  Every rank has its neighbours' vectors mapped into its address space
// (cudaIpcOpenMemHandle, host/pmg_p2p.c) and PUSHES its boundary planes into their ghost planes (peer stores are posted: a
// first version that pulled with peer loads paid the NVLink round trip per load and lost to NCCL, 0.235 against 0.225 ms per
// fused step on 2 B200s).  A mailbox of four 64-bit flags per rank, written by the neighbours through the same mapping, orders
// the exchange:
//   1. "ready" (only when this vector was also the one of the previous exchange): my earlier kernels on this stream are
//      complete (stream order), so nothing of mine reads my ghost planes any more -- tell both neighbours; wait until both have
//      told me the same.  Otherwise step 3 of the previous exchange already orders things: a neighbour that is one exchange
//      ahead has seen my "pushed" of the previous exchange, sent after every kernel of mine that read THIS vector's ghost
//      planes, and it cannot be two ahead (it waits for my "pushed" of this one);
//   2. store my boundary planes into the neighbours' ghost planes (coalesced peer stores over NVLink);
//   3. "pushed": after a system-wide fence, tell both neighbours that their ghost planes are filled; wait for theirs.  When the
//      kernel ends, this rank's ghost planes are current.
// The epoch lives in device memory and advances by one per exchange on every rank (the ranks run the same sequence of
// exchanges), so the kernel is captured into CUDA graphs like any other.
#include <stdint.h>
#include "pmg_cuda_common.h"
#include "pmg_kernels.h"

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// mailbox words: 0 ready_from_lower, 1 ready_from_upper, 2 pushed_from_lower, 3 pushed_from_upper, 4 epoch (local), 5 ticket (local)
__global__ void __launch_bounds__(256)
k_halo_push(const double *mine, double *peer_lower, double *peer_upper, int64_t n_to_lower, int64_t src_to_lower, int64_t dst_in_lower,
            int64_t n_to_upper, int64_t src_to_upper, int64_t dst_in_upper, unsigned long long *mb, unsigned long long *mb_lower,
            unsigned long long *mb_upper, int need_ready)
{
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) {
    const unsigned long long epoch = ld_acquire_sys(mb + 4) + 1; // every CTA reads it before the last one to finish bumps it
    if (need_ready) {
      if (blockIdx.x == 0) {
        __threadfence_system();
        if (mb_lower) st_release_sys(mb_lower + 1, epoch);
        if (mb_upper) st_release_sys(mb_upper + 0, epoch);
      }
      if (mb_lower) while (ld_acquire_sys(mb + 0) < epoch) { }
      if (mb_upper) while (ld_acquire_sys(mb + 1) < epoch) { }
    }
    s_epoch = epoch;
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (peer_lower) {
    const double *src = mine + src_to_lower;
    double *dst = peer_lower + dst_in_lower;
#pragma unroll 4
    for (int64_t i = t0; i < n_to_lower; i += stride) dst[i] = src[i];
  }
  if (peer_upper) {
    const double *src = mine + src_to_upper;
    double *dst = peer_upper + dst_in_upper;
#pragma unroll 4
    for (int64_t i = t0; i < n_to_upper; i += stride) dst[i] = src[i];
  }
  __threadfence_system(); // this thread's peer stores are visible system-wide before its CTA takes a ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long epoch = s_epoch;
    const unsigned long long ticket = atomicAdd(mb + 5, 1ull);
    if (ticket == gridDim.x - 1) { // the last CTA: every store of this exchange has been fenced
      mb[5] = 0;
      __threadfence_system();
      if (mb_lower) st_release_sys(mb_lower + 3, epoch);
      if (mb_upper) st_release_sys(mb_upper + 2, epoch);
      if (mb_lower) while (ld_acquire_sys(mb + 2) < epoch) { }
      if (mb_upper) while (ld_acquire_sys(mb + 3) < epoch) { }
      st_release_sys(mb + 4, epoch);
    }
  }
}

} // namespace

extern "C" int pmgk_halo_push(const double *mine, double *peer_lower, double *peer_upper, int64_t n_to_lower, int64_t src_to_lower,
                              int64_t dst_in_lower, int64_t n_to_upper, int64_t src_to_upper, int64_t dst_in_upper, void *mailbox,
                              void *mailbox_lower, void *mailbox_upper, int need_ready, void *stream)
{
  if (!mine || !mailbox) return PMG_ERR_ARG;
  const int64_t n = (peer_lower ? n_to_lower : 0) + (peer_upper ? n_to_upper : 0);
  int grid = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (grid < 1) grid = 1;
  if (grid > 148) grid = 148;
  k_halo_push<<<grid, 256, 0, (cudaStream_t)stream>>>(mine, peer_lower, peer_upper, n_to_lower, src_to_lower, dst_in_lower, n_to_upper,
                                                     src_to_upper, dst_in_upper, (unsigned long long *)mailbox,
                                                     (unsigned long long *)mailbox_lower, (unsigned long long *)mailbox_upper, need_ready);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}
