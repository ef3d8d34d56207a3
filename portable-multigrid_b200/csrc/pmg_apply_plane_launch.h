// pmg_apply_plane_launch.h -- kernel entry + launcher of the plane-per-step apply kernel (csrc/pmg_apply_plane.h) for ONE
// epilogue mode.  Included by pmg_apply_plane_m{0..4}.cu, each of which defines PMG_PLANE_TU_MODE (0..3 = PmgApplyMode,
// 4 = PMG_MODE_CHEB_STEP without x_old) and so provides pmg_plane_dispatch_m<mode>(); the translation units compile in parallel.
#include "pmg_apply_plane.h"
#include "pmg_cuda_common.h"
#include "pmg_kernels.h"

namespace {

__device__ __forceinline__ void pmg_st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pmg_ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <class Tile>
struct PmgPlaneDeviceExec {
  typename Tile::ThreadState st;
  template <class F> __device__ __forceinline__ void for_each_thread(F f) { f((int)threadIdx.x, st); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  // block until mb[flag] >= mb[epoch] (system scope: the flag word is written by a neighbour GPU over NVLink).  The wait is
  // bounded: a neighbour that never arrives (a rank that died) must not hang this GPU -- after ~4 s the CTA gives up, raises
  // the mailbox's error word (read by pmgk_fused_error) and goes on with whatever its ghost planes hold.
  __device__ __forceinline__ void wait_flag(unsigned long long *mb, int flag, int epoch)
  {
    if (threadIdx.x == 0) {
      const unsigned long long e = pmg_ld_acquire_sys(mb + epoch);
      if (pmg_ld_acquire_sys(mb + flag) < e) {
        const long long t0 = clock64();
        // (once a wait has timed out nobody waits any more: the run is lost, it must still come to an end)
        while (pmg_ld_acquire_sys(mb + flag) < e && pmg_ld_acquire_sys(mb + PMG_FUSED_ERROR) == 0) {
          __nanosleep(100);
          if (clock64() - t0 > 8000000000ll) { atomicExch(mb + PMG_FUSED_ERROR, 1ull); break; } // ~4 s at 2 GHz
        }
      }
    }
    __syncthreads();
  }
};

template <int P, int BX, int BY, int NT, int MINB, int UZ, int FM, int PR, int XS, int PUSH>
__global__ void __launch_bounds__(NT, MINB)
pmg_plane_kernel(const __grid_constant__ PmgSweepParams<P> p, int chunk_first, int chunk_stride)
{
  using Tile = PmgPlaneTile<P, BX, BY, NT, FM, UZ, 0, 3, 0, PR, XS, PUSH>;
  extern __shared__ __align__(128) double pmg_plane_smem[];
  PmgPlaneDeviceExec<Tile> ex;
  const int b = blockIdx.x;
  const int tile_x = b % p.tiles_x;
  const int tile_y = (b / p.tiles_x) % p.tiles_y;
  // the launch's CTAs work on z-chunks chunk_first + i * chunk_stride (all chunks: 0, 1; the two chunks that read the slab's
  // ghost planes: 0, n_chunks - 1; the others: 1, 1)
  const int group = b / (p.tiles_x * p.tiles_y);
  int chunk = chunk_first + group * chunk_stride;
  // ---- fused ghost exchange (p.mb: whole launches of slabs with neighbours) ---------------------------------------------
  // The bottom chunk's CTAs (at_lo) need the lower neighbour's word before their march (they start in its ghost planes), the top
  // chunk's CTAs (at_hi) only before they fetch the upper ghost plane, the last plane of their march (ex.wait_flag in
  // Tile::layer); after the march both copy their part of the slab's boundary planes to the neighbours (Tile::push_boundary),
  // and the last CTA of each group tells the neighbour that faces it.  Interior chunks never look at a flag.  Chunk launch
  // order: natural (bit 1 of p.consume, the default) or top chunk first / bottom chunk last, which gives every flag at least a
  // chunk march of slack but measured slower (host/pmg_operator.c).  Flag words: PMG_FUSED_* (csrc/pmg_apply_plane.h); every
  // rank issues the same sequence of fused launches.
  bool at_lo = false, at_hi = false;
  if (PUSH && p.mb) {
    const int last = p.n_chunks - 1;
    if (!(p.consume & 2)) chunk = group == 0 ? last : group == last ? 0 : group;
    at_lo = chunk == 0; at_hi = chunk == last;
    // u's lower ghost planes were pushed by the lower neighbour's previous fused launch: wait until its top chunk is complete
    // (which also says that it has stopped reading the ghost plane this chunk's first epilogue pushes into)
    if (at_lo && (p.consume & 1) && p.mb_lo) ex.wait_flag(p.mb, PMG_FUSED_FROM_LO, PMG_FUSED_EPOCH_LO);
  }
  Tile::run(p, ex, pmg_plane_smem, tile_x, tile_y, chunk);
  if (!PUSH || !(at_lo || at_hi)) return;
  Tile::push_boundary(p, ex, tile_x, tile_y, at_lo, at_hi);
  // the CTA's stores (peer stores included) are ordered before the ticket by the barrier + the fence of the thread that takes it
  // (fences are cumulative: the pattern of a grid-wide barrier)
  __syncthreads();
  if (threadIdx.x == 0) {
    if (p.consume & 8) __threadfence(); else if (!(p.consume & 4)) __threadfence_system(); // (bits 2, 3: experiments only)
    const unsigned long long tiles = (unsigned long long)p.tiles_x * p.tiles_y;
    if (at_lo && atomicAdd(p.mb + PMG_FUSED_TICKET_LO, 1ull) == tiles - 1) { // the last CTA of the bottom chunk
      p.mb[PMG_FUSED_TICKET_LO] = 0;
      const unsigned long long e = pmg_ld_acquire_sys(p.mb + PMG_FUSED_EPOCH_LO) + 1;
      __threadfence_system();
      if (p.mb_lo) pmg_st_release_sys(p.mb_lo + PMG_FUSED_FROM_HI, e); // I am my lower neighbour's upper neighbour
      pmg_st_release_sys(p.mb + PMG_FUSED_EPOCH_LO, e);
    }
    if (at_hi && atomicAdd(p.mb + PMG_FUSED_TICKET_HI, 1ull) == tiles - 1) { // the last CTA of the top chunk
      p.mb[PMG_FUSED_TICKET_HI] = 0;
      const unsigned long long e = pmg_ld_acquire_sys(p.mb + PMG_FUSED_EPOCH_HI) + 1;
      __threadfence_system();
      if (p.mb_hi) pmg_st_release_sys(p.mb_hi + PMG_FUSED_FROM_LO, e);
      pmg_st_release_sys(p.mb + PMG_FUSED_EPOCH_HI, e);
    }
  }
}

// z-chunks of a launch.  Measured on B200 (profiles/r02_plane_chunk_sweep.txt, Q4): one wave of CTAs that each march the
// whole column is the SLOWEST choice at 100 M DoFs (96 GDoF/s against 130 with 8 chunks per tile; a random start delay per CTA
// changes nothing, so it is not the lockstep of the phases -- the u-read and the out-write streams of all CTAs then sit a fixed
// few planes apart in memory); about eight CTAs per slot on staggered z ranges is where the curve is flat.  A chunk costs
// P + 2 steps on top of its own (the recomputed layer below it, fill and drain), so chunks keep >= ~32 steps -- unless that
// leaves SMs without a CTA (mid-size levels), where filling the machine comes first.
inline void choose_plane_chunks(int tiles, int layers, int slots, int degree, int *n_chunks, int *layers_per_chunk)
{
  const long c_target = (8L * slots + tiles - 1) / tiles;
  const long c_fill = (slots + tiles - 1) / tiles;
  const int lpc_min = (32 + degree - 1) / degree;
  long c_cap = layers / lpc_min;
  long c_fill_cap = layers / 2;
  if (c_cap < 1) c_cap = 1;
  if (c_fill_cap < 1) c_fill_cap = 1;
  long c_floor = c_fill < c_fill_cap ? c_fill : c_fill_cap; // chunks needed to give every slot a CTA, at >= 2 layers each
  if (c_floor < c_cap) c_floor = c_cap;
  long c = c_target < c_floor ? c_target : c_floor;
  if (c < 1) c = 1;
  const int lpc = (layers + (int)c - 1) / (int)c;
  *layers_per_chunk = lpc;
  *n_chunks = (layers + lpc - 1) / lpc;
}

template <int P, int BX, int BY, int NT, int MINB, int UZ, int FM, int PR, int XS, int PUSH>
int launch_plane(const pmgk_level *lv, const double *u, const double *b, const double *xold, double *out, double f1, double f2,
                 cudaStream_t stream, int *geom, int part)
{
  using Tile = PmgPlaneTile<P, BX, BY, NT, FM, UZ, 0, 3, 0, PR, XS, PUSH>;
  auto kernel = pmg_plane_kernel<P, BX, BY, NT, MINB, UZ, FM, PR, XS, PUSH>;
  PmgSweepParams<P> p;
  p.nx = lv->nx; p.ny = lv->ny; p.nz = lv->nz;
  p.Nx = lv->Nx; p.Ny = lv->Ny; p.Nz = lv->Nz;
  p.faces = lv->faces;
  p.z0 = lv->z0; p.nzl = lv->nzl;
  p.cz_lo = lv->cz_lo; p.cz_hi = lv->cz_hi;
  p.z_own_lo = lv->z_own_lo; p.z_own_hi = lv->z_own_hi;
  p.tiles_x = Tile::tiles_of(lv->nx, lv->faces >> 1 & 1u, BX);
  p.tiles_y = Tile::tiles_of(lv->ny, lv->faces >> 3 & 1u, BY);
  const int smem_bytes = Tile::SMEM_DOUBLES * (int)sizeof(double);
  // per device: the shared-memory opt-in and the occupancy belong to the device the launch goes to
  enum { MAXDEV = 64 };
  static int ctas_per_sm[MAXDEV];
  int dev = 0;
  PMG_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAXDEV) return PMG_ERR_UNSUPPORTED;
  int per_sm = __atomic_load_n(&ctas_per_sm[dev], __ATOMIC_ACQUIRE);
  if (per_sm == 0) {
    PMG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    PMG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NT, smem_bytes));
    if (per_sm < 1) return PMG_ERR_CUDA;
    __atomic_store_n(&ctas_per_sm[dev], per_sm, __ATOMIC_RELEASE);
  }
  const int slots = pmgk_device_sm_count() * per_sm;
  choose_plane_chunks(p.tiles_x * p.tiles_y, lv->cz_hi - lv->cz_lo, slots, P, &p.n_chunks, &p.layers_per_chunk);
  pmg_sweep_fill_matrices<P>(p, lv->Mref, lv->Kref, lv->h);
  p.mode = (FM == 4) ? PMG_MODE_CHEB_STEP : FM; p.u = u; p.b = b; p.xold = xold; p.out = out; p.f1 = f1; p.f2 = f2;
  p.dinv_vec = lv->dinv_vec; p.dinv_tab = lv->dinv_tab;
  // fused ghost push (pmgk_apply_push): whole launches of slabs with at least two cell layers
  p.push_lo = p.push_hi = nullptr; p.mb = p.mb_lo = p.mb_hi = nullptr; p.consume = 0;
  if (const pmgk_push *ps = PUSH ? pmg_tl_push : nullptr) {
    if (part != PMGK_PART_ALL || lv->cz_hi - lv->cz_lo < 2 || !ps->mailbox) return PMG_ERR_UNSUPPORTED;
    const int64_t plane = (int64_t)lv->Nx * lv->Ny;
    if (ps->push && ps->out_lower) p.push_lo = ps->out_lower + plane * (lv->z0 - ps->lower_z0);
    if (ps->push && ps->out_upper) p.push_hi = ps->out_upper + plane * (lv->z0 - ps->upper_z0);
    p.mb = (unsigned long long *)ps->mailbox;
    p.mb_lo = (unsigned long long *)ps->mailbox_lower; p.mb_hi = (unsigned long long *)ps->mailbox_upper;
    p.consume = ps->consume;
  }
  // a launch in parts (halo exchange overlapped with the chunks that read no ghost plane, host/pmg_operator.c): the first and
  // the last chunk read the slab's ghost planes, the others run while those are in flight
  int chunk_first = 0, chunk_stride = 1, chunk_count = p.n_chunks;
  if (part != PMGK_PART_ALL) {
    if (p.n_chunks < 3) return PMG_ERR_UNSUPPORTED;
    if (part == PMGK_PART_INTERIOR) { chunk_first = 1; chunk_count = p.n_chunks - 2; }
    else { chunk_stride = p.n_chunks - 1; chunk_count = 2; }
  }
  const int grid = p.tiles_x * p.tiles_y * chunk_count;
  if (geom) { geom[0] = grid; geom[1] = NT; geom[2] = smem_bytes; geom[3] = p.n_chunks; return 0; }
  /* the kernel indexes inside a dof plane with 32-bit element offsets (planes themselves are 64-bit offsets apart) */
  if ((int64_t)lv->Nx * lv->Ny * 4 >= (int64_t)1 << 31) return PMG_ERR_UNSUPPORTED;
  kernel<<<grid, NT, smem_bytes, stream>>>(p, chunk_first, chunk_stride);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

} // namespace

#define PMG_PLANE_CAT2(a, b) a##b
#define PMG_PLANE_CAT(a, b) PMG_PLANE_CAT2(a, b)

int PMG_PLANE_CAT(pmg_plane_dispatch_m, PMG_PLANE_TU_MODE)(const pmgk_level *lv, const double *u, const double *b, const double *xold,
                                                           double *out, double f1, double f2, cudaStream_t s, int *geom, int part)
{
  switch (lv->degree) {
// the residual and the Chebyshev steps (modes 1 .. 4) come in a second instance with the fused ghost exchange (pmgk_apply_push)
#if PMG_PLANE_TU_MODE >= 1
#define PMG_PLANE_LAUNCH(P, BX, BY, NT, MINB, UZ, PR, XS)                                                                              \
  case P:                                                                                                                             \
    return pmg_tl_push ? launch_plane<P, BX, BY, NT, MINB, UZ, PMG_PLANE_TU_MODE, PR, XS, 1>(lv, u, b, xold, out, f1, f2, s, geom, part) \
                       : launch_plane<P, BX, BY, NT, MINB, UZ, PMG_PLANE_TU_MODE, PR, XS, 0>(lv, u, b, xold, out, f1, f2, s, geom, part);
#else
#define PMG_PLANE_LAUNCH(P, BX, BY, NT, MINB, UZ, PR, XS) \
  case P: return pmg_tl_push ? PMG_ERR_UNSUPPORTED : launch_plane<P, BX, BY, NT, MINB, UZ, PMG_PLANE_TU_MODE, PR, XS, 0>(lv, u, b, xold, out, f1, f2, s, geom, part);
#endif
#if PMG_PLANE_TU_MODE == 0
#define PMG_PLANE_CASE(P, BX, BY, NT, MINB, UZ, PR, XS) PMG_PLANE_LAUNCH(P, BX, BY, NT, MINB, UZ, PR, XS)
#define PMG_PLANE_CASE_F(P, BX, BY, NT, MINB, UZ, PR, XS)
#else
#define PMG_PLANE_CASE(P, BX, BY, NT, MINB, UZ, PR, XS)
#define PMG_PLANE_CASE_F(P, BX, BY, NT, MINB, UZ, PR, XS) PMG_PLANE_LAUNCH(P, BX, BY, NT, MINB, UZ, PR, XS)
#endif
#include "pmg_apply_plane_tiles.inc"
#undef PMG_PLANE_CASE
#undef PMG_PLANE_CASE_F
#undef PMG_PLANE_LAUNCH
    default: return PMG_ERR_UNSUPPORTED;
  }
}
