// pmg_apply_sweep_launch.h -- kernel entry + launcher of the line-marching apply kernel (csrc/pmg_apply_sweep.h) for ONE
// epilogue mode.  Included by pmg_apply_sweep_m{0,1,2,3}.cu, each of which defines PMG_SWEEP_TU_MODE (the PmgApplyMode the
// translation unit is compiled for) and so provides pmg_sweep_dispatch_m<mode>(); four translation units compile in parallel.
#include "pmg_apply_sweep.h"
#include "pmg_cuda_common.h"
#include "pmg_kernels.h"

namespace {

template <class Tile>
struct PmgSweepDeviceExec {
  typename Tile::ThreadState st;
  template <class F> __device__ __forceinline__ void for_each_thread(F f) { f((int)threadIdx.x, st); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void sync_some(int n) { if ((int)threadIdx.x < n) asm volatile("bar.sync 1, %0;\n" ::"r"(n) : "memory"); }
};

// chunk_first, chunk_stride: the launch's CTAs work on z-chunks chunk_first + i * chunk_stride (all chunks: 0, 1; the two
// chunks that touch the slab's ghost planes: 0, n_chunks - 1; the others: 1, 1)
template <int P, int BX, int BY, int LZ, int NT, int MINB, int US, int FM, int RL>
__global__ void __launch_bounds__(NT, MINB)
pmg_sweep_kernel(const __grid_constant__ PmgSweepParams<P> p, int chunk_first, int chunk_stride)
{
  using Tile = PmgSweepTile<P, BX, BY, LZ, NT, US, FM, 1, RL>;
  extern __shared__ __align__(128) double pmg_sweep_smem[];
  PmgSweepDeviceExec<Tile> ex;
  const int b = blockIdx.x;
  const int tile_x = b % p.tiles_x;
  const int tile_y = (b / p.tiles_x) % p.tiles_y;
  const int chunk = chunk_first + (b / (p.tiles_x * p.tiles_y)) * chunk_stride;
  Tile::run(p, ex, pmg_sweep_smem, tile_x, tile_y, chunk);
}

// number of z-chunks: minimise waves * (layers + recomputed layer and plane below the chunk)
void choose_sweep_chunks(int tiles, int layers, int slots, int degree, int min_chunks, int *n_chunks, int *layers_per_chunk)
{
  double best_cost = -1;
  int best_c = 1;
  if (min_chunks > layers) min_chunks = layers;
  for (int c = min_chunks; c <= layers; ++c) {
    const int lpc = (layers + c - 1) / c;
    const int used = (layers + lpc - 1) / lpc;
    if (used != c) continue;
    const long waves = ((long)tiles * c + slots - 1) / slots;
    const double cost = waves * (lpc + (c > 1 ? 1.0 + 1.0 / degree : 0.0));
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_c = c; }
  }
  *n_chunks = best_c;
  *layers_per_chunk = (layers + best_c - 1) / best_c;
}

// Per-mode overrides of the tile table (csrc/pmg_apply_sweep_tiles.inc).  The plain apply of Q4 keeps the cell loops of the y
// and x sweeps rolled (RL = 1): 106 registers instead of 163, so 4 CTAs of its 52.6 KB fit an SM instead of 3 -- measured on
// B200 at 100 M DoFs: 0.825 ms against 0.879 ms (122 against 114 GDoF/s, profiles/r01_v6_ablation_experiment.txt).  The fused
// modes stay at 3 CTAs per SM by shared memory and gain nothing from rolled loops; Q2 / Q3 / Q5 measured 3-5 % slower rolled
// (profiles/r01_v6_rolled_loops_experiment.txt).
template <int P, int FM> struct PmgSweepModeTune { static constexpr int roll = 0, min_ctas = 0; };
template <> struct PmgSweepModeTune<4, PMG_MODE_APPLY> { static constexpr int roll = 1, min_ctas = 4; };

template <int P, int BX, int BY, int LZ, int NT, int MINB, int US, int FM, int RL>
int launch_sweep(const pmgk_level *lv, const double *u, const double *b, const double *xold, double *out, double f1,
                 double f2, cudaStream_t stream, int *geom, int part)
{
  using Tile = PmgSweepTile<P, BX, BY, LZ, NT, US, FM, 1, RL>;
  void (*kernel)(const PmgSweepParams<P>, int, int) = pmg_sweep_kernel<P, BX, BY, LZ, NT, MINB, US, FM, RL>;
  PmgSweepParams<P> p;
  p.nx = lv->nx; p.ny = lv->ny; p.nz = lv->nz;
  p.Nx = lv->Nx; p.Ny = lv->Ny; p.Nz = lv->Nz;
  p.faces = lv->faces;
  p.z0 = lv->z0; p.nzl = lv->nzl;
  p.cz_lo = lv->cz_lo; p.cz_hi = lv->cz_hi;
  p.z_own_lo = lv->z_own_lo; p.z_own_hi = lv->z_own_hi;
  p.tiles_x = (lv->nx + BX - 1) / BX;
  p.tiles_y = (lv->ny + BY - 1) / BY;
  // APPLY stages only u; the other modes also stage the epilogue's b / x_old rows
  constexpr bool epi = (FM != PMG_MODE_APPLY);
  const int smem_bytes = Tile::smem_doubles(epi) * (int)sizeof(double);
  // per device: the shared-memory opt-in and the occupancy belong to the device the launch goes to
  enum { MAXDEV = 64 };
  static int ctas_per_sm_dev[MAXDEV];
  int dev = 0;
  PMG_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAXDEV) return PMG_ERR_UNSUPPORTED;
  int ctas_per_sm = __atomic_load_n(&ctas_per_sm_dev[dev], __ATOMIC_ACQUIRE);
  if (ctas_per_sm == 0) {
    PMG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    PMG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, Tile::NT, smem_bytes));
    if (ctas_per_sm < 1) return PMG_ERR_CUDA;
    __atomic_store_n(&ctas_per_sm_dev[dev], ctas_per_sm, __ATOMIC_RELEASE);
  }
  const int slots = pmgk_device_sm_count() * ctas_per_sm;
  // a launch in parts (PMG_HALO_OVERLAP=1, host/pmg_operator.c) needs at least three chunks: the first and the last read the
  // slab's ghost planes, the others run while those are in flight.  A whole launch takes the cheapest chunking (thin slabs:
  // 16 layers of 32 x 32 tiles at 8 GPUs cost 7 waves x 7.25 layers in three chunks, 5 x 9.25 in two)
  choose_sweep_chunks(p.tiles_x * p.tiles_y, lv->cz_hi - lv->cz_lo, slots, P, part != PMGK_PART_ALL ? 3 : 1, &p.n_chunks, &p.layers_per_chunk);
  pmg_sweep_fill_matrices<P>(p, lv->Mref, lv->Kref, lv->h);
  p.mode = FM; p.u = u; p.b = b; p.xold = xold; p.out = out; p.f1 = f1; p.f2 = f2;
  p.dinv_vec = lv->dinv_vec; p.dinv_tab = lv->dinv_tab;
  p.push_lo = p.push_hi = nullptr; p.mb = p.mb_lo = p.mb_hi = nullptr; p.consume = 0; // fused ghost push: plane kernel only
  int chunk_first = 0, chunk_stride = 1, chunk_count = p.n_chunks;
  if (part != PMGK_PART_ALL) {
    if (p.n_chunks < 3) return PMG_ERR_UNSUPPORTED;
    if (part == PMGK_PART_INTERIOR) { chunk_first = 1; chunk_count = p.n_chunks - 2; }
    else { chunk_stride = p.n_chunks - 1; chunk_count = 2; }
  }
  const int grid = p.tiles_x * p.tiles_y * chunk_count;
  if (geom) { geom[0] = grid; geom[1] = Tile::NT; geom[2] = smem_bytes; geom[3] = p.n_chunks; return 0; }
  if (((uintptr_t)u | (uintptr_t)b | (uintptr_t)xold) & 15) return PMG_ERR_ARG; /* bulk copies: 16-byte aligned vectors */
  kernel<<<grid, Tile::NT, smem_bytes, stream>>>(p, chunk_first, chunk_stride);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

} // namespace

#define PMG_SWEEP_CAT2(a, b) a##b
#define PMG_SWEEP_CAT(a, b) PMG_SWEEP_CAT2(a, b)

int PMG_SWEEP_CAT(pmg_sweep_dispatch_m, PMG_SWEEP_TU_MODE)(const pmgk_level *lv, const double *u, const double *b, const double *xold,
                                                           double *out, double f1, double f2, cudaStream_t s, int *geom, int part)
{
  switch (lv->degree) {
#define PMG_SWEEP_CASE(P, BX, BY, LZ, NT, MINB, US) \
  case P: return launch_sweep<P, BX, BY, LZ, NT, (PmgSweepModeTune<P, PMG_SWEEP_TU_MODE>::min_ctas ? PmgSweepModeTune<P, PMG_SWEEP_TU_MODE>::min_ctas : MINB), \
                              US, PMG_SWEEP_TU_MODE, PmgSweepModeTune<P, PMG_SWEEP_TU_MODE>::roll>(lv, u, b, xold, out, f1, f2, s, geom, part);
#include "pmg_apply_sweep_tiles.inc"
#undef PMG_SWEEP_CASE
    default: return PMG_ERR_UNSUPPORTED;
  }
}
