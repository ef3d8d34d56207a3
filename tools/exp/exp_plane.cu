// tools/exp/exp_plane.cu -- stand-alone timing harness for the plane-per-step apply kernel (development tool).
// Builds ONE tile configuration (-DC_P -DC_BX -DC_BY -DC_NT -DMINB [-DC_UZ]) and times APPLY and CHEB_STEP launches on an
// n^3-cell cube; with -DWITH_REF the result is compared with the line-marching kernel's (shipped tile of that degree).
//   exp_plane <cells> [reps] [chunks]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "pmg_apply_plane.h"
extern "C" void pmg_fe_pencil(int p, double *M, double *K);
#ifndef C_UZ
#define C_UZ 1
#endif
#ifndef C_NU
#define C_NU 3
#endif
#ifndef C_EPF
#define C_EPF 0
#endif
#ifndef C_PR
#define C_PR 1
#endif
#ifndef C_XS
#define C_XS 1
#endif
#ifndef C_FM
#define C_FM -1
#endif
#ifndef MINB
#define MINB 1
#endif
#define STR2(x) #x
#define STR(x) STR2(x)
constexpr int P = C_P;
template <int FM> using TileT = PmgPlaneTile<C_P, C_BX, C_BY, C_NT, FM, C_UZ, 0, C_NU, C_EPF, C_PR, C_XS>;
template <class Tile> struct Ex {
  typename Tile::ThreadState st;
  template <class F> __device__ __forceinline__ void for_each_thread(F f) { f((int)threadIdx.x, st); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
};
template <int FM>
__global__ void __launch_bounds__(C_NT, MINB) kern(const __grid_constant__ PmgSweepParams<P> p)
{
  extern __shared__ __align__(128) double sm[];
  Ex<TileT<FM>> ex;
  const int b = blockIdx.x;
#ifdef C_STAG
  __nanosleep(((blockIdx.x * 2654435761u) >> 24) * C_STAG); // experiment: random start delay of up to 255 * C_STAG ns
#endif
  TileT<FM>::run(p, ex, sm, b % p.tiles_x, (b / p.tiles_x) % p.tiles_y, b / (p.tiles_x * p.tiles_y));
}
#ifdef WITH_REF
struct RefEx {
  template <class Tile> struct E {
    typename Tile::ThreadState st;
    template <class F> __device__ __forceinline__ void for_each_thread(F f) { f((int)threadIdx.x, st); }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void sync_some(int n) { if ((int)threadIdx.x < n) asm volatile("bar.sync 1, %0;\n" ::"r"(n) : "memory"); }
  };
};
template <int Q> struct RefSel;
#define PMG_SWEEP_CASE(PP, BX, BY, LZ, NT, MB, US) \
  template <> struct RefSel<PP> { using T0 = PmgSweepTile<PP, BX, BY, LZ, NT, US, 0>; using T3 = PmgSweepTile<PP, BX, BY, LZ, NT, US, 3>; };
#include "pmg_apply_sweep_tiles.inc"
#undef PMG_SWEEP_CASE
template <class Tile>
__global__ void __launch_bounds__(Tile::NT, 1) refkern(const __grid_constant__ PmgSweepParams<P> p)
{
  extern __shared__ __align__(128) double sm[];
  RefEx::E<Tile> ex;
  const int b = blockIdx.x;
  Tile::run(p, ex, sm, b % p.tiles_x, (b / p.tiles_x) % p.tiles_y, b / (p.tiles_x * p.tiles_y));
}
#endif
__global__ void fill(double *a, size_t n, unsigned mul, unsigned add, unsigned mod)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    a[i] = (double)(((unsigned)i * mul + add) % mod) / (double)mod - 0.5;
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

static int choose_chunks(int tiles, int n, int slots, int chunks_arg, int *lpc)
{
  int c_best = chunks_arg;
  if (c_best <= 0) {
    double best = -1;
    for (int c = 1; c <= n; ++c) {
      int l = (n + c - 1) / c; if ((n + l - 1) / l != c) continue;
      long waves = ((long)tiles * c + slots - 1) / slots; double cost = waves * (l + (c > 1 ? 1.0 + 1.0 / P : 0.0));
      if (best < 0 || cost < best) { best = cost; c_best = c; }
    }
  }
  *lpc = (n + c_best - 1) / c_best;
  return (n + *lpc - 1) / *lpc;
}

template <int FM>
static float run_mode(PmgSweepParams<P> p, int n, int reps, int chunks_arg, const char *name, size_t N)
{
  using Tile = TileT<FM>;
  const int smem = Tile::SMEM_DOUBLES * 8;
  CK(cudaFuncSetAttribute(kern<FM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern<FM>, C_NT, smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern<FM>));
  p.tiles_x = Tile::tiles_of(n, true, C_BX); p.tiles_y = Tile::tiles_of(n, true, C_BY);
  int dev_sms = 148; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  p.n_chunks = choose_chunks(p.tiles_x * p.tiles_y, n, dev_sms * per_sm, chunks_arg, &p.layers_per_chunk);
  const int grid = p.tiles_x * p.tiles_y * p.n_chunks;
  p.mode = FM;
  for (int i = 0; i < 2; ++i) kern<FM><<<grid, C_NT, smem>>>(p);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) kern<FM><<<grid, C_NT, smem>>>(p);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  printf("%s P=%d n=%d N=%zu tile=%dx%d nt=%d uz=%d nu=%d epf=%d minb=%d regs=%d smem=%dKB ctas/sm=%d grid=%d chunks=%d mode=%d: %.3f ms %.1f GDoF/s\n",
         name, P, n, N, C_BX, C_BY, C_NT, C_UZ, C_NU, C_EPF, MINB, fa.numRegs, smem / 1024, per_sm, grid, p.n_chunks, FM, ms, N / ms / 1e6);
  return ms;
}

int main(int argc, char **argv)
{
  int n = argc > 1 ? atoi(argv[1]) : 0; if (n <= 0) n = (464 + P / 2) / P;
  const int reps = argc > 2 ? atoi(argv[2]) : 10;
  const int chunks = argc > 3 ? atoi(argv[3]) : 0;
  PmgSweepParams<P> p{};
  p.nx = p.ny = p.nz = n; p.Nx = p.Ny = p.Nz = n * P + 1; p.faces = 0x3F;
  p.z0 = 0; p.nzl = p.Nz; p.cz_lo = 0; p.cz_hi = n; p.z_own_lo = 0; p.z_own_hi = p.Nz;
  double M[100], K[100], h[3] = {1.0 / n, 1.0 / n, 1.0 / n};
  pmg_fe_pencil(P, M, K);
  pmg_sweep_fill_matrices<P>(p, M, K, h);
  const size_t N = (size_t)p.Nx * p.Ny * p.Nz;
  double *u, *b, *xo, *out, *tab, *ref;
  const size_t bytes = N * 8 + 8 * (size_t)p.Nx + 16;
  CK(cudaMalloc(&u, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&xo, bytes)); CK(cudaMalloc(&out, bytes)); CK(cudaMalloc(&ref, bytes));
  std::vector<double> htab(1000);
  for (int i = 0; i < 1000; ++i) htab[i] = 0.5 + 0.001 * i;
  CK(cudaMalloc(&tab, 1000 * 8)); CK(cudaMemcpy(tab, htab.data(), 8000, cudaMemcpyHostToDevice));
  fill<<<1184, 256>>>(u, N, 2654435761u, 0u, 1000u); fill<<<1184, 256>>>(b, N, 40503u, 17u, 997u); fill<<<1184, 256>>>(xo, N, 69069u, 5u, 991u);
  CK(cudaDeviceSynchronize());
  p.u = u; p.b = b; p.xold = xo; p.out = out; p.f1 = 0.3; p.f2 = 0.1; p.dinv_tab = tab; p.dinv_vec = nullptr;
  const char *name = STR(C_NAME);
  run_mode<0>(p, n, reps, chunks, name, N);
#ifdef WITH_REF
  {
    std::vector<double> a(N), r(N);
    CK(cudaMemcpy(a.data(), out, N * 8, cudaMemcpyDeviceToHost));
    using T0 = RefSel<P>::T0;
    PmgSweepParams<P> q = p; q.out = ref; q.mode = 0;
    constexpr int BXc = (T0::CW - 1) / P, BYc = (T0::RW - 1) / P;
    q.tiles_x = (n + BXc - 1) / BXc; q.tiles_y = (n + BYc - 1) / BYc; q.n_chunks = 1; q.layers_per_chunk = n;
    const int smem = T0::smem_doubles(false) * 8;
    CK(cudaFuncSetAttribute(refkern<T0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    refkern<T0><<<q.tiles_x * q.tiles_y, T0::NT, smem>>>(q);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(r.data(), ref, N * 8, cudaMemcpyDeviceToHost));
    double num = 0, den = 0;
    for (size_t i = 0; i < N; ++i) { num += (a[i] - r[i]) * (a[i] - r[i]); den += r[i] * r[i]; }
    printf("%s APPLY rel l2 difference to the line-marching kernel: %.3e\n", name, sqrt(num / den));
  }
#endif
  // fused step: out = x_old buffer (in place), as the smoother calls it
  {
    PmgSweepParams<P> q = p; q.out = xo;
    run_mode<3>(q, n, reps, chunks, name, N);
  }
#ifdef WITH_REF
  {
    // one clean fused step into `out` from fresh x_old, compared with the line-marching kernel
    fill<<<1184, 256>>>(xo, N, 69069u, 5u, 991u);
    PmgSweepParams<P> q = p; q.out = out; q.mode = 3;
    using Tile = TileT<3>;
    q.tiles_x = Tile::tiles_of(n, true, C_BX); q.tiles_y = Tile::tiles_of(n, true, C_BY); q.n_chunks = 2; q.layers_per_chunk = (n + 1) / 2;
    kern<3><<<q.tiles_x * q.tiles_y * 2, C_NT, Tile::SMEM_DOUBLES * 8>>>(q);
    CK(cudaDeviceSynchronize());
    std::vector<double> a(N), r(N);
    CK(cudaMemcpy(a.data(), out, N * 8, cudaMemcpyDeviceToHost));
    using T3 = RefSel<P>::T3;
    PmgSweepParams<P> w = p; w.out = ref; w.mode = 3;
    constexpr int BXc = (T3::CW - 1) / P, BYc = (T3::RW - 1) / P;
    w.tiles_x = (n + BXc - 1) / BXc; w.tiles_y = (n + BYc - 1) / BYc; w.n_chunks = 1; w.layers_per_chunk = n;
    const int smem = T3::smem_doubles(true) * 8;
    CK(cudaFuncSetAttribute(refkern<T3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    refkern<T3><<<w.tiles_x * w.tiles_y, T3::NT, smem>>>(w);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(r.data(), ref, N * 8, cudaMemcpyDeviceToHost));
    double num = 0, den = 0;
    for (size_t i = 0; i < N; ++i) { num += (a[i] - r[i]) * (a[i] - r[i]); den += r[i] * r[i]; }
    printf("%s CHEB_STEP rel l2 difference to the line-marching kernel: %.3e\n", name, sqrt(num / den));
  }
#endif
  return 0;
}
