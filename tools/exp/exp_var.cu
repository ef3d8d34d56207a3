// tools/exp/exp_var.cu -- stand-alone timing harness for the variable-coefficient apply kernel (development tool).
// Builds ONE tile configuration (-DC_P -DC_BX -DC_BY -DMINB) and times APPLY and CHEB_STEP launches on an n^3-cell cube:
//   exp_var <cells> [reps]      (coefficient = 1 everywhere: timing only)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "pmg_apply_var.h"
extern "C" void pmg_fe_shape_tables(int p, double *Sq, double *Dco, double *G, double *gq, double *gw);
#define STR2(x) #x
#define STR(x) STR2(x)
#ifndef MINB
#define MINB 1
#endif
using Tile = PmgVarTile<C_P, C_BX, C_BY>;
struct Ex {
  Tile::ThreadState st;
  template <class F> __device__ __forceinline__ void for_each_thread(F f) { f((int)threadIdx.x, st); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
};
__global__ void __launch_bounds__(Tile::NT, MINB) kern(const __grid_constant__ PmgVarParams<C_P> p)
{
  extern __shared__ double sm[];
  Ex ex;
  const int b = blockIdx.x;
  Tile::run(p, ex, sm, b % p.tiles_x, (b / p.tiles_x) % p.tiles_y, b / (p.tiles_x * p.tiles_y));
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
int main(int argc, char **argv)
{
  constexpr int P = C_P, N1 = P + 1;
  int n = argc > 1 ? atoi(argv[1]) : 0; if (n <= 0) n = (310 + P / 2) / P; const int reps = argc > 2 ? atoi(argv[2]) : 10;
  PmgVarParams<P> p{};
  p.nx = p.ny = p.nz = n; p.Nx = p.Ny = p.Nz = n * P + 1; p.faces = 0x3F;
  p.z0 = 0; p.nzl = p.Nz; p.cz_lo = 0; p.cz_hi = n; p.z_own_lo = 0; p.z_own_hi = p.Nz;
  p.tiles_x = (n + C_BX - 1) / C_BX; p.tiles_y = (n + C_BY - 1) / C_BY;
  pmg_fe_shape_tables(P, p.S, p.D, nullptr, nullptr, nullptr);
  p.c[0] = p.c[1] = p.c[2] = 1.0 / n;
  const size_t N = (size_t)p.Nx * p.Ny * p.Nz, NQ = (size_t)n * N1 * n * N1 * n * N1;
  std::vector<double> hu(N), hc(NQ, 1.0);
  for (size_t i = 0; i < N; ++i) hu[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
  double *u, *b, *xo, *out, *dv, *coef;
  CK(cudaMalloc(&u, N * 8)); CK(cudaMalloc(&b, N * 8)); CK(cudaMalloc(&xo, N * 8)); CK(cudaMalloc(&out, N * 8)); CK(cudaMalloc(&dv, N * 8));
  CK(cudaMalloc(&coef, NQ * 8));
  CK(cudaMemcpy(u, hu.data(), N * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(b, hu.data(), N * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, hu.data(), N * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(coef, hc.data(), NQ * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(xo, 0, N * 8));
  p.u = u; p.b = b; p.xold = xo; p.out = out; p.f1 = 0.3; p.f2 = 0.1; p.dinv_tab = nullptr; p.dinv_vec = dv; p.coef = coef; p.coef_cz0 = 0;
  const int smem = Tile::SMEM_DOUBLES * 8;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Tile::NT, smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
  int chunks = 1;
  { const int slots = 148 * per_sm, tiles = p.tiles_x * p.tiles_y; long best = -1;
    for (int c = 1; c <= n; ++c) { int lpc = (n + c - 1) / c; if ((n + lpc - 1) / lpc != c) continue;
      long waves = ((long)tiles * c + slots - 1) / slots; long cost = waves * (lpc + (c > 1 ? 1 : 0));
      if (best < 0 || cost < best) { best = cost; chunks = c; } } }
  p.layers_per_chunk = (n + chunks - 1) / chunks; p.n_chunks = (n + p.layers_per_chunk - 1) / p.layers_per_chunk;
  const int grid = p.tiles_x * p.tiles_y * p.n_chunks;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode : {0, 3}) {
    p.mode = mode; p.out = (mode == 3) ? xo : out;
    for (int i = 0; i < 2; ++i) kern<<<grid, Tile::NT, smem>>>(p);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) kern<<<grid, Tile::NT, smem>>>(p);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("%s P=%d n=%d N=%zu tile=%dx%d nt=%d minb=%d regs=%d spill=%zuB smem=%dKB ctas/sm=%d grid=%d chunks=%d mode=%d: %.3f ms %.1f GDoF/s\n",
           STR(C_NAME), P, n, N, C_BX, C_BY, Tile::NT, MINB, fa.numRegs, (size_t)fa.localSizeBytes, smem / 1024, per_sm, grid, p.n_chunks, mode, ms, N / ms / 1e6);
  }
  return 0;
}
