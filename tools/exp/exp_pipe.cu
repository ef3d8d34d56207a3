// tools/exp/exp_pipe.cu -- timing + self-check harness for the PIPELINED line-marching apply kernel (csrc/pmg_apply_sweep_pipe.h):
// times it like exp_sweep.cu and compares its output with the plain kernel's (same tile, NT = NG) on the same input.
// build: tools/exp/build_exp.sh NAME P BX BY LZ NG MINB -DEXP_SRC_PIPE ...   (C_NT is the group size NG; the CTA has 2 NG threads)
// Builds ONE tile configuration (-DCFG=P,BX,BY,LZ,NT -DPP=P -DMINB=n) with optional -DPMG_EXP_* switches and times
// APPLY and CHEB_STEP launches on an n^3-cell cube:  exp_sweep <cells> [reps] [chunks]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include "pmg_apply_sweep_pipe.h"
extern "C" void pmg_fe_pencil(int p, double *M, double *K);
#define PP C_P
#ifndef C_US
#define C_US 0
#endif
#ifndef C_FM
#define C_FM -1
#endif
#ifndef C_SG
#define C_SG 1
#endif
#ifndef C_RL
#define C_RL 0
#endif
#ifndef C_A2
#define C_A2 0
#endif
#ifndef C_EG
#define C_EG 0
#endif
#define CFG C_P, C_BX, C_BY, C_LZ, C_NT, C_US, C_FM, C_SG, C_RL, C_A2, C_EG
#define STR2(x) #x
#define STR(x) STR2(x)
#define CFGSTR STR(C_P) "," STR(C_BX) "," STR(C_BY) "," STR(C_LZ) "," STR(C_NT)
#define EXPNAME STR(C_NAME)
#ifndef MINB
#define MINB 3
#endif
using Tile = PmgSweepPipe<C_P, C_BX, C_BY, C_LZ, C_NT, C_US, C_FM, C_RL, C_EG>;
using RefTile = PmgSweepTile<C_P, C_BX, C_BY, C_LZ, C_NT, C_US, C_FM, 1, 0>;
struct Ex {
  Tile::ThreadState st;
#ifdef PMG_EXP_ROT
  // rotate the warps' roles from CTA to CTA: the CTAs resident on one SM (b, b + 148, b + 296 in the first wave) then put
  // their busy warps (phases 1 / 2 occupy the first ones) on different SM sub-partitions
  template <class F> __device__ __forceinline__ void for_each_thread(F f)
  {
    constexpr int NW = Tile::NT / 32;
    const int rot = (blockIdx.x / 148 + blockIdx.x) % NW;
    f((int)((((threadIdx.x >> 5) + rot) % NW) * 32 + (threadIdx.x & 31)), st);
  }
#else
  template <class F> __device__ __forceinline__ void for_each_thread(F f) { f((int)threadIdx.x, st); }
#endif
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void sync_some(int n) { if ((int)threadIdx.x < n) asm volatile("bar.sync 1, %0;\n" ::"r"(n) : "memory"); }
};
struct RefEx {
  RefTile::ThreadState st;
  template <class F> __device__ __forceinline__ void for_each_thread(F f) { f((int)threadIdx.x, st); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void sync_some(int n) { if ((int)threadIdx.x < n) asm volatile("bar.sync 1, %0;\n" ::"r"(n) : "memory"); }
};
__global__ void __launch_bounds__(RefTile::NT) ref_kern(const __grid_constant__ PmgSweepParams<PP> p)
{
  extern __shared__ __align__(128) double sm[];
  RefEx ex;
  const int b = blockIdx.x;
  RefTile::run(p, ex, sm, b % p.tiles_x, (b / p.tiles_x) % p.tiles_y, b / (p.tiles_x * p.tiles_y));
}
__global__ void maxdiff(const double *a, const double *b, size_t n, double *res)
{
  double m = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double d = fabs(a[i] - b[i]);
    if (!(d <= 1e300)) d = 1e300; // NaN (an unwritten or poisoned dof) counts as a mismatch, fmax would drop it
    m = fmax(m, d);
  }
  atomicMax((unsigned long long *)res, (unsigned long long)__double_as_longlong(m)); // non-negative doubles order like integers
}
__global__ void __launch_bounds__(Tile::NT, MINB) kern(const __grid_constant__ PmgSweepParams<PP> p)
{
  extern __shared__ __align__(128) double sm[];
  Ex ex;
  const int b = blockIdx.x;
  Tile::run(p, ex, sm, b % p.tiles_x, (b / p.tiles_x) % p.tiles_y, b / (p.tiles_x * p.tiles_y));
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
int main(int argc, char **argv)
{
  int n = argc > 1 ? atoi(argv[1]) : 0; if (n <= 0) n = (464 + PP / 2) / PP; const int reps = argc > 2 ? atoi(argv[2]) : 10;
  int chunks = argc > 3 ? atoi(argv[3]) : 0;
  constexpr int P = PP;
  PmgSweepParams<P> p{};
  p.nx = p.ny = p.nz = n; p.Nx = p.Ny = p.Nz = n * P + 1; p.faces = 0x3F;
  p.z0 = 0; p.nzl = p.Nz; p.cz_lo = 0; p.cz_hi = n; p.z_own_lo = 0; p.z_own_hi = p.Nz;
  constexpr int BXc = (Tile::CW - 1) / P, BYc = (Tile::RW - 1) / P;
  p.tiles_x = (n + BXc - 1) / BXc; p.tiles_y = (n + BYc - 1) / BYc;
  double M[100], K[100], h[3] = {1.0 / n, 1.0 / n, 1.0 / n};
  pmg_fe_pencil(P, M, K);
  pmg_sweep_fill_matrices<P>(p, M, K, h);
  const size_t N = (size_t)p.Nx * p.Ny * p.Nz;
  std::vector<double> hu(N);
  for (size_t i = 0; i < N; ++i) hu[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
  double *u, *b, *xo, *out, *tab;
  CK(cudaMalloc(&u, N * 8 + 8 * (size_t)p.Nx + 16)); CK(cudaMalloc(&b, N * 8 + 8 * (size_t)p.Nx + 16)); CK(cudaMalloc(&xo, N * 8 + 8 * (size_t)p.Nx + 16)); CK(cudaMalloc(&out, N * 8 + 8 * (size_t)p.Nx + 16));
  CK(cudaMalloc(&tab, 1000 * 8));
  CK(cudaMemcpy(u, hu.data(), N * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(b, hu.data(), N * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(xo, 0, N * 8)); CK(cudaMemset(tab, 0, 8000));
  p.u = u; p.b = b; p.xold = xo; p.out = out; p.f1 = 0.3; p.f2 = 0.1; p.dinv_tab = tab; p.dinv_vec = nullptr;
  int smem = Tile::SMEM_DOUBLES * 8;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Tile::NT, smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
  const int chunks_arg = chunks;
  // z-chunks: waves * (layers + 1 + 1/P) minimal for the CTA slots of the launch at hand (chosen per mode: the modes differ in
  // shared memory, hence in resident CTAs per SM)
  auto choose_chunks = [&](int per_sm_now) {
    int c_best = chunks_arg;
    if (c_best <= 0) {
      const int slots = 148 * per_sm_now, tiles = p.tiles_x * p.tiles_y; double best = -1;
      for (int c = 1; c <= n; ++c) { int lpc = (n + c - 1) / c; if ((n + lpc - 1) / lpc != c) continue;
        long waves = ((long)tiles * c + slots - 1) / slots; double cost = waves * (lpc + (c > 1 ? 1.0 + 1.0 / P : 0.0));
        if (best < 0 || cost < best) { best = cost; c_best = c; } }
    }
    p.layers_per_chunk = (n + c_best - 1) / c_best; p.n_chunks = (n + p.layers_per_chunk - 1) / p.layers_per_chunk;
    return p.tiles_x * p.tiles_y * p.n_chunks;
  };
  int grid = choose_chunks(per_sm);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#if C_FM >= 0
  for (int mode : {C_FM}) {
#else
  for (int mode : {0, 3}) {
#endif
    p.mode = mode; p.out = (mode == 3) ? xo : out;
    smem = Tile::smem_doubles(mode != 0) * 8;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Tile::NT, smem));
    grid = choose_chunks(per_sm);
    for (int i = 0; i < 3; ++i) kern<<<grid, Tile::NT, smem>>>(p);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) kern<<<grid, Tile::NT, smem>>>(p);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("%s P=%d n=%d N=%zu cfg=%s minb=%d regs=%d smem=%dKB ctas/sm=%d grid=%d chunks=%d mode=%d: %.3f ms %.1f GDoF/s\n",
           EXPNAME, P, n, N, CFGSTR, MINB, fa.numRegs, smem / 1024, per_sm, grid, p.n_chunks, mode, ms, N / ms / 1e6);
  }
  // self-check: same input through the plain kernel (same tile, NG threads) and the pipelined one, separate output vectors
  {
    const int mode = (C_FM >= 0) ? C_FM : 3;
    double *out2, *res; CK(cudaMalloc(&out2, N * 8 + 8 * (size_t)p.Nx + 16)); CK(cudaMalloc(&res, 8)); CK(cudaMemset(res, 0, 8));
    CK(cudaMemset(out, 0xFF, N * 8)); CK(cudaMemset(out2, 0x7F, N * 8)); // different garbage: an unwritten dof shows up
    p.mode = mode;
    const int rsmem = RefTile::smem_doubles(mode != 0) * 8;
    CK(cudaFuncSetAttribute(ref_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, rsmem));
    p.out = out2; ref_kern<<<grid, RefTile::NT, rsmem>>>(p);
    p.out = out; kern<<<grid, Tile::NT, Tile::smem_doubles(mode != 0) * 8>>>(p);
    maxdiff<<<592, 256>>>(out, out2, N, res);
    CK(cudaDeviceSynchronize());
    double hres; CK(cudaMemcpy(&hres, res, 8, cudaMemcpyDeviceToHost));
    printf("%s self-check mode=%d: max |pipelined - plain| = %.3e %s\n", EXPNAME, mode, hres, hres == 0.0 ? "(bitwise equal)" : "");
  }
  return 0;
}
