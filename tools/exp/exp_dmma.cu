// tools/exp/exp_dmma.cu -- can the FP64 tensor-core MMA (mma.sync.m8n8k4.f64, SASS DMMA) carry the sum-factorised sweeps?
// (BASELINE config 3: "HBM vs FP64-DMMA roofline crossover"; development tool, not part of the library.)
//
// Q7 is the degree the instruction fits exactly: 8 nodes per direction = the MMA's M and N, two k-steps of 4.  One warp holds
// one cell's 8 x 8 x 8 values in registers and applies the cell operator
//     A_cell u = Mz [ Kx My u + Mx Ky u ] + Kz Mx My u        (c = My u, d = Ky u;  g = Kx c + Mx d, m = Mx c;  out = Mz g + Kz m)
// plane by plane in z: the y sweep and the x sweep of a plane are 4 + 6 DMMAs whose operand / accumulator fragments chain
// WITHOUT any data movement -- with the 1-D matrix as the A operand, the result of the y sweep (rows = y_out, columns = x) is,
// lane for lane, a B operand of the x sweep (k = x in the order 2 (lane % 4) + j, which only permutes the matrix constants) --
// and the z sweep is output-stationary in 16 thread-local accumulators (32 DFMAs per plane).  Per cell: 80 DMMA + 256 DFMA
// warp instructions for 7 x 8^4 = 28 672 FMAs, i.e. 336 issue slots instead of 896.
//
// Cell-local input and output blocks (no gather from a shared-dof vector, no assembly of the cell results, no halo): this is
// the UPPER BOUND of any DMMA-based apply at Q7.  Printed: cells/s and the DoF rate it would correspond to (343 dofs per
// cell), next to the FP64 time of the instruction mix at the measured pipe rates.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/exp/bin/exp_dmma tools/exp/exp_dmma.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__constant__ double cM[64], cK[64]; // 1-D cell matrices, row-major 8 x 8 (the z sweep reads them as uniform operands)

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ZSWEEP = 0: without the z sweep's DFMAs (out = g + m of the last plane: what the DMMA part alone costs)
template <int ZSWEEP>
__global__ void __launch_bounds__(128, 4) k_cell_dmma(const double *__restrict__ U, double *__restrict__ O, int n_cells)
{
  const int lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  // A-operand fragments of the 1-D matrices: lane holds Mat[row = r][column = the k index of this lane in this k-step]
  double aMy[2], aKy[2], aMx[2], aKx[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    aMy[s] = cM[r * 8 + 4 * s + q]; aKy[s] = cK[r * 8 + 4 * s + q]; // y sweep: k-step s covers y = 4 s + q
    aMx[s] = cM[r * 8 + 2 * q + s]; aKx[s] = cK[r * 8 + 2 * q + s]; // x sweep: k-step j covers x = 2 q + j
  }
  for (int cell = warp; cell < n_cells; cell += n_warps) {
    const double *u = U + (size_t)cell * 512;
    double uz[8][2];
#pragma unroll
    for (int z = 0; z < 8; ++z)
#pragma unroll
      for (int s = 0; s < 2; ++s) uz[z][s] = u[z * 64 + (4 * s + q) * 8 + r]; // B fragment: [k = y = 4 s + q][n = x = r]
    double out[8][2];
#pragma unroll
    for (int zo = 0; zo < 8; ++zo) { out[zo][0] = 0.0; out[zo][1] = 0.0; }
#pragma unroll
    for (int z = 0; z < 8; ++z) {
      double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0; // c, d [y_out = r][x = 2 q + j]
      dmma(c0, c1, aMy[0], uz[z][0]); dmma(c0, c1, aMy[1], uz[z][1]);
      dmma(d0, d1, aKy[0], uz[z][0]); dmma(d0, d1, aKy[1], uz[z][1]);
      double g0 = 0.0, g1 = 0.0, m0 = 0.0, m1 = 0.0; // g, m [x_out = r][y = 2 q + j]
      dmma(g0, g1, aKx[0], c0); dmma(g0, g1, aKx[1], c1);
      dmma(g0, g1, aMx[0], d0); dmma(g0, g1, aMx[1], d1);
      dmma(m0, m1, aMx[0], c0); dmma(m0, m1, aMx[1], c1);
      if (ZSWEEP) {
#pragma unroll
        for (int zo = 0; zo < 8; ++zo) {
          out[zo][0] = fma(cM[zo * 8 + z], g0, fma(cK[zo * 8 + z], m0, out[zo][0]));
          out[zo][1] = fma(cM[zo * 8 + z], g1, fma(cK[zo * 8 + z], m1, out[zo][1]));
        }
      } else {
        out[z][0] = g0 + m0; out[z][1] = g1 + m1;
      }
    }
    double *o = O + (size_t)cell * 512;
#pragma unroll
    for (int zo = 0; zo < 8; ++zo)
#pragma unroll
      for (int j = 0; j < 2; ++j) o[zo * 64 + (2 * q + j) * 8 + r] = out[zo][j]; // [z][y = 2 q + j][x = r]
  }
}

// streaming copy of the same blocks: what the memory side of the experiment costs alone
__global__ void __launch_bounds__(128, 4) k_cell_copy(const double *__restrict__ U, double *__restrict__ O, int n_cells)
{
  const int lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int cell = warp; cell < n_cells; cell += n_warps) {
    const double *u = U + (size_t)cell * 512;
    double *o = O + (size_t)cell * 512;
    double uz[8][2];
#pragma unroll
    for (int z = 0; z < 8; ++z)
#pragma unroll
      for (int s = 0; s < 2; ++s) uz[z][s] = u[z * 64 + (4 * s + q) * 8 + r];
#pragma unroll
    for (int z = 0; z < 8; ++z)
#pragma unroll
      for (int s = 0; s < 2; ++s) o[z * 64 + (4 * s + q) * 8 + r] = uz[z][s];
  }
}

static void host_cell(const double *M, const double *K, const double *u, double *out)
{
  static double c[512], d[512], g[512], m[512];
  for (int z = 0; z < 8; ++z) for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
    double sc = 0, sd = 0;
    for (int k = 0; k < 8; ++k) { sc += M[y * 8 + k] * u[z * 64 + k * 8 + x]; sd += K[y * 8 + k] * u[z * 64 + k * 8 + x]; }
    c[z * 64 + y * 8 + x] = sc; d[z * 64 + y * 8 + x] = sd;
  }
  for (int z = 0; z < 8; ++z) for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
    double sg = 0, sm = 0;
    for (int k = 0; k < 8; ++k) { sg += K[x * 8 + k] * c[z * 64 + y * 8 + k] + M[x * 8 + k] * d[z * 64 + y * 8 + k]; sm += M[x * 8 + k] * c[z * 64 + y * 8 + k]; }
    g[z * 64 + y * 8 + x] = sg; m[z * 64 + y * 8 + x] = sm;
  }
  for (int z = 0; z < 8; ++z) for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
    double s = 0;
    for (int k = 0; k < 8; ++k) s += M[z * 8 + k] * g[k * 64 + y * 8 + x] + K[z * 8 + k] * m[k * 64 + y * 8 + x];
    out[z * 64 + y * 8 + x] = s;
  }
}

template <class F> static float time_ms(F f, int reps)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main(int argc, char **argv)
{
  const int n_cells = argc > 1 ? atoi(argv[1]) : 262144; // 2 x 1 GiB of cell blocks: far beyond L2
  const int ctas_per_sm = argc > 2 ? atoi(argv[2]) : 4;
  std::vector<double> M(64), K(64);
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) { // symmetric test matrices (the kernel does not use more structure)
    M[i * 8 + j] = 1.0 / (1.0 + abs(i - j)) + (i == j ? 2.0 : 0.0);
    K[i * 8 + j] = (i == j ? 3.0 : -0.5 / (1.0 + (i - j) * (i - j)));
  }
  CK(cudaMemcpyToSymbol(cM, M.data(), 64 * sizeof(double)));
  CK(cudaMemcpyToSymbol(cK, K.data(), 64 * sizeof(double)));
  const size_t n = (size_t)n_cells * 512;
  std::vector<double> hu(n);
  unsigned long long sd = 12345;
  for (size_t i = 0; i < n; ++i) { sd = sd * 6364136223846793005ull + 1442695040888963407ull; hu[i] = (double)(sd >> 11) / 9007199254740992.0 - 0.5; }
  double *U, *O;
  CK(cudaMalloc(&U, n * sizeof(double))); CK(cudaMalloc(&O, n * sizeof(double)));
  CK(cudaMemcpy(U, hu.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = sms * ctas_per_sm;
  k_cell_dmma<1><<<grid, 128>>>(U, O, n_cells);
  CK(cudaDeviceSynchronize());
  // parity of the fragment chaining: three cells against the plain triple loops
  double worst = 0.0;
  for (int cell : {0, n_cells / 2 + 1, n_cells - 1}) {
    std::vector<double> got(512), ref(512);
    CK(cudaMemcpy(got.data(), O + (size_t)cell * 512, 512 * sizeof(double), cudaMemcpyDeviceToHost));
    host_cell(M.data(), K.data(), hu.data() + (size_t)cell * 512, ref.data());
    double num = 0, den = 0;
    for (int i = 0; i < 512; ++i) { num += (got[i] - ref[i]) * (got[i] - ref[i]); den += ref[i] * ref[i]; }
    worst = fmax(worst, sqrt(num / den));
  }
  printf("DMMA cell operator Q7: rel l2 difference to the triple loops %.3e (%s)\n", worst, worst < 1e-13 ? "ok" : "MISMATCH");
  const float t_full = time_ms([&] { k_cell_dmma<1><<<grid, 128>>>(U, O, n_cells); }, 10);
  const float t_mma = time_ms([&] { k_cell_dmma<0><<<grid, 128>>>(U, O, n_cells); }, 10);
  const float t_copy = time_ms([&] { k_cell_copy<<<grid, 128>>>(U, O, n_cells); }, 10);
  CK(cudaGetLastError());
  const double dofs = 343.0 * n_cells;
  const double fma_slots = (80.0 * 256 + 256.0 * 32) * n_cells; // per cell: 80 DMMA x 256 + 256 DFMA x 32 lanes = 28 672 FMAs
  printf("cells %d (%.0f M dofs at 343 per cell), grid %d x 128 threads\n", n_cells, dofs / 1e6, grid);
  printf("y+x sweeps as DMMA, z sweep as DFMA : %.3f ms  %.1f GDoF/s-equivalent  %.2f TFLOP/s FP64\n", t_full, dofs / t_full / 1e6, 2 * fma_slots / t_full / 1e9);
  printf("DMMA part only (no z sweep)         : %.3f ms  %.1f GDoF/s-equivalent  %.2f TFLOP/s FP64\n", t_mma, dofs / t_mma / 1e6, 2 * 80.0 * 256 * n_cells / t_mma / 1e9);
  printf("copy of the cell blocks only        : %.3f ms  %.1f GB/s\n", t_copy, 2.0 * n * 8 / t_copy / 1e6);
  return worst < 1e-13 ? 0 : 1;
}
