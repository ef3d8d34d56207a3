// tools/exp/tma_test.cu -- minimal TMA 2-D tile load test (development tool): tma_test <dtype 0=f64 1=u64 2=f32x2> <box0> <box1>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "pmg_apply_sweep.h"
__global__ void k(const __grid_constant__ PmgTmap map, const PmgTmap *gmap, double *out, int x, int y, int n, int variant)
{
  extern __shared__ __align__(128) double sm[];
  uint64_t *bar = (uint64_t *)(sm + 4096);
  if (threadIdx.x == 0) {
    pmg_mbar_init(bar, 1);
    pmg_mbar_init_fence();
    pmg_mbar_arrive_expect(bar, n * 8);
    const PmgTmap *m = (variant & 1) ? gmap : &map;
    if (variant & 2) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               ::"r"((unsigned)__cvta_generic_to_shared(sm)), "l"(m), "r"(x), "r"(y),
                 "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
    } else pmg_tma_load_2d(sm, m, x, y, bar);
  }
  __syncthreads();
  pmg_mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char **argv)
{
  const int dt = argc > 1 ? atoi(argv[1]) : 0, box0 = argc > 2 ? atoi(argv[2]) : 22, box1 = argc > 3 ? atoi(argv[3]) : 11;
  const int Nx = 65, rows = 40;
  double *v, *out; cudaMalloc(&v, (Nx * rows + Nx + 2) * 8); cudaMalloc(&out, 4096 * 8);
  double h[Nx * rows + Nx + 2]; for (int i = 0; i < Nx * rows + Nx + 2; ++i) h[i] = i;
  cudaMemcpy(v, h, sizeof(h), cudaMemcpyHostToDevice);
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PmgTmap map;
  const int f = (dt == 2) ? 2 : 1;
  const cuuint64_t gdim[2] = {(cuuint64_t)2 * Nx * f, (cuuint64_t)((rows + 1) / 2)};
  const cuuint64_t gstride[1] = {(cuuint64_t)2 * Nx * 8};
  const cuuint32_t box[2] = {(cuuint32_t)box0 * f, (cuuint32_t)box1}, estr[2] = {1, 1};
  const CUtensorMapDataType type = dt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : dt == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult rc = ((encode_fn)fn)((CUtensorMap *)&map, type, 2, (void *)v, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("dtype %d box %dx%d encode rc=%d\n", dt, box0, box1, (int)rc);
  if (rc) return 1;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4200 * 8);
  const int variant = argc > 4 ? atoi(argv[4]) : 0;
  PmgTmap *gmap; cudaMalloc(&gmap, 128); cudaMemcpy(gmap, &map, 128, cudaMemcpyHostToDevice);
  printf("variant %d\n", variant);
  k<<<1, 128, 4200 * 8>>>(map, gmap, out, 3 * f, 2, box0 * box1, variant);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e) return 1;
  double r[4096]; cudaMemcpy(r, out, box0 * box1 * 8, cudaMemcpyDeviceToHost);
  printf("row0: %g %g %g ... row1: %g %g (expect %d %d %d ... %d %d)\n", r[0], r[1], r[2], r[box0], r[box0 + 1], 2 * 2 * Nx + 3, 2 * 2 * Nx + 4,
         2 * 2 * Nx + 5, 3 * 2 * Nx + 3, 3 * 2 * Nx + 4);
  return 0;
}
