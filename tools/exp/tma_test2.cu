// known-good pattern from the CUDA programming guide (libcu++), to compare with the hand-written PTX
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void k(const __grid_constant__ CUtensorMap tensor_map, double *out, int x, int y)
{
  __shared__ alignas(128) double smem_buffer[16 * 4];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < 64; i += blockDim.x) out[i] = smem_buffer[i];
}
int main()
{
  const int Nx = 65, rows = 40;
  double *v, *out; cudaMalloc(&v, (Nx * rows + Nx + 2) * 8); cudaMalloc(&out, 4096 * 8);
  double h[Nx * rows + Nx + 2]; for (int i = 0; i < Nx * rows + Nx + 2; ++i) h[i] = i;
  cudaMemcpy(v, h, sizeof(h), cudaMemcpyHostToDevice);
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)2 * Nx, (cuuint64_t)((rows + 1) / 2)};
  const cuuint64_t gstride[1] = {(cuuint64_t)2 * Nx * 8};
  const cuuint32_t box[2] = {16, 4}, estr[2] = {1, 1};
  CUresult rc = ((encode_fn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)v, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d q=%d\n", (int)rc, (int)q);
  k<<<1, 128>>>(map, out, 3, 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e) return 1;
  double r[64]; cudaMemcpy(r, out, 64 * 8, cudaMemcpyDeviceToHost);
  printf("row0: %g %g %g ... row1: %g %g (expect %d %d %d ... %d %d)\n", r[0], r[1], r[2], r[16], r[17], 2 * 2 * Nx + 3, 2 * 2 * Nx + 4, 2 * 2 * Nx + 5, 3 * 2 * Nx + 3, 3 * 2 * Nx + 4);
  return 0;
}
