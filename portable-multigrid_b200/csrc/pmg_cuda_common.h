// pmg_cuda_common.h -- shared helpers of the CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "pmg.h"

extern "C" void pmg_set_error(const char *fmt, ...);
extern "C" void pmg_count_launch(int n);

#define PMG_CUDA_CHECK(call)                                                                 \
  do {                                                                                       \
    cudaError_t pmg_e_ = (call);                                                             \
    if (pmg_e_ != cudaSuccess) {                                                             \
      pmg_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(pmg_e_)); \
      return PMG_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

/* the fused-push descriptor of the pmgk_apply_push call in progress on this thread (csrc/pmg_apply.cu), read by the plane
   kernel's launcher; NULL for every other launch */
#include "pmg_kernels.h"
extern thread_local const pmgk_push *pmg_tl_push;
