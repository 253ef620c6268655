// Shared helpers for the pvo_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define PVO_OK 0
#define PVO_ERR_INVALID 1
#define PVO_ERR_CUDA 2
#define PVO_ERR_UNSUPPORTED 3
#define PVO_ERR_WORKSPACE 4

#define PVO_DTYPE_F32 0
#define PVO_DTYPE_F16 1

#define PVO_SM_COUNT 148

static inline int pvo_check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? PVO_OK : PVO_ERR_CUDA;
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
