// Shared helpers for the cpsd_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
// the public prototypes: every extern "C" definition in csrc/ is compiled against them, so a
// drift between the header and an implementation is a compile error
#include "../../include/cpsd_b200.h"

#define CPSD_OK 0
#define CPSD_ERR_INVALID 1
#define CPSD_ERR_CUDA 2
#define CPSD_ERR_UNSUPPORTED 3

extern "C" void cpsd_set_error(const char* msg);
// counts kernel launches made through the library (bench.py reads it)
extern "C" void cpsd_count_launch(int n);

#define CPSD_CHECK_ARG(cond, msg)            \
  do {                                       \
    if (!(cond)) {                           \
      cpsd_set_error(msg);                   \
      return CPSD_ERR_INVALID;               \
    }                                        \
  } while (0)

#define CPSD_LAUNCH_CHECK()                                   \
  do {                                                        \
    cpsd_count_launch(1);                                     \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) {                                 \
      cpsd_set_error(cudaGetErrorString(e__));                \
      return CPSD_ERR_CUDA;                                   \
    }                                                         \
  } while (0)

#define CPSD_CUDA(call)                                       \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) {                                 \
      cpsd_set_error(cudaGetErrorString(e__));                \
      return CPSD_ERR_CUDA;                                   \
    }                                                         \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of doubles; `red` must hold >= 33 doubles of shared memory.
// All threads get the result.  blockDim.x must be a multiple of 32.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = (lane < nw) ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
