// Shared helpers for the SRGAN B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "srgan_b200.h"

namespace srgan {

void set_error(const char* fmt, ...);

#define SRGAN_CHECK_ARG(cond, msg)                                   \
  do {                                                               \
    if (!(cond)) {                                                   \
      ::srgan::set_error("%s: %s", __func__, msg);                   \
      return SRGAN_E_BADARG;                                         \
    }                                                                \
  } while (0)

// Launch-error check that never synchronises.
#define SRGAN_RETURN_LAUNCH()                                        \
  do {                                                               \
    cudaError_t e__ = cudaGetLastError();                            \
    if (e__ != cudaSuccess) {                                        \
      ::srgan::set_error("%s: %s", __func__, cudaGetErrorString(e__)); \
      return (int)e__;                                               \
    }                                                                \
    return SRGAN_OK;                                                 \
  } while (0)

constexpr int kNumSMs = 148;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE function attribute: a process that drives several GPUs
// must set it once on each.  `done` is a per-call-site bitmask indexed by device ordinal (devices >= 64: set always).
template <typename F>
inline cudaError_t ensure_dyn_smem(F* fn, int bytes, unsigned long long* done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && ((*done >> dev) & 1ull)) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev < 64) *done |= 1ull << dev;
  return e;
}

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case SRGAN_ACT_RELU:  return v > 0.f ? v : 0.f;
    case SRGAN_ACT_LRELU: return v > 0.f ? v : v * slope;
    case SRGAN_ACT_TANH:  return tanhf(v);
    default:              return v;
  }
}
// derivative of the activation expressed through its pre-activation value v
__device__ __forceinline__ float act_grad_pre(float v, int act, float slope) {
  switch (act) {
    case SRGAN_ACT_RELU:  return v > 0.f ? 1.f : 0.f;
    case SRGAN_ACT_LRELU: return v > 0.f ? 1.f : slope;
    case SRGAN_ACT_TANH:  { float t = tanhf(v); return 1.f - t * t; }
    default:              return 1.f;
  }
}
// derivative expressed through the activation OUTPUT y (what autograd saves)
__device__ __forceinline__ float act_grad_out(float y, int act, float slope) {
  switch (act) {
    case SRGAN_ACT_RELU:  return y > 0.f ? 1.f : 0.f;
    case SRGAN_ACT_LRELU: return y > 0.f ? 1.f : slope;
    case SRGAN_ACT_TANH:  return 1.f - y * y;
    default:              return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum with a fixed combination order (deterministic). `red` needs 32 floats.
// Result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();               // protect `red` from a previous use
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0) {
    r = lane < nw ? red[lane] : 0.f;
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

}  // namespace srgan
