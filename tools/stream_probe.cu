// Streaming probe (bring-up tool): what does y = x * k + o reach on a 67 MB plane with 128-bit vs 256-bit
// accesses, different loads in flight per thread and grid sizes, with the L2 flushed before every launch?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/stream_probe tools/stream_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld8(const float* p) {
  f8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ void st8(float* p, const f8& r) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
               "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}

template <int U>
__global__ void __launch_bounds__(256) k_v4(const float4* __restrict__ x, float4* __restrict__ y, size_t n4, float k, float o) {
  size_t i = ((size_t)blockIdx.x * 256 * U) + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * 256 * U;
  for (; i < n4; i += stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 256 < n4) v[u] = __ldg(x + i + u * 256);
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 256 < n4) {
      float4 t = v[u]; t.x = fmaxf(t.x * k + o, 0.f); t.y = fmaxf(t.y * k + o, 0.f); t.z = fmaxf(t.z * k + o, 0.f); t.w = fmaxf(t.w * k + o, 0.f);
      y[i + u * 256] = t;
    }
  }
}
template <int U>
__global__ void __launch_bounds__(256) k_v8(const float* __restrict__ x, float* __restrict__ y, size_t n8, float k, float o) {
  size_t i = ((size_t)blockIdx.x * 256 * U) + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * 256 * U;
  for (; i < n8; i += stride) {
    f8 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 256 < n8) v[u] = ld8(x + (i + u * 256) * 8);
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 256 < n8) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[u].v[e] = fmaxf(v[u].v[e] * k + o, 0.f);
      st8(y + (i + u * 256) * 8, v[u]);
    }
  }
}
template <int U>
__global__ void __launch_bounds__(256) r_v8(const float* __restrict__ x, float* __restrict__ out, size_t n8) {
  size_t i = ((size_t)blockIdx.x * 256 * U) + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * 256 * U;
  float s = 0.f;
  for (; i < n8; i += stride) {
    f8 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 256 < n8) v[u] = ld8(x + (i + u * 256) * 8);
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 256 < n8) {
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[u].v[e];
    }
  }
  if (s == 12345.f) out[0] = s;
}

static float* flushbuf;
template <class F> static float timed(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ts;
  f(); f(); cudaDeviceSynchronize();
  for (int r = 0; r < 9; ++r) {
    cudaMemsetAsync(flushbuf, 0, 256 << 20);
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ts.push_back(ms * 1e3f);
  }
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

int main() {
  cudaMalloc(&flushbuf, 256 << 20);
  for (double mb : {67.1, 268.4}) {
    const size_t n = (size_t)(mb * 1e6 / 4) / 2048 * 2048;
    float *x, *y; cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4); cudaMemset(x, 0, n * 4);
    printf("plane %.1f MB\n", mb);
    for (int cps : {2, 4, 8, 16, 0}) {
      const size_t n4 = n / 4, n8 = n / 8;
      auto g = [&](size_t items, int U) { size_t full = (items + 256 * U - 1) / (256 * U); return (unsigned)(cps ? std::min<size_t>(full, 148 * cps) : full); };
      float t;
      t = timed([&] { k_v4<4><<<g(n4, 4), 256>>>((const float4*)x, (float4*)y, n4, 1.1f, 0.1f); });
      printf("  rw v4 U4 ctas/sm %2d: %6.1f us %5.0f GB/s\n", cps, t, 2 * n * 4 / t * 1e-3);
      t = timed([&] { k_v8<2><<<g(n8, 2), 256>>>(x, y, n8, 1.1f, 0.1f); });
      printf("  rw v8 U2 ctas/sm %2d: %6.1f us %5.0f GB/s\n", cps, t, 2 * n * 4 / t * 1e-3);
      t = timed([&] { k_v8<4><<<g(n8, 4), 256>>>(x, y, n8, 1.1f, 0.1f); });
      printf("  rw v8 U4 ctas/sm %2d: %6.1f us %5.0f GB/s\n", cps, t, 2 * n * 4 / t * 1e-3);
      t = timed([&] { r_v8<4><<<g(n8, 4), 256>>>(x, y, n8); });
      printf("  r  v8 U4 ctas/sm %2d: %6.1f us %5.0f GB/s\n", cps, t, n * 4 / t * 1e-3);
    }
    cudaFree(x); cudaFree(y);
  }
  return 0;
}
