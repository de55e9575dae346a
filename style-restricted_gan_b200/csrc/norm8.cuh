// 8-channel vector access and activation helpers shared by norm8.cu (two-kernel path) and norm8c.cu (one-pass
// cluster kernels).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace srgan {

// ---- 8-channel vector access
template <typename T> struct Raw8;
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<__nv_bfloat16> { uint4 q; };

__device__ __forceinline__ Raw8<float> ld_raw(const float* p) {
  Raw8<float> r;
  r.a = __ldg(reinterpret_cast<const float4*>(p));
  r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  return r;
}
__device__ __forceinline__ Raw8<__nv_bfloat16> ld_raw(const __nv_bfloat16* p) {
  Raw8<__nv_bfloat16> r;
  r.q = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}
__device__ __forceinline__ void unpack(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void unpack(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
  const uint32_t w[4] = {r.q.x, r.q.y, r.q.z, r.q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { v[2 * e] = __uint_as_float(w[e] << 16); v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    w[e] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void ld8f(const float* p, float (&v)[8]) {      // per-(n,c) tables and parameters
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T> struct Flight { static constexpr int kRows = sizeof(T) == 2 ? 8 : 4; };   // 128 B per thread

template <int ACT>
__device__ __forceinline__ float n8_act(float v, float slope) {
  if (ACT == SRGAN_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == SRGAN_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (ACT == SRGAN_ACT_TANH) return tanhf(v);
  return v;
}
template <int ACT>
__device__ __forceinline__ float n8_act_grad(float v, float slope) {
  if (ACT == SRGAN_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  if (ACT == SRGAN_ACT_LRELU) return v > 0.f ? 1.f : slope;
  if (ACT == SRGAN_ACT_TANH) { const float t = tanhf(v); return 1.f - t * t; }
  return 1.f;
}

// ---- per-channel constants, shared by the two-kernel and the one-pass kernels (explicit roundings: both paths must
// produce the same bits from the same sums)
// S1 = sum (x - pv), S2 = sum (x - pv)^2 over the plane
__device__ __forceinline__ void n8_mean_rstd(double S1, double S2, float pv, float inv_hw, float eps, float* mu,
                                             float* rs) {
  const float m = __fmul_rn((float)S1, inv_hw);
  const float var = fmaxf(__fmaf_rn(-m, m, __fmul_rn((float)S2, inv_hw)), 0.f);
  *mu = __fadd_rn(pv, m);
  *rs = rsqrtf(__fadd_rn(var, eps));
}
// v = x*k + o, xh = x*rs + cc
__device__ __forceinline__ void n8_consts(float mu, float rs, float g, float b, float tb, float* k, float* o,
                                          float* cc) {
  const float c = -__fmul_rn(mu, rs);
  *cc = c;
  *k = __fmul_rn(rs, g);
  *o = __fmaf_rn(__fadd_rn(tb, c), g, b);
}
// dx = k*dv + A*x + B  (= rstd*gamma*(dv - m1/HW - xh*m2/HW))
__device__ __forceinline__ void n8_bwd_consts(float k, float rs, float cc, float m1, float m2, float inv_hw, float* A,
                                              float* B) {
  const float a1 = __fmul_rn(m1, inv_hw), a2 = __fmul_rn(m2, inv_hw);
  *A = __fmul_rn(__fmul_rn(-k, a2), rs);
  *B = __fmul_rn(-k, __fmaf_rn(a2, cc, a1));
}

// norm8c.cu: one-pass cluster kernels.  Return false when the plane is not eligible (the caller runs the two-kernel
// path); otherwise the launch status is in *status.
bool norm8c_fwd(const void* x, bool xb, void* y, bool yb, float* mean, float* rstd, const float* gamma,
                const float* beta, const float* cbias, const void* residual, int N, int HW, int C, float eps, int act,
                float slope, cudaStream_t st, cudaError_t* status);
bool norm8c_bwd(const void* dy, bool yb, const void* x, bool xb, const float* mean, const float* rstd,
                const float* gamma, const float* beta, const float* cbias, void* dx, float* s1, float* s2, int N,
                int HW, int C, int act, float slope, cudaStream_t st, cudaError_t* status);

}  // namespace srgan
