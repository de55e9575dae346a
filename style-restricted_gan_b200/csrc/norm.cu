// Fused (conditional) instance normalisation, forward and backward, NHWC fp32.
//
//   xh = (x - mean_hw) * rstd_hw ; v = (xh + cbias[n][c]) * gamma[c] + beta[c] ; y = act(v) (+ residual)
//
// HBM-bound streaming design: every pass is a grid of (pixel slice, image, channel chunk) CTAs whose threads
// read whole pixel rows with float4 loads (consecutive threads -> consecutive 16 B: every 128-byte line is
// used completely), 4 independent loads in flight per thread and several CTAs per SM, no shared-memory
// staging, so the memory system sees tens of KB in flight per SM.
//   forward : (1) statistics  - per-slice partial sums of (x - p), (x - p)^2 about a per-channel pivot p
//                               (the first pixel; keeps the one-pass variance well conditioned);
//             (2) apply       - every CTA folds the <= 64 slice partials of its channels in a fixed order
//                               (deterministic), then normalises its slice: read x, write y.
//   backward: (1) reduce      - partial sums of dv and dv*xh (dv = dy * act'(v));
//             (2) apply       - dx = rstd*gamma * (dv - mean(dv) - xh * mean(dv*xh)).
// Algorithmic traffic: forward 2 reads + 1 write of the plane (the second read is an L2 hit for planes that
// fit the 126 MB L2), backward 4 reads + 1 write.
#include <cuda_bf16.h>
#include "common.cuh"
#include <stdlib.h>

namespace srgan {

constexpr int kNormThreads = 256;
constexpr int kNormMaxSlices = 64;

struct NormP {
  int N, HW, C;
  int q4;         // float4 per pixel row (C / 4)
  int TPR;        // threads per pixel row == float4 per channel chunk
  int RPP;        // pixel rows per pass
  int SL, slice;  // pixel slices per image, pixels per slice
  float eps, slope;
  int act;
  int given;      // statistics (forward) / correction coefficients (backward) are read from the per-(n,c) arrays
};

struct NormIdx { int cg, row, c4, px0, px1; bool active; };

__device__ __forceinline__ NormIdx norm_idx(const NormP& p) {
  NormIdx i;
  i.cg = threadIdx.x % p.TPR;
  i.row = threadIdx.x / p.TPR;
  i.active = i.row < p.RPP;
  i.c4 = blockIdx.z * p.TPR + i.cg;
  i.px0 = blockIdx.x * p.slice;
  i.px1 = min(p.HW, i.px0 + p.slice);
  return i;
}

__device__ __forceinline__ float4 f4(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ void acc4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// Partial sums in double (the default): each thread sums an ATOM - its 4 (forward) or 2 (backward) rows of one
// aligned group of 4 * RPP pixel rows - in fp32, in an order fixed by the row index alone, and everything above the
// atom (the thread's running sum, the CTA's row reduction, the slice partials, their fold) is added in fp64.  Slices
// start on atom boundaries, so the statistics do not depend on how many slices an image is cut into - i.e. on how
// many images the rank holds (plan_norm sizes the grid to one wave) - up to a 2^-53 reassociation error that survives
// the final rounding to fp32 about once in 10^9 values.  That is what makes an N-rank data-parallel step reproduce
// the 1-GPU step on the same global batch (tools/dp_check.py); with fp32 partials the summation order leaks into the
// statistics and TF32 operand truncation downstream amplifies it to 3e-5 on the losses.
struct alignas(16) D4 { double x, y, z, w; };
__device__ __forceinline__ void acc4(D4& a, const D4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void acc4(D4& a, const float4& b) {
  a.x += (double)b.x; a.y += (double)b.y; a.z += (double)b.z; a.w += (double)b.w;
}
template <typename V> __device__ __forceinline__ V vzero();
template <> __device__ __forceinline__ float4 vzero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ D4 vzero<D4>() { return D4{0., 0., 0., 0.}; }
__device__ __forceinline__ float4 to_f4(const float4& v) { return v; }
__device__ __forceinline__ float4 to_f4(const D4& v) { return make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w); }
template <typename V> __device__ __forceinline__ V ld_part(const V* q) { return *q; }
template <> __device__ __forceinline__ float4 ld_part<float4>(const float4* q) { return __ldg(q); }

// Sum (a, b) over the pixel rows of the CTA in a fixed order; threads with row == 0 get the result.
template <typename V>
__device__ __forceinline__ void rows_reduce(const NormP& p, const NormIdx& i, V& a, V& b, V* sa, V* sb) {
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  if (i.row == 0) {
    V ra = vzero<V>(), rb = vzero<V>();
    for (int r = 0; r < p.RPP; ++r) { acc4(ra, sa[r * p.TPR + i.cg]); acc4(rb, sb[r * p.TPR + i.cg]); }
    a = ra; b = rb;
  }
}

// Fold the SL slice partials of (image n, float4 column c4) in slice order.
template <typename V>
__device__ __forceinline__ void fold_parts(const NormP& p, const V* __restrict__ part, int n, int c4, float4& s1,
                                           float4& s2) {
  V t1 = vzero<V>(), t2 = vzero<V>();
  const V* p1 = part + (size_t)n * p.SL * p.q4 + c4;
  const V* p2 = p1 + (size_t)p.N * p.SL * p.q4;
  for (int s = 0; s < p.SL; ++s) { acc4(t1, ld_part(p1 + (size_t)s * p.q4)); acc4(t2, ld_part(p2 + (size_t)s * p.q4)); }
  s1 = to_f4(t1); s2 = to_f4(t2);
}

// Storage-type views of an activation tensor as groups of 4 channels (16 B of fp32, 8 B of bf16): the kernels index
// pixels / channel groups, the view does the load / store and the conversion.  Arithmetic is fp32 either way.
template <typename T> struct In4;
template <> struct In4<float> {
  const float* p;
  __device__ __forceinline__ In4 operator+(size_t i) const { return In4{p + 4 * i}; }
  __device__ __forceinline__ float4 ld() const { return __ldg(reinterpret_cast<const float4*>(p)); }
};
template <> struct In4<__nv_bfloat16> {
  const __nv_bfloat16* p;
  __device__ __forceinline__ In4 operator+(size_t i) const { return In4{p + 4 * i}; }
  __device__ __forceinline__ float4 ld() const {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                       __uint_as_float(q.y & 0xffff0000u));
  }
};
template <typename T> struct Out4;
template <> struct Out4<float> {
  float* p;
  __device__ __forceinline__ Out4 operator+(size_t i) const { return Out4{p + 4 * i}; }
  __device__ __forceinline__ void st(const float4& v) const { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Out4<__nv_bfloat16> {
  __nv_bfloat16* p;
  __device__ __forceinline__ Out4 operator+(size_t i) const { return Out4{p + 4 * i}; }
  __device__ __forceinline__ void st(const float4& v) const {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
};

// partial layout: part[(n * SL + slice) * q4 + c4], two planes (first, second moment) of N*SL*q4 V each
template <typename V, typename T = float>
__global__ void __launch_bounds__(kNormThreads, 4) inorm_stats_kernel(NormP p, const T* __restrict__ x,
                                                                   V* __restrict__ part) {
  __shared__ V sa[kNormThreads], sb[kNormThreads];
  const NormIdx i = norm_idx(p);
  const int n = blockIdx.y;
  const In4<T> xg = In4<T>{x} + ((size_t)n * p.HW * p.q4 + i.c4);
  V s1 = vzero<V>(), s2 = vzero<V>();
  if (i.active) {
    const float4 pv = xg.ld();                       // pivot: first pixel of the plane
    const int step = p.RPP;
    int r = i.px0 + i.row;
#define SRGAN_ACC(v, t1, t2) { float a = v.x - pv.x, b = v.y - pv.y, c = v.z - pv.z, d = v.w - pv.w; \
                       t1.x += a; t1.y += b; t1.z += c; t1.w += d; t2.x += a * a; t2.y += b * b; t2.z += c * c; t2.w += d * d; }
    if constexpr (sizeof(V) == sizeof(float4)) {
      for (; r + 3 * step < i.px1; r += 4 * step) {
        float4 v0 = (xg + (size_t)r * p.q4).ld(), v1 = (xg + (size_t)(r + step) * p.q4).ld();
        float4 v2 = (xg + (size_t)(r + 2 * step) * p.q4).ld(), v3 = (xg + (size_t)(r + 3 * step) * p.q4).ld();
        SRGAN_ACC(v0, s1, s2) SRGAN_ACC(v1, s1, s2) SRGAN_ACC(v2, s1, s2) SRGAN_ACC(v3, s1, s2)
      }
      for (; r < i.px1; r += step) { float4 v0 = (xg + (size_t)r * p.q4).ld(); SRGAN_ACC(v0, s1, s2) }
    } else {
      for (; r + 3 * step < i.px1; r += 4 * step) {      // one atom per iteration
        float4 v0 = (xg + (size_t)r * p.q4).ld(), v1 = (xg + (size_t)(r + step) * p.q4).ld();
        float4 v2 = (xg + (size_t)(r + 2 * step) * p.q4).ld(), v3 = (xg + (size_t)(r + 3 * step) * p.q4).ld();
        float4 a1 = f4(0.f), a2 = f4(0.f);
        SRGAN_ACC(v0, a1, a2) SRGAN_ACC(v1, a1, a2) SRGAN_ACC(v2, a1, a2) SRGAN_ACC(v3, a1, a2)
        acc4(s1, a1); acc4(s2, a2);
      }
      if (r < i.px1) {                                   // the image's last, incomplete atom
        float4 a1 = f4(0.f), a2 = f4(0.f);
        for (; r < i.px1; r += step) { float4 v0 = (xg + (size_t)r * p.q4).ld(); SRGAN_ACC(v0, a1, a2) }
        acc4(s1, a1); acc4(s2, a2);
      }
    }
#undef SRGAN_ACC
  }
  rows_reduce(p, i, s1, s2, sa, sb);
  if (i.row == 0) {
    const size_t o = ((size_t)n * p.SL + blockIdx.x) * p.q4 + i.c4;
    part[o] = s1;
    part[(size_t)p.N * p.SL * p.q4 + o] = s2;
  }
}

// T: storage type of x (the producing convolution's output), TY: storage type of y and of the residual
template <typename V, typename T = float, typename TY = T>
__global__ void __launch_bounds__(kNormThreads, 4) inorm_apply_kernel(
    NormP p, const T* __restrict__ x, const V* __restrict__ part, TY* __restrict__ y,
    float* mean_out, float* rstd_out, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ cbias, const TY* __restrict__ residual) {
  const NormIdx i = norm_idx(p);
  if (!i.active) return;
  const int n = blockIdx.y;
  const size_t plane = (size_t)n * p.HW * p.q4 + i.c4;
  const In4<T> xg = In4<T>{x} + plane;
  const int c = i.c4 * 4;
  float4 mu, rs;
  if (p.given) {
    mu = __ldg(reinterpret_cast<const float4*>(mean_out + (size_t)n * p.C + c));
    rs = __ldg(reinterpret_cast<const float4*>(rstd_out + (size_t)n * p.C + c));
  } else {
    float4 s1, s2;
    fold_parts(p, part, n, i.c4, s1, s2);
    const float4 pv = xg.ld();
    const float inv = 1.f / (float)p.HW;
#define SRGAN_STAT(f) { float m = s1.f * inv; float var = fmaxf(s2.f * inv - m * m, 0.f); mu.f = pv.f + m; rs.f = rsqrtf(var + p.eps); }
    SRGAN_STAT(x) SRGAN_STAT(y) SRGAN_STAT(z) SRGAN_STAT(w)
#undef SRGAN_STAT
  }
  if (!p.given && blockIdx.x == 0 && i.row == 0) {
    reinterpret_cast<float4*>(mean_out + (size_t)n * p.C + c)[0] = mu;
    reinterpret_cast<float4*>(rstd_out + (size_t)n * p.C + c)[0] = rs;
  }
  const float4 g = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : f4(1.f);
  const float4 b = beta ? __ldg(reinterpret_cast<const float4*>(beta + c)) : f4(0.f);
  const float4 tb = cbias ? __ldg(reinterpret_cast<const float4*>(cbias + (size_t)n * p.C + c)) : f4(0.f);
  // y = act(x * k + o):  k = rstd*gamma, o = (cbias - mean*rstd)*gamma + beta
  const float4 k = make_float4(rs.x * g.x, rs.y * g.y, rs.z * g.z, rs.w * g.w);
  const float4 o = make_float4((tb.x - mu.x * rs.x) * g.x + b.x, (tb.y - mu.y * rs.y) * g.y + b.y,
                               (tb.z - mu.z * rs.z) * g.z + b.z, (tb.w - mu.w * rs.w) * g.w + b.w);
  const Out4<TY> yg = Out4<TY>{y} + plane;
  const bool has_res = residual != nullptr;
  const In4<TY> rg = In4<TY>{residual} + (has_res ? plane : 0);
  auto one = [&](float4 v, int r) {
    float4 t;
    t.x = apply_act(fmaf(v.x, k.x, o.x), p.act, p.slope);
    t.y = apply_act(fmaf(v.y, k.y, o.y), p.act, p.slope);
    t.z = apply_act(fmaf(v.z, k.z, o.z), p.act, p.slope);
    t.w = apply_act(fmaf(v.w, k.w, o.w), p.act, p.slope);
    if (has_res) acc4(t, (rg + (size_t)r * p.q4).ld());
    (yg + (size_t)r * p.q4).st(t);
  };
  const int step = p.RPP;
  int r = i.px0 + i.row;
  for (; r + 3 * step < i.px1; r += 4 * step) {
    float4 v0 = (xg + (size_t)r * p.q4).ld(), v1 = (xg + (size_t)(r + step) * p.q4).ld();
    float4 v2 = (xg + (size_t)(r + 2 * step) * p.q4).ld(), v3 = (xg + (size_t)(r + 3 * step) * p.q4).ld();
    one(v0, r); one(v1, r + step); one(v2, r + 2 * step); one(v3, r + 3 * step);
  }
  for (; r < i.px1; r += step) one((xg + (size_t)r * p.q4).ld(), r);
}

struct NormBwdConsts { float4 mu, rs, g, b, tb; };

__device__ __forceinline__ NormBwdConsts norm_bwd_consts(const NormP& p, int n, int c, const float* mean,
                                                         const float* rstd, const float* gamma, const float* beta,
                                                         const float* cbias) {
  NormBwdConsts k;
  k.mu = __ldg(reinterpret_cast<const float4*>(mean + (size_t)n * p.C + c));
  k.rs = __ldg(reinterpret_cast<const float4*>(rstd + (size_t)n * p.C + c));
  k.g = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : f4(1.f);
  k.b = beta ? __ldg(reinterpret_cast<const float4*>(beta + c)) : f4(0.f);
  k.tb = cbias ? __ldg(reinterpret_cast<const float4*>(cbias + (size_t)n * p.C + c)) : f4(0.f);
  return k;
}

__device__ __forceinline__ void norm_dv_xh(const NormP& p, const NormBwdConsts& k, float4 xv, float4 dy, float4& xh,
                                           float4& dv) {
  xh.x = (xv.x - k.mu.x) * k.rs.x; xh.y = (xv.y - k.mu.y) * k.rs.y;
  xh.z = (xv.z - k.mu.z) * k.rs.z; xh.w = (xv.w - k.mu.w) * k.rs.w;
  dv.x = dy.x * act_grad_pre((xh.x + k.tb.x) * k.g.x + k.b.x, p.act, p.slope);
  dv.y = dy.y * act_grad_pre((xh.y + k.tb.y) * k.g.y + k.b.y, p.act, p.slope);
  dv.z = dy.z * act_grad_pre((xh.z + k.tb.z) * k.g.z + k.b.z, p.act, p.slope);
  dv.w = dy.w * act_grad_pre((xh.w + k.tb.w) * k.g.w + k.b.w, p.act, p.slope);
}

// T: storage type of x and dx, TY: storage type of dy (= the forward output's)
template <typename V, typename T = float, typename TY = T>
__global__ void __launch_bounds__(kNormThreads, 4) inorm_bwd_reduce_kernel(
    NormP p, const TY* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ cbias, V* __restrict__ part) {
  __shared__ V sa[kNormThreads], sb[kNormThreads];
  const NormIdx i = norm_idx(p);
  const int n = blockIdx.y;
  V a1 = vzero<V>(), a2 = vzero<V>();
  if (i.active) {
    const size_t plane = (size_t)n * p.HW * p.q4 + i.c4;
    const In4<T> xg = In4<T>{x} + plane;
    const In4<TY> dg = In4<TY>{dy} + plane;
    const NormBwdConsts k = norm_bwd_consts(p, n, i.c4 * 4, mean, rstd, gamma, beta, cbias);
    auto one = [&](float4 xv, float4 dv_in, auto& t1, auto& t2) {
      float4 xh, dv;
      norm_dv_xh(p, k, xv, dv_in, xh, dv);
      t1.x += dv.x; t1.y += dv.y; t1.z += dv.z; t1.w += dv.w;
      t2.x += dv.x * xh.x; t2.y += dv.y * xh.y; t2.z += dv.z * xh.z; t2.w += dv.w * xh.w;
    };
    const int step = p.RPP;
    int r = i.px0 + i.row;
    if constexpr (sizeof(V) == sizeof(float4)) {
      for (; r + step < i.px1; r += 2 * step) {
        float4 x0 = (xg + (size_t)r * p.q4).ld(), d0 = (dg + (size_t)r * p.q4).ld();
        float4 x1 = (xg + (size_t)(r + step) * p.q4).ld(), d1 = (dg + (size_t)(r + step) * p.q4).ld();
        one(x0, d0, a1, a2); one(x1, d1, a1, a2);
      }
      for (; r < i.px1; r += step) one((xg + (size_t)r * p.q4).ld(), (dg + (size_t)r * p.q4).ld(), a1, a2);
    } else {
      for (; r + step < i.px1; r += 2 * step) {          // one atom (2 rows of this thread) per iteration
        float4 x0 = (xg + (size_t)r * p.q4).ld(), d0 = (dg + (size_t)r * p.q4).ld();
        float4 x1 = (xg + (size_t)(r + step) * p.q4).ld(), d1 = (dg + (size_t)(r + step) * p.q4).ld();
        float4 b1 = f4(0.f), b2 = f4(0.f);
        one(x0, d0, b1, b2); one(x1, d1, b1, b2);
        acc4(a1, b1); acc4(a2, b2);
      }
      if (r < i.px1) {
        float4 b1 = f4(0.f), b2 = f4(0.f);
        one((xg + (size_t)r * p.q4).ld(), (dg + (size_t)r * p.q4).ld(), b1, b2);
        acc4(a1, b1); acc4(a2, b2);
      }
    }
  }
  rows_reduce(p, i, a1, a2, sa, sb);
  if (i.row == 0) {
    const size_t o = ((size_t)n * p.SL + blockIdx.x) * p.q4 + i.c4;
    part[o] = a1;
    part[(size_t)p.N * p.SL * p.q4 + o] = a2;
  }
}

template <typename V, typename T = float, typename TY = T>
__global__ void __launch_bounds__(kNormThreads, 4) inorm_bwd_apply_kernel(
    NormP p, const TY* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ cbias, const V* __restrict__ part, T* __restrict__ dx,
    float* s1_out, float* s2_out) {
  const NormIdx i = norm_idx(p);
  if (!i.active) return;
  const int n = blockIdx.y;
  const int c = i.c4 * 4;
  const NormBwdConsts k = norm_bwd_consts(p, n, c, mean, rstd, gamma, beta, cbias);
  const float4 kk = make_float4(k.rs.x * k.g.x, k.rs.y * k.g.y, k.rs.z * k.g.z, k.rs.w * k.g.w);
  float4 m1, m2;
  if (p.given) {                       // dx = rstd*gamma * (dv - m1 - xh * m2) with caller-provided m1, m2
    m1 = __ldg(reinterpret_cast<const float4*>(s1_out + (size_t)n * p.C + c));
    m2 = __ldg(reinterpret_cast<const float4*>(s2_out + (size_t)n * p.C + c));
  } else {
    float4 S1, S2;
    fold_parts(p, part, n, i.c4, S1, S2);
    if (blockIdx.x == 0 && i.row == 0) {
      reinterpret_cast<float4*>(s1_out + (size_t)n * p.C + c)[0] = S1;
      reinterpret_cast<float4*>(s2_out + (size_t)n * p.C + c)[0] = S2;
    }
    const float inv = 1.f / (float)p.HW;
    m1 = make_float4(S1.x * inv, S1.y * inv, S1.z * inv, S1.w * inv);
    m2 = make_float4(S2.x * inv, S2.y * inv, S2.z * inv, S2.w * inv);
  }
  const size_t plane = (size_t)n * p.HW * p.q4 + i.c4;
  const In4<T> xg = In4<T>{x} + plane;
  const In4<TY> dg = In4<TY>{dy} + plane;
  const Out4<T> og = Out4<T>{dx} + plane;
  auto one = [&](float4 xv, float4 dv_in, int r) {
    float4 xh, dv, o;
    norm_dv_xh(p, k, xv, dv_in, xh, dv);
    o.x = kk.x * (dv.x - m1.x - xh.x * m2.x);
    o.y = kk.y * (dv.y - m1.y - xh.y * m2.y);
    o.z = kk.z * (dv.z - m1.z - xh.z * m2.z);
    o.w = kk.w * (dv.w - m1.w - xh.w * m2.w);
    (og + (size_t)r * p.q4).st(o);
  };
  const int step = p.RPP;
  int r = i.px0 + i.row;
  for (; r + step < i.px1; r += 2 * step) {
    float4 x0 = (xg + (size_t)r * p.q4).ld(), d0 = (dg + (size_t)r * p.q4).ld();
    float4 x1 = (xg + (size_t)(r + step) * p.q4).ld(), d1 = (dg + (size_t)(r + step) * p.q4).ld();
    one(x0, d0, r); one(x1, d1, r + step);
  }
  for (; r < i.px1; r += step) one((xg + (size_t)r * p.q4).ld(), (dg + (size_t)r * p.q4).ld(), r);
}

// ------------------------------------------------------------------------------------------------------------
// Fused two-phase kernels: ONE launch whose HBM traffic is the algorithmic minimum (forward: read x, write y;
// backward: read dy and x, write dx).  Persistent CTAs pull work from two queues:
//   phase-1 items (statistics of one pixel slice of one image x channel chunk): stream the slice, publish the partial
//     sums, bump the group's arrival counter; the CTA whose arrival completes the group folds the SL partials in
//     slice order (bit-identical whoever folds), writes the group's statistics and raises its flag;
//   phase-2 items (normalise one slice): taken - with priority - as soon as the flag of the group at the head of the
//     queue is up; they re-read the slice while it is still in L2 (a group is consumed microseconds after it was
//     produced: the working set is ~resident CTAs x slice, far below 126 MB) and write the result.
// No CTA keeps a slice on chip and no CTA blocks while work is available (a phase-2 ticket is only taken with a
// compare-and-swap once its group is ready), so registers stay low, 5-6 CTAs per SM hide the latency of the atomics
// and fences, and reads and writes overlap all the time.  When phase-1 items run out, idle CTAs wait for the groups
// still being produced by running CTAs (bounded spin, trap on timeout).
struct FusedIdx { int n, z, cg, row, c4, px0, px1, g, s; bool active; };

__device__ __forceinline__ FusedIdx fused_idx(const NormP& p, int item) {
  FusedIdx i;
  const int nchunk = p.q4 / p.TPR;
  i.g = item / p.SL; i.s = item - i.g * p.SL;
  i.n = i.g / nchunk; i.z = i.g - i.n * nchunk;
  i.cg = threadIdx.x % p.TPR;
  i.row = threadIdx.x / p.TPR;
  i.active = i.row < p.RPP;
  i.c4 = i.z * p.TPR + i.cg;
  i.px0 = i.s * p.slice;
  i.px1 = min(p.HW, i.px0 + p.slice);
  return i;
}

// L2 eviction hints for the fused kernels: a slice read in phase 1 is read again a few microseconds later (keep it),
// after phase 2 it is dead (let it go first)
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_drop() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
#define ld_keep(p) ld_hint(p, pol_keep)
#define ld_drop(p) ld_hint(p, pol_drop)

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ctl[0] = phase-1 tickets, ctl[1] = phase-2 tickets.  Phase-2 tickets are handed out with a plain fetch-add (a
// compare-and-swap retry loop is quadratic in the number of CTAs that see a group become ready at the same moment)
// once the group at the head of the queue is up; a ticket that lands in a group which is NOT ready yet is stashed
// - the CTA carries on with phase-1 work and looks at the stash first every time.  A CTA blocks only when it holds
// no phase-1 item and phase 1 has run out, i.e. when every missing group is in the hands of running CTAs.
struct FusedSched { int* ctl; const int* flag; int total, SL; bool phase1_left; int stash; };

// returns kind (0 phase 1, 1 phase 2, 2 all done, 3 nothing available right now - only when !may_block) and the
// item.  Called by thread 0.
__device__ __forceinline__ int fused_next(FusedSched& q, int& item, bool may_block) {
  const long long t0 = clock64();
  for (;;) {
    if (q.stash >= 0) {
      if (ld_acquire_gpu(q.flag + q.stash / q.SL) != 0) { item = q.stash; q.stash = -1; return 1; }
    } else {
      const int h = ld_relaxed_gpu(q.ctl + 1);
      if (h >= q.total) {
        if (!q.phase1_left) return 2;
      } else if (ld_acquire_gpu(q.flag + h / q.SL) != 0) {
        const int t = atomicAdd(q.ctl + 1, 1);
        if (t < q.total) {
          if (t / q.SL == h / q.SL || ld_acquire_gpu(q.flag + t / q.SL) != 0) { item = t; return 1; }
          q.stash = t;
        }
      }
    }
    if (q.phase1_left) {
      const int t = atomicAdd(q.ctl, 1);
      if (t < q.total) { item = t; return 0; }
      q.phase1_left = false;
      continue;
    }
    if (!may_block) return 3;
    __nanosleep(100);
    if (clock64() - t0 > (1ll << 32)) __trap();
  }
}

// Work loop scaffolding shared by the forward and backward kernels: thread 0 picks the NEXT item while the CTA
// streams the current one (the L2 round trips of the scheduler stay off the critical path).  The look-ahead never
// blocks - the group it would wait for may need the item this CTA is holding.
__device__ __forceinline__ void fused_fetch(FusedSched& q, int* s_kind, int* s_item, int slot, bool may_block) {
  int item = 0;
  s_kind[slot] = fused_next(q, item, may_block);
  s_item[slot] = item;
}
#define SRGAN_FUSED_BEGIN()                                                          \
  __shared__ int s_kind[2], s_item[2];                                               \
  FusedSched sched = {ctl, flag, total, p.SL, true, -1};                                 \
  if (threadIdx.x == 0) fused_fetch(sched, s_kind, s_item, 0, true);                 \
  __syncthreads();                                                                   \
  for (int cur = 0;; cur ^= 1) {                                                     \
    int kind = s_kind[cur];                                                          \
    if (kind == 3) {                   /* nothing was available at look-ahead time */ \
      __syncthreads();                                                               \
      if (threadIdx.x == 0) fused_fetch(sched, s_kind, s_item, cur, true);           \
      __syncthreads();                                                               \
      kind = s_kind[cur];                                                            \
    }                                                                                \
    if (kind == 2) break;                                                            \
    const int item = s_item[cur];                                                    \
    if (threadIdx.x == 0) fused_fetch(sched, s_kind, s_item, cur ^ 1, false);
#define SRGAN_FUSED_END() \
    __syncthreads();      \
  }

// After this CTA's partials are stored: arrive on the group; true (every thread) for the CTA that completes it.
__device__ __forceinline__ bool fused_arrive(int* arrive, int g, int SL, int* s_flag) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int old = atomicAdd(arrive + g, 1);
    __threadfence();
    *s_flag = old == SL - 1;
  }
  __syncthreads();
  return *s_flag != 0;
}

__global__ void __launch_bounds__(kNormThreads, 4) inorm_fwd_fused_kernel(
    const NormP p, const float* __restrict__ x, int* __restrict__ ctl, float4* __restrict__ part,
    float* __restrict__ y, float* mean_out, float* rstd_out, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ cbias, const float* __restrict__ residual) {
  __shared__ float4 sa[kNormThreads], sb[kNormThreads];
  __shared__ int s_last;
  const uint64_t pol_keep = l2_policy_keep(), pol_drop = l2_policy_drop();
  const int nchunk = p.q4 / p.TPR;
  const int groups = p.N * nchunk, total = groups * p.SL;
  int* arrive = ctl + 2;
  int* flag = arrive + groups;
  const size_t plane2 = (size_t)p.N * p.SL * p.q4;
  SRGAN_FUSED_BEGIN()
    const FusedIdx i = fused_idx(p, item);
    const size_t plane = (size_t)i.n * p.HW * p.q4 + i.c4;
    const float4* xg = reinterpret_cast<const float4*>(x) + plane;
    const int step = p.RPP;
    if (kind == 0) {
      // ---------------------------------------------------------------- statistics of one slice
      float4 s1 = f4(0.f), s2 = f4(0.f);
      if (i.active) {
        const float4 pv = __ldg(xg);                     // pivot: first pixel of the plane
        int r = i.px0 + i.row;
        for (; r + 3 * step < i.px1; r += 4 * step) {
          float4 v0 = ld_keep(xg + (size_t)r * p.q4), v1 = ld_keep(xg + (size_t)(r + step) * p.q4);
          float4 v2 = ld_keep(xg + (size_t)(r + 2 * step) * p.q4), v3 = ld_keep(xg + (size_t)(r + 3 * step) * p.q4);
#define SRGAN_ACC(v) { float a = v.x - pv.x, b = v.y - pv.y, c = v.z - pv.z, d = v.w - pv.w; \
                       s1.x += a; s1.y += b; s1.z += c; s1.w += d; s2.x += a * a; s2.y += b * b; s2.z += c * c; s2.w += d * d; }
          SRGAN_ACC(v0) SRGAN_ACC(v1) SRGAN_ACC(v2) SRGAN_ACC(v3)
        }
        for (; r < i.px1; r += step) { float4 v0 = ld_keep(xg + (size_t)r * p.q4); SRGAN_ACC(v0) }
#undef SRGAN_ACC
      }
      sa[threadIdx.x] = s1; sb[threadIdx.x] = s2;
      __syncthreads();
      if (i.row == 0) {
        float4 ra = f4(0.f), rb = f4(0.f);
        for (int r = 0; r < p.RPP; ++r) { acc4(ra, sa[r * p.TPR + i.cg]); acc4(rb, sb[r * p.TPR + i.cg]); }
        const size_t o = ((size_t)i.n * p.SL + i.s) * p.q4 + i.c4;
        __stcg(part + o, ra);
        __stcg(part + plane2 + o, rb);
      }
      if (fused_arrive(arrive, i.g, p.SL, &s_last)) {
        // this CTA completed the group: fold the slices in order, publish mean / rstd, raise the flag
        if (i.row == 0) {
          float4 t1 = f4(0.f), t2 = f4(0.f);
          const float4* p1 = part + (size_t)i.n * p.SL * p.q4 + i.c4;
          for (int s = 0; s < p.SL; ++s) { acc4(t1, __ldcg(p1 + (size_t)s * p.q4)); acc4(t2, __ldcg(p1 + plane2 + (size_t)s * p.q4)); }
          const float4 pv = __ldg(xg);
          const float inv = 1.f / (float)p.HW;
          float4 mu, rs;
#define SRGAN_STAT(f) { float m = t1.f * inv; float var = fmaxf(t2.f * inv - m * m, 0.f); mu.f = pv.f + m; rs.f = rsqrtf(var + p.eps); }
          SRGAN_STAT(x) SRGAN_STAT(y) SRGAN_STAT(z) SRGAN_STAT(w)
#undef SRGAN_STAT
          __stcg(reinterpret_cast<float4*>(mean_out + (size_t)i.n * p.C) + i.c4, mu);
          __stcg(reinterpret_cast<float4*>(rstd_out + (size_t)i.n * p.C) + i.c4, rs);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          __threadfence();
          st_release_gpu(flag + i.g, 1);
        }
      }
    } else if (i.active) {
      // ---------------------------------------------------------------- normalise one slice (x from L2)
      const int c = i.c4 * 4;
      const float4 mu = __ldcg(reinterpret_cast<const float4*>(mean_out + (size_t)i.n * p.C) + i.c4);
      const float4 rs = __ldcg(reinterpret_cast<const float4*>(rstd_out + (size_t)i.n * p.C) + i.c4);
      const float4 g = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : f4(1.f);
      const float4 b = beta ? __ldg(reinterpret_cast<const float4*>(beta + c)) : f4(0.f);
      const float4 tb = cbias ? __ldg(reinterpret_cast<const float4*>(cbias + (size_t)i.n * p.C + c)) : f4(0.f);
      const float4 k = make_float4(rs.x * g.x, rs.y * g.y, rs.z * g.z, rs.w * g.w);
      const float4 o = make_float4((tb.x - mu.x * rs.x) * g.x + b.x, (tb.y - mu.y * rs.y) * g.y + b.y,
                                   (tb.z - mu.z * rs.z) * g.z + b.z, (tb.w - mu.w * rs.w) * g.w + b.w);
      float4* yg = reinterpret_cast<float4*>(y) + plane;
      const float4* rg = residual ? reinterpret_cast<const float4*>(residual) + plane : nullptr;
      auto one = [&](float4 v, int r) {
        float4 t;
        t.x = apply_act(fmaf(v.x, k.x, o.x), p.act, p.slope);
        t.y = apply_act(fmaf(v.y, k.y, o.y), p.act, p.slope);
        t.z = apply_act(fmaf(v.z, k.z, o.z), p.act, p.slope);
        t.w = apply_act(fmaf(v.w, k.w, o.w), p.act, p.slope);
        if (rg) acc4(t, __ldg(rg + (size_t)r * p.q4));
        yg[(size_t)r * p.q4] = t;
      };
      int r = i.px0 + i.row;
      for (; r + 3 * step < i.px1; r += 4 * step) {
        float4 v0 = ld_drop(xg + (size_t)r * p.q4), v1 = ld_drop(xg + (size_t)(r + step) * p.q4);
        float4 v2 = ld_drop(xg + (size_t)(r + 2 * step) * p.q4), v3 = ld_drop(xg + (size_t)(r + 3 * step) * p.q4);
        one(v0, r); one(v1, r + step); one(v2, r + 2 * step); one(v3, r + 3 * step);
      }
      for (; r < i.px1; r += step) one(ld_drop(xg + (size_t)r * p.q4), r);
    }
  SRGAN_FUSED_END()
}

__global__ void __launch_bounds__(kNormThreads, 4) inorm_bwd_fused_kernel(
    const NormP p, const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ cbias, int* __restrict__ ctl, float4* __restrict__ part, float* __restrict__ dx,
    float* s1_out, float* s2_out) {
  __shared__ float4 sa[kNormThreads], sb[kNormThreads];
  __shared__ int s_last;
  const uint64_t pol_keep = l2_policy_keep(), pol_drop = l2_policy_drop();
  const int nchunk = p.q4 / p.TPR;
  const int groups = p.N * nchunk, total = groups * p.SL;
  int* arrive = ctl + 2;
  int* flag = arrive + groups;
  const size_t plane2 = (size_t)p.N * p.SL * p.q4;
  SRGAN_FUSED_BEGIN()
    const FusedIdx i = fused_idx(p, item);
    const int c = i.c4 * 4;
    const size_t plane = (size_t)i.n * p.HW * p.q4 + i.c4;
    const float4* xg = reinterpret_cast<const float4*>(x) + plane;
    const float4* dg = reinterpret_cast<const float4*>(dy) + plane;
    const int step = p.RPP;
    if (kind == 0) {
      float4 a1 = f4(0.f), a2 = f4(0.f);
      if (i.active) {
        const NormBwdConsts k = norm_bwd_consts(p, i.n, c, mean, rstd, gamma, beta, cbias);
        auto one = [&](float4 xv, float4 dv_in) {
          float4 xh, dv;
          norm_dv_xh(p, k, xv, dv_in, xh, dv);
          acc4(a1, dv);
          a2.x += dv.x * xh.x; a2.y += dv.y * xh.y; a2.z += dv.z * xh.z; a2.w += dv.w * xh.w;
        };
        int r = i.px0 + i.row;
        for (; r + step < i.px1; r += 2 * step) {
          float4 x0 = ld_keep(xg + (size_t)r * p.q4), d0 = ld_keep(dg + (size_t)r * p.q4);
          float4 x1 = ld_keep(xg + (size_t)(r + step) * p.q4), d1 = ld_keep(dg + (size_t)(r + step) * p.q4);
          one(x0, d0); one(x1, d1);
        }
        for (; r < i.px1; r += step) one(ld_keep(xg + (size_t)r * p.q4), ld_keep(dg + (size_t)r * p.q4));
      }
      sa[threadIdx.x] = a1; sb[threadIdx.x] = a2;
      __syncthreads();
      if (i.row == 0) {
        float4 ra = f4(0.f), rb = f4(0.f);
        for (int r = 0; r < p.RPP; ++r) { acc4(ra, sa[r * p.TPR + i.cg]); acc4(rb, sb[r * p.TPR + i.cg]); }
        const size_t o = ((size_t)i.n * p.SL + i.s) * p.q4 + i.c4;
        __stcg(part + o, ra);
        __stcg(part + plane2 + o, rb);
      }
      if (fused_arrive(arrive, i.g, p.SL, &s_last)) {
        if (i.row == 0) {
          float4 t1 = f4(0.f), t2 = f4(0.f);
          const float4* p1 = part + (size_t)i.n * p.SL * p.q4 + i.c4;
          for (int s = 0; s < p.SL; ++s) { acc4(t1, __ldcg(p1 + (size_t)s * p.q4)); acc4(t2, __ldcg(p1 + plane2 + (size_t)s * p.q4)); }
          __stcg(reinterpret_cast<float4*>(s1_out + (size_t)i.n * p.C) + i.c4, t1);
          __stcg(reinterpret_cast<float4*>(s2_out + (size_t)i.n * p.C) + i.c4, t2);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          __threadfence();
          st_release_gpu(flag + i.g, 1);
        }
      }
    } else if (i.active) {
      const NormBwdConsts k = norm_bwd_consts(p, i.n, c, mean, rstd, gamma, beta, cbias);
      const float4 kk = make_float4(k.rs.x * k.g.x, k.rs.y * k.g.y, k.rs.z * k.g.z, k.rs.w * k.g.w);
      const float4 S1 = __ldcg(reinterpret_cast<const float4*>(s1_out + (size_t)i.n * p.C) + i.c4);
      const float4 S2 = __ldcg(reinterpret_cast<const float4*>(s2_out + (size_t)i.n * p.C) + i.c4);
      const float inv = 1.f / (float)p.HW;
      const float4 m1 = make_float4(S1.x * inv, S1.y * inv, S1.z * inv, S1.w * inv);
      const float4 m2 = make_float4(S2.x * inv, S2.y * inv, S2.z * inv, S2.w * inv);
      float4* og = reinterpret_cast<float4*>(dx) + plane;
      auto one = [&](float4 xv, float4 dv_in, int r) {
        float4 xh, dv, o;
        norm_dv_xh(p, k, xv, dv_in, xh, dv);
        o.x = kk.x * (dv.x - m1.x - xh.x * m2.x);
        o.y = kk.y * (dv.y - m1.y - xh.y * m2.y);
        o.z = kk.z * (dv.z - m1.z - xh.z * m2.z);
        o.w = kk.w * (dv.w - m1.w - xh.w * m2.w);
        og[(size_t)r * p.q4] = o;
      };
      int r = i.px0 + i.row;
      for (; r + step < i.px1; r += 2 * step) {
        float4 x0 = ld_drop(xg + (size_t)r * p.q4), d0 = ld_drop(dg + (size_t)r * p.q4);
        float4 x1 = ld_drop(xg + (size_t)(r + step) * p.q4), d1 = ld_drop(dg + (size_t)(r + step) * p.q4);
        one(x0, d0, r); one(x1, d1, r + step);
      }
      for (; r < i.px1; r += step) one(ld_drop(xg + (size_t)r * p.q4), ld_drop(dg + (size_t)r * p.q4), r);
    }
  SRGAN_FUSED_END()
}

// dgamma[c] = sum_n (s2 + cbias*s1) ; dbeta[c] = sum_n s1 ; dcbias[n][c] = gamma[c]*s1[n][c]
// one warp per channel: lanes stride over the images, fixed-order butterfly reduction
__global__ void inorm_param_grads_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                         const float* __restrict__ gamma, const float* __restrict__ cbias,
                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                         float* __restrict__ dcbias, int N, int C) {
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  const float g = gamma ? gamma[c] : 1.f;
  float dg = 0.f, db = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float a = s1[(size_t)n * C + c], b = s2[(size_t)n * C + c];
    const float t = cbias ? cbias[(size_t)n * C + c] : 0.f;
    dg += b + t * a;
    db += a;
    if (dcbias) dcbias[(size_t)n * C + c] = g * a;
  }
  dg = warp_sum(dg);
  db = warp_sum(db);
  if (lane == 0) {
    if (dgamma) dgamma[c] = dg;
    if (dbeta) dbeta[c] = db;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Batch-statistics norms: nn.BatchNorm2d (norm_type="batch") and CBBNorm2d (ref pyfiles/model.py:75-171):
//   CBBNorm2d:   out = BN_batch(x);  y = (out - mean_hw(out) + tanh(Linear(con))) * w + b
//              = ((x - m_nc) * r_c + cbias_nc) * w_c + b_c       (m_nc instance mean, r_c batch rstd)
//   BatchNorm2d: y = (x - mu_c) * r_c * w_c + b_c
// Both are the instance-norm apply kernel with per-(n,c) statistics GIVEN: (m_nc, r_c) resp. (mu_c, r_c).  The
// small per-(image, channel) tables are produced / consumed by the kernels below, in stages, so that data-parallel
// ranks can all-gather them between stages and evaluate the batch statistics in the single-GPU order.

// slice partials -> per-(n,c) tables.  pivot != nullptr: a = instance mean, b = sum of squared deviations (partials
// are sums about the first pixel of the plane); else a, b = plain sums of the two partial planes.
template <typename V>
__global__ void norm_fold_kernel(NormP p, const V* __restrict__ part, const float* __restrict__ pivot_x,
                                 float* __restrict__ a_out, float* __restrict__ b_out) {
  const int c4 = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
  if (c4 >= p.q4) return;
  float4 s1, s2;
  fold_parts(p, part, n, c4, s1, s2);
  if (pivot_x) {
    const float4 pv = __ldg(reinterpret_cast<const float4*>(pivot_x) + (size_t)n * p.HW * p.q4 + c4);
    const float inv = 1.f / (float)p.HW;
#define SRGAN_CONV(f) { const float m = s1.f * inv; s2.f = fmaxf(s2.f - s1.f * m, 0.f); s1.f = pv.f + m; }
    SRGAN_CONV(x) SRGAN_CONV(y) SRGAN_CONV(z) SRGAN_CONV(w)
#undef SRGAN_CONV
  }
  reinterpret_cast<float4*>(a_out + (size_t)n * p.C)[c4] = s1;
  reinterpret_cast<float4*>(b_out + (size_t)n * p.C)[c4] = s2;
}

// One warp per channel over the (all-gathered) tables of N_all images; rows [n0, n0 + N_loc) are this rank's.
// training: batch mean / biased variance (Chan's combination of the per-image moments, fixed butterfly order),
// running statistics updated with the unbiased variance; else the running statistics are used.
__global__ void bnorm_batch_stats_kernel(const float* __restrict__ mean_nc, const float* __restrict__ m2_nc,
                                         int N_all, int n0, int N_loc, int HW, int C, float eps, int cond,
                                         int training, float* __restrict__ running_mean,
                                         float* __restrict__ running_var, float momentum,
                                         float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                         float* __restrict__ batch_mean) {
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  float mu, var;
  if (training) {
    float sm = 0.f;
    for (int n = lane; n < N_all; n += 32) sm += mean_nc[(size_t)n * C + c];
    mu = warp_sum(sm) / (float)N_all;
    float sq = 0.f;
    for (int n = lane; n < N_all; n += 32) {
      const float d = mean_nc[(size_t)n * C + c] - mu;
      sq += m2_nc[(size_t)n * C + c] + (float)HW * d * d;
    }
    const float M2 = warp_sum(sq);
    const float cnt = (float)N_all * (float)HW;
    var = M2 / cnt;
    if (lane == 0 && running_mean && running_var) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (M2 / fmaxf(cnt - 1.f, 1.f));
    }
  } else {
    mu = running_mean[c];
    var = running_var[c];
  }
  const float rs = rsqrtf(var + eps);
  if (lane == 0) batch_mean[c] = mu;
  for (int n = lane; n < N_loc; n += 32) {
    mean_out[(size_t)n * C + c] = cond ? mean_nc[(size_t)(n0 + n) * C + c] : mu;
    rstd_out[(size_t)n * C + c] = rs;
  }
}

// dx = rstd*gamma * (dv - m1 - xh * m2): per-(n,c) coefficients from the (all-gathered) sums s1 = sum dv,
// s2 = sum dv*xh.   CBBNorm2d: m1 = s1_nc/HW + delta_nc * M, m2 = M, M = sum_n s2 / (N*HW),
// delta_nc = (m_nc - mu_c) * r_c;  BatchNorm2d: m1 = sum_n s1 / (N*HW), m2 = M.  Evaluation mode (constant r_c):
// CBBNorm2d m1 = s1_nc/HW, m2 = 0; BatchNorm2d m1 = m2 = 0.
__global__ void bnorm_bwd_coeffs_kernel(const float* __restrict__ s1_all, const float* __restrict__ s2_all,
                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                        const float* __restrict__ batch_mean, int N_all, int n0, int N_loc, int HW,
                                        int C, int cond, int training, float* __restrict__ m1_out,
                                        float* __restrict__ m2_out) {
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  float S1 = 0.f, S2 = 0.f;
  if (training) {
    for (int n = lane; n < N_all; n += 32) { S1 += s1_all[(size_t)n * C + c]; S2 += s2_all[(size_t)n * C + c]; }
    S1 = warp_sum(S1);
    S2 = warp_sum(S2);
  }
  const float inv_all = 1.f / ((float)N_all * (float)HW), inv = 1.f / (float)HW;
  const float M = training ? S2 * inv_all : 0.f;
  const float mu = batch_mean[c];
  for (int n = lane; n < N_loc; n += 32) {
    const size_t o = (size_t)n * C + c;
    float m1;
    if (cond) {
      const float delta = (mean[o] - mu) * rstd[o];
      m1 = s1_all[(size_t)(n0 + n) * C + c] * inv + delta * M;
    } else {
      m1 = training ? S1 * inv_all : 0.f;
    }
    m1_out[o] = m1;
    m2_out[o] = M;
  }
}

// fp64 partial sums over fixed atoms (see D4 above).  SRGAN_DBG_NORM_F32_PARTIALS=1 (bring-up, A/B timing) restores
// fp32 partials, whose value depends on the slice count.
static bool norm_f64() {
  static const bool on = !(getenv("SRGAN_DBG_NORM_F32_PARTIALS") && atoi(getenv("SRGAN_DBG_NORM_F32_PARTIALS")) != 0);
  return on;
}

// channel chunking + pixel slicing; returns false when C cannot be mapped onto 256 threads
static bool plan_norm(int N, int HW, int C, NormP* out) {
  NormP p = {};
  p.N = N; p.HW = HW; p.C = C; p.q4 = C / 4;
  int nchunk = 1;
  while (nchunk <= p.q4 && (p.q4 % nchunk || p.q4 / nchunk > 64)) ++nchunk;
  if (nchunk > p.q4) {                       // no divisor gives <= 64 float4 per chunk: take rows of up to 256
    if (p.q4 > kNormThreads) return false;
    nchunk = 1;
  }
  p.TPR = p.q4 / nchunk;
  p.RPP = kNormThreads / p.TPR;
  // One full wave: the streaming kernels keep 4 CTAs of 256 threads per SM (registers), and a grid of 2.05 waves -
  // what "8 CTAs per SM worth of slices" produced at batch 64 - leaves the SMs idle for a quarter of the launch
  // (ncu: sm__cycles_active 57 k of 77 k elapsed).  SRGAN_DBG_NORM_CTAS_PER_SM overrides the 4.
  static const int per_sm = getenv("SRGAN_DBG_NORM_CTAS_PER_SM") ? atoi(getenv("SRGAN_DBG_NORM_CTAS_PER_SM")) : 4;
  long long want = ((long long)per_sm * kNumSMs) / ((long long)N * nchunk);
  long long cap = ceil_div(HW, p.RPP * 4);   // at least 4 passes per CTA
  long long SL = want < cap ? want : cap;
  if (SL < 1) SL = 1;
  if (SL > kNormMaxSlices) SL = kNormMaxSlices;
  p.slice = ceil_div(HW, (int)SL);
  if (norm_f64()) p.slice = ceil_div(p.slice, 4 * p.RPP) * (4 * p.RPP);   // slices start on atom boundaries
  p.SL = ceil_div(HW, p.slice);
  *out = p;
  return true;
}

static size_t norm_ws_bytes(const NormP& p) {
  return (size_t)2 * p.N * p.SL * p.q4 * (norm_f64() ? sizeof(D4) : sizeof(float4));
}
// fused kernels: [2 ticket counters | arrivals | flags] (zeroed before every launch) in front of the partials
static size_t fused_ctl_bytes(const NormP& p) {
  const size_t groups = (size_t)p.N * (p.q4 / p.TPR);
  return ((2 + 2 * groups) * sizeof(int) + 255) / 256 * 256;
}
static size_t fused_ws_bytes(const NormP& p) { return fused_ctl_bytes(p) + norm_ws_bytes(p); }
// The fused single-launch path is opt-in (SRGAN_NORM_FUSED=1): measured at batch 64 it is 5-10 % faster than the
// two-kernel path on the large planes and slower on the small ones, and the training step as a whole does not move
// (100.4 vs 100.8 ms) - see DESIGN.md 2.4.
static bool fused_enabled() {
  static const bool on = getenv("SRGAN_NORM_FUSED") && atoi(getenv("SRGAN_NORM_FUSED")) != 0;
  return on;
}
static unsigned fused_grid(const void* kernel, const NormP& p) {
  static const void* fn[2] = {nullptr, nullptr};
  static int occ[2] = {0, 0};
  int k = fn[0] == kernel ? 0 : 1;
  if (fn[k] != kernel) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, kNormThreads, 0) != cudaSuccess || o < 1) o = 1;
    fn[k] = kernel; occ[k] = o;
  }
  const long long total = (long long)p.N * (p.q4 / p.TPR) * p.SL;
  const long long cap = (long long)occ[k] * kNumSMs;
  return (unsigned)(total < cap ? total : cap);
}
static dim3 norm_grid(const NormP& p) { return dim3(p.SL, p.N, p.q4 / p.TPR); }

}  // namespace srgan

using namespace srgan;

extern "C" int srgan_norm_partials_fp64(void) { return norm_f64() ? 1 : 0; }

extern "C" size_t srgan_inorm_workspace(int N, int HW, int C) {
  NormP p;
  if (N <= 0 || HW <= 0 || C <= 0 || C % 4 || !plan_norm(N, HW, C, &p)) return 0;
  return fused_ws_bytes(p);
}

extern "C" int srgan_inorm_fwd(const float* x, float* y, float* mean, float* rstd, const float* gamma,
                               const float* beta, const float* cbias, const float* residual, int N, int HW,
                               int C, float eps, int act, float slope, void* ws, size_t ws_bytes, void* stream) {
  SRGAN_CHECK_ARG(x && y && mean && rstd, "null pointer");
  SRGAN_CHECK_ARG(N >= 0 && HW > 0 && C > 0 && C % 8 == 0, "need C % 8 == 0, HW > 0");
  SRGAN_CHECK_ARG(((uintptr_t)x | (uintptr_t)y | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma |
                   (uintptr_t)beta | (uintptr_t)cbias | (uintptr_t)residual | (uintptr_t)ws) % 16 == 0,
                  "pointers must be 16-byte aligned");
  SRGAN_CHECK_ARG(N <= 65535, "N too large for grid.y");
  if (N == 0) return SRGAN_OK;
  NormP p;
  SRGAN_CHECK_ARG(plan_norm(N, HW, C, &p), "channel count cannot be mapped");
  p.eps = eps; p.slope = slope; p.act = act;
  cudaStream_t st = (cudaStream_t)stream;
  if (fused_enabled()) {
    if (!ws || ws_bytes < fused_ws_bytes(p)) { set_error("inorm_fwd: workspace %zu < %zu", ws_bytes, fused_ws_bytes(p)); return SRGAN_E_WORKSPACE; }
    cudaError_t me = cudaMemsetAsync(ws, 0, fused_ctl_bytes(p), st);
    if (me != cudaSuccess) { set_error("inorm_fwd: %s", cudaGetErrorString(me)); return (int)me; }
    inorm_fwd_fused_kernel<<<fused_grid((const void*)inorm_fwd_fused_kernel, p), kNormThreads, 0, st>>>(
        p, x, (int*)ws, (float4*)((uint8_t*)ws + fused_ctl_bytes(p)), y, mean, rstd, gamma, beta, cbias, residual);
    SRGAN_RETURN_LAUNCH();
  }
  if (!ws || ws_bytes < norm_ws_bytes(p)) { set_error("inorm_fwd: workspace %zu < %zu", ws_bytes, norm_ws_bytes(p)); return SRGAN_E_WORKSPACE; }
  if (norm_f64()) {
    inorm_stats_kernel<D4><<<norm_grid(p), kNormThreads, 0, st>>>(p, x, (D4*)ws);
    inorm_apply_kernel<D4><<<norm_grid(p), kNormThreads, 0, st>>>(p, x, (const D4*)ws, y, mean, rstd, gamma, beta,
                                                                  cbias, residual);
  } else {
    inorm_stats_kernel<float4><<<norm_grid(p), kNormThreads, 0, st>>>(p, x, (float4*)ws);
    inorm_apply_kernel<float4><<<norm_grid(p), kNormThreads, 0, st>>>(p, x, (const float4*)ws, y, mean, rstd, gamma,
                                                                      beta, cbias, residual);
  }
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_inorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                               const float* gamma, const float* beta, const float* cbias, float* dx, float* s1,
                               float* s2, int N, int HW, int C, int act, float slope, void* ws, size_t ws_bytes,
                               void* stream) {
  SRGAN_CHECK_ARG(dy && x && mean && rstd && dx && s1 && s2, "null pointer");
  SRGAN_CHECK_ARG(N >= 0 && HW > 0 && C > 0 && C % 8 == 0, "need C % 8 == 0, HW > 0");
  SRGAN_CHECK_ARG(((uintptr_t)dy | (uintptr_t)x | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma |
                   (uintptr_t)beta | (uintptr_t)cbias | (uintptr_t)dx | (uintptr_t)s1 | (uintptr_t)s2 |
                   (uintptr_t)ws) % 16 == 0, "pointers must be 16-byte aligned");
  SRGAN_CHECK_ARG(N <= 65535, "N too large for grid.y");
  if (N == 0) return SRGAN_OK;
  NormP p;
  SRGAN_CHECK_ARG(plan_norm(N, HW, C, &p), "channel count cannot be mapped");
  p.eps = 0.f; p.slope = slope; p.act = act;
  cudaStream_t st = (cudaStream_t)stream;
  if (fused_enabled()) {
    if (!ws || ws_bytes < fused_ws_bytes(p)) { set_error("inorm_bwd: workspace %zu < %zu", ws_bytes, fused_ws_bytes(p)); return SRGAN_E_WORKSPACE; }
    cudaError_t me = cudaMemsetAsync(ws, 0, fused_ctl_bytes(p), st);
    if (me != cudaSuccess) { set_error("inorm_bwd: %s", cudaGetErrorString(me)); return (int)me; }
    inorm_bwd_fused_kernel<<<fused_grid((const void*)inorm_bwd_fused_kernel, p), kNormThreads, 0, st>>>(
        p, dy, x, mean, rstd, gamma, beta, cbias, (int*)ws, (float4*)((uint8_t*)ws + fused_ctl_bytes(p)), dx, s1, s2);
    SRGAN_RETURN_LAUNCH();
  }
  if (!ws || ws_bytes < norm_ws_bytes(p)) { set_error("inorm_bwd: workspace %zu < %zu", ws_bytes, norm_ws_bytes(p)); return SRGAN_E_WORKSPACE; }
  if (norm_f64()) {
    inorm_bwd_reduce_kernel<D4><<<norm_grid(p), kNormThreads, 0, st>>>(p, dy, x, mean, rstd, gamma, beta, cbias, (D4*)ws);
    inorm_bwd_apply_kernel<D4><<<norm_grid(p), kNormThreads, 0, st>>>(p, dy, x, mean, rstd, gamma, beta, cbias,
                                                                      (const D4*)ws, dx, s1, s2);
  } else {
    inorm_bwd_reduce_kernel<float4><<<norm_grid(p), kNormThreads, 0, st>>>(p, dy, x, mean, rstd, gamma, beta, cbias,
                                                                           (float4*)ws);
    inorm_bwd_apply_kernel<float4><<<norm_grid(p), kNormThreads, 0, st>>>(p, dy, x, mean, rstd, gamma, beta, cbias,
                                                                          (const float4*)ws, dx, s1, s2);
  }
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_inorm_param_grads(const float* s1, const float* s2, const float* gamma,
                                       const float* cbias, float* dgamma, float* dbeta, float* dcbias, int N,
                                       int C, void* stream) {
  SRGAN_CHECK_ARG(s1 && s2, "null pointer");
  if (C == 0) return SRGAN_OK;
  inorm_param_grads_kernel<<<ceil_div(C * 32, 256), 256, 0, (cudaStream_t)stream>>>(s1, s2, gamma, cbias, dgamma,
                                                                                  dbeta, dcbias, N, C);
  SRGAN_RETURN_LAUNCH();
}

// ---------------------------------------------------------------------------------- batch-statistics norms (ABI)
static int bn_common_checks(const char* fn, int N, int HW, int C, NormP* p) {
  if (!(N >= 0 && HW > 0 && C > 0 && C % 8 == 0)) { set_error("%s: need C %% 8 == 0, HW > 0", fn); return SRGAN_E_BADARG; }
  if (N > 65535) { set_error("%s: N too large for grid.y", fn); return SRGAN_E_BADARG; }
  if (!plan_norm(N, HW, C, p)) { set_error("%s: channel count cannot be mapped", fn); return SRGAN_E_BADARG; }
  return SRGAN_OK;
}

extern "C" int srgan_bnorm_image_stats(const float* x, float* mean_nc, float* m2_nc, int N, int HW, int C, void* ws,
                                       size_t ws_bytes, void* stream) {
  SRGAN_CHECK_ARG(x && mean_nc && m2_nc, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)x | (uintptr_t)mean_nc | (uintptr_t)m2_nc | (uintptr_t)ws) % 16 == 0, "pointers must be 16-byte aligned");
  NormP p;
  if (int e = bn_common_checks(__func__, N, HW, C, &p)) return e;
  if (N == 0) return SRGAN_OK;
  if (!ws || ws_bytes < norm_ws_bytes(p)) { set_error("bnorm_image_stats: workspace %zu < %zu", ws_bytes, norm_ws_bytes(p)); return SRGAN_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (norm_f64()) {
    inorm_stats_kernel<D4><<<norm_grid(p), kNormThreads, 0, st>>>(p, x, (D4*)ws);
    norm_fold_kernel<D4><<<dim3(ceil_div(p.q4, 64), N), 64, 0, st>>>(p, (const D4*)ws, x, mean_nc, m2_nc);
  } else {
    inorm_stats_kernel<float4><<<norm_grid(p), kNormThreads, 0, st>>>(p, x, (float4*)ws);
    norm_fold_kernel<float4><<<dim3(ceil_div(p.q4, 64), N), 64, 0, st>>>(p, (const float4*)ws, x, mean_nc, m2_nc);
  }
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_bnorm_batch_stats(const float* mean_nc_all, const float* m2_nc_all, int N_all, int n0, int N_loc,
                                       int HW, int C, float eps, int cond, int training, float* running_mean,
                                       float* running_var, float momentum, float* mean, float* rstd,
                                       float* batch_mean, void* stream) {
  SRGAN_CHECK_ARG(mean_nc_all && m2_nc_all && mean && rstd && batch_mean, "null pointer");
  SRGAN_CHECK_ARG(N_all > 0 && n0 >= 0 && N_loc >= 0 && n0 + N_loc <= N_all && HW > 0 && C > 0, "bad extents");
  SRGAN_CHECK_ARG(training || (running_mean && running_var), "evaluation mode needs running statistics");
  bnorm_batch_stats_kernel<<<ceil_div(C * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      mean_nc_all, m2_nc_all, N_all, n0, N_loc, HW, C, eps, cond, training, running_mean, running_var, momentum, mean,
      rstd, batch_mean);
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_bnorm_apply(const float* x, float* y, const float* mean, const float* rstd, const float* gamma,
                                 const float* beta, const float* cbias, const float* residual, int N, int HW, int C,
                                 int act, float slope, void* stream) {
  SRGAN_CHECK_ARG(x && y && mean && rstd, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)x | (uintptr_t)y | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma |
                   (uintptr_t)beta | (uintptr_t)cbias | (uintptr_t)residual) % 16 == 0, "pointers must be 16-byte aligned");
  NormP p;
  if (int e = bn_common_checks(__func__, N, HW, C, &p)) return e;
  if (N == 0) return SRGAN_OK;
  p.eps = 0.f; p.slope = slope; p.act = act; p.given = 1;
  inorm_apply_kernel<float4><<<norm_grid(p), kNormThreads, 0, (cudaStream_t)stream>>>(
      p, x, nullptr, y, const_cast<float*>(mean), const_cast<float*>(rstd), gamma, beta, cbias, residual);
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_bnorm_bwd_sums(const float* dy, const float* x, const float* mean, const float* rstd,
                                    const float* gamma, const float* beta, const float* cbias, float* s1, float* s2,
                                    int N, int HW, int C, int act, float slope, void* ws, size_t ws_bytes,
                                    void* stream) {
  SRGAN_CHECK_ARG(dy && x && mean && rstd && s1 && s2, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)dy | (uintptr_t)x | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma |
                   (uintptr_t)beta | (uintptr_t)cbias | (uintptr_t)s1 | (uintptr_t)s2 | (uintptr_t)ws) % 16 == 0,
                  "pointers must be 16-byte aligned");
  NormP p;
  if (int e = bn_common_checks(__func__, N, HW, C, &p)) return e;
  if (N == 0) return SRGAN_OK;
  p.eps = 0.f; p.slope = slope; p.act = act;
  if (!ws || ws_bytes < norm_ws_bytes(p)) { set_error("bnorm_bwd_sums: workspace %zu < %zu", ws_bytes, norm_ws_bytes(p)); return SRGAN_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (norm_f64()) {
    inorm_bwd_reduce_kernel<D4><<<norm_grid(p), kNormThreads, 0, st>>>(p, dy, x, mean, rstd, gamma, beta, cbias, (D4*)ws);
    norm_fold_kernel<D4><<<dim3(ceil_div(p.q4, 64), N), 64, 0, st>>>(p, (const D4*)ws, nullptr, s1, s2);
  } else {
    inorm_bwd_reduce_kernel<float4><<<norm_grid(p), kNormThreads, 0, st>>>(p, dy, x, mean, rstd, gamma, beta, cbias,
                                                                           (float4*)ws);
    norm_fold_kernel<float4><<<dim3(ceil_div(p.q4, 64), N), 64, 0, st>>>(p, (const float4*)ws, nullptr, s1, s2);
  }
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_bnorm_bwd_coeffs(const float* s1_all, const float* s2_all, const float* mean, const float* rstd,
                                      const float* batch_mean, int N_all, int n0, int N_loc, int HW, int C, int cond,
                                      int training, float* m1, float* m2, void* stream) {
  SRGAN_CHECK_ARG(s1_all && s2_all && mean && rstd && batch_mean && m1 && m2, "null pointer");
  SRGAN_CHECK_ARG(N_all > 0 && n0 >= 0 && N_loc >= 0 && n0 + N_loc <= N_all && HW > 0 && C > 0, "bad extents");
  bnorm_bwd_coeffs_kernel<<<ceil_div(C * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      s1_all, s2_all, mean, rstd, batch_mean, N_all, n0, N_loc, HW, C, cond, training, m1, m2);
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_bnorm_bwd_apply(const float* dy, const float* x, const float* mean, const float* rstd,
                                     const float* gamma, const float* beta, const float* cbias, const float* m1,
                                     const float* m2, float* dx, int N, int HW, int C, int act, float slope,
                                     void* stream) {
  SRGAN_CHECK_ARG(dy && x && mean && rstd && m1 && m2 && dx, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)dy | (uintptr_t)x | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma |
                   (uintptr_t)beta | (uintptr_t)cbias | (uintptr_t)m1 | (uintptr_t)m2 | (uintptr_t)dx) % 16 == 0,
                  "pointers must be 16-byte aligned");
  NormP p;
  if (int e = bn_common_checks(__func__, N, HW, C, &p)) return e;
  if (N == 0) return SRGAN_OK;
  p.eps = 0.f; p.slope = slope; p.act = act; p.given = 1;
  inorm_bwd_apply_kernel<float4><<<norm_grid(p), kNormThreads, 0, (cudaStream_t)stream>>>(
      p, dy, x, mean, rstd, gamma, beta, cbias, nullptr, dx, const_cast<float*>(m1), const_cast<float*>(m2));
  SRGAN_RETURN_LAUNCH();
}
