// Conditional instance normalisation, second generation: the kernels behind srgan_inorm_{fwd,bwd}_mixed.
//
//   xh = (x - mean_hw) * rstd_hw ; v = (xh + cbias[n][c]) * gamma[c] + beta[c] ; y = act(v) (+ residual)
//   ref: CBINorm2d.forward pyfiles/model.py:54-67, nn.InstanceNorm2d(affine=False) :178, the ReLU / add that follow.
//
// What round 1's kernels (norm.cu) left on the table, measured with ncu on the shipped kernels (profiles/r2c_*): they
// move 8 (bf16) or 16 (fp32) bytes per thread and load, 4 loads in flight, and every apply CTA re-folds the slice
// partials of its image before it starts: 25 us for a 33 MB plane that HBM delivers in 6.  The kernels are bound by
// latency x bytes in flight, not by bytes.  Here:
//   * a thread owns 8 channels (16 B of bf16, 32 B of fp32) of a pixel row and keeps 8 (bf16) / 4 (fp32) rows in
//     flight: 128 B per thread, 64-128 KB per SM;
//   * the statistics kernel finishes the job itself: the CTA whose arrival completes an image (ticket counter) folds
//     the slice partials in slice order and writes mean / rstd (backward: the two sums), so the apply kernels start
//     streaming at once and carry 4 constants per channel (y = act(x*k + o); dx = k*dv + A*x + B);
//   * storage types of x / dx and y / dy / residual are independent template parameters (fp32 | bf16).
// Determinism and split invariance are those of norm.cu: fp32 sums over fixed ATOMS (4 rows of one thread forward, 2
// backward, aligned to 4*RPP / 2*RPP pixels from the start of the image), fp64 above the atom, slices start on atom
// group boundaries, partials folded in slice order - the statistics do not depend on the slice count, i.e. on how many
// images a rank holds.
// Algorithmic HBM bytes: forward read x + write y, backward read dy, x + write dx; the second read of x
// (apply after statistics) and of dy / x (backward apply) is an L2 hit for planes below the 126 MB L2.
#include "norm8.cuh"
#include <stdlib.h>

namespace srgan {

constexpr int kN8Threads = 256;
constexpr int kN8MaxSlices = 64;

struct Norm8P {
  int N, HW, C;
  int C8;          // 8-channel vectors per pixel row
  int TPR, RPP;    // threads per pixel row (vectors per channel chunk), pixel rows per CTA pass
  int chunks;      // channel chunks (gridDim.z)
  int SL, slice;   // pixel slices per image, pixels per slice (statistics kernels)
  int SLa, slice_a;  // the same for the apply kernels (no partials: any slicing)
  float eps, slope;
  int act;
};


struct N8Idx { int cg, row, v, px0, px1; bool active; };
__device__ __forceinline__ N8Idx n8_idx(const Norm8P& p, int slice) {
  N8Idx i;
  i.cg = threadIdx.x % p.TPR;
  i.row = threadIdx.x / p.TPR;
  i.active = i.row < p.RPP;
  i.v = blockIdx.z * p.TPR + i.cg;
  i.px0 = blockIdx.x * slice;
  i.px1 = min(p.HW, i.px0 + slice);
  return i;
}

// Sum the 16 running sums of every thread over the pixel rows of the CTA in row order (fixed), write the slice
// partial, and - in the CTA whose arrival completes (image, channel chunk) - fold the SL partials in slice order into
// `fold` (shared, TPR * 16 doubles).  Returns true in that CTA, with `fold` valid after the call.
__device__ __forceinline__ bool n8_finish(const Norm8P& p, const N8Idx& i, const double (&s)[16], double* sred,
                                          double* __restrict__ part, int* __restrict__ counters, int* s_last) {
  const int n = blockIdx.y;
#pragma unroll
  for (int j = 0; j < 16; ++j) sred[j * kN8Threads + threadIdx.x] = s[j];
  __syncthreads();
  if (i.row == 0) {
    double* dst = part + (((size_t)n * p.SL + blockIdx.x) * p.C8 + i.v) * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      double acc = 0.;
      for (int r = 0; r < p.RPP; ++r) acc += sred[j * kN8Threads + r * p.TPR + i.cg];
      dst[j] = acc;
    }
  }
  __threadfence();
  __syncthreads();
  int* ctr = counters + n * p.chunks + blockIdx.z;
  if (threadIdx.x == 0) *s_last = (atomicAdd(ctr, 1) == p.SL - 1);
  __syncthreads();
  if (!*s_last) return false;
  __threadfence();
  for (int idx = threadIdx.x; idx < p.TPR * 16; idx += kN8Threads) {
    const int cgi = idx >> 4, j = idx & 15;
    const double* src = part + (((size_t)n * p.SL) * p.C8 + blockIdx.z * p.TPR + cgi) * 16 + j;
    double acc = 0.;
    for (int sl = 0; sl < p.SL; ++sl) acc += __ldcg(src + (size_t)sl * p.C8 * 16);
    sred[idx] = acc;
  }
  if (threadIdx.x == 0) *ctr = 0;                 // ready for the next launch
  __syncthreads();
  return true;
}

// ------------------------------------------------------------------------------------------------ forward
template <typename TX>
__global__ void __launch_bounds__(kN8Threads, 2) inorm8_stats_kernel(Norm8P p, const TX* __restrict__ x,
                                                                    double* __restrict__ part,
                                                                    int* __restrict__ counters,
                                                                    float* __restrict__ mean, float* __restrict__ rstd) {
  __shared__ double sred[16 * kN8Threads];
  __shared__ int s_last;
  const N8Idx i = n8_idx(p, p.slice);
  const int n = blockIdx.y;
  const TX* xg = x + ((size_t)n * p.HW * p.C8 + i.v) * 8;
  const size_t rs = (size_t)p.C8 * 8;               // elements per pixel row
  double s[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) s[j] = 0.;
  if (i.active) {
    float pv[8];
    unpack(ld_raw(xg), pv);                         // pivot: first pixel of the plane
    const int step = p.RPP;
    int r = i.px0 + i.row;
    auto atom = [&](const Raw8<TX>* raw, int cnt) {
      float t1[8], t2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { t1[e] = 0.f; t2[e] = 0.f; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < cnt) {
          float v[8];
          unpack(raw[j], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) { const float d = v[e] - pv[e]; t1[e] += d; t2[e] += d * d; }
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] += (double)t1[e]; s[8 + e] += (double)t2[e]; }
    };
    if (Flight<TX>::kRows == 8) {
      for (; r + 7 * step < i.px1; r += 8 * step) {          // two atoms in flight
        Raw8<TX> raw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) raw[j] = ld_raw(xg + (size_t)(r + j * step) * rs);
        atom(raw, 4);
        atom(raw + 4, 4);
      }
    }
    for (; r + 3 * step < i.px1; r += 4 * step) {
      Raw8<TX> raw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) raw[j] = ld_raw(xg + (size_t)(r + j * step) * rs);
      atom(raw, 4);
    }
    if (r < i.px1) {                                         // the image's last, incomplete atom
      Raw8<TX> raw[4];
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r + j * step < i.px1) { raw[j] = ld_raw(xg + (size_t)(r + j * step) * rs); cnt = j + 1; }
      atom(raw, cnt);
    }
  }
  if (!n8_finish(p, i, s, sred, part, counters, &s_last)) return;
  // last CTA of (image, chunk): statistics of its TPR * 8 channels
  for (int idx = threadIdx.x; idx < p.TPR * 8; idx += kN8Threads) {
    const int cgi = idx >> 3, e = idx & 7;
    const int c = (blockIdx.z * p.TPR + cgi) * 8 + e;
    float pv[8];
    unpack(ld_raw(x + ((size_t)n * p.HW * p.C8 + blockIdx.z * p.TPR + cgi) * 8), pv);
    float mu, rs;
    n8_mean_rstd(sred[cgi * 16 + e], sred[cgi * 16 + 8 + e], pv[e], 1.f / (float)p.HW, p.eps, &mu, &rs);
    mean[(size_t)n * p.C + c] = mu;
    rstd[(size_t)n * p.C + c] = rs;
  }
}


// y = act(x * k + o) (+ residual):  k = rstd*gamma, o = (cbias - mean*rstd)*gamma + beta
template <typename TX, typename TY, int ACT>
__global__ void __launch_bounds__(kN8Threads, 3) inorm8_apply_kernel(
    Norm8P p, const TX* __restrict__ x, TY* __restrict__ y, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ cbias, const TY* __restrict__ residual) {
  const N8Idx i = n8_idx(p, p.slice_a);
  if (!i.active) return;
  const int n = blockIdx.y, c = i.v * 8;
  float k[8], o[8];
  {
    float mu[8], rs[8], g[8], b[8], tb[8];
    ld8f(mean + (size_t)n * p.C + c, mu);
    ld8f(rstd + (size_t)n * p.C + c, rs);
#pragma unroll
    for (int e = 0; e < 8; ++e) { g[e] = 1.f; b[e] = 0.f; tb[e] = 0.f; }
    if (gamma) ld8f(gamma + c, g);
    if (beta) ld8f(beta + c, b);
    if (cbias) ld8f(cbias + (size_t)n * p.C + c, tb);
#pragma unroll
    for (int e = 0; e < 8; ++e) { float cc; n8_consts(mu[e], rs[e], g[e], b[e], tb[e], &k[e], &o[e], &cc); }
  }
  const size_t rs_ = (size_t)p.C8 * 8;
  const size_t base = ((size_t)n * p.HW * p.C8 + i.v) * 8;
  const TX* xg = x + base;
  TY* yg = y + base;
  const TY* rg = residual ? residual + base : nullptr;
  auto one = [&](const Raw8<TX>& raw, int r) {
    float v[8];
    unpack(raw, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = n8_act<ACT>(fmaf(v[e], k[e], o[e]), p.slope);
    if (rg) {
      float q[8];
      unpack(ld_raw(rg + (size_t)r * rs_), q);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += q[e];
    }
    st8(yg + (size_t)r * rs_, v);
  };
  constexpr int NF = Flight<TX>::kRows;
  const int step = p.RPP;
  int r = i.px0 + i.row;
  for (; r + (NF - 1) * step < i.px1; r += NF * step) {
    Raw8<TX> raw[NF];
#pragma unroll
    for (int j = 0; j < NF; ++j) raw[j] = ld_raw(xg + (size_t)(r + j * step) * rs_);
#pragma unroll
    for (int j = 0; j < NF; ++j) one(raw[j], r + j * step);
  }
  for (; r < i.px1; r += step) one(ld_raw(xg + (size_t)r * rs_), r);
}

// ------------------------------------------------------------------------------------------------ backward
// dv = dy * act'(v), v = x*k + o ; xh = x*rs + c (c = -mean*rs) ; partial sums of dv and dv*xh over atoms of 2 rows
template <typename TX, typename TY, int ACT>
__global__ void __launch_bounds__(kN8Threads, 2) inorm8_bwd_reduce_kernel(
    Norm8P p, const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ cbias, double* __restrict__ part, int* __restrict__ counters,
    float* __restrict__ s1_out, float* __restrict__ s2_out) {
  __shared__ double sred[16 * kN8Threads];
  __shared__ int s_last;
  const N8Idx i = n8_idx(p, p.slice);
  const int n = blockIdx.y, c = i.v * 8;
  double s[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) s[j] = 0.;
  if (i.active) {
    float k[8], o[8], rs[8], cc[8];
    {
      float mu[8], g[8], b[8], tb[8];
      ld8f(mean + (size_t)n * p.C + c, mu);
      ld8f(rstd + (size_t)n * p.C + c, rs);
#pragma unroll
      for (int e = 0; e < 8; ++e) { g[e] = 1.f; b[e] = 0.f; tb[e] = 0.f; }
      if (gamma) ld8f(gamma + c, g);
      if (beta) ld8f(beta + c, b);
      if (cbias) ld8f(cbias + (size_t)n * p.C + c, tb);
#pragma unroll
      for (int e = 0; e < 8; ++e) n8_consts(mu[e], rs[e], g[e], b[e], tb[e], &k[e], &o[e], &cc[e]);
    }
    const size_t rs_ = (size_t)p.C8 * 8;
    const size_t base = ((size_t)n * p.HW * p.C8 + i.v) * 8;
    const TX* xg = x + base;
    const TY* dg = dy + base;
    auto atom = [&](const Raw8<TX>* rx, const Raw8<TY>* rd, int cnt) {
      float t1[8], t2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { t1[e] = 0.f; t2[e] = 0.f; }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j < cnt) {
          float xv[8], dv[8];
          unpack(rx[j], xv);
          unpack(rd[j], dv);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float d = dv[e] * n8_act_grad<ACT>(fmaf(xv[e], k[e], o[e]), p.slope);
            t1[e] += d; t2[e] += d * fmaf(xv[e], rs[e], cc[e]);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] += (double)t1[e]; s[8 + e] += (double)t2[e]; }
    };
    const int step = p.RPP;
    int r = i.px0 + i.row;
    constexpr bool kTwo = sizeof(TX) == 2 && sizeof(TY) == 2;       // two atoms (4 rows x 2 tensors) in flight
    if (kTwo) {
      for (; r + 3 * step < i.px1; r += 4 * step) {
        Raw8<TX> rx[4];
        Raw8<TY> rd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { rx[j] = ld_raw(xg + (size_t)(r + j * step) * rs_); rd[j] = ld_raw(dg + (size_t)(r + j * step) * rs_); }
        atom(rx, rd, 2);
        atom(rx + 2, rd + 2, 2);
      }
    }
    for (; r + step < i.px1; r += 2 * step) {
      Raw8<TX> rx[2];
      Raw8<TY> rd[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) { rx[j] = ld_raw(xg + (size_t)(r + j * step) * rs_); rd[j] = ld_raw(dg + (size_t)(r + j * step) * rs_); }
      atom(rx, rd, 2);
    }
    if (r < i.px1) {
      Raw8<TX> rx[2];
      Raw8<TY> rd[2];
      rx[0] = ld_raw(xg + (size_t)r * rs_); rd[0] = ld_raw(dg + (size_t)r * rs_);
      atom(rx, rd, 1);
    }
  }
  if (!n8_finish(p, i, s, sred, part, counters, &s_last)) return;
  for (int idx = threadIdx.x; idx < p.TPR * 8; idx += kN8Threads) {
    const int cgi = idx >> 3, e = idx & 7;
    const int ch = (blockIdx.z * p.TPR + cgi) * 8 + e;
    s1_out[(size_t)n * p.C + ch] = (float)sred[cgi * 16 + e];
    s2_out[(size_t)n * p.C + ch] = (float)sred[cgi * 16 + 8 + e];
  }
}

// dx = rstd*gamma * (dv - m1 - xh*m2), m = sums / HW   ==  k*dv + A*x + B  with A = -k*m2*rs, B = -k*(m1 + m2*c)
template <typename TX, typename TY, int ACT>
__global__ void __launch_bounds__(kN8Threads, 3) inorm8_bwd_apply_kernel(
    Norm8P p, const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ cbias, const float* __restrict__ s1, const float* __restrict__ s2,
    TX* __restrict__ dx) {
  const N8Idx i = n8_idx(p, p.slice_a);
  if (!i.active) return;
  const int n = blockIdx.y, c = i.v * 8;
  float k[8], o[8], A[8], B[8];
  {
    float mu[8], rs[8], g[8], b[8], tb[8], m1[8], m2[8];
    ld8f(mean + (size_t)n * p.C + c, mu);
    ld8f(rstd + (size_t)n * p.C + c, rs);
    ld8f(s1 + (size_t)n * p.C + c, m1);
    ld8f(s2 + (size_t)n * p.C + c, m2);
#pragma unroll
    for (int e = 0; e < 8; ++e) { g[e] = 1.f; b[e] = 0.f; tb[e] = 0.f; }
    if (gamma) ld8f(gamma + c, g);
    if (beta) ld8f(beta + c, b);
    if (cbias) ld8f(cbias + (size_t)n * p.C + c, tb);
    const float inv = 1.f / (float)p.HW;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float cc;
      n8_consts(mu[e], rs[e], g[e], b[e], tb[e], &k[e], &o[e], &cc);
      n8_bwd_consts(k[e], rs[e], cc, m1[e], m2[e], inv, &A[e], &B[e]);
    }
  }
  const size_t rs_ = (size_t)p.C8 * 8;
  const size_t base = ((size_t)n * p.HW * p.C8 + i.v) * 8;
  const TX* xg = x + base;
  const TY* dg = dy + base;
  TX* og = dx + base;
  auto one = [&](const Raw8<TX>& rx, const Raw8<TY>& rd, int r) {
    float xv[8], dv[8];
    unpack(rx, xv);
    unpack(rd, dv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = dv[e] * n8_act_grad<ACT>(fmaf(xv[e], k[e], o[e]), p.slope);
      xv[e] = fmaf(k[e], d, fmaf(A[e], xv[e], B[e]));
    }
    st8(og + (size_t)r * rs_, xv);
  };
  constexpr int NF = (sizeof(TX) == 2 && sizeof(TY) == 2) ? 4 : 2;
  const int step = p.RPP;
  int r = i.px0 + i.row;
  for (; r + (NF - 1) * step < i.px1; r += NF * step) {
    Raw8<TX> rx[NF];
    Raw8<TY> rd[NF];
#pragma unroll
    for (int j = 0; j < NF; ++j) { rx[j] = ld_raw(xg + (size_t)(r + j * step) * rs_); rd[j] = ld_raw(dg + (size_t)(r + j * step) * rs_); }
#pragma unroll
    for (int j = 0; j < NF; ++j) one(rx[j], rd[j], r + j * step);
  }
  for (; r < i.px1; r += step) one(ld_raw(xg + (size_t)r * rs_), ld_raw(dg + (size_t)r * rs_), r);
}

// ------------------------------------------------------------------------------------------------ host
// atom_rows: rows of one thread per atom (4 forward, 2 backward): slices of the statistics kernels start on multiples
// of atom_rows * RPP pixels
static bool plan_norm8(int N, int HW, int C, int atom_rows, Norm8P* out) {
  Norm8P p = {};
  p.N = N; p.HW = HW; p.C = C; p.C8 = C / 8;
  int chunks = 1;
  while (chunks <= p.C8 && (p.C8 % chunks || p.C8 / chunks > 32)) ++chunks;
  if (chunks > p.C8) return false;
  p.chunks = chunks;
  p.TPR = p.C8 / chunks;
  p.RPP = kN8Threads / p.TPR;
  static const int per_sm_s = getenv("SRGAN_DBG_NORM8_STATS_CTAS") ? atoi(getenv("SRGAN_DBG_NORM8_STATS_CTAS")) : 2;
  static const int per_sm_a = getenv("SRGAN_DBG_NORM8_APPLY_CTAS") ? atoi(getenv("SRGAN_DBG_NORM8_APPLY_CTAS")) : 3;
  auto slices = [&](int per_sm, int align, int* slice) {
    long long want = ((long long)per_sm * kNumSMs) / ((long long)N * chunks);
    long long cap = ceil_div(HW, p.RPP * 4);
    long long SL = want < cap ? want : cap;
    if (SL < 1) SL = 1;
    if (SL > kN8MaxSlices) SL = kN8MaxSlices;
    int sl = ceil_div(HW, (int)SL);
    sl = ceil_div(sl, align) * align;
    *slice = sl;
    return ceil_div(HW, sl);
  };
  p.SL = slices(per_sm_s, atom_rows * p.RPP, &p.slice);
  p.SLa = slices(per_sm_a, p.RPP, &p.slice_a);
  *out = p;
  return true;
}

static size_t norm8_ws_bytes(const Norm8P& p) { return (size_t)p.N * p.SL * p.C8 * 16 * sizeof(double); }

// mean / rstd of every (image, channel) from the per-tile sums a convolution epilogue wrote (conv_umma.cu STATS):
// rows of one image are added in row order in fp64: the result does not depend on the batch.
__global__ void inorm_stats_from_tiles_kernel(const float2* __restrict__ tiles, int rows, int N, int C, float inv_hw,
                                              float eps, float* __restrict__ mean, float* __restrict__ rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C;
  const float2* src = tiles + (size_t)n * rows * C + c;
  double s1 = 0., s2 = 0.;
  int r = 0;
  for (; r + 3 < rows; r += 4) {                       // 4 independent loads in flight, fixed summation order
    const float2 a = __ldg(src + (size_t)r * C), b = __ldg(src + (size_t)(r + 1) * C);
    const float2 c2 = __ldg(src + (size_t)(r + 2) * C), d = __ldg(src + (size_t)(r + 3) * C);
    s1 += (double)a.x; s2 += (double)a.y; s1 += (double)b.x; s2 += (double)b.y;
    s1 += (double)c2.x; s2 += (double)c2.y; s1 += (double)d.x; s2 += (double)d.y;
  }
  for (; r < rows; ++r) { const float2 a = __ldg(src + (size_t)r * C); s1 += (double)a.x; s2 += (double)a.y; }
  const double m = s1 * (double)inv_hw;
  double var = s2 * (double)inv_hw - m * m;
  if (var < 0.) var = 0.;
  mean[i] = (float)m;
  rstd[i] = rsqrtf((float)var + eps);
}

template <typename TX, typename TY, int ACT>
static void launch_fwd8(const Norm8P& p, const void* x, void* y, float* mean, float* rstd, const float* gamma,
                        const float* beta, const float* cbias, const void* residual, void* ws, int* counters,
                        cudaStream_t st, bool stats_given) {
  if (!stats_given)
    inorm8_stats_kernel<TX><<<dim3(p.SL, p.N, p.chunks), kN8Threads, 0, st>>>(p, (const TX*)x, (double*)ws, counters,
                                                                              mean, rstd);
  inorm8_apply_kernel<TX, TY, ACT><<<dim3(p.SLa, p.N, p.chunks), kN8Threads, 0, st>>>(
      p, (const TX*)x, (TY*)y, mean, rstd, gamma, beta, cbias, (const TY*)residual);
}
template <typename TX, typename TY, int ACT>
static void launch_bwd8(const Norm8P& p, const void* dy, const void* x, const float* mean, const float* rstd,
                        const float* gamma, const float* beta, const float* cbias, void* dx, float* s1, float* s2,
                        void* ws, int* counters, cudaStream_t st) {
  inorm8_bwd_reduce_kernel<TX, TY, ACT><<<dim3(p.SL, p.N, p.chunks), kN8Threads, 0, st>>>(
      p, (const TY*)dy, (const TX*)x, mean, rstd, gamma, beta, cbias, (double*)ws, counters, s1, s2);
  inorm8_bwd_apply_kernel<TX, TY, ACT><<<dim3(p.SLa, p.N, p.chunks), kN8Threads, 0, st>>>(
      p, (const TY*)dy, (const TX*)x, mean, rstd, gamma, beta, cbias, s1, s2, (TX*)dx);
}

#define SRGAN_N8_ACT(FN, TX, TY, ...)                                               \
  switch (act) {                                                                    \
    case SRGAN_ACT_RELU:  FN<TX, TY, SRGAN_ACT_RELU>(__VA_ARGS__); break;           \
    case SRGAN_ACT_LRELU: FN<TX, TY, SRGAN_ACT_LRELU>(__VA_ARGS__); break;          \
    case SRGAN_ACT_TANH:  FN<TX, TY, SRGAN_ACT_TANH>(__VA_ARGS__); break;           \
    default:              FN<TX, TY, SRGAN_ACT_NONE>(__VA_ARGS__); break;           \
  }
#define SRGAN_N8_TYPES(FN, ...)                                                     \
  do {                                                                              \
    using B = __nv_bfloat16;                                                        \
    if (xb && yb) { SRGAN_N8_ACT(FN, B, B, __VA_ARGS__) }                           \
    else if (xb) { SRGAN_N8_ACT(FN, B, float, __VA_ARGS__) }                        \
    else if (yb) { SRGAN_N8_ACT(FN, float, B, __VA_ARGS__) }                        \
    else { SRGAN_N8_ACT(FN, float, float, __VA_ARGS__) }                            \
  } while (0)

}  // namespace srgan

using namespace srgan;

extern "C" size_t srgan_inorm_mixed_workspace(int N, int HW, int C) {
  Norm8P p, q;
  if (N <= 0 || HW <= 0 || C <= 0 || C % 8 || !plan_norm8(N, HW, C, 4, &p) || !plan_norm8(N, HW, C, 2, &q)) return 0;
  const size_t a = norm8_ws_bytes(p), b = norm8_ws_bytes(q);
  return a > b ? a : b;
}
extern "C" size_t srgan_inorm_mixed_counters(int N, int C) { return (size_t)(N > 0 ? N : 0) * (C / 8 + 1) * sizeof(int); }

static int check8(const void* a, const void* b, int N, int HW, int C, int x_dtype, int y_dtype, const int* counters) {
  if (!a || !b || !counters) { set_error("inorm_mixed: null pointer"); return SRGAN_E_BADARG; }
  if (!(N >= 0 && HW > 0 && C > 0 && C % 8 == 0)) { set_error("inorm_mixed: need C %% 8 == 0, HW > 0"); return SRGAN_E_BADARG; }
  if (N > 65535) { set_error("inorm_mixed: N too large for grid.y"); return SRGAN_E_BADARG; }
  if (!((x_dtype == SRGAN_DT_F32 || x_dtype == SRGAN_DT_BF16) && (y_dtype == SRGAN_DT_F32 || y_dtype == SRGAN_DT_BF16))) {
    set_error("inorm_mixed: unknown dtype code");
    return SRGAN_E_BADARG;
  }
  return SRGAN_OK;
}

extern "C" int srgan_inorm_fwd_mixed(const void* x, int x_dtype, void* y, int y_dtype, float* mean, float* rstd,
                                     const float* gamma, const float* beta, const float* cbias, const void* residual,
                                     int N, int HW, int C, float eps, int act, float slope, int stats_given, void* ws,
                                     size_t ws_bytes, int* counters, void* stream) {
  if (int e = check8(x, y, N, HW, C, x_dtype, y_dtype, counters)) return e;
  SRGAN_CHECK_ARG(mean && rstd, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)x | (uintptr_t)y | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma |
                   (uintptr_t)beta | (uintptr_t)cbias | (uintptr_t)residual | (uintptr_t)ws) % 16 == 0,
                  "pointers must be 16-byte aligned");
  if (N == 0) return SRGAN_OK;
  Norm8P p;
  SRGAN_CHECK_ARG(plan_norm8(N, HW, C, 4, &p), "channel count cannot be mapped");
  p.eps = eps; p.slope = slope; p.act = act;
  if (!ws || ws_bytes < norm8_ws_bytes(p)) { set_error("inorm_fwd_mixed: workspace %zu < %zu", ws_bytes, norm8_ws_bytes(p)); return SRGAN_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const bool xb = x_dtype == SRGAN_DT_BF16, yb = y_dtype == SRGAN_DT_BF16;
  if (!stats_given) {                    // one pass over HBM when a cluster can hold the image (norm8c.cu)
    cudaError_t ce;
    if (norm8c_fwd(x, xb, y, yb, mean, rstd, gamma, beta, cbias, residual, N, HW, C, eps, act, slope, st, &ce)) {
      if (ce != cudaSuccess) { set_error("inorm_fwd_mixed (one pass): %s", cudaGetErrorString(ce)); return (int)ce; }
      SRGAN_RETURN_LAUNCH();
    }
  }
  SRGAN_N8_TYPES(launch_fwd8, p, x, y, mean, rstd, gamma, beta, cbias, residual, ws, counters, st, stats_given != 0);
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_inorm_stats_from_tiles(const float* tile_stats, int rows, int N, int HW, int C, float eps,
                                            float* mean, float* rstd, void* stream) {
  SRGAN_CHECK_ARG(tile_stats && mean && rstd, "null pointer");
  SRGAN_CHECK_ARG(rows > 0 && N >= 0 && HW > 0 && C > 0, "bad sizes");
  SRGAN_CHECK_ARG((uintptr_t)tile_stats % 8 == 0, "tile statistics must be 8-byte aligned");
  if (N == 0) return SRGAN_OK;
  const int total = N * C;
  inorm_stats_from_tiles_kernel<<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(
      (const float2*)tile_stats, rows, N, C, 1.f / (float)HW, eps, mean, rstd);
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_inorm_bwd_mixed(const void* dy, int y_dtype, const void* x, int x_dtype, const float* mean,
                                     const float* rstd, const float* gamma, const float* beta, const float* cbias,
                                     void* dx, float* s1, float* s2, int N, int HW, int C, int act, float slope,
                                     void* ws, size_t ws_bytes, int* counters, void* stream) {
  if (int e = check8(dy, x, N, HW, C, x_dtype, y_dtype, counters)) return e;
  SRGAN_CHECK_ARG(mean && rstd && dx && s1 && s2, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)dy | (uintptr_t)x | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma |
                   (uintptr_t)beta | (uintptr_t)cbias | (uintptr_t)dx | (uintptr_t)s1 | (uintptr_t)s2 |
                   (uintptr_t)ws) % 16 == 0, "pointers must be 16-byte aligned");
  if (N == 0) return SRGAN_OK;
  Norm8P p;
  SRGAN_CHECK_ARG(plan_norm8(N, HW, C, 2, &p), "channel count cannot be mapped");
  p.eps = 0.f; p.slope = slope; p.act = act;
  if (!ws || ws_bytes < norm8_ws_bytes(p)) { set_error("inorm_bwd_mixed: workspace %zu < %zu", ws_bytes, norm8_ws_bytes(p)); return SRGAN_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const bool xb = x_dtype == SRGAN_DT_BF16, yb = y_dtype == SRGAN_DT_BF16;
  {
    cudaError_t ce;
    if (norm8c_bwd(dy, yb, x, xb, mean, rstd, gamma, beta, cbias, dx, s1, s2, N, HW, C, act, slope, st, &ce)) {
      if (ce != cudaSuccess) { set_error("inorm_bwd_mixed (one pass): %s", cudaGetErrorString(ce)); return (int)ce; }
      SRGAN_RETURN_LAUNCH();
    }
  }
  SRGAN_N8_TYPES(launch_bwd8, p, dy, x, mean, rstd, gamma, beta, cbias, dx, s1, s2, ws, counters, st);
  SRGAN_RETURN_LAUNCH();
}
