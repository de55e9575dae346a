// fp32 FFMA implicit-GEMM convolution: fprop / dgrad / wgrad for arbitrary geometry.
//
// Role in the design (DESIGN.md "conv engines"): this is the exact-fp32 engine
// (SRGAN_CONV_FP32).  It serves (1) the tight parity tests against the fp32 CPU oracle,
// (2) the shapes tensor cores cannot take (C=3 stems, K=1/3/4 heads, tiny M), and (3) the
// on-device reference the tcgen05 engine is validated against at full size.
// The tensor-core engine lives in conv_umma.cu.
//
// GEMM views (NHWC activations, KRSC filters):
//   fprop : Y[m=(n,p,q)][k]    = sum_{(r,s,c)} X[n, p*st-pad+r, q*st-pad+s, c] * W[k][r][s][c]
//   dgrad : dX[m=(n,h,w)][c]   = sum_{(r,s,k)} dY[n,(h+pad-r)/st,(w+pad-s)/st,k] * W[k][r][s][c]
//   wgrad : dW[k][(r,s,c)]     = sum_{(n,p,q)} dY[n,p,q,k] * X[n, p*st-pad+r, q*st-pad+s, c]
#include "common.cuh"

namespace srgan {

template <int BM, int BN, int BK, int TM, int TN>
struct Cfg {
  static constexpr int kBM = BM, kBN = BN, kBK = BK, kTM = TM, kTN = TN;
  static constexpr int kThreads = (BM / TM) * (BN / TN);
  static constexpr int kAPer = BM * BK / kThreads;  // A elements per thread per k-tile
  static constexpr int kBPer = BN * BK / kThreads;
  static constexpr int kPad = 4;
  static_assert(kThreads == 256, "tile configs assume 256 threads");
  static_assert(TM % 4 == 0 || TM == 1 || TM == 2, "TM");
};

// Shared tile compute: acc[TM][TN] += As[kk][ty*TM+i] * Bs[kk][tx*TN+j]
template <class C>
__device__ __forceinline__ void tile_fma(const float (*As)[C::kBM + C::kPad],
                                         const float (*Bs)[C::kBN + C::kPad],
                                         float (&acc)[C::kTM][C::kTN], int ty, int tx) {
#pragma unroll
  for (int kk = 0; kk < C::kBK; ++kk) {
    float a[C::kTM], b[C::kTN];
#pragma unroll
    for (int i = 0; i < C::kTM; ++i) a[i] = As[kk][ty * C::kTM + i];
#pragma unroll
    for (int j = 0; j < C::kTN; ++j) b[j] = Bs[kk][tx * C::kTN + j];
#pragma unroll
    for (int i = 0; i < C::kTM; ++i)
#pragma unroll
      for (int j = 0; j < C::kTN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

struct ConvP {
  int N, H, W, C, K, R, S, P, Q, stride, pad;
  long long xs_n, xs_h, xs_w, xs_c;
};

// ------------------------------------------------------------------------------------ fprop
template <class C>
__global__ void __launch_bounds__(256) conv_fprop_ffma(ConvP d, const float* __restrict__ x,
                                                       const float* __restrict__ w,
                                                       const float* __restrict__ bias,
                                                       float* __restrict__ y, int act, float slope) {
  __shared__ float As[C::kBK][C::kBM + C::kPad];
  __shared__ float Bs[C::kBK][C::kBN + C::kPad];
  const int tid = threadIdx.x;
  const long long M = (long long)d.N * d.P * d.Q;
  const int Kg = d.R * d.S * d.C;
  const long long m0 = (long long)blockIdx.x * C::kBM;
  const int n0 = blockIdx.y * C::kBN;

  // loader mapping: k fastest across threads (coalesced along c)
  const int lk = tid % C::kBK;
  const int lrow = tid / C::kBK;                  // 0 .. 256/BK-1
  constexpr int kRowStep = 256 / C::kBK;
  // per-row decode (fixed over the k loop)
  long long abase[C::kAPer];
  int aih0[C::kAPer], aiw0[C::kAPer];
#pragma unroll
  for (int i = 0; i < C::kAPer; ++i) {
    long long m = m0 + lrow + i * kRowStep;
    if (m < M) {
      int q = (int)(m % d.Q);
      long long t = m / d.Q;
      int p = (int)(t % d.P);
      int n = (int)(t / d.P);
      abase[i] = (long long)n * d.xs_n;
      aih0[i] = p * d.stride - d.pad;
      aiw0[i] = q * d.stride - d.pad;
    } else {
      abase[i] = -1;
      aih0[i] = 0; aiw0[i] = 0;
    }
  }
  float ra[C::kAPer], rb[C::kBPer];
  auto load_tile = [&](int k0) {
    int k = k0 + lk;
    bool kin = k < Kg;
    int tap = kin ? k / d.C : 0;
    int ci = k - tap * d.C;
    int r = tap / d.S, s = tap - r * d.S;
#pragma unroll
    for (int i = 0; i < C::kAPer; ++i) {
      float v = 0.f;
      int ih = aih0[i] + r, iw = aiw0[i] + s;
      if (kin && abase[i] >= 0 && (unsigned)ih < (unsigned)d.H && (unsigned)iw < (unsigned)d.W)
        v = __ldg(x + abase[i] + ih * d.xs_h + iw * d.xs_w + ci * d.xs_c);
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < C::kBPer; ++i) {
      int n = n0 + lrow + i * kRowStep;
      rb[i] = (kin && n < d.K) ? __ldg(w + (long long)n * Kg + k) : 0.f;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < C::kAPer; ++i) As[lk][lrow + i * kRowStep] = ra[i];
#pragma unroll
    for (int i = 0; i < C::kBPer; ++i) Bs[lk][lrow + i * kRowStep] = rb[i];
  };

  const int tx = tid % (C::kBN / C::kTN), ty = tid / (C::kBN / C::kTN);
  float acc[C::kTM][C::kTN];
#pragma unroll
  for (int i = 0; i < C::kTM; ++i)
#pragma unroll
    for (int j = 0; j < C::kTN; ++j) acc[i][j] = 0.f;

  load_tile(0);
  for (int k0 = 0; k0 < Kg; k0 += C::kBK) {
    store_tile();
    __syncthreads();
    if (k0 + C::kBK < Kg) load_tile(k0 + C::kBK);   // prefetch into registers
    tile_fma<C>(As, Bs, acc, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < C::kTM; ++i) {
    long long m = m0 + ty * C::kTM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < C::kTN; ++j) {
      int n = n0 + tx * C::kTN + j;
      if (n < d.K) {
        float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
        y[m * d.K + n] = apply_act(v, act, slope);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ dgrad
template <class C>
__global__ void __launch_bounds__(256) conv_dgrad_ffma(ConvP d, const float* __restrict__ dy,
                                                       const float* __restrict__ w,
                                                       float* __restrict__ dx) {
  __shared__ float As[C::kBK][C::kBM + C::kPad];
  __shared__ float Bs[C::kBK][C::kBN + C::kPad];
  const int tid = threadIdx.x;
  const long long M = (long long)d.N * d.H * d.W;
  const int Kg = d.R * d.S * d.K;   // reduction: (tap, k) with k fastest
  const long long m0 = (long long)blockIdx.x * C::kBM;
  const int n0 = blockIdx.y * C::kBN;

  // A loader: k fastest across threads (coalesced along output channel k of dy)
  const int lk = tid % C::kBK;
  const int lrow = tid / C::kBK;
  constexpr int kRowStep = 256 / C::kBK;
  int an[C::kAPer], ahp[C::kAPer], awp[C::kAPer];
#pragma unroll
  for (int i = 0; i < C::kAPer; ++i) {
    long long m = m0 + lrow + i * kRowStep;
    if (m < M) {
      int iw = (int)(m % d.W);
      long long t = m / d.W;
      int ih = (int)(t % d.H);
      an[i] = (int)(t / d.H);
      ahp[i] = ih + d.pad;
      awp[i] = iw + d.pad;
    } else {
      an[i] = -1; ahp[i] = 0; awp[i] = 0;
    }
  }
  // B loader: n (input channel c) fastest across threads
  const int ln = tid % C::kBN;
  const int lkb = tid / C::kBN;
  constexpr int kKStep = 256 / C::kBN;
  float ra[C::kAPer], rb[C::kBPer];
  auto load_tile = [&](int k0) {
    {
      int k = k0 + lk;
      bool kin = k < Kg;
      int tap = kin ? k / d.K : 0;
      int co = k - tap * d.K;
      int r = tap / d.S, s = tap - r * d.S;
#pragma unroll
      for (int i = 0; i < C::kAPer; ++i) {
        float v = 0.f;
        int th = ahp[i] - r, tw = awp[i] - s;
        if (kin && an[i] >= 0 && th >= 0 && tw >= 0) {
          int p = th / d.stride, q = tw / d.stride;
          if (p * d.stride == th && q * d.stride == tw && p < d.P && q < d.Q)
            v = __ldg(dy + (((long long)an[i] * d.P + p) * d.Q + q) * d.K + co);
        }
        ra[i] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < C::kBPer; ++i) {
      int k = k0 + lkb + i * kKStep;
      int n = n0 + ln;
      float v = 0.f;
      if (k < Kg && n < d.C) {
        int tap = k / d.K;
        int co = k - tap * d.K;
        v = __ldg(w + ((long long)co * d.R * d.S + tap) * d.C + n);
      }
      rb[i] = v;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < C::kAPer; ++i) As[lk][lrow + i * kRowStep] = ra[i];
#pragma unroll
    for (int i = 0; i < C::kBPer; ++i) Bs[lkb + i * kKStep][ln] = rb[i];
  };

  const int tx = tid % (C::kBN / C::kTN), ty = tid / (C::kBN / C::kTN);
  float acc[C::kTM][C::kTN];
#pragma unroll
  for (int i = 0; i < C::kTM; ++i)
#pragma unroll
    for (int j = 0; j < C::kTN; ++j) acc[i][j] = 0.f;

  load_tile(0);
  for (int k0 = 0; k0 < Kg; k0 += C::kBK) {
    store_tile();
    __syncthreads();
    if (k0 + C::kBK < Kg) load_tile(k0 + C::kBK);
    tile_fma<C>(As, Bs, acc, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < C::kTM; ++i) {
    long long m = m0 + ty * C::kTM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < C::kTN; ++j) {
      int n = n0 + tx * C::kTN + j;
      if (n < d.C) dx[m * d.C + n] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------ wgrad
// grid = (ceil(K/BM), ceil(RSC/BN), splits); split z reduces pixels [z*chunk, (z+1)*chunk).
template <class C>
__global__ void __launch_bounds__(256) conv_wgrad_ffma(ConvP d, const float* __restrict__ x,
                                                       const float* __restrict__ dy,
                                                       float* __restrict__ out, long long chunk) {
  __shared__ float As[C::kBK][C::kBM + C::kPad];
  __shared__ float Bs[C::kBK][C::kBN + C::kPad];
  const int tid = threadIdx.x;
  const long long Mpix = (long long)d.N * d.P * d.Q;
  const int Ng = d.R * d.S * d.C;
  const int m0 = blockIdx.x * C::kBM;   // output channel k
  const int n0 = blockIdx.y * C::kBN;   // (tap, c)
  const long long kbeg = (long long)blockIdx.z * chunk;
  const long long kend = min(Mpix, kbeg + chunk);

  // A loader: m (=k channel of dy) fastest
  const int lm = tid % C::kBM;
  const int lka = tid / C::kBM;
  constexpr int kKStepA = 256 / C::kBM;
  // B loader: n (=(tap,c)) fastest; this thread's n is fixed for the whole kernel
  const int ln = tid % C::kBN;
  const int lkb = tid / C::kBN;
  constexpr int kKStepB = 256 / C::kBN;
  const int ng = n0 + ln;
  const bool nin = ng < Ng;
  int br = 0, bs = 0, bc = 0;
  if (nin) {
    int tap = ng / d.C;
    bc = ng - tap * d.C;
    br = tap / d.S;
    bs = tap - br * d.S;
  }
  float ra[C::kAPer], rb[C::kBPer];
  auto load_tile = [&](long long k0) {
#pragma unroll
    for (int i = 0; i < C::kAPer; ++i) {
      long long k = k0 + lka + i * kKStepA;
      int m = m0 + lm;
      ra[i] = (k < kend && m < d.K) ? __ldg(dy + k * d.K + m) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < C::kBPer; ++i) {
      long long k = k0 + lkb + i * kKStepB;
      float v = 0.f;
      if (k < kend && nin) {
        int q = (int)(k % d.Q);
        long long t = k / d.Q;
        int p = (int)(t % d.P);
        int n = (int)(t / d.P);
        int ih = p * d.stride - d.pad + br, iw = q * d.stride - d.pad + bs;
        if ((unsigned)ih < (unsigned)d.H && (unsigned)iw < (unsigned)d.W)
          v = __ldg(x + (long long)n * d.xs_n + ih * d.xs_h + iw * d.xs_w + bc * d.xs_c);
      }
      rb[i] = v;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < C::kAPer; ++i) As[lka + i * kKStepA][lm] = ra[i];
#pragma unroll
    for (int i = 0; i < C::kBPer; ++i) Bs[lkb + i * kKStepB][ln] = rb[i];
  };

  const int tx = tid % (C::kBN / C::kTN), ty = tid / (C::kBN / C::kTN);
  float acc[C::kTM][C::kTN];
#pragma unroll
  for (int i = 0; i < C::kTM; ++i)
#pragma unroll
    for (int j = 0; j < C::kTN; ++j) acc[i][j] = 0.f;

  if (kbeg < kend) load_tile(kbeg);
  for (long long k0 = kbeg; k0 < kend; k0 += C::kBK) {
    store_tile();
    __syncthreads();
    if (k0 + C::kBK < kend) load_tile(k0 + C::kBK);
    tile_fma<C>(As, Bs, acc, ty, tx);
    __syncthreads();
  }
  float* o = out + (long long)blockIdx.z * d.K * Ng;
#pragma unroll
  for (int i = 0; i < C::kTM; ++i) {
    int m = m0 + ty * C::kTM + i;
    if (m >= d.K) continue;
#pragma unroll
    for (int j = 0; j < C::kTN; ++j) {
      int n = n0 + tx * C::kTN + j;
      if (n < Ng) o[(long long)m * Ng + n] = acc[i][j];
    }
  }
}

// out[i] = sum_z part[z][i], z ascending (deterministic)
__global__ void splitk_reduce(const float* __restrict__ part, float* __restrict__ out, long long n, int splits) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(long long)z * n + i];
  out[i] = s;
}

// column sums of [rows][C]: grid.x = ceil(C/32), block (32, 8). Fixed order.
__global__ void colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int C) {
  __shared__ float red[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < C)
    for (long long r = threadIdx.y; r < rows; r += 8) s += x[r * C + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    out[c] = t;
  }
}

// stage 1 of the two-stage column sum: block (x = column tile, y = row block) sums its row range
__global__ void colsum_partial_kernel(const float* __restrict__ x, float* __restrict__ part, long long rows, int C,
                                      long long rows_per_block) {
  __shared__ float red[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += x[r * C + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    part[(long long)blockIdx.y * C + c] = t;
  }
}

using CfgWide = Cfg<128, 64, 16, 8, 4>;    // general shapes
using CfgThin = Cfg<256, 16, 16, 8, 2>;    // few output columns (K = 1,3,4 heads; C=3 dgrad)
using CfgHalf = Cfg<64, 64, 16, 4, 4>;     // wgrad of layers with <= 64 filters (stems)

static ConvP to_p(const srgan_conv_desc* d) {
  ConvP p;
  p.N = d->N; p.H = d->H; p.W = d->W; p.C = d->C; p.K = d->K; p.R = d->R; p.S = d->S;
  p.P = d->P; p.Q = d->Q; p.stride = d->stride; p.pad = d->pad;
  if (d->xs_n == 0 && d->xs_h == 0 && d->xs_w == 0 && d->xs_c == 0) {
    p.xs_c = 1; p.xs_w = d->C; p.xs_h = (long long)d->W * d->C; p.xs_n = (long long)d->H * d->W * d->C;
  } else {
    p.xs_n = d->xs_n; p.xs_h = d->xs_h; p.xs_w = d->xs_w; p.xs_c = d->xs_c;
  }
  return p;
}

int wgrad_ffma_splits(const srgan_conv_desc* d) {
  long long Mpix = (long long)d->N * d->P * d->Q;
  int Ng = d->R * d->S * d->C;
  int tiles = ceil_div(d->K, d->K <= 64 ? 64 : 128) * ceil_div(Ng, 64);
  int splits = ceil_div(2 * kNumSMs, tiles);
  long long maxs = ceil_div64(Mpix, 64);   // keep >= 64 pixels per split
  if (splits > maxs) splits = (int)maxs;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  return splits;
}

constexpr int kColsumBlocksF = 64;

size_t conv_ffma_workspace(const srgan_conv_desc* d, int pass) {
  if (pass != 2) return 0;
  int splits = wgrad_ffma_splits(d);
  size_t b = splits == 1 ? 0 : (size_t)splits * d->K * d->R * d->S * d->C * sizeof(float);
  return b + (size_t)kColsumBlocksF * d->K * sizeof(float);
}

int conv_fprop_ffma_launch(const srgan_conv_desc* d, const float* x, const float* w, const float* bias,
                           float* y, int act, float slope, cudaStream_t st) {
  ConvP p = to_p(d);
  long long M = (long long)d->N * d->P * d->Q;
  if (M == 0) return SRGAN_OK;
  if (d->K <= 16) {
    dim3 g((unsigned)ceil_div64(M, CfgThin::kBM), ceil_div(d->K, CfgThin::kBN));
    conv_fprop_ffma<CfgThin><<<g, 256, 0, st>>>(p, x, w, bias, y, act, slope);
  } else {
    dim3 g((unsigned)ceil_div64(M, CfgWide::kBM), ceil_div(d->K, CfgWide::kBN));
    conv_fprop_ffma<CfgWide><<<g, 256, 0, st>>>(p, x, w, bias, y, act, slope);
  }
  SRGAN_RETURN_LAUNCH();
}

int conv_dgrad_ffma_launch(const srgan_conv_desc* d, const float* dy, const float* w, float* dx,
                           cudaStream_t st) {
  ConvP p = to_p(d);
  long long M = (long long)d->N * d->H * d->W;
  if (M == 0) return SRGAN_OK;
  if (d->C <= 16) {
    dim3 g((unsigned)ceil_div64(M, CfgThin::kBM), ceil_div(d->C, CfgThin::kBN));
    conv_dgrad_ffma<CfgThin><<<g, 256, 0, st>>>(p, dy, w, dx);
  } else {
    dim3 g((unsigned)ceil_div64(M, CfgWide::kBM), ceil_div(d->C, CfgWide::kBN));
    conv_dgrad_ffma<CfgWide><<<g, 256, 0, st>>>(p, dy, w, dx);
  }
  SRGAN_RETURN_LAUNCH();
}

// ------------------------------------------------------------------------------------ small-K heads
// Discriminator heads (K = 1 patch logit, K = 4 class logits) have a long reduction (up to 64 taps x 512
// channels) and very few output pixels: one block per output pixel, threads stride over the (tap, channel/4)
// index space with float4 loads, fixed-order block reduction (deterministic).  Exact fp32.
template <int KMAX>
__global__ void __launch_bounds__(256) conv_fprop_head_kernel(ConvP d, const float* __restrict__ x,
                                                              const float* __restrict__ w,
                                                              const float* __restrict__ bias,
                                                              float* __restrict__ y, int act, float slope) {
  __shared__ float red[32];
  const int pix = blockIdx.x;
  const int q = pix % d.Q, pp = (pix / d.Q) % d.P, n = pix / (d.Q * d.P);
  const int C4 = d.C >> 2, T = d.R * d.S;
  float acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
  for (int i = threadIdx.x; i < T * C4; i += blockDim.x) {
    const int t = i / C4, c4 = i - t * C4;
    const int r = t / d.S, s = t - r * d.S;
    const int h = pp * d.stride - d.pad + r, ww = q * d.stride - d.pad + s;
    if (h < 0 || h >= d.H || ww < 0 || ww >= d.W) continue;
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * d.H + h) * d.W + ww) * d.C) + c4);
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < d.K) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + ((size_t)k * T + t) * d.C) + c4);
        acc[k] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[k]))));
      }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < d.K) {
      const float v = block_sum(acc[k], red);
      if (threadIdx.x == 0) y[(size_t)pix * d.K + k] = apply_act(v + (bias ? bias[k] : 0.f), act, slope);
    }
}

bool conv_head_supported(const srgan_conv_desc* d) {
  return d->K <= 4 && d->C % 4 == 0 && (long long)d->N * d->P * d->Q <= 65536 && d->R * d->S * d->C >= 1024;
}

int conv_fprop_head_launch(const srgan_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                           int act, float slope, cudaStream_t st) {
  ConvP p = to_p(d);
  const unsigned pixels = (unsigned)(d->N * d->P * d->Q);
  if (pixels == 0) return SRGAN_OK;
  if (((uintptr_t)x | (uintptr_t)w) % 16) { set_error("head conv: tensors must be 16-byte aligned"); return SRGAN_E_BADARG; }
  conv_fprop_head_kernel<4><<<pixels, 256, 0, st>>>(p, x, w, bias, y, act, slope);
  SRGAN_RETURN_LAUNCH();
}

void splitk_reduce_launch(const float* part, float* out, long long n, int splits, cudaStream_t st) {
  splitk_reduce<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(part, out, n, splits);
}

// out[c] = sum_r x[r][c].  With scratch (>= blocks*C floats) large inputs use two fixed-order stages.
int colsum_launch(const float* x, float* out, long long rows, int C, float* scratch, int scratch_blocks,
                  cudaStream_t st) {
  if (C == 0) return SRGAN_OK;
  if (scratch && scratch_blocks > 1 && rows >= 4096) {
    long long rpb = ceil_div64(rows, scratch_blocks);
    int nb = (int)ceil_div64(rows, rpb);
    colsum_partial_kernel<<<dim3(ceil_div(C, 32), nb), dim3(32, 8), 0, st>>>(x, scratch, rows, C, rpb);
    colsum_kernel<<<ceil_div(C, 32), dim3(32, 8), 0, st>>>(scratch, out, nb, C);
  } else {
    colsum_kernel<<<ceil_div(C, 32), dim3(32, 8), 0, st>>>(x, out, rows, C);
  }
  SRGAN_RETURN_LAUNCH();
}

int conv_wgrad_ffma_launch(const srgan_conv_desc* d, const float* x, const float* dy, float* dw,
                           float* dbias, void* ws, size_t ws_bytes, cudaStream_t st) {
  ConvP p = to_p(d);
  long long Mpix = (long long)d->N * d->P * d->Q;
  int Ng = d->R * d->S * d->C;
  int splits = wgrad_ffma_splits(d);
  size_t need = conv_ffma_workspace(d, 2);
  if (need > ws_bytes || !ws) { set_error("conv_wgrad: workspace %zu < %zu", ws_bytes, need); return SRGAN_E_WORKSPACE; }
  float* csum = (float*)ws + (splits == 1 ? 0 : (size_t)splits * d->K * Ng);
  if (dw) {
    long long chunk = ceil_div64(ceil_div64(Mpix, splits), CfgWide::kBK) * CfgWide::kBK;
    float* out = splits == 1 ? dw : (float*)ws;
    if (d->K <= 64) {
      dim3 g(ceil_div(d->K, CfgHalf::kBM), ceil_div(Ng, CfgHalf::kBN), splits);
      conv_wgrad_ffma<CfgHalf><<<g, 256, 0, st>>>(p, x, dy, out, chunk);
    } else {
      dim3 g(ceil_div(d->K, CfgWide::kBM), ceil_div(Ng, CfgWide::kBN), splits);
      conv_wgrad_ffma<CfgWide><<<g, 256, 0, st>>>(p, x, dy, out, chunk);
    }
    if (splits > 1) {
      long long n = (long long)d->K * Ng;
      splitk_reduce<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>((const float*)ws, dw, n, splits);
    }
  }
  if (dbias) return colsum_launch(dy, dbias, Mpix, d->K, csum, kColsumBlocksF, st);
  SRGAN_RETURN_LAUNCH();
}

}  // namespace srgan
