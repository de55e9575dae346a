// Memory-bound glue kernels: layout changes, reflect padding, pooling, activations,
// conditional bias, softmax, reparametrisation.  NHWC fp32; vectorised (float4) where the
// channel count allows, grid-stride loops sized to a multiple of the SM count.
#include <cuda_bf16.h>
#include "common.cuh"

namespace srgan {

static inline unsigned grid_for(size_t work, int threads) {
  size_t blocks = (work + threads - 1) / threads;
  size_t cap = (size_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}
#define GRID_STRIDE(i, n) \
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

// element type of the streaming kernels below: float, or float4 when channels and pointers allow 16-byte accesses
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float4 vadd(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float vscale(float a, float s) { return a * s; }
__device__ __forceinline__ float4 vscale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
template <typename T> __device__ __forceinline__ T vzero();
template <> __device__ __forceinline__ float vzero<float>() { return 0.f; }
template <> __device__ __forceinline__ float4 vzero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
static inline bool vec4_ok(int C, const void* a, const void* b, const void* c = nullptr) {
  return C % 4 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) % 16) == 0;
}

// ------------------------------------------------------------------ NCHW <-> NHWC (smem transpose)
// One block transposes a [C x 32 pixels] panel; coalesced on both sides.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xs = x + (size_t)n * C * HW;
  float* ys = y + (size_t)n * C * HW;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? xs[(size_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int p = p0 + j, c = c0 + threadIdx.x;
    if (c < C && p < HW) ys[(size_t)p * C + c] = tile[threadIdx.x][j];
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xs = x + (size_t)n * C * HW;
  float* ys = y + (size_t)n * C * HW;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int p = p0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? xs[(size_t)p * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, p = p0 + threadIdx.x;
    if (c < C && p < HW) ys[(size_t)c * HW + p] = tile[threadIdx.x][j];
  }
}

// ------------------------------------------------------------------ reflect padding
__device__ __forceinline__ int reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
// y[N][H+2p][W+2p][C]   (C counts elements of T)
template <typename T>
__global__ void reflect_pad_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C,
                                       int pad) {
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const size_t total = (size_t)N * Hp * Wp * C;
  GRID_STRIDE(i, total) {
    int c = (int)(i % (unsigned)C);
    size_t t = i / (unsigned)C;
    int w = (int)(t % (unsigned)Wp); t /= (unsigned)Wp;
    int h = (int)(t % (unsigned)Hp);
    int n = (int)(t / (unsigned)Hp);
    int sh = reflect(h - pad, H), sw = reflect(w - pad, W);
    y[i] = __ldg(x + (((size_t)n * H + sh) * W + sw) * C + c);
  }
}
// dx[n][h][w][c] = sum over padded positions that mirror onto (h,w); gather form, fixed order
template <typename T>
__global__ void reflect_pad_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C,
                                       int pad) {
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const size_t total = (size_t)N * H * W * C;
  GRID_STRIDE(i, total) {
    int c = (int)(i % (unsigned)C);
    size_t t = i / (unsigned)C;
    int w = (int)(t % (unsigned)W); t /= (unsigned)W;
    int h = (int)(t % (unsigned)H);
    int n = (int)(t / (unsigned)H);
    // candidate padded rows: h+pad (interior), pad-h (top mirror, 1<=h<=pad), 2(H-1)-h+pad (bottom mirror)
    int hs[3], ws[3], nh = 0, nw = 0;
    hs[nh++] = h + pad;
    if (h >= 1 && h <= pad) hs[nh++] = pad - h;
    if (h <= H - 2 && h >= H - 1 - pad) hs[nh++] = 2 * (H - 1) - h + pad;
    ws[nw++] = w + pad;
    if (w >= 1 && w <= pad) ws[nw++] = pad - w;
    if (w <= W - 2 && w >= W - 1 - pad) ws[nw++] = 2 * (W - 1) - w + pad;
    T s = vzero<T>();
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) s = vadd(s, __ldg(dy + (((size_t)n * Hp + hs[a]) * Wp + ws[b]) * C + c));
    dx[i] = s;
  }
}

// ------------------------------------------------------------------ pooling
// AvgPool2d(2,2): floor output size, trailing row/col dropped.   (C counts elements of T)
template <typename T>
__global__ void avgpool2_fwd_kernel(const T* __restrict__ x, const T* __restrict__ addend, T* __restrict__ y, int N,
                                    int H, int W, int C) {
  const int P = H / 2, Q = W / 2;
  const size_t total = (size_t)N * P * Q * C;
  GRID_STRIDE(i, total) {
    int c = (int)(i % (unsigned)C);
    size_t t = i / (unsigned)C;
    int q = (int)(t % (unsigned)Q); t /= (unsigned)Q;
    int p = (int)(t % (unsigned)P);
    int n = (int)(t / (unsigned)P);
    const T* b = x + (((size_t)n * H + 2 * p) * W + 2 * q) * C + c;
    T v = vscale(vadd(vadd(__ldg(b), __ldg(b + C)), vadd(__ldg(b + (size_t)W * C), __ldg(b + (size_t)W * C + C))), 0.25f);
    if (addend) v = vadd(v, __ldg(addend + i));
    y[i] = v;
  }
}
template <typename T>
__global__ void avgpool2_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C) {
  const int P = H / 2, Q = W / 2;
  const size_t total = (size_t)N * H * W * C;
  GRID_STRIDE(i, total) {
    int c = (int)(i % (unsigned)C);
    size_t t = i / (unsigned)C;
    int w = (int)(t % (unsigned)W); t /= (unsigned)W;
    int h = (int)(t % (unsigned)H);
    int n = (int)(t / (unsigned)H);
    int p = h >> 1, q = w >> 1;
    dx[i] = (p < P && q < Q) ? vscale(__ldg(dy + (((size_t)n * P + p) * Q + q) * C + c), 0.25f) : vzero<T>();
  }
}
// ---- bf16 storage (the encoder's trunk under the bf16 engine): 8 channels = 16 bytes per thread and access
__device__ __forceinline__ void bf8_unpack(const uint4& q, float (&v)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { v[2 * e] = __uint_as_float(w[e] << 16); v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
}
__device__ __forceinline__ uint4 bf8_pack(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    w[e] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// reflect pad backward on bf16: the (up to 4) mirrored contributions are added in fp32 and rounded once (C8 = C / 8)
__global__ void reflect_pad_bwd_bf16_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int N, int H, int W,
                                            int C8, int pad) {
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const size_t total = (size_t)N * H * W * C8;
  GRID_STRIDE(i, total) {
    int c = (int)(i % (unsigned)C8);
    size_t t = i / (unsigned)C8;
    int w = (int)(t % (unsigned)W); t /= (unsigned)W;
    int h = (int)(t % (unsigned)H);
    int n = (int)(t / (unsigned)H);
    int hs[3], ws[3], nh = 0, nw = 0;
    hs[nh++] = h + pad;
    if (h >= 1 && h <= pad) hs[nh++] = pad - h;
    if (h <= H - 2 && h >= H - 1 - pad) hs[nh++] = 2 * (H - 1) - h + pad;
    ws[nw++] = w + pad;
    if (w >= 1 && w <= pad) ws[nw++] = pad - w;
    if (w <= W - 2 && w >= W - 1 - pad) ws[nw++] = 2 * (W - 1) - w + pad;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) {
        float v[8];
        bf8_unpack(__ldg(dy + (((size_t)n * Hp + hs[a]) * Wp + ws[b]) * C8 + c), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += v[e];
      }
    dx[i] = bf8_pack(s);
  }
}
// y (fp32) = avgpool2(a: bf16) + b (fp32): the tail of an encoder block, big tensor in bf16, block output in fp32
__global__ void avgpool2_add_mixed_kernel(const uint4* __restrict__ a, const float4* __restrict__ b,
                                          float4* __restrict__ y, int N, int H, int W, int C8) {
  const int P = H / 2, Q = W / 2;
  const size_t total = (size_t)N * P * Q * C8;
  GRID_STRIDE(i, total) {
    int c = (int)(i % (unsigned)C8);
    size_t t = i / (unsigned)C8;
    int q = (int)(t % (unsigned)Q); t /= (unsigned)Q;
    int p = (int)(t % (unsigned)P);
    int n = (int)(t / (unsigned)P);
    const uint4* s = a + (((size_t)n * H + 2 * p) * W + 2 * q) * C8 + c;
    float v0[8], v1[8], v2[8], v3[8], o[8];
    bf8_unpack(__ldg(s), v0); bf8_unpack(__ldg(s + C8), v1);
    bf8_unpack(__ldg(s + (size_t)W * C8), v2); bf8_unpack(__ldg(s + (size_t)W * C8 + C8), v3);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = ((v0[e] + v1[e]) + (v2[e] + v3[e])) * 0.25f;
    const float4 b0 = __ldg(b + 2 * i), b1 = __ldg(b + 2 * i + 1);
    y[2 * i] = make_float4(o[0] + b0.x, o[1] + b0.y, o[2] + b0.z, o[3] + b0.w);
    y[2 * i + 1] = make_float4(o[4] + b1.x, o[5] + b1.y, o[6] + b1.z, o[7] + b1.w);
  }
}
// dx (bf16) = 0.25 * dy (fp32) broadcast over the 2x2 window (zero in a dropped trailing row / column)
__global__ void avgpool2_bwd_mixed_kernel(const float4* __restrict__ dy, uint4* __restrict__ dx, int N, int H, int W,
                                          int C8) {
  const int P = H / 2, Q = W / 2;
  const size_t total = (size_t)N * H * W * C8;
  GRID_STRIDE(i, total) {
    int c = (int)(i % (unsigned)C8);
    size_t t = i / (unsigned)C8;
    int w = (int)(t % (unsigned)W); t /= (unsigned)W;
    int h = (int)(t % (unsigned)H);
    int n = (int)(t / (unsigned)H);
    int p = h >> 1, q = w >> 1;
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p < P && q < Q) {
      const float4* s = dy + ((((size_t)n * P + p) * Q + q) * C8 + c) * 2;
      const float4 a = __ldg(s), b = __ldg(s + 1);
      o[0] = a.x * 0.25f; o[1] = a.y * 0.25f; o[2] = a.z * 0.25f; o[3] = a.w * 0.25f;
      o[4] = b.x * 0.25f; o[5] = b.y * 0.25f; o[6] = b.z * 0.25f; o[7] = b.w * 0.25f;
    }
    dx[i] = bf8_pack(o);
  }
}
// AvgPool2d(3, stride 2, padding 1, count_include_pad=False)
__global__ void avgpool3s2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W,
                                      int C) {
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)N * P * Q * C;
  GRID_STRIDE(i, total) {
    int c = (int)(i % C);
    size_t t = i / C;
    int q = (int)(t % Q); t /= Q;
    int p = (int)(t % P);
    int n = (int)(t / P);
    int h0 = max(2 * p - 1, 0), h1 = min(2 * p + 2, H), w0 = max(2 * q - 1, 0), w1 = min(2 * q + 2, W);
    float s = 0.f;
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) s += __ldg(x + (((size_t)n * H + h) * W + w) * C + c);
    y[i] = s / (float)((h1 - h0) * (w1 - w0));
  }
}
__global__ void avgpool3s2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N, int H, int W,
                                      int C) {
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)N * H * W * C;
  GRID_STRIDE(i, total) {
    int c = (int)(i % C);
    size_t t = i / C;
    int w = (int)(t % W); t /= W;
    int h = (int)(t % H);
    int n = (int)(t / H);
    // windows p with 2p-1 <= h <= 2p+1
    float s = 0.f;
    for (int p = max((h - 1 + 1) / 2, 0); p <= min((h + 1) / 2, P - 1); ++p) {
      int h0 = max(2 * p - 1, 0), h1 = min(2 * p + 2, H);
      if (h < h0 || h >= h1) continue;
      for (int q = max(w / 2, 0); q <= min((w + 1) / 2, Q - 1); ++q) {
        int w0 = max(2 * q - 1, 0), w1 = min(2 * q + 2, W);
        if (w < w0 || w >= w1) continue;
        s += __ldg(dy + (((size_t)n * P + p) * Q + q) * C + c) / (float)((h1 - h0) * (w1 - w0));
      }
    }
    dx[i] = s;
  }
}

// f[n][c] = mean_hw lrelu(x).  One warp-row per (n, 32 channels); block (32, 8) strides pixels.
__global__ void lrelu_gap_fwd_kernel(const float* __restrict__ x, float* __restrict__ f, int HW, int C,
                                     float slope) {
  __shared__ float red[8][33];
  const int n = blockIdx.y, c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < C)
    for (int p = threadIdx.y; p < HW; p += 8) {
      float v = __ldg(x + ((size_t)n * HW + p) * C + c);
      s += v > 0.f ? v : v * slope;
    }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    f[(size_t)n * C + c] = t / (float)HW;
  }
}
__global__ void lrelu_gap_bwd_kernel(const float* __restrict__ df, const float* __restrict__ x,
                                     float* __restrict__ dx, int N, int HW, int C, float slope) {
  const size_t total = (size_t)N * HW * C;
  const float inv = 1.f / (float)HW;
  GRID_STRIDE(i, total) {
    int c = (int)(i % C);
    int n = (int)(i / ((size_t)HW * C));
    float v = __ldg(x + i);
    dx[i] = __ldg(df + (size_t)n * C + c) * inv * (v > 0.f ? 1.f : slope);
  }
}

// ------------------------------------------------------------------ elementwise
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                               size_t n, int act, float slope) {
  size_t n4 = n / 4;
  const float4* d4 = reinterpret_cast<const float4*>(dy);
  const float4* y4 = reinterpret_cast<const float4*>(y);
  float4* o4 = reinterpret_cast<float4*>(dx);
  GRID_STRIDE(i, n4) {
    float4 a = __ldg(d4 + i), b = __ldg(y4 + i), o;
    o.x = a.x * act_grad_out(b.x, act, slope);
    o.y = a.y * act_grad_out(b.y, act, slope);
    o.z = a.z * act_grad_out(b.z, act, slope);
    o.w = a.w * act_grad_out(b.w, act, slope);
    o4[i] = o;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dx[i] = dy[i] * act_grad_out(y[i], act, slope);
}
// dst = bf16(src), 8 elements (32 B in, 16 B out) per thread and trip
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  const size_t n8 = n / 8;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  GRID_STRIDE(i, n8) {
    const float4 a = __ldg(s4 + 2 * i), b = __ldg(s4 + 2 * i + 1);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    d4[i] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                       *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
  }
  for (size_t i = n8 * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}
// bf16 forms for the bf16 discriminator tower: dz = dy * act'(y) on bf16 tensors (8 elements per thread and trip), and
// the widening cast at the tower's fp32 boundary
__device__ __forceinline__ void bf16x8_to_f32(const uint4& q, float* o) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { o[2 * e] = __uint_as_float(w[e] << 16); o[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
}
__global__ void act_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                    __nv_bfloat16* __restrict__ dx, size_t n, int act, float slope) {
  const size_t n8 = n / 8;
  const uint4* d4 = reinterpret_cast<const uint4*>(dy);
  const uint4* y4 = reinterpret_cast<const uint4*>(y);
  uint4* o4 = reinterpret_cast<uint4*>(dx);
  GRID_STRIDE(i, n8) {
    float a[8], b[8];
    bf16x8_to_f32(__ldg(d4 + i), a);
    bf16x8_to_f32(__ldg(y4 + i), b);
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 h = __floats2bfloat162_rn(a[2 * e] * act_grad_out(b[2 * e], act, slope),
                                               a[2 * e + 1] * act_grad_out(b[2 * e + 1], act, slope));
      w[e] = *reinterpret_cast<uint32_t*>(&h);
    }
    o4[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (size_t i = n8 * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dx[i] = __float2bfloat16_rn(__bfloat162float(dy[i]) * act_grad_out(__bfloat162float(y[i]), act, slope));
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t n8 = n / 8;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  GRID_STRIDE(i, n8) {
    float a[8];
    bf16x8_to_f32(__ldg(s4 + i), a);
    d4[2 * i] = make_float4(a[0], a[1], a[2], a[3]);
    d4[2 * i + 1] = make_float4(a[4], a[5], a[6], a[7]);
  }
  for (size_t i = n8 * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                           size_t n) {
  size_t n4 = n / 4;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float4* o4 = reinterpret_cast<float4*>(y);
  GRID_STRIDE(i, n4) {
    float4 p = __ldg(a4 + i), q = __ldg(b4 + i);
    o4[i] = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = a[i] + b[i];
}

// ------------------------------------------------------------------ conditional bias
// t[n][c] = tanh(b[c] + sum_j con[n][j] w[c][j]); J is small (<= 64).
__global__ void condbias_fwd_kernel(const float* __restrict__ con, const float* __restrict__ w,
                                    const float* __restrict__ b, float* __restrict__ t, int N, int J, int C) {
  const size_t total = (size_t)N * C;
  GRID_STRIDE(i, total) {
    int c = (int)(i % C), n = (int)(i / C);
    float s = b ? __ldg(b + c) : 0.f;
    for (int j = 0; j < J; ++j) s = fmaf(__ldg(con + (size_t)n * J + j), __ldg(w + (size_t)c * J + j), s);
    t[i] = tanhf(s);
  }
}
// One WARP per output element (dw[c][j], db[c], dcon[n][j]): lanes stride over the reduction index and a
// butterfly shuffle adds them in a fixed order (deterministic); the loads of a warp are independent, so the
// kernel is no longer a chain of 64-256 dependent loads per thread.
__global__ void condbias_bwd_kernel(const float* __restrict__ dt, const float* __restrict__ t,
                                    const float* __restrict__ con, const float* __restrict__ w,
                                    float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dcon,
                                    int N, int J, int C) {
  const size_t nw = (size_t)C * J, nb = C, nc = (size_t)N * J;
  const size_t total = nw + nb + nc;
  const int lane = threadIdx.x & 31;
  const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < total; i += warps) {
    float s = 0.f;
    if (i < nw) {
      if (!dw) continue;
      const int c = (int)(i / J), j = (int)(i % J);
      for (int n = lane; n < N; n += 32) {
        const float tv = __ldg(t + (size_t)n * C + c);
        s = fmaf(__ldg(dt + (size_t)n * C + c) * (1.f - tv * tv), __ldg(con + (size_t)n * J + j), s);
      }
      s = warp_sum(s);
      if (lane == 0) dw[i] = s;
    } else if (i < nw + nb) {
      if (!db) continue;
      const int c = (int)(i - nw);
      for (int n = lane; n < N; n += 32) {
        const float tv = __ldg(t + (size_t)n * C + c);
        s += __ldg(dt + (size_t)n * C + c) * (1.f - tv * tv);
      }
      s = warp_sum(s);
      if (lane == 0) db[c] = s;
    } else {
      if (!dcon) continue;
      const size_t k = i - nw - nb;
      const int n = (int)(k / J), j = (int)(k % J);
      for (int c = lane; c < C; c += 32) {
        const float tv = __ldg(t + (size_t)n * C + c);
        s = fmaf(__ldg(dt + (size_t)n * C + c) * (1.f - tv * tv), __ldg(w + (size_t)c * J + j), s);
      }
      s = warp_sum(s);
      if (lane == 0) dcon[k] = s;
    }
  }
}

// ------------------------------------------------------------------ softmax over the last dim (J small)
__global__ void softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int J) {
  GRID_STRIDE(n, (size_t)N) {
    const float* r = x + n * J;
    float m = r[0];
    for (int j = 1; j < J; ++j) m = fmaxf(m, r[j]);
    float s = 0.f;
    for (int j = 0; j < J; ++j) s += expf(r[j] - m);
    for (int j = 0; j < J; ++j) y[n * J + j] = expf(r[j] - m) / s;
  }
}
__global__ void softmax_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                   float* __restrict__ dx, int N, int J) {
  GRID_STRIDE(n, (size_t)N) {
    float dot = 0.f;
    for (int j = 0; j < J; ++j) dot += dy[n * J + j] * y[n * J + j];
    for (int j = 0; j < J; ++j) dx[n * J + j] = y[n * J + j] * (dy[n * J + j] - dot);
  }
}

// ------------------------------------------------------------------ reparametrisation
__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                   const float* __restrict__ eps, float* __restrict__ z, size_t n) {
  GRID_STRIDE(i, n) z[i] = fmaf(eps[i], expf(0.5f * lv[i]), mu[i]);
}
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ lv,
                                   const float* __restrict__ eps, float* __restrict__ dmu,
                                   float* __restrict__ dlv, size_t n) {
  GRID_STRIDE(i, n) {
    float g = dz[i];
    if (dmu) dmu[i] = g;
    if (dlv) dlv[i] = g * eps[i] * 0.5f * expf(0.5f * lv[i]);
  }
}

// g += p1 (+ p2), then p1 = p2 = 0: the later gradient contributions of a backward pass, parked in "pending" copies of
// the optimizer's flat gradient buffer, are folded in ARRIVAL order - bit-identical to autograd's one-add-per-tensor
// accumulation - and the parking buffers are left zero for the next pass.
__global__ void grad_fold_kernel(float4* __restrict__ g, float4* __restrict__ p1, float4* __restrict__ p2, size_t n4) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // two elements per trip: six independent 16-byte loads in flight per thread (the one-element form reached 1.1 TB/s)
  for (; i + stride < n4; i += 2 * stride) {
    float4 a0 = g[i], a1 = g[i + stride];
    const float4 b0 = p1[i], b1 = p1[i + stride];
    float4 c0 = z, c1 = z;
    if (p2) { c0 = p2[i]; c1 = p2[i + stride]; }
    a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
    a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
    if (p2) {
      a0.x += c0.x; a0.y += c0.y; a0.z += c0.z; a0.w += c0.w;
      a1.x += c1.x; a1.y += c1.y; a1.z += c1.z; a1.w += c1.w;
      p2[i] = z; p2[i + stride] = z;
    }
    p1[i] = z; p1[i + stride] = z;
    g[i] = a0; g[i + stride] = a1;
  }
  for (; i < n4; i += stride) {
    float4 a = g[i];
    const float4 b = p1[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    p1[i] = z;
    if (p2) {
      const float4 c = p2[i];
      a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
      p2[i] = z;
    }
    g[i] = a;
  }
}
}  // namespace srgan

using namespace srgan;
#define ST ((cudaStream_t)stream)

extern "C" int srgan_nchw_to_nhwc(const float* x, float* y, int N, int C, int H, int W, void* stream) {
  SRGAN_CHECK_ARG(x && y && N >= 0 && C > 0 && H > 0 && W > 0, "bad argument");
  if (N == 0) return SRGAN_OK;
  dim3 g(ceil_div(H * W, 32), ceil_div(C, 32), N);
  nchw_to_nhwc_kernel<<<g, dim3(32, 8), 0, ST>>>(x, y, C, H * W);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_nhwc_to_nchw(const float* x, float* y, int N, int C, int H, int W, void* stream) {
  SRGAN_CHECK_ARG(x && y && N >= 0 && C > 0 && H > 0 && W > 0, "bad argument");
  if (N == 0) return SRGAN_OK;
  dim3 g(ceil_div(H * W, 32), ceil_div(C, 32), N);
  nhwc_to_nchw_kernel<<<g, dim3(32, 8), 0, ST>>>(x, y, C, H * W);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_reflect_pad_fwd(const float* x, float* y, int N, int H, int W, int C, int pad, void* stream) {
  SRGAN_CHECK_ARG(x && y && pad >= 0 && pad < H && pad < W, "reflect pad needs pad < H,W");
  size_t total = (size_t)N * (H + 2 * pad) * (W + 2 * pad) * C;
  if (total == 0) return SRGAN_OK;
  if (vec4_ok(C, x, y))
    reflect_pad_fwd_kernel<float4><<<grid_for(total / 4, 256), 256, 0, ST>>>((const float4*)x, (float4*)y, N, H, W, C / 4, pad);
  else
    reflect_pad_fwd_kernel<float><<<grid_for(total, 256), 256, 0, ST>>>(x, y, N, H, W, C, pad);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_reflect_pad_bwd(const float* dy, float* dx, int N, int H, int W, int C, int pad, void* stream) {
  SRGAN_CHECK_ARG(dy && dx && pad >= 0 && pad < H && pad < W, "reflect pad needs pad < H,W");
  size_t total = (size_t)N * H * W * C;
  if (total == 0) return SRGAN_OK;
  if (vec4_ok(C, dy, dx))
    reflect_pad_bwd_kernel<float4><<<grid_for(total / 4, 256), 256, 0, ST>>>((const float4*)dy, (float4*)dx, N, H, W, C / 4, pad);
  else
    reflect_pad_bwd_kernel<float><<<grid_for(total, 256), 256, 0, ST>>>(dy, dx, N, H, W, C, pad);
  SRGAN_RETURN_LAUNCH();
}
// ---- bf16 storage variants (C % 8 == 0, 16-byte aligned tensors)
extern "C" int srgan_reflect_pad_fwd_bf16(const void* x, void* y, int N, int H, int W, int C, int pad, void* stream) {
  SRGAN_CHECK_ARG(x && y && pad >= 0 && pad < H && pad < W, "reflect pad needs pad < H,W");
  SRGAN_CHECK_ARG(C % 8 == 0 && ((uintptr_t)x | (uintptr_t)y) % 16 == 0, "bf16 reflect pad: C % 8 == 0, 16-byte aligned");
  size_t total = (size_t)N * (H + 2 * pad) * (W + 2 * pad) * (C / 8);
  if (total == 0) return SRGAN_OK;
  reflect_pad_fwd_kernel<uint4><<<grid_for(total, 256), 256, 0, ST>>>((const uint4*)x, (uint4*)y, N, H, W, C / 8, pad);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_reflect_pad_bwd_bf16(const void* dy, void* dx, int N, int H, int W, int C, int pad, void* stream) {
  SRGAN_CHECK_ARG(dy && dx && pad >= 0 && pad < H && pad < W, "reflect pad needs pad < H,W");
  SRGAN_CHECK_ARG(C % 8 == 0 && ((uintptr_t)dy | (uintptr_t)dx) % 16 == 0, "bf16 reflect pad: C % 8 == 0, 16-byte aligned");
  size_t total = (size_t)N * H * W * (C / 8);
  if (total == 0) return SRGAN_OK;
  reflect_pad_bwd_bf16_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const uint4*)dy, (uint4*)dx, N, H, W, C / 8, pad);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_avgpool2_add_fwd_mixed(const void* a_bf16, const float* b, float* y, int N, int H, int W, int C,
                                            void* stream) {
  SRGAN_CHECK_ARG(a_bf16 && b && y, "null pointer");
  SRGAN_CHECK_ARG(C % 8 == 0 && ((uintptr_t)a_bf16 | (uintptr_t)b | (uintptr_t)y) % 16 == 0, "C % 8 == 0, 16-byte aligned");
  size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return SRGAN_OK;
  avgpool2_add_mixed_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const uint4*)a_bf16, (const float4*)b, (float4*)y, N, H, W, C / 8);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_avgpool2_bwd_mixed(const float* dy, void* dx_bf16, int N, int H, int W, int C, void* stream) {
  SRGAN_CHECK_ARG(dy && dx_bf16, "null pointer");
  SRGAN_CHECK_ARG(C % 8 == 0 && ((uintptr_t)dy | (uintptr_t)dx_bf16) % 16 == 0, "C % 8 == 0, 16-byte aligned");
  size_t total = (size_t)N * H * W * (C / 8);
  if (total == 0) return SRGAN_OK;
  avgpool2_bwd_mixed_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const float4*)dy, (uint4*)dx_bf16, N, H, W, C / 8);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_avgpool2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream) {
  SRGAN_CHECK_ARG(x && y, "null pointer");
  size_t total = (size_t)N * (H / 2) * (W / 2) * C;
  if (total == 0) return SRGAN_OK;
  if (vec4_ok(C, x, y))
    avgpool2_fwd_kernel<float4><<<grid_for(total / 4, 256), 256, 0, ST>>>((const float4*)x, nullptr, (float4*)y, N, H, W, C / 4);
  else
    avgpool2_fwd_kernel<float><<<grid_for(total, 256), 256, 0, ST>>>(x, nullptr, y, N, H, W, C);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_avgpool2_add_fwd(const float* a, const float* b, float* y, int N, int H, int W, int C,
                                      void* stream) {
  SRGAN_CHECK_ARG(a && b && y, "null pointer");
  size_t total = (size_t)N * (H / 2) * (W / 2) * C;
  if (total == 0) return SRGAN_OK;
  if (vec4_ok(C, a, b, y))
    avgpool2_fwd_kernel<float4><<<grid_for(total / 4, 256), 256, 0, ST>>>((const float4*)a, (const float4*)b, (float4*)y, N, H, W, C / 4);
  else
    avgpool2_fwd_kernel<float><<<grid_for(total, 256), 256, 0, ST>>>(a, b, y, N, H, W, C);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_avgpool2_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream) {
  SRGAN_CHECK_ARG(dy && dx, "null pointer");
  size_t total = (size_t)N * H * W * C;
  if (total == 0) return SRGAN_OK;
  if (vec4_ok(C, dy, dx))
    avgpool2_bwd_kernel<float4><<<grid_for(total / 4, 256), 256, 0, ST>>>((const float4*)dy, (float4*)dx, N, H, W, C / 4);
  else
    avgpool2_bwd_kernel<float><<<grid_for(total, 256), 256, 0, ST>>>(dy, dx, N, H, W, C);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_avgpool3s2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream) {
  SRGAN_CHECK_ARG(x && y, "null pointer");
  size_t total = (size_t)N * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1) * C;
  if (total == 0) return SRGAN_OK;
  avgpool3s2_fwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>(x, y, N, H, W, C);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_avgpool3s2_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream) {
  SRGAN_CHECK_ARG(dy && dx, "null pointer");
  size_t total = (size_t)N * H * W * C;
  if (total == 0) return SRGAN_OK;
  avgpool3s2_bwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dy, dx, N, H, W, C);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_lrelu_gap_fwd(const float* x, float* f, int N, int HW, int C, float slope, void* stream) {
  SRGAN_CHECK_ARG(x && f && HW > 0 && N <= 65535, "bad argument");
  if (N == 0 || C == 0) return SRGAN_OK;
  lrelu_gap_fwd_kernel<<<dim3(ceil_div(C, 32), N), dim3(32, 8), 0, ST>>>(x, f, HW, C, slope);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_lrelu_gap_bwd(const float* df, const float* x, float* dx, int N, int HW, int C, float slope,
                                   void* stream) {
  SRGAN_CHECK_ARG(df && x && dx && HW > 0, "bad argument");
  size_t total = (size_t)N * HW * C;
  if (total == 0) return SRGAN_OK;
  lrelu_gap_bwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>(df, x, dx, N, HW, C, slope);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_act_bwd(const float* dy, const float* y, float* dx, size_t n, int act, float slope,
                             void* stream) {
  SRGAN_CHECK_ARG(dy && y && dx, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)dy | (uintptr_t)y | (uintptr_t)dx) % 16 == 0, "pointers must be 16-byte aligned");
  if (n == 0) return SRGAN_OK;
  act_bwd_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, ST>>>(dy, y, dx, n, act, slope);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_act_bwd_bf16(const void* dy, const void* y, void* dx, size_t n, int act, float slope, void* stream) {
  SRGAN_CHECK_ARG(dy && y && dx, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)dy | (uintptr_t)y | (uintptr_t)dx) % 16 == 0, "pointers must be 16-byte aligned");
  if (n == 0) return SRGAN_OK;
  act_bwd_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, ST>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y,
                                                              (__nv_bfloat16*)dx, n, act, slope);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_cast_bf16_f32(const void* src, float* dst, size_t n, void* stream) {
  SRGAN_CHECK_ARG(src && dst, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)src | (uintptr_t)dst) % 16 == 0, "pointers must be 16-byte aligned");
  if (n == 0) return SRGAN_OK;
  cast_bf16_f32_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, ST>>>((const __nv_bfloat16*)src, dst, n);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_grad_fold(float* g, float* p1, float* p2, size_t n, void* stream) {
  SRGAN_CHECK_ARG(g && p1, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)g | (uintptr_t)p1 | (uintptr_t)p2) % 16 == 0 && n % 4 == 0, "buffers must be 16-byte aligned, n % 4 == 0");
  if (n == 0) return SRGAN_OK;
  grad_fold_kernel<<<grid_for(n / 4, 256), 256, 0, ST>>>((float4*)g, (float4*)p1, (float4*)p2, n / 4);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_cast_f32_bf16(const float* src, void* dst, size_t n, void* stream) {
  SRGAN_CHECK_ARG(src && dst, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)src | (uintptr_t)dst) % 16 == 0, "pointers must be 16-byte aligned");
  if (n == 0) return SRGAN_OK;
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, ST>>>(src, (__nv_bfloat16*)dst, n);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_add(const float* a, const float* b, float* y, size_t n, void* stream) {
  SRGAN_CHECK_ARG(a && b && y, "null pointer");
  SRGAN_CHECK_ARG(((uintptr_t)a | (uintptr_t)b | (uintptr_t)y) % 16 == 0, "pointers must be 16-byte aligned");
  if (n == 0) return SRGAN_OK;
  add_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, ST>>>(a, b, y, n);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_condbias_fwd(const float* con, const float* w, const float* b, float* t, int N, int J, int C,
                                  void* stream) {
  SRGAN_CHECK_ARG(con && w && t && J > 0 && C > 0, "bad argument");
  if (N == 0) return SRGAN_OK;
  condbias_fwd_kernel<<<grid_for((size_t)N * C, 128), 128, 0, ST>>>(con, w, b, t, N, J, C);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_condbias_bwd(const float* dt, const float* t, const float* con, const float* w, float* dw,
                                  float* db, float* dcon, int N, int J, int C, void* stream) {
  SRGAN_CHECK_ARG(dt && t && con && w && J > 0 && C > 0, "bad argument");
  size_t total = (size_t)C * J + C + (size_t)N * J;            // one warp per output
  condbias_bwd_kernel<<<grid_for(total * 32, 256), 256, 0, ST>>>(dt, t, con, w, dw, db, dcon, N, J, C);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_softmax_fwd(const float* x, float* y, int N, int J, void* stream) {
  SRGAN_CHECK_ARG(x && y && J > 0, "bad argument");
  if (N == 0) return SRGAN_OK;
  softmax_fwd_kernel<<<grid_for(N, 128), 128, 0, ST>>>(x, y, N, J);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_softmax_bwd(const float* dy, const float* y, float* dx, int N, int J, void* stream) {
  SRGAN_CHECK_ARG(dy && y && dx && J > 0, "bad argument");
  if (N == 0) return SRGAN_OK;
  softmax_bwd_kernel<<<grid_for(N, 128), 128, 0, ST>>>(dy, y, dx, N, J);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z, size_t n,
                                 void* stream) {
  SRGAN_CHECK_ARG(mu && logvar && eps && z, "null pointer");
  if (n == 0) return SRGAN_OK;
  reparam_fwd_kernel<<<grid_for(n, 128), 128, 0, ST>>>(mu, logvar, eps, z, n);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_reparam_bwd(const float* dz, const float* logvar, const float* eps, float* dmu,
                                 float* dlogvar, size_t n, void* stream) {
  SRGAN_CHECK_ARG(dz && logvar && eps, "null pointer");
  if (n == 0) return SRGAN_OK;
  reparam_bwd_kernel<<<grid_for(n, 128), 128, 0, ST>>>(dz, logvar, eps, dmu, dlogvar, n);
  SRGAN_RETURN_LAUNCH();
}
