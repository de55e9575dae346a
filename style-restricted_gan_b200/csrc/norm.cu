// Fused (conditional) instance normalisation, forward and backward, NHWC fp32.
//
//   xh = (x - mean_hw) * rstd_hw ; v = (xh + cbias[n][c]) * gamma[c] + beta[c] ; y = act(v) (+ residual)
//
// HBM-bound.  One CTA owns (image n, CH consecutive channels, one slice of the HW pixels).  The
// pixel dimension is split over a thread-block CLUSTER of SP CTAs: each CTA keeps its slice of
// the plane in shared memory, the per-channel partial sums are exchanged through distributed
// shared memory (DSMEM) in rank order (deterministic), and the slice is normalised from shared
// memory -> the plane is read from HBM exactly once and written once (2*4 bytes / element fwd,
// 3*4 bytes / element bwd: the algorithmic minimum).  When a slice cannot fit in shared memory
// the same kernel re-reads global memory (CACHED=false).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace srgan {

constexpr int kNormThreads = 256;
constexpr size_t kNormSmemMax = 200 * 1024;      // leave room for static smem
constexpr size_t kNormSmemTwoCtas = 100 * 1024;  // slice size that still lets 2 CTAs share an SM

struct NormP {
  int N, HW, C;
  int slice;      // pixels per CTA
  int SP;         // cluster size along z
  float eps, slope;
  int act;
};

// Sum `v` (one float4 per thread: 4 channels) over all threads with the same channel group,
// then over the cluster.  Result (per channel of this thread's group) is returned to every thread.
template <int CH>
__device__ __forceinline__ float4 chunk_allreduce(float4 v, float4* part /*[256]*/, float* xchg /*[CH]*/,
                                                  float* tot /*[CH]*/, int SP) {
  constexpr int TPR = CH / 4;
  const int tid = threadIdx.x;
  part[tid] = v;
  __syncthreads();
  if (tid < TPR) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = tid; r < kNormThreads; r += TPR) {
      float4 p = part[r];
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    reinterpret_cast<float4*>(xchg)[tid] = s;
  }
  if (SP > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                       // xchg of every CTA is written
    if (tid < CH) {
      float s = 0.f;
      for (int r = 0; r < SP; ++r) s += *cluster.map_shared_rank(xchg + tid, r);
      tot[tid] = s;
    }
    cluster.sync();                       // nobody still reads my xchg; tot visible
  } else {
    __syncthreads();
    if (tid < CH) tot[tid] = xchg[tid];
    __syncthreads();
  }
  return reinterpret_cast<float4*>(tot)[tid % TPR];
}

template <int CH, bool CACHED>
__global__ void __launch_bounds__(kNormThreads) inorm_fwd_kernel(
    NormP p, const float* __restrict__ x, float* __restrict__ y, float* __restrict__ mean_out,
    float* __restrict__ rstd_out, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ cbias, const float* __restrict__ residual) {
  constexpr int TPR = CH / 4;                 // threads per pixel row
  constexpr int RPP = kNormThreads / TPR;     // pixel rows per pass
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float4 part[kNormThreads];
  __shared__ __align__(16) float xchg[CH];
  __shared__ __align__(16) float tot[CH];
  float4* cache = reinterpret_cast<float4*>(smem_raw);

  const int tid = threadIdx.x;
  const int cg4 = tid % TPR;                  // which float4 of the chunk
  const int row = tid / TPR;
  const int n = blockIdx.y;
  const int c0 = blockIdx.x * CH;
  const int px0 = blockIdx.z * p.slice;
  const int px1 = min(p.HW, px0 + p.slice);
  const int npx = max(0, px1 - px0);
  const float4* xg = reinterpret_cast<const float4*>(x + ((size_t)n * p.HW + px0) * p.C + c0) + cg4;
  const int rs4 = p.C / 4;                    // float4 stride between pixels

  // pass 0: sum
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = row; r < npx; r += RPP) {
    float4 v = __ldg(xg + (size_t)r * rs4);
    if (CACHED) cache[r * TPR + cg4] = v;
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float4 t4 = chunk_allreduce<CH>(s, part, xchg, tot, p.SP);
  const float inv = 1.f / (float)p.HW;
  const float4 mu = make_float4(t4.x * inv, t4.y * inv, t4.z * inv, t4.w * inv);

  // pass 1: centred second moment
  s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = row; r < npx; r += RPP) {
    float4 v = CACHED ? cache[r * TPR + cg4] : __ldg(xg + (size_t)r * rs4);
    float a = v.x - mu.x, b = v.y - mu.y, c = v.z - mu.z, d = v.w - mu.w;
    s.x += a * a; s.y += b * b; s.z += c * c; s.w += d * d;
  }
  t4 = chunk_allreduce<CH>(s, part, xchg, tot, p.SP);
  float4 rs;
  rs.x = rsqrtf(t4.x * inv + p.eps); rs.y = rsqrtf(t4.y * inv + p.eps);
  rs.z = rsqrtf(t4.z * inv + p.eps); rs.w = rsqrtf(t4.w * inv + p.eps);

  const int c = c0 + cg4 * 4;
  if (blockIdx.z == 0 && row == 0) {
    reinterpret_cast<float4*>(mean_out + (size_t)n * p.C + c)[0] = mu;
    reinterpret_cast<float4*>(rstd_out + (size_t)n * p.C + c)[0] = rs;
  }
  float4 g = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
  float4 b = beta ? __ldg(reinterpret_cast<const float4*>(beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 tb = cbias ? __ldg(reinterpret_cast<const float4*>(cbias + (size_t)n * p.C + c))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
  // pass 2: normalise + conditional bias + affine + activation (+ residual)
  float4* yg = reinterpret_cast<float4*>(y + ((size_t)n * p.HW + px0) * p.C + c0) + cg4;
  const float4* rg = residual
      ? reinterpret_cast<const float4*>(residual + ((size_t)n * p.HW + px0) * p.C + c0) + cg4 : nullptr;
  for (int r = row; r < npx; r += RPP) {
    float4 v = CACHED ? cache[r * TPR + cg4] : __ldg(xg + (size_t)r * rs4);
    float4 o;
    o.x = apply_act(((v.x - mu.x) * rs.x + tb.x) * g.x + b.x, p.act, p.slope);
    o.y = apply_act(((v.y - mu.y) * rs.y + tb.y) * g.y + b.y, p.act, p.slope);
    o.z = apply_act(((v.z - mu.z) * rs.z + tb.z) * g.z + b.z, p.act, p.slope);
    o.w = apply_act(((v.w - mu.w) * rs.w + tb.w) * g.w + b.w, p.act, p.slope);
    if (rg) {
      float4 q = __ldg(rg + (size_t)r * rs4);
      o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
    }
    yg[(size_t)r * rs4] = o;
  }
}

template <int CH, bool CACHED>
__global__ void __launch_bounds__(kNormThreads) inorm_bwd_kernel(
    NormP p, const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ cbias, float* __restrict__ dx, float* __restrict__ s1_out,
    float* __restrict__ s2_out) {
  constexpr int TPR = CH / 4;
  constexpr int RPP = kNormThreads / TPR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float4 part[kNormThreads];
  __shared__ __align__(16) float xchg[CH];
  __shared__ __align__(16) float tot[CH];
  float4* cache_dv = reinterpret_cast<float4*>(smem_raw);
  float4* cache_xh = cache_dv + (size_t)p.slice * TPR;

  const int tid = threadIdx.x;
  const int cg4 = tid % TPR;
  const int row = tid / TPR;
  const int n = blockIdx.y;
  const int c0 = blockIdx.x * CH;
  const int c = c0 + cg4 * 4;
  const int px0 = blockIdx.z * p.slice;
  const int px1 = min(p.HW, px0 + p.slice);
  const int npx = max(0, px1 - px0);
  const int rs4 = p.C / 4;
  const size_t base = ((size_t)n * p.HW + px0) * p.C + c0;
  const float4* xg = reinterpret_cast<const float4*>(x + base) + cg4;
  const float4* dg = reinterpret_cast<const float4*>(dy + base) + cg4;
  float4* og = reinterpret_cast<float4*>(dx + base) + cg4;

  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + (size_t)n * p.C + c));
  const float4 rs = __ldg(reinterpret_cast<const float4*>(rstd + (size_t)n * p.C + c));
  const float4 g = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 b = beta ? __ldg(reinterpret_cast<const float4*>(beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 tb = cbias ? __ldg(reinterpret_cast<const float4*>(cbias + (size_t)n * p.C + c))
                          : make_float4(0.f, 0.f, 0.f, 0.f);

  auto dv_xh = [&](float4 xv, float4 dv_in, float4& xh, float4& dv) {
    xh.x = (xv.x - mu.x) * rs.x; xh.y = (xv.y - mu.y) * rs.y;
    xh.z = (xv.z - mu.z) * rs.z; xh.w = (xv.w - mu.w) * rs.w;
    dv.x = dv_in.x * act_grad_pre((xh.x + tb.x) * g.x + b.x, p.act, p.slope);
    dv.y = dv_in.y * act_grad_pre((xh.y + tb.y) * g.y + b.y, p.act, p.slope);
    dv.z = dv_in.z * act_grad_pre((xh.z + tb.z) * g.z + b.z, p.act, p.slope);
    dv.w = dv_in.w * act_grad_pre((xh.w + tb.w) * g.w + b.w, p.act, p.slope);
  };

  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = row; r < npx; r += RPP) {
    float4 xh, dv;
    dv_xh(__ldg(xg + (size_t)r * rs4), __ldg(dg + (size_t)r * rs4), xh, dv);
    if (CACHED) { cache_dv[r * TPR + cg4] = dv; cache_xh[r * TPR + cg4] = xh; }
    a1.x += dv.x; a1.y += dv.y; a1.z += dv.z; a1.w += dv.w;
    a2.x += dv.x * xh.x; a2.y += dv.y * xh.y; a2.z += dv.z * xh.z; a2.w += dv.w * xh.w;
  }
  const float4 S1 = chunk_allreduce<CH>(a1, part, xchg, tot, p.SP);
  const float4 S2 = chunk_allreduce<CH>(a2, part, xchg, tot, p.SP);
  if (blockIdx.z == 0 && row == 0) {
    reinterpret_cast<float4*>(s1_out + (size_t)n * p.C + c)[0] = S1;
    reinterpret_cast<float4*>(s2_out + (size_t)n * p.C + c)[0] = S2;
  }
  const float inv = 1.f / (float)p.HW;
  const float4 k = make_float4(rs.x * g.x, rs.y * g.y, rs.z * g.z, rs.w * g.w);
  const float4 m1 = make_float4(S1.x * inv, S1.y * inv, S1.z * inv, S1.w * inv);
  const float4 m2 = make_float4(S2.x * inv, S2.y * inv, S2.z * inv, S2.w * inv);
  for (int r = row; r < npx; r += RPP) {
    float4 xh, dv;
    if (CACHED) { dv = cache_dv[r * TPR + cg4]; xh = cache_xh[r * TPR + cg4]; }
    else dv_xh(__ldg(xg + (size_t)r * rs4), __ldg(dg + (size_t)r * rs4), xh, dv);
    float4 o;
    o.x = k.x * (dv.x - m1.x - xh.x * m2.x);
    o.y = k.y * (dv.y - m1.y - xh.y * m2.y);
    o.z = k.z * (dv.z - m1.z - xh.z * m2.z);
    o.w = k.w * (dv.w - m1.w - xh.w * m2.w);
    og[(size_t)r * rs4] = o;
  }
}

// dgamma[c] = sum_n (s2 + cbias*s1) ; dbeta[c] = sum_n s1 ; dcbias[n][c] = gamma[c]*s1[n][c]
__global__ void inorm_param_grads_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                         const float* __restrict__ gamma, const float* __restrict__ cbias,
                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                         float* __restrict__ dcbias, int N, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float g = gamma ? gamma[c] : 1.f;
  float dg = 0.f, db = 0.f;
  for (int n = 0; n < N; ++n) {
    float a = s1[(size_t)n * C + c], b = s2[(size_t)n * C + c];
    float t = cbias ? cbias[(size_t)n * C + c] : 0.f;
    dg += b + t * a;
    db += a;
    if (dcbias) dcbias[(size_t)n * C + c] = g * a;
  }
  if (dgamma) dgamma[c] = dg;
  if (dbeta) dbeta[c] = db;
}

struct NormPlan { int CH, SP, slice; bool cached; size_t smem; };

// arrays = number of float planes kept in shared memory per pixel-channel (1 fwd, 2 bwd)
static NormPlan plan_norm(int N, int HW, int C, int arrays) {
  const int chs[3] = {32, 16, 8};
  NormPlan best{0, 1, HW, false, 0};
  for (int ci = 0; ci < 3; ++ci) {
    int CH = chs[ci];
    if (C % CH) continue;
    if (!best.CH) best.CH = CH;
    for (int SP = 1; SP <= 8; SP *= 2) {
      int slice = ceil_div(HW, SP);
      size_t bytes = (size_t)slice * CH * 4 * arrays;
      long long ctas = (long long)N * (C / CH) * SP;
      bool fits = bytes <= kNormSmemMax;
      bool enough = ctas >= 2 * kNumSMs || SP == 8 || slice <= 256;
      if (fits && (bytes <= kNormSmemTwoCtas || SP == 8) && enough) {
        return NormPlan{CH, SP, slice, true, bytes};
      }
      if (fits && SP == 8) return NormPlan{CH, SP, slice, true, bytes};
    }
  }
  // streaming fallback: widest chunk, split for parallelism only
  int CH = best.CH;
  int SP = 1;
  while (SP < 8 && (long long)N * (C / CH) * SP < 2 * kNumSMs && HW / (SP * 2) >= 256) SP *= 2;
  return NormPlan{CH, SP, ceil_div(HW, SP), false, 0};
}

template <class K, class... Args>
static int launch_cluster(K kernel, dim3 grid, int SP, size_t smem, cudaStream_t st, Args... args) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNormSmemMax);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kNormThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = SP;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) { set_error("inorm launch: %s", cudaGetErrorString(e)); return (int)e; }
  return SRGAN_OK;
}

template <int CH>
static int inorm_fwd_ch(const NormPlan& pl, NormP p, cudaStream_t st, const float* x, float* y, float* mean,
                        float* rstd, const float* gamma, const float* beta, const float* cbias,
                        const float* residual) {
  dim3 grid(p.C / CH, p.N, pl.SP);
  if (pl.cached)
    return launch_cluster(inorm_fwd_kernel<CH, true>, grid, pl.SP, pl.smem, st, p, x, y, mean, rstd, gamma, beta,
                          cbias, residual);
  return launch_cluster(inorm_fwd_kernel<CH, false>, grid, pl.SP, (size_t)0, st, p, x, y, mean, rstd, gamma, beta,
                        cbias, residual);
}

template <int CH>
static int inorm_bwd_ch(const NormPlan& pl, NormP p, cudaStream_t st, const float* dy, const float* x,
                        const float* mean, const float* rstd, const float* gamma, const float* beta,
                        const float* cbias, float* dx, float* s1, float* s2) {
  dim3 grid(p.C / CH, p.N, pl.SP);
  if (pl.cached)
    return launch_cluster(inorm_bwd_kernel<CH, true>, grid, pl.SP, pl.smem, st, p, dy, x, mean, rstd, gamma, beta,
                          cbias, dx, s1, s2);
  return launch_cluster(inorm_bwd_kernel<CH, false>, grid, pl.SP, (size_t)0, st, p, dy, x, mean, rstd, gamma,
                        beta, cbias, dx, s1, s2);
}

}  // namespace srgan

using namespace srgan;

extern "C" int srgan_inorm_fwd(const float* x, float* y, float* mean, float* rstd, const float* gamma,
                               const float* beta, const float* cbias, const float* residual, int N, int HW,
                               int C, float eps, int act, float slope, void* stream) {
  SRGAN_CHECK_ARG(x && y && mean && rstd, "null pointer");
  SRGAN_CHECK_ARG(N >= 0 && HW > 0 && C > 0 && C % 8 == 0, "need C % 8 == 0, HW > 0");
  SRGAN_CHECK_ARG(((uintptr_t)x | (uintptr_t)y | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma |
                   (uintptr_t)beta | (uintptr_t)cbias | (uintptr_t)residual) % 16 == 0, "pointers must be 16-byte aligned");
  SRGAN_CHECK_ARG(N <= 65535, "N too large for grid.y");
  if (N == 0) return SRGAN_OK;
  NormPlan pl = plan_norm(N, HW, C, 1);
  NormP p{N, HW, C, pl.slice, pl.SP, eps, slope, act};
  cudaStream_t st = (cudaStream_t)stream;
  switch (pl.CH) {
    case 32: return inorm_fwd_ch<32>(pl, p, st, x, y, mean, rstd, gamma, beta, cbias, residual);
    case 16: return inorm_fwd_ch<16>(pl, p, st, x, y, mean, rstd, gamma, beta, cbias, residual);
    default: return inorm_fwd_ch<8>(pl, p, st, x, y, mean, rstd, gamma, beta, cbias, residual);
  }
}

extern "C" int srgan_inorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                               const float* gamma, const float* beta, const float* cbias, float* dx, float* s1,
                               float* s2, int N, int HW, int C, int act, float slope, void* stream) {
  SRGAN_CHECK_ARG(dy && x && mean && rstd && dx && s1 && s2, "null pointer");
  SRGAN_CHECK_ARG(N >= 0 && HW > 0 && C > 0 && C % 8 == 0, "need C % 8 == 0, HW > 0");
  SRGAN_CHECK_ARG(((uintptr_t)dy | (uintptr_t)x | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma |
                   (uintptr_t)beta | (uintptr_t)cbias | (uintptr_t)dx | (uintptr_t)s1 | (uintptr_t)s2) % 16 == 0,
                  "pointers must be 16-byte aligned");
  SRGAN_CHECK_ARG(N <= 65535, "N too large for grid.y");
  if (N == 0) return SRGAN_OK;
  NormPlan pl = plan_norm(N, HW, C, 2);
  NormP p{N, HW, C, pl.slice, pl.SP, 0.f, slope, act};
  cudaStream_t st = (cudaStream_t)stream;
  switch (pl.CH) {
    case 32: return inorm_bwd_ch<32>(pl, p, st, dy, x, mean, rstd, gamma, beta, cbias, dx, s1, s2);
    case 16: return inorm_bwd_ch<16>(pl, p, st, dy, x, mean, rstd, gamma, beta, cbias, dx, s1, s2);
    default: return inorm_bwd_ch<8>(pl, p, st, dy, x, mean, rstd, gamma, beta, cbias, dx, s1, s2);
  }
}

extern "C" int srgan_inorm_param_grads(const float* s1, const float* s2, const float* gamma,
                                       const float* cbias, float* dgamma, float* dbeta, float* dcbias, int N,
                                       int C, void* stream) {
  SRGAN_CHECK_ARG(s1 && s2, "null pointer");
  if (C == 0) return SRGAN_OK;
  inorm_param_grads_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(s1, s2, gamma, cbias, dgamma,
                                                                                  dbeta, dcbias, N, C);
  SRGAN_RETURN_LAUNCH();
}
