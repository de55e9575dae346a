// Loss kernels: full-tensor reductions (L1 / MSE) and the SRGAN latent batch losses
// (batch-KL, correlation, soft-histogram imitation, conventional KL).
//
// Reductions are two-level with a fixed combination order: per-thread strided partials ->
// warp shuffle -> per-block partial in scratch -> the LAST block to finish (atomic ticket) adds
// the per-block partials in index order.  Bit-reproducible for a given (n, grid).
#include <math.h>
#include "common.cuh"

namespace srgan {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 4 * kNumSMs;   // 592
// scratch layout: float partial[kRedMaxBlocks]; unsigned ticket (must start at 0; self-resetting)
constexpr size_t kRedScratchBytes = (kRedMaxBlocks + 4) * sizeof(float);

static inline unsigned red_grid(size_t n4) {
  size_t b = (n4 + kRedThreads * 4 - 1) / (kRedThreads * 4);
  if (b > (size_t)kRedMaxBlocks) b = kRedMaxBlocks;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// mode 0: |a-b| ; 1: (a-target)^2 ; 2: (a-b)^2
template <int MODE>
__global__ void __launch_bounds__(kRedThreads) reduce_mean_kernel(const float* __restrict__ a,
                                                                  const float* __restrict__ b, float target,
                                                                  size_t n, float* __restrict__ out,
                                                                  float* __restrict__ scratch) {
  __shared__ float red[32];
  __shared__ bool last;
  auto f = [&](float x, float y) -> float {
    if (MODE == 0) return fabsf(x - y);
    float d = x - y;
    return d * d;
  };
  float s = 0.f;
  const size_t n4 = n / 4;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 p = __ldg(a4 + i);
    float4 q = MODE == 1 ? make_float4(target, target, target, target) : __ldg(b4 + i);
    s += (f(p.x, q.x) + f(p.y, q.y)) + (f(p.z, q.z) + f(p.w, q.w));
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += f(a[i], MODE == 1 ? target : b[i]);
  s = block_sum(s, red);
  unsigned* ticket = reinterpret_cast<unsigned*>(scratch + kRedMaxBlocks);
  if (threadIdx.x == 0) {
    scratch[blockIdx.x] = s;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float t = 0.f;
    // fixed order: thread j sums partials j, j+256, ...; then block_sum
    for (unsigned j = threadIdx.x; j < gridDim.x; j += blockDim.x) t += __ldcg(scratch + j);
    t = block_sum(t, red);
    if (threadIdx.x == 0) {
      out[0] = t / (float)n;
      *ticket = 0u;
    }
  }
}

__global__ void l1_mean_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   const float* __restrict__ g, float* __restrict__ da, float* __restrict__ db,
                                   size_t n) {
  const float s = __ldg(g) / (float)n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float d = a[i] - b[i];
    float v = d > 0.f ? s : (d < 0.f ? -s : 0.f);
    if (da) da[i] = v;
    if (db) db[i] = -v;
  }
}
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float target,
                               const float* __restrict__ g, float* __restrict__ da, float* __restrict__ db,
                               size_t n) {
  const float s = 2.f * __ldg(g) / (float)n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = s * (a[i] - (b ? b[i] : target));
    if (da) da[i] = v;
    if (db) db[i] = -v;
  }
}

// =====================================================================================
// Latent batch losses.  mu[n][D], D <= 16, bins <= 64.  One CTA of 512 threads.
// out layout (floats): [0..4) losses | mean[D] | var[D] | cdiag[D] | corr[D*D] | hist[D*bins] | hsum[D]
// =====================================================================================
constexpr int kLatThreads = 512;
constexpr int kLatMaxD = 16;
constexpr int kLatMaxBins = 64;

struct LatOff {
  int mean, var, cdiag, corr, hist, hsum, total;
  __host__ __device__ LatOff(int D, int bins) {
    mean = 4; var = mean + D; cdiag = var + D; corr = cdiag + D; hist = corr + D * D; hsum = hist + D * bins;
    total = hsum + D;
  }
};

__device__ __forceinline__ float hist_kernel_val(float x, float center, float sigma, float denom, float delta) {
  float u = (x - center) / sigma;
  return expf(-0.5f * (u * u)) / denom * delta;
}

__global__ void __launch_bounds__(kLatThreads) latent_fwd_kernel(
    const float* __restrict__ mu, const float* __restrict__ logvar, int n, int D, float n_cfg,
    const float* __restrict__ target, int bins, float hmin, float hmax, float sigma, int flags,
    float* __restrict__ out) {
  __shared__ float part[kLatThreads];
  __shared__ float red[32];
  __shared__ float s_mean[kLatMaxD], s_cov[kLatMaxD * kLatMaxD], s_hist[kLatMaxD * kLatMaxBins];
  const int tid = threadIdx.x;
  const LatOff o(D, bins);
  const int DD = D * D;

  // ---- mean_d : thread t -> column t % D, row group t / D
  {
    const int G = kLatThreads / D;
    const int d = tid % D, gq = tid / D;
    float s = 0.f;
    if (gq < G)
      for (int i = gq; i < n; i += G) s += mu[(size_t)i * D + d];
    part[tid] = s;
    __syncthreads();
    if (tid < D) {
      float t = 0.f;
      for (int j = 0; j < G; ++j) t += part[j * D + tid];
      s_mean[tid] = t / (float)n;
    }
    __syncthreads();
  }
  // ---- centred cross products cov_de = sum_i (x_id-m_d)(x_ie-m_e)
  {
    const int G = kLatThreads / DD;     // D<=16 -> DD<=256 -> G>=2
    const int pr = tid % DD, gq = tid / DD;
    const int d = pr / D, e = pr % D;
    const float md = s_mean[d], me = s_mean[e];
    float s = 0.f;
    if (gq < G)
      for (int i = gq; i < n; i += G) s = fmaf(mu[(size_t)i * D + d] - md, mu[(size_t)i * D + e] - me, s);
    part[tid] = s;
    __syncthreads();
    if (tid < DD) {
      float t = 0.f;
      for (int j = 0; j < G; ++j) t += part[j * DD + tid];
      s_cov[tid] = t;
    }
    __syncthreads();
  }
  const float nm1 = (float)(n - 1);
  float l_bkl = 0.f, l_corr = 0.f, l_hist = 0.f, l_kl = 0.f;
  if (tid < D) {
    out[o.mean + tid] = s_mean[tid];
    float cd = s_cov[tid * D + tid] / nm1;
    out[o.cdiag + tid] = cd;
    out[o.var + tid] = cd * n_cfg / (n_cfg - 1.f);
  }
  // ---- batch KL (thread 0, D terms in order)
  if ((flags & SRGAN_LATENT_BKL) && tid == 0) {
    float s = 0.f;
    for (int d = 0; d < D; ++d) {
      float var = s_cov[d * D + d] / nm1 * n_cfg / (n_cfg - 1.f);
      float m = s_mean[d];
      s += 1.f + logf(var) - m * m - var;
    }
    l_bkl = -0.5f * s;
  }
  // ---- correlation matrix + loss
  if (flags & SRGAN_LATENT_CORR) {
    if (tid < DD) {
      int d = tid / D, e = tid % D;
      float c = s_cov[tid] / nm1;
      float sd = sqrtf(s_cov[d * D + d] / nm1), se = sqrtf(s_cov[e * D + e] / nm1);
      float r = c / se / sd;
      r = fminf(fmaxf(r, -1.f), 1.f);
      out[o.corr + tid] = r;
      part[tid] = fabsf(r - (d == e ? 1.f : 0.f));
    }
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int j = 0; j < DD; ++j) s += part[j];
      l_corr = s / (float)(D * (D - 1));
    }
    __syncthreads();
  }
  // ---- soft histogram h[d][b] and KL(target || p)
  if (flags & SRGAN_LATENT_HIST) {
    const float delta = (hmax - hmin) / (float)bins;
    const float denom = (float)((double)sigma * sqrt(2.0 * M_PI));
    const int DB = D * bins;
    // warp w handles entries w, w+16, ...; lanes stride samples; fixed shuffle order
    const int lane = tid & 31, wid = tid >> 5, nw = kLatThreads / 32;
    for (int e = wid; e < DB; e += nw) {
      int d = e / bins, b = e % bins;
      float center = hmin + delta * ((float)b + 0.5f);
      float s = 0.f;
      for (int i = lane; i < n; i += 32) s += hist_kernel_val(mu[(size_t)i * D + d], center, sigma, denom, delta);
      s = warp_sum(s);
      if (lane == 0) { s_hist[e] = s; out[o.hist + e] = s; }
    }
    __syncthreads();
    if (tid < D) {
      float S = 0.f;
      for (int b = 0; b < bins; ++b) S += s_hist[tid * bins + b];
      out[o.hsum + tid] = S;
      float kl = 0.f;
      if (target)
        for (int b = 0; b < bins; ++b) {
          float p = s_hist[tid * bins + b] / S + 1e-8f;
          float t = target[b];
          kl += t * (logf(t) - logf(p));
        }
      part[tid] = kl;
    }
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int d = 0; d < D; ++d) s += part[d];
      l_hist = s;
    }
    __syncthreads();
  }
  // ---- conventional KL: -0.5 * sum(1 + lv - mu^2 - exp(lv))
  if (flags & SRGAN_LATENT_KL) {
    float s = 0.f;
    for (int i = tid; i < n * D; i += kLatThreads) {
      float m = mu[i], lv = logvar[i];
      s += 1.f + lv - m * m - expf(lv);
    }
    s = block_sum(s, red);
    l_kl = -0.5f * s;
  }
  if (tid == 0) {
    out[0] = l_bkl; out[1] = l_corr; out[2] = l_hist; out[3] = l_kl;
  }
}

// Per-block prologue: Gamma = dL/dcov-matrix from an upstream gradient on the (clamped) corrcoef.
//   Hm[d][e]   upstream gradient wrt corr entry (already masked)
//   gam[d][e]  off-diagonal: Hm/(sd*se) ; diagonal: -0.5*sum_{e!=d}(Hm_de*C_de + Hm_ed*C_ed)/c_dd
__device__ void corr_gamma(const float* Hm, const float* C, const float* cdiag, int D, float* gam) {
  for (int t = threadIdx.x; t < D * D; t += blockDim.x) {
    int d = t / D, e = t % D;
    if (d != e) {
      gam[t] = Hm[t] / (sqrtf(cdiag[d]) * sqrtf(cdiag[e]));
    } else {
      float s = 0.f;
      for (int k = 0; k < D; ++k)
        if (k != d) s += Hm[d * D + k] * C[d * D + k] + Hm[k * D + d] * C[k * D + d];
      gam[t] = -0.5f * s / cdiag[d];
    }
  }
}

struct LatW { float w[4]; };

__global__ void __launch_bounds__(256) latent_bwd_kernel(
    const float* __restrict__ mu, const float* __restrict__ logvar, int n, int D, float n_cfg,
    const float* __restrict__ target, int bins, float hmin, float hmax, float sigma, int flags,
    const float* __restrict__ stats, const float* __restrict__ g4 /* device upstream grads or NULL (=> corr only) */,
    const float* __restrict__ dcorr_ext /* optional explicit upstream on corr */,
    float* __restrict__ dmu, float* __restrict__ dlogvar, int row0, int rows) {
  __shared__ float s_mean[kLatMaxD], s_var[kLatMaxD], s_cd[kLatMaxD];
  __shared__ float s_C[kLatMaxD * kLatMaxD], s_H[kLatMaxD * kLatMaxD], s_gam[kLatMaxD * kLatMaxD];
  __shared__ float s_g[kLatMaxD * kLatMaxBins];
  const LatOff o(D, bins);
  const int tid = threadIdx.x;
  const int DD = D * D;
  LatW lw;
  for (int k = 0; k < 4; ++k) lw.w[k] = g4 ? __ldg(g4 + k) : (k == 1 ? 1.f : 0.f);
  for (int t = tid; t < D; t += blockDim.x) {
    s_mean[t] = stats[o.mean + t]; s_var[t] = stats[o.var + t]; s_cd[t] = stats[o.cdiag + t];
  }
  const bool do_corr = (flags & SRGAN_LATENT_CORR) != 0;
  if (do_corr) {
    for (int t = tid; t < DD; t += blockDim.x) {
      int d = t / D, e = t % D;
      float c = stats[o.corr + t];
      s_C[t] = c;
      float h;
      if (dcorr_ext) {
        h = dcorr_ext[t];
      } else {
        float df = c - (d == e ? 1.f : 0.f);
        h = (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)) / (float)(D * (D - 1));
      }
      // clamp passes gradient only strictly inside (-1,1); the diagonal is identically 1
      if (d == e || fabsf(c) >= 1.f) h = 0.f;
      s_H[t] = h;
    }
  }
  __syncthreads();
  if (do_corr) corr_gamma(s_H, s_C, s_cd, D, s_gam);
  const bool do_hist = (flags & SRGAN_LATENT_HIST) != 0;
  if (do_hist) {
    // g[d][b] = dL/dh_db = -t_b/(p_b S) + sum_b' t_b' h_b' / (p_b' S^2)
    for (int d = tid; d < D; d += blockDim.x) {
      float S = stats[o.hsum + d];
      float acc = 0.f;
      for (int b = 0; b < bins; ++b) {
        float h = stats[o.hist + d * bins + b];
        float p = h / S + 1e-8f;
        acc += target[b] * h / (p * S * S);
      }
      for (int b = 0; b < bins; ++b) {
        float h = stats[o.hist + d * bins + b];
        float p = h / S + 1e-8f;
        s_g[d * bins + b] = -target[b] / (p * S) + acc;
      }
    }
  }
  __syncthreads();
  const float nm1 = (float)(n - 1);
  const float a = n_cfg / (nm1 * (n_cfg - 1.f));
  const float delta = (hmax - hmin) / (float)bins;
  const float denom = (float)((double)sigma * sqrt(2.0 * M_PI));
  const float inv_s2 = 1.f / (sigma * sigma);
  const int total = rows * D;
  for (int idx = blockIdx.x * blockDim.x + tid; idx < total; idx += gridDim.x * blockDim.x) {
    const int i = row0 + idx / D, d = idx % D;
    const float x = mu[(size_t)i * D + d];
    const float xm = x - s_mean[d];
    float g = 0.f;
    if (flags & SRGAN_LATENT_BKL) g += lw.w[0] * (-(1.f / s_var[d] - 1.f) * a * xm + s_mean[d] / (float)n);
    if (do_corr) {
      float s = 0.f;
      for (int e = 0; e < D; ++e)
        s += (s_gam[d * D + e] + s_gam[e * D + d]) * (mu[(size_t)i * D + e] - s_mean[e]);
      g += lw.w[1] * s / nm1;
    }
    if (do_hist) {
      float s = 0.f;
      for (int b = 0; b < bins; ++b) {
        float center = hmin + delta * ((float)b + 0.5f);
        float u = x - center;
        s += s_g[d * bins + b] * hist_kernel_val(x, center, sigma, denom, delta) * (-u * inv_s2);
      }
      g += lw.w[2] * s;
    }
    if (flags & SRGAN_LATENT_KL) {
      g += lw.w[3] * x;
      if (dlogvar) dlogvar[(size_t)idx] = lw.w[3] * (-0.5f) * (1.f - expf(logvar[(size_t)i * D + d]));
    }
    dmu[(size_t)idx] = g;
  }
}

// stand-alone soft histogram of a vector x[n]: one block per bin
__global__ void softhist_fwd_kernel(const float* __restrict__ x, int n, int bins, float hmin, float hmax,
                                    float sigma, float* __restrict__ h) {
  __shared__ float red[32];
  const float delta = (hmax - hmin) / (float)bins;
  const float denom = (float)((double)sigma * sqrt(2.0 * M_PI));
  const float center = hmin + delta * ((float)blockIdx.x + 0.5f);
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += hist_kernel_val(x[i], center, sigma, denom, delta);
  s = block_sum(s, red);
  if (threadIdx.x == 0) h[blockIdx.x] = s;
}
__global__ void softhist_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dh, int n, int bins,
                                    float hmin, float hmax, float sigma, float* __restrict__ dx) {
  const float delta = (hmax - hmin) / (float)bins;
  const float denom = (float)((double)sigma * sqrt(2.0 * M_PI));
  const float inv_s2 = 1.f / (sigma * sigma);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float xv = x[i], s = 0.f;
    for (int b = 0; b < bins; ++b) {
      float center = hmin + delta * ((float)b + 0.5f);
      s += __ldg(dh + b) * hist_kernel_val(xv, center, sigma, denom, delta) * (-(xv - center) * inv_s2);
    }
    dx[i] = s;
  }
}

// One launch over a flat buffer; torch.optim.Adam semantics (no amsgrad, no weight decay).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt) {
  const float step_size = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float wgt = 1.f - b1;                          // torch lerp: two-sided form
    float mi = wgt < 0.5f ? m[i] + wgt * (gi - m[i]) : gi - (gi - m[i]) * (1.f - wgt);
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float den = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / den);
  }
}

// Same update with the step-dependent scalars read from device memory ([lr, beta1, beta2, eps, 1 - beta1^t,
// sqrt(1 - beta2^t)]): the launch can then live in a CUDA graph that is replayed every step.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, size_t n, const float* __restrict__ hyper) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], bc1 = hyper[4], bc2_sqrt = hyper[5];
  const float step_size = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float wgt = 1.f - b1;
    float mi = wgt < 0.5f ? m[i] + wgt * (gi - m[i]) : gi - (gi - m[i]) * (1.f - wgt);
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float den = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / den);
  }
}

}  // namespace srgan

using namespace srgan;
#define ST ((cudaStream_t)stream)

extern "C" size_t srgan_reduce_scratch_bytes(size_t) { return kRedScratchBytes; }

static int check_red(const void* a, const void* out, const void* scratch) {
  if (!a || !out || !scratch) { set_error("reduce: null pointer"); return SRGAN_E_BADARG; }
  if ((uintptr_t)a % 16) { set_error("reduce: input must be 16-byte aligned"); return SRGAN_E_BADARG; }
  return SRGAN_OK;
}
extern "C" int srgan_l1_mean_fwd(const float* a, const float* b, size_t n, float* out, void* scratch, void* stream) {
  if (int e = check_red(a, out, scratch)) return e;
  SRGAN_CHECK_ARG(b && (uintptr_t)b % 16 == 0 && n > 0, "bad argument");
  reduce_mean_kernel<0><<<red_grid(n / 4), kRedThreads, 0, ST>>>(a, b, 0.f, n, out, (float*)scratch);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_mse_const_fwd(const float* x, float target, size_t n, float* out, void* scratch, void* stream) {
  if (int e = check_red(x, out, scratch)) return e;
  SRGAN_CHECK_ARG(n > 0, "empty input");
  reduce_mean_kernel<1><<<red_grid(n / 4), kRedThreads, 0, ST>>>(x, x, target, n, out, (float*)scratch);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_mse_fwd(const float* a, const float* b, size_t n, float* out, void* scratch, void* stream) {
  if (int e = check_red(a, out, scratch)) return e;
  SRGAN_CHECK_ARG(b && (uintptr_t)b % 16 == 0 && n > 0, "bad argument");
  reduce_mean_kernel<2><<<red_grid(n / 4), kRedThreads, 0, ST>>>(a, b, 0.f, n, out, (float*)scratch);
  SRGAN_RETURN_LAUNCH();
}
static unsigned ew_grid(size_t n) {
  size_t b = (n + 255) / 256;
  size_t cap = (size_t)kNumSMs * 16;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}
extern "C" int srgan_l1_mean_bwd(const float* a, const float* b, const float* g, float* da, float* db, size_t n,
                                 void* stream) {
  SRGAN_CHECK_ARG(a && b && g && n > 0, "bad argument");
  l1_mean_bwd_kernel<<<ew_grid(n), 256, 0, ST>>>(a, b, g, da, db, n);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_mse_const_bwd(const float* x, float target, const float* g, float* dx, size_t n, void* stream) {
  SRGAN_CHECK_ARG(x && g && dx && n > 0, "bad argument");
  mse_bwd_kernel<<<ew_grid(n), 256, 0, ST>>>(x, nullptr, target, g, dx, nullptr, n);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_mse_bwd(const float* a, const float* b, const float* g, float* da, float* db, size_t n,
                             void* stream) {
  SRGAN_CHECK_ARG(a && b && g && n > 0, "bad argument");
  mse_bwd_kernel<<<ew_grid(n), 256, 0, ST>>>(a, b, 0.f, g, da, db, n);
  SRGAN_RETURN_LAUNCH();
}

static int check_lat(const float* mu, int n, int D, int bins, int flags, const float* logvar,
                     const float* target) {
  if (!mu || n < 2 || D < 1 || D > kLatMaxD) { set_error("latent: need n >= 2, 1 <= D <= 16"); return SRGAN_E_BADARG; }
  if ((flags & SRGAN_LATENT_HIST) && (bins < 1 || bins > kLatMaxBins)) {
    set_error("latent: need 1 <= bins <= 64"); return SRGAN_E_BADARG;
  }
  if ((flags & SRGAN_LATENT_KL) && !logvar) { set_error("latent: KL needs logvar"); return SRGAN_E_BADARG; }
  (void)target;
  return SRGAN_OK;
}
extern "C" int srgan_latent_losses_fwd(const float* mu, const float* logvar, int n, int D, float n_cfg,
                                       const float* hist_target, int bins, float hist_min, float hist_max,
                                       float sigma, int flags, float* out, void* stream) {
  if (int e = check_lat(mu, n, D, bins, flags, logvar, hist_target)) return e;
  SRGAN_CHECK_ARG(out, "null out");
  latent_fwd_kernel<<<1, kLatThreads, 0, ST>>>(mu, logvar, n, D, n_cfg, hist_target, bins, hist_min, hist_max,
                                               sigma, flags, out);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_latent_losses_bwd(const float* mu, const float* logvar, int n, int D, float n_cfg,
                                       const float* hist_target, int bins, float hist_min, float hist_max,
                                       float sigma, int flags, const float* out, const float* g4,
                                       float* dmu, float* dlogvar, int row0, int rows, void* stream) {
  if (int e = check_lat(mu, n, D, bins, flags, logvar, hist_target)) return e;
  SRGAN_CHECK_ARG(out && g4 && dmu && row0 >= 0 && rows >= 0 && row0 + rows <= n, "bad argument");
  SRGAN_CHECK_ARG(!(flags & SRGAN_LATENT_HIST) || hist_target, "hist needs a target");
  if (rows == 0) return SRGAN_OK;
  int blocks = ceil_div(rows * D, 256);
  if (blocks > kNumSMs) blocks = kNumSMs;
  latent_bwd_kernel<<<blocks, 256, 0, ST>>>(mu, logvar, n, D, n_cfg, hist_target, bins, hist_min, hist_max, sigma,
                                            flags, out, g4, nullptr, dmu, dlogvar, row0, rows);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_corrcoef_bwd(const float* mu, int n, int D, const float* stats, const float* dcorr,
                                  float* dmu, void* stream) {
  if (int e = check_lat(mu, n, D, 1, 0, nullptr, nullptr)) return e;
  SRGAN_CHECK_ARG(stats && dcorr && dmu, "null pointer");
  int blocks = ceil_div(n * D, 256);
  if (blocks > kNumSMs) blocks = kNumSMs;
  // bins is only used for offsets before the corr block, which do not depend on it
  latent_bwd_kernel<<<blocks, 256, 0, ST>>>(mu, nullptr, n, D, 2.f, nullptr, 1, 0.f, 1.f, 1.f, SRGAN_LATENT_CORR,
                                            stats, nullptr, dcorr, dmu, nullptr, 0, n);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_softhist_fwd(const float* x, int n, int bins, float hist_min, float hist_max, float sigma,
                                  float* h, void* stream) {
  SRGAN_CHECK_ARG(x && h && n > 0 && bins > 0, "bad argument");
  softhist_fwd_kernel<<<bins, 256, 0, ST>>>(x, n, bins, hist_min, hist_max, sigma, h);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_softhist_bwd(const float* x, const float* dh, int n, int bins, float hist_min,
                                  float hist_max, float sigma, float* dx, void* stream) {
  SRGAN_CHECK_ARG(x && dh && dx && n > 0 && bins > 0, "bad argument");
  int blocks = ceil_div(n, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  softhist_bwd_kernel<<<blocks, 256, 0, ST>>>(x, dh, n, bins, hist_min, hist_max, sigma, dx);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                               float beta2, float eps, int step, void* stream) {
  SRGAN_CHECK_ARG(p && g && m && v && step >= 1, "bad argument");
  if (n == 0) return SRGAN_OK;
  float bc1 = 1.f - (float)pow((double)beta1, step);
  float bc2 = (float)sqrt(1.0 - pow((double)beta2, step));
  adam_kernel<<<ew_grid(n), 256, 0, ST>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2);
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_adam_step_dev(float* p, const float* g, float* m, float* v, size_t n, const float* hyper,
                                   void* stream) {
  SRGAN_CHECK_ARG(p && g && m && v && hyper, "bad argument");
  if (n == 0) return SRGAN_OK;
  adam_dev_kernel<<<ew_grid(n), 256, 0, ST>>>(p, g, m, v, n, hyper);
  SRGAN_RETURN_LAUNCH();
}
