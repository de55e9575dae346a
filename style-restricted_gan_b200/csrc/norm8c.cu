// Conditional instance normalisation in ONE pass over HBM: thread-block clusters hold an image plane in shared memory.
//
//   xh = (x - mean_hw) * rstd_hw ; v = (xh + cbias[n][c]) * gamma[c] + beta[c] ; y = act(v) (+ residual)
//   ref: CBINorm2d.forward pyfiles/model.py:54-67, nn.InstanceNorm2d(affine=False) :178, the ReLU / add that follow.
//
// The two-kernel path of norm8.cu reads x twice forward (statistics, apply) and dy / x twice backward.  Here a cluster
// of CL <= 8 CTAs owns one image: CTA `rank` pulls its slice of pixel rows (<= 64 KB, contiguous in NHWC) into shared
// memory with a few bulk copies (cp.async.bulk, one mbarrier per 16 / 32 KB chunk: 64 KB in flight per CTA from one
// thread, three CTAs per SM), sums the chunks as they land, publishes its per-channel partial sums in its own shared
// memory, and after ONE cluster barrier every CTA folds the CL partials through distributed shared memory in rank
// order and streams the result out of its resident slice.  Forward: x is read once and y written once - the
// algorithmic traffic of SURVEY 8d.  Backward: dy is resident, x streams through registers twice (the second time an
// L2 hit a few microseconds after the first), dx is written once.
// Determinism / split invariance: the ATOMS of norm8.cu are kept - fp32 sums over rows {q, q+RPP, q+2RPP, q+3RPP}
// (forward) or {q, q+RPP} (backward) of an aligned group of 4*RPP pixels, RPP = 256 / (C/8), everything above in fp64
// in a fixed order - so the statistics equal those of the two-kernel path up to fp64 reassociation (2^-53), and the
// choice between the two paths depends on the plane (HW, C, storage type) only, never on the batch.
#include "norm8.cuh"
#include "umma_ptx.cuh"
#include <stdlib.h>

namespace srgan {

constexpr int kCThreads = 256;
constexpr int kCMaxData = 65536;     // bytes of the resident slice
constexpr int kCRedBytes = 4096;     // CTA reduction scratch, later the per-channel constants
constexpr int kCPartBytes = 4096;    // C x {S1, S2} doubles, read by the other CTAs of the cluster
constexpr int kCMaxChunks = 8;
constexpr int kCMaxCluster = 8;

struct Norm8CP {
  int N, HW, C;
  int TPR, RPP;        // reference mapping (8 channels per thread): defines the atoms; mapping of the output phase
  int TPR4, RPP4;      // 4 channels per thread: mapping of the summation phase
  int chunk_px;        // 4 * RPP pixels = one bulk copy
  int CL, slice;       // cluster size, pixels per CTA (a multiple of chunk_px)
  int nchunks;         // chunks per CTA
  int data_bytes;      // resident bytes per CTA
  float eps, slope, inv_hw;
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ double2 ld_dsmem_f64x2(uint32_t cluster_addr) {
  double2 v;
  asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// ---- 4-channel access (summation phase) and shared-memory 8-channel access
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 a = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(a.x << 16); v[1] = __uint_as_float(a.x & 0xffff0000u);
  v[2] = __uint_as_float(a.y << 16); v[3] = __uint_as_float(a.y & 0xffff0000u);
}
struct Raw4f { float4 a; };
struct Raw4h { uint2 a; };
__device__ __forceinline__ Raw4f ldg4(const float* p) { Raw4f r; r.a = __ldg(reinterpret_cast<const float4*>(p)); return r; }
__device__ __forceinline__ Raw4h ldg4(const __nv_bfloat16* p) { Raw4h r; r.a = __ldg(reinterpret_cast<const uint2*>(p)); return r; }
__device__ __forceinline__ void unpack4(const Raw4f& r, float (&v)[4]) { v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; }
__device__ __forceinline__ void unpack4(const Raw4h& r, float (&v)[4]) {
  v[0] = __uint_as_float(r.a.x << 16); v[1] = __uint_as_float(r.a.x & 0xffff0000u);
  v[2] = __uint_as_float(r.a.y << 16); v[3] = __uint_as_float(r.a.y & 0xffff0000u);
}
template <typename T> struct Raw4Of;
template <> struct Raw4Of<float> { using type = Raw4f; };
template <> struct Raw4Of<__nv_bfloat16> { using type = Raw4h; };

__device__ __forceinline__ Raw8<float> lds_raw(const float* p) {
  Raw8<float> r;
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *(reinterpret_cast<const float4*>(p) + 1);
  return r;
}
__device__ __forceinline__ Raw8<__nv_bfloat16> lds_raw(const __nv_bfloat16* p) {
  Raw8<__nv_bfloat16> r;
  r.q = *reinterpret_cast<const uint4*>(p);
  return r;
}
__device__ __forceinline__ float ldg1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldg1(const __nv_bfloat16* p) {
  return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
}

struct C8Smem {
  unsigned char* data;
  double* red;
  double* part;
  uint64_t* bars;
};
__device__ __forceinline__ C8Smem c8_smem(unsigned char* sm, const Norm8CP& p) {
  C8Smem s;
  s.data = sm;
  s.red = reinterpret_cast<double*>(sm + p.data_bytes);
  s.part = reinterpret_cast<double*>(sm + p.data_bytes + kCRedBytes);
  s.bars = reinterpret_cast<uint64_t*>(sm + p.data_bytes + kCRedBytes + kCPartBytes);
  return s;
}

// thread 0: one bulk copy per chunk of the CTA's slice, each on its own mbarrier
template <typename T>
__device__ __forceinline__ void c8_issue_loads(const Norm8CP& p, const C8Smem& s, const T* src, int rows) {
  for (int c = 0; c < p.nchunks; ++c) mbar_init(&s.bars[c], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  for (int c = 0; c < p.nchunks; ++c) {
    const int cnt = min(rows - c * p.chunk_px, p.chunk_px);
    if (cnt <= 0) break;
    const uint32_t bytes = (uint32_t)cnt * p.C * sizeof(T);
    mbar_expect_tx(&s.bars[c], bytes);
    bulk_g2s(reinterpret_cast<T*>(s.data) + (size_t)c * p.chunk_px * p.C, src + (size_t)c * p.chunk_px * p.C, bytes,
             &s.bars[c]);
  }
}

// the 8 running sums of every thread -> part[c] = {S1, S2} of the CTA's slice, rows added in row order
__device__ __forceinline__ void c8_cta_reduce(const Norm8CP& p, const C8Smem& s, const double (&acc)[8]) {
  const int tid = threadIdx.x;
  double2* red2 = reinterpret_cast<double2*>(s.red);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red2[tid] = make_double2(acc[j], acc[4 + j]);
    __syncthreads();
    if (tid < 2 * p.TPR4) {
      const int which = tid / p.TPR4, cqi = tid - which * p.TPR4;
      double a = 0.;
      for (int r = 0; r < p.RPP4; ++r) a += s.red[(r * p.TPR4 + cqi) * 2 + which];
      s.part[(cqi * 4 + j) * 2 + which] = a;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void c8_fold(const Norm8CP& p, const C8Smem& s, int c, double* S1, double* S2) {
  const uint32_t mine = smem_u32(s.part + c * 2);
  double2 v[kCMaxCluster];
#pragma unroll
  for (int rk = 0; rk < kCMaxCluster; ++rk)        // all remote loads in flight at once
    v[rk] = rk < p.CL ? ld_dsmem_f64x2(mapa_shared(mine, (uint32_t)rk)) : make_double2(0., 0.);
  double a = 0., b = 0.;
#pragma unroll
  for (int rk = 0; rk < kCMaxCluster; ++rk) { a += v[rk].x; b += v[rk].y; }   // rank order (+0.0 beyond CL)
  *S1 = a;
  *S2 = b;
}

// ------------------------------------------------------------------------------------------------ forward
template <typename TX, typename TY, int ACT>
__global__ void __launch_bounds__(kCThreads, 3) inorm8c_fwd_kernel(
    Norm8CP p, const TX* __restrict__ x, TY* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ cbias,
    const TY* __restrict__ residual) {
  extern __shared__ __align__(128) unsigned char sm[];
  const C8Smem s = c8_smem(sm, p);
  const TX* data = reinterpret_cast<const TX*>(s.data);
  const int tid = threadIdx.x, n = blockIdx.y;
  const int rank = (int)cluster_ctarank();
  const int px0 = rank * p.slice;
  const int rows = min(p.HW, px0 + p.slice) - px0;
  const TX* ximg = x + (size_t)n * p.HW * p.C;
  const size_t slice_off = ((size_t)n * p.HW + px0) * p.C;
  if (tid == 0) {
    c8_issue_loads<TX>(p, s, x + slice_off, rows);
    if (residual) bulk_prefetch_l2(residual + slice_off, (uint32_t)rows * p.C * sizeof(TY));
  }
  __syncthreads();
  // per-channel inputs of the fold, requested before the data arrives
  float f_pv = 0.f, f_g = 1.f, f_b = 0.f, f_tb = 0.f;
  if (tid < p.C) {
    f_pv = ldg1(ximg + tid);
    if (gamma) f_g = __ldg(gamma + tid);
    if (beta) f_b = __ldg(beta + tid);
    if (cbias) f_tb = __ldg(cbias + (size_t)n * p.C + tid);
  }
  // ---- sums of (x - pivot), (x - pivot)^2 : 4 channels per thread, two atom columns per chunk
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.;
  {
    const int q4 = tid / p.TPR4, cq = tid - q4 * p.TPR4;
    if (q4 < p.RPP4) {
      float pv[4];
      unpack4(ldg4(ximg + cq * 4), pv);                   // pivot: first pixel of the plane
      for (int c = 0; c < p.nchunks; ++c) {
        const int base = c * p.chunk_px;
        if (base >= rows) break;
        mbar_wait(&s.bars[c], 0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r0 = base + q4 + h * p.RPP4;
          float t1[4], t2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) { t1[e] = 0.f; t2[e] = 0.f; }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = r0 + j * p.RPP;
            if (r < rows) {
              float v[4];
              ld4(data + (size_t)r * p.C + cq * 4, v);
#pragma unroll
              for (int e = 0; e < 4; ++e) { const float d = v[e] - pv[e]; t1[e] += d; t2[e] += d * d; }
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) { acc[e] += (double)t1[e]; acc[4 + e] += (double)t2[e]; }
        }
      }
    }
  }
  c8_cta_reduce(p, s, acc);
  cluster_arrive();
  cluster_wait();
  // ---- every CTA folds the cluster's partials (rank order) and builds y = act(x*k + o) for all channels
  float2* ko = reinterpret_cast<float2*>(s.red);
  if (tid < p.C) {
    double S1, S2;
    c8_fold(p, s, tid, &S1, &S2);
    float mu, rs;
    n8_mean_rstd(S1, S2, f_pv, p.inv_hw, p.eps, &mu, &rs);
    if (rank == 0) {
      mean[(size_t)n * p.C + tid] = mu;
      rstd[(size_t)n * p.C + tid] = rs;
    }
    float k, o, cc;
    n8_consts(mu, rs, f_g, f_b, f_tb, &k, &o, &cc);
    ko[tid] = make_float2(k, o);
  }
  __syncthreads();
  cluster_arrive();                       // this CTA is done with the others' shared memory
  {
    const int row = tid / p.TPR, cg = tid - row * p.TPR;
    if (row < p.RPP) {
      float k[8], o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float2 t = ko[cg * 8 + e]; k[e] = t.x; o[e] = t.y; }
      TY* yg = y + slice_off + cg * 8;
      const TY* rg = residual ? residual + slice_off + cg * 8 : nullptr;
      for (int c = 0; c < p.nchunks; ++c) {
        const int base = c * p.chunk_px + row;
        if (base >= rows) break;
        Raw8<TY> res[4];
        if (rg) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = base + j * p.RPP;
            if (r < rows) res[j] = ld_raw(rg + (size_t)r * p.C);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = base + j * p.RPP;
          if (r < rows) {
            float v[8];
            unpack(lds_raw(data + (size_t)r * p.C + cg * 8), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = n8_act<ACT>(fmaf(v[e], k[e], o[e]), p.slope);
            if (rg) {
              float q[8];
              unpack(res[j], q);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] += q[e];
            }
            st8(yg + (size_t)r * p.C, v);
          }
        }
      }
    }
  }
  cluster_wait();                         // nobody leaves while a neighbour may still read its partials
}

// ------------------------------------------------------------------------------------------------ backward
// dv = dy * act'(x*k + o) ; xh = x*rs + cc ; S1 = sum dv, S2 = sum dv*xh ; dx = k*dv + A*x + B
template <typename TX, typename TY, int ACT>
__global__ void __launch_bounds__(kCThreads, 3) inorm8c_bwd_kernel(
    Norm8CP p, const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ cbias, float* __restrict__ s1_out, float* __restrict__ s2_out, TX* __restrict__ dx) {
  extern __shared__ __align__(128) unsigned char sm[];
  const C8Smem s = c8_smem(sm, p);
  const TY* data = reinterpret_cast<const TY*>(s.data);
  const int tid = threadIdx.x, n = blockIdx.y;
  const int rank = (int)cluster_ctarank();
  const int px0 = rank * p.slice;
  const int rows = min(p.HW, px0 + p.slice) - px0;
  const size_t slice_off = ((size_t)n * p.HW + px0) * p.C;
  if (tid == 0) c8_issue_loads<TY>(p, s, dy + slice_off, rows);
  __syncthreads();
  float f_mu = 0.f, f_rs = 1.f, f_g = 1.f, f_b = 0.f, f_tb = 0.f;      // per-channel inputs of the fold
  if (tid < p.C) {
    f_mu = __ldg(mean + (size_t)n * p.C + tid);
    f_rs = __ldg(rstd + (size_t)n * p.C + tid);
    if (gamma) f_g = __ldg(gamma + tid);
    if (beta) f_b = __ldg(beta + tid);
    if (cbias) f_tb = __ldg(cbias + (size_t)n * p.C + tid);
  }
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.;
  {
    const int q4 = tid / p.TPR4, cq = tid - q4 * p.TPR4;
    if (q4 < p.RPP4) {
      float k[4], o[4], rs[4], cc[4];
      {
        const size_t nc = (size_t)n * p.C + cq * 4;
        const float4 mu4 = __ldg(reinterpret_cast<const float4*>(mean + nc));
        const float4 rs4 = __ldg(reinterpret_cast<const float4*>(rstd + nc));
        const float4 g4 = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + cq * 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 b4 = beta ? __ldg(reinterpret_cast<const float4*>(beta + cq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 t4 = cbias ? __ldg(reinterpret_cast<const float4*>(cbias + nc)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float mu_[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rs_[4] = {rs4.x, rs4.y, rs4.z, rs4.w};
        const float g_[4] = {g4.x, g4.y, g4.z, g4.w}, b_[4] = {b4.x, b4.y, b4.z, b4.w}, t_[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { rs[e] = rs_[e]; n8_consts(mu_[e], rs_[e], g_[e], b_[e], t_[e], &k[e], &o[e], &cc[e]); }
      }
      const TX* xg = x + slice_off + cq * 4;
      for (int c = 0; c < p.nchunks; ++c) {
        const int base = c * p.chunk_px;
        if (base >= rows) break;
        typename Raw4Of<TX>::type rx[8];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = base + q4 + h * p.RPP4 + j * p.RPP;
            if (r < rows) rx[h * 4 + j] = ldg4(xg + (size_t)r * p.C);
          }
        mbar_wait(&s.bars[c], 0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r0 = base + q4 + h * p.RPP4;
#pragma unroll
          for (int a = 0; a < 2; ++a) {            // two atoms of two rows
            float t1[4], t2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { t1[e] = 0.f; t2[e] = 0.f; }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int r = r0 + (2 * a + j) * p.RPP;
              if (r < rows) {
                float xv[4], dv[4];
                unpack4(rx[h * 4 + 2 * a + j], xv);
                ld4(data + (size_t)r * p.C + cq * 4, dv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float d = dv[e] * n8_act_grad<ACT>(fmaf(xv[e], k[e], o[e]), p.slope);
                  t1[e] += d; t2[e] += d * fmaf(xv[e], rs[e], cc[e]);
                }
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) { acc[e] += (double)t1[e]; acc[4 + e] += (double)t2[e]; }
          }
        }
      }
    }
  }
  c8_cta_reduce(p, s, acc);
  cluster_arrive();
  cluster_wait();
  float4* tab = reinterpret_cast<float4*>(s.red);     // k, o, A, B per channel
  if (tid < p.C) {
    double S1, S2;
    c8_fold(p, s, tid, &S1, &S2);
    const float m1 = (float)S1, m2 = (float)S2;
    const size_t nc = (size_t)n * p.C + tid;
    if (rank == 0) { s1_out[nc] = m1; s2_out[nc] = m2; }
    float k, o, cc, A, B;
    n8_consts(f_mu, f_rs, f_g, f_b, f_tb, &k, &o, &cc);
    n8_bwd_consts(k, f_rs, cc, m1, m2, p.inv_hw, &A, &B);
    tab[tid] = make_float4(k, o, A, B);
  }
  __syncthreads();
  cluster_arrive();
  {
    const int row = tid / p.TPR, cg = tid - row * p.TPR;
    if (row < p.RPP) {
      float k[8], o[8], A[8], B[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float4 t = tab[cg * 8 + e]; k[e] = t.x; o[e] = t.y; A[e] = t.z; B[e] = t.w; }
      const TX* xg = x + slice_off + cg * 8;
      TX* og = dx + slice_off + cg * 8;
      for (int c = 0; c < p.nchunks; ++c) {
        const int base = c * p.chunk_px + row;
        if (base >= rows) break;
        Raw8<TX> rx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = base + j * p.RPP;
          if (r < rows) rx[j] = ld_raw(xg + (size_t)r * p.C);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = base + j * p.RPP;
          if (r < rows) {
            float xv[8], dv[8];
            unpack(rx[j], xv);
            unpack(lds_raw(data + (size_t)r * p.C + cg * 8), dv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float d = dv[e] * n8_act_grad<ACT>(fmaf(xv[e], k[e], o[e]), p.slope);
              xv[e] = fmaf(k[e], d, fmaf(A[e], xv[e], B[e]));
            }
            st8(og + (size_t)r * p.C, xv);
          }
        }
      }
    }
  }
  cluster_wait();
}

// ------------------------------------------------------------------------------------------------ host
// resident_bytes: element size of the tensor held in shared memory (x forward, dy backward)
static int& onepass_enabled() {
  // opt-in: measured slower than the two-kernel path at the production batch (64 images = 1.33 waves of the 48
  // clusters a B200 holds; 36.9 vs 33.8 us forward, 59.4 vs 46.1 us backward, whole step 57.4 vs 56.3 ms), faster
  // from 48 images down (profiles/r2t_onepass_*).  The choice must not depend on the batch, so it is a switch.
  static int enabled = getenv("SRGAN_NORM_ONEPASS") ? atoi(getenv("SRGAN_NORM_ONEPASS")) : 0;
  return enabled;
}
static bool plan_norm8c(int N, int HW, int C, int resident_bytes, Norm8CP* out) {
  if (!onepass_enabled() || N <= 0 || HW <= 0 || C <= 0 || C % 8 || C > 256) return false;
  Norm8CP p = {};
  p.N = N; p.HW = HW; p.C = C;
  p.TPR = C / 8;
  p.RPP = kCThreads / p.TPR;
  if (p.RPP < 2 || (p.RPP & 1)) return false;
  p.TPR4 = 2 * p.TPR;
  p.RPP4 = p.RPP / 2;
  p.chunk_px = 4 * p.RPP;
  const long long chunk_bytes = (long long)p.chunk_px * C * resident_bytes;
  const int max_chunks = (int)(kCMaxData / chunk_bytes) < kCMaxChunks ? (int)(kCMaxData / chunk_bytes) : kCMaxChunks;
  if (max_chunks < 1) return false;
  const int total = ceil_div(HW, p.chunk_px);
  int CL = ceil_div(total, max_chunks);
  if (CL > kCMaxCluster) return false;
  const int per = ceil_div(total, CL);
  CL = ceil_div(total, per);
  p.CL = CL;
  p.nchunks = per;
  p.slice = per * p.chunk_px;
  p.data_bytes = (int)(per * chunk_bytes);
  p.inv_hw = 1.f / (float)HW;
  *out = p;
  return true;
}
static int norm8c_smem(const Norm8CP& p) { return p.data_bytes + kCRedBytes + kCPartBytes + kCMaxChunks * 8; }

template <typename K, typename... Args>
static cudaError_t launch_cluster(K kernel, unsigned long long* attr_done, const Norm8CP& p, cudaStream_t st,
                                  Args... args) {
  cudaError_t e = ensure_dyn_smem(kernel, kCMaxData + kCRedBytes + kCPartBytes + kCMaxChunks * 8, attr_done);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.CL, p.N, 1);
  cfg.blockDim = dim3(kCThreads, 1, 1);
  cfg.dynamicSmemBytes = norm8c_smem(p);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, p, args...);
}

template <typename TX, typename TY, int ACT>
static cudaError_t launch_fwd8c(const Norm8CP& p, const void* x, void* y, float* mean, float* rstd, const float* gamma,
                                const float* beta, const float* cbias, const void* residual, cudaStream_t st) {
  static unsigned long long done = 0;
  return launch_cluster(inorm8c_fwd_kernel<TX, TY, ACT>, &done, p, st, (const TX*)x, (TY*)y, mean, rstd, gamma, beta,
                        cbias, (const TY*)residual);
}
template <typename TX, typename TY, int ACT>
static cudaError_t launch_bwd8c(const Norm8CP& p, const void* dy, const void* x, const float* mean, const float* rstd,
                                const float* gamma, const float* beta, const float* cbias, float* s1, float* s2,
                                void* dx, cudaStream_t st) {
  static unsigned long long done = 0;
  return launch_cluster(inorm8c_bwd_kernel<TX, TY, ACT>, &done, p, st, (const TY*)dy, (const TX*)x, mean, rstd, gamma,
                        beta, cbias, s1, s2, (TX*)dx);
}

#define SRGAN_N8C_ACT(FN, TX, TY, ...)                                                  \
  switch (act) {                                                                        \
    case SRGAN_ACT_RELU:  err = FN<TX, TY, SRGAN_ACT_RELU>(__VA_ARGS__); break;         \
    case SRGAN_ACT_LRELU: err = FN<TX, TY, SRGAN_ACT_LRELU>(__VA_ARGS__); break;        \
    case SRGAN_ACT_TANH:  err = FN<TX, TY, SRGAN_ACT_TANH>(__VA_ARGS__); break;         \
    default:              err = FN<TX, TY, SRGAN_ACT_NONE>(__VA_ARGS__); break;         \
  }
#define SRGAN_N8C_TYPES(FN, ...)                                                        \
  do {                                                                                  \
    using B = __nv_bfloat16;                                                            \
    if (xb && yb) { SRGAN_N8C_ACT(FN, B, B, __VA_ARGS__) }                              \
    else if (xb) { SRGAN_N8C_ACT(FN, B, float, __VA_ARGS__) }                           \
    else if (yb) { SRGAN_N8C_ACT(FN, float, B, __VA_ARGS__) }                           \
    else { SRGAN_N8C_ACT(FN, float, float, __VA_ARGS__) }                               \
  } while (0)

bool norm8c_fwd(const void* x, bool xb, void* y, bool yb, float* mean, float* rstd, const float* gamma,
                const float* beta, const float* cbias, const void* residual, int N, int HW, int C, float eps, int act,
                float slope, cudaStream_t st, cudaError_t* status) {
  Norm8CP p;
  if (!plan_norm8c(N, HW, C, xb ? 2 : 4, &p)) return false;
  p.eps = eps; p.slope = slope;
  cudaError_t err = cudaSuccess;
  SRGAN_N8C_TYPES(launch_fwd8c, p, x, y, mean, rstd, gamma, beta, cbias, residual, st);
  *status = err;
  return true;
}

bool norm8c_bwd(const void* dy, bool yb, const void* x, bool xb, const float* mean, const float* rstd,
                const float* gamma, const float* beta, const float* cbias, void* dx, float* s1, float* s2, int N,
                int HW, int C, int act, float slope, cudaStream_t st, cudaError_t* status) {
  Norm8CP p;
  if (!plan_norm8c(N, HW, C, yb ? 2 : 4, &p)) return false;
  p.eps = 0.f; p.slope = slope;
  cudaError_t err = cudaSuccess;
  SRGAN_N8C_TYPES(launch_bwd8c, p, dy, x, mean, rstd, gamma, beta, cbias, s1, s2, dx, st);
  *status = err;
  return true;
}

}  // namespace srgan

using namespace srgan;

extern "C" int srgan_inorm_onepass_plan(int HW, int C, int resident_dtype, int* cluster, int* slice_px) {
  Norm8CP p;
  if (!(resident_dtype == SRGAN_DT_F32 || resident_dtype == SRGAN_DT_BF16)) return 0;
  if (!plan_norm8c(1, HW, C, resident_dtype == SRGAN_DT_BF16 ? 2 : 4, &p)) return 0;
  if (cluster) *cluster = p.CL;
  if (slice_px) *slice_px = p.slice;
  return 1;
}

extern "C" int srgan_inorm_onepass_enable(int on) {
  const int before = onepass_enabled();
  if (on >= 0) onepass_enabled() = on != 0;
  return before;
}

// introspection (tools/norm_onepass_probe.py): clusters of the forward bf16 -> bf16 ReLU kernel the device
// can hold at once for this plane, from cudaOccupancyMaxActiveClusters
extern "C" int srgan_inorm_onepass_max_clusters(int HW, int C, int resident_dtype) {
  Norm8CP p;
  if (!(resident_dtype == SRGAN_DT_F32 || resident_dtype == SRGAN_DT_BF16)) return -1;
  if (!plan_norm8c(1, HW, C, resident_dtype == SRGAN_DT_BF16 ? 2 : 4, &p)) return 0;
  auto kernel = inorm8c_fwd_kernel<__nv_bfloat16, __nv_bfloat16, SRGAN_ACT_RELU>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           kCMaxData + kCRedBytes + kCPartBytes + kCMaxChunks * 8) != cudaSuccess)
    return -1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.CL, 1024, 1);
  cfg.blockDim = dim3(kCThreads, 1, 1);
  cfg.dynamicSmemBytes = norm8c_smem(p);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}
