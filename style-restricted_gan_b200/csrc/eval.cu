// Notebook-04 classifier loss and PRDC evaluation.
//
//   cross entropy   ref: nn.CrossEntropyLoss() on the classifier output, notebook 04 cells 18 / 22
//                   (loss = mean_n (logsumexp(x_n) - x_n[label_n]); the reference feeds it softmax OUTPUTS, which is
//                   kept: the kernel takes whatever [N, J] scores it is given)
//   PRDC            ref: GAN_evaluation.get_prdc pyfiles/evaluation.py:98-110 -> prdc.compute_prdc (prdc==0.2,
//                   Docker/requirements.txt:13; Naeem et al., "Reliable Fidelity and Diversity Metrics for Generative
//                   Models", ICML 2020):
//                     r_real[i]  = distance from real i to its k-th nearest OTHER real sample
//                                  (= (k+1)-th smallest entry of row i of the real x real distance matrix, self included)
//                     precision  = mean_j  any_i  d(real_i, fake_j) < r_real[i]
//                     recall     = mean_i  any_j  d(real_i, fake_j) < r_fake[j]
//                     density    = 1 / (k M) * sum_j sum_i [d(real_i, fake_j) < r_real[i]]
//                     coverage   = mean_i  min_j d(real_i, fake_j) < r_real[i]
//                   Everything the metrics are made of is a COUNT of comparisons; the kernels return the integer counts
//                   (bit-exact against the CPU oracle), the four ratios are formed on the host.
// Distances are SQUARED Euclidean distances in fp64 (features fp32, differences and sums fp64, k ascending): the square
// root of the reference is monotone, so every comparison has the same outcome; fp64 keeps the comparisons away from
// the rounding of a float32 GEMM-style distance.  Byte / integer work otherwise: HBM-bound, no tensor cores.
#include "common.cuh"

namespace srgan {

// ------------------------------------------------------------------------------------------------ cross entropy
// one warp per row; row losses are written to loss_rows, a second fixed-order kernel averages them
__global__ void xent_rows_kernel(const float* __restrict__ x, const long long* __restrict__ label, float* __restrict__ loss_rows,
                                 int N, int J) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* xr = x + (size_t)row * J;
  float m = -INFINITY;
  for (int j = lane; j < J; j += 32) m = fmaxf(m, xr[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int j = lane; j < J; j += 32) s += expf(xr[j] - m);
  s = warp_sum(s);
  if (lane == 0) {
    const long long l = label[row];
    loss_rows[row] = (l >= 0 && l < J) ? (m + logf(s)) - xr[l] : 0.f;
  }
}
__global__ void mean_fixed_kernel(const float* __restrict__ v, float* __restrict__ out, int N) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += v[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s / (float)N;
}
// dx[n][j] = (softmax(x_n)[j] - [j == label_n]) * gout / N
__global__ void xent_bwd_kernel(const float* __restrict__ x, const long long* __restrict__ label, const float* __restrict__ gout,
                                float* __restrict__ dx, int N, int J) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* xr = x + (size_t)row * J;
  float m = -INFINITY;
  for (int j = lane; j < J; j += 32) m = fmaxf(m, xr[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int j = lane; j < J; j += 32) s += expf(xr[j] - m);
  s = warp_sum(s);
  const float g = __ldg(gout) / (float)N, inv = 1.f / s;
  const long long l = label[row];
  for (int j = lane; j < J; j += 32) dx[(size_t)row * J + j] = (expf(xr[j] - m) * inv - (j == l ? 1.f : 0.f)) * g;
}

// ------------------------------------------------------------------------------------------------ PRDC
// d2[i][j] = sum_k (a[i][k] - b[j][k])^2 in fp64, k ascending.  64 x 64 outputs per CTA, 16 x 16 threads, 4 x 4 each;
// k is staged through shared memory 16 at a time (rows of a / b are read once per 64 columns / rows of the tile).
constexpr int kPT = 64, kPK = 16;
__global__ void __launch_bounds__(256) pairdist2_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ d2,
                                                         int N, int M, int D) {
  __shared__ float sa[kPK][kPT + 1], sb[kPK][kPT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * kPT, j0 = blockIdx.x * kPT;
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.;
  for (int k0 = 0; k0 < D; k0 += kPK) {
    for (int e = threadIdx.x; e < kPT * kPK; e += 256) {
      const int r = e / kPK, k = e - r * kPK;
      sa[k][r] = (i0 + r < N && k0 + k < D) ? __ldg(a + (size_t)(i0 + r) * D + k0 + k) : 0.f;
      sb[k][r] = (j0 + r < M && k0 + k < D) ? __ldg(b + (size_t)(j0 + r) * D + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPK; ++k) {
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { av[u] = (double)sa[k][ty * 4 + u]; bv[u] = (double)sb[k][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) { const double df = av[u] - bv[v]; acc[u][v] = fma(df, df, acc[u][v]); }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < N && j < M) d2[(size_t)i * M + j] = acc[u][v];
    }
}

// radius[i] = (k+1)-th smallest entry of row i of d2 [N][N] (k >= 0; the row contains the zero self distance).  One CTA
// per row: k+1 rounds, each finds the smallest (value, index) pair that is lexicographically larger than the previous
// one - duplicates are taken one by one, like a sort would.
__global__ void __launch_bounds__(256) kth_smallest_kernel(const double* __restrict__ d2, double* __restrict__ radius, int N, int k) {
  __shared__ double sv[256];
  __shared__ int si[256];
  const double* row = d2 + (size_t)blockIdx.x * N;
  double pv = -1.;
  int pi = -1;
  for (int round = 0; round <= k; ++round) {
    double bv = INFINITY;
    int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < N; j += 256) {
      const double v = row[j];
      const bool after = v > pv || (v == pv && j > pi);
      if (after && (v < bv || (v == bv && j < bi))) { bv = v; bi = j; }
    }
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) {
        const double ov = sv[threadIdx.x + s];
        const int oi = si[threadIdx.x + s];
        if (ov < sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
      }
      __syncthreads();
    }
    pv = sv[0];
    pi = si[0];
    __syncthreads();
  }
  if (threadIdx.x == 0) radius[blockIdx.x] = pv;
}

// per real row i of d2_rf [N][M]:  row_hits_fake[i] = #{j : d < r_fake[j]},  row_min_in[i] = [min_j d < r_real[i]]
__global__ void __launch_bounds__(256) prdc_rows_kernel(const double* __restrict__ d2, const double* __restrict__ r_real,
                                                         const double* __restrict__ r_fake, int* __restrict__ row_hits_fake,
                                                         int* __restrict__ row_min_in, int N, int M) {
  __shared__ double sv[256];
  __shared__ int sc[256];
  const int i = blockIdx.x;
  const double* row = d2 + (size_t)i * M;
  double mn = INFINITY;
  int c = 0;
  for (int j = threadIdx.x; j < M; j += 256) {
    const double v = row[j];
    mn = fmin(mn, v);
    c += v < r_fake[j];
  }
  sv[threadIdx.x] = mn;
  sc[threadIdx.x] = c;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sv[threadIdx.x] = fmin(sv[threadIdx.x], sv[threadIdx.x + s]); sc[threadIdx.x] += sc[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { row_hits_fake[i] = sc[0]; row_min_in[i] = sv[0] < r_real[i]; }
}
// per fake column j:  col_hits_real[j] = #{i : d2[i][j] < r_real[i]}.  Threads of a warp read consecutive j (coalesced);
// the rows are cut into gridDim.y slabs whose integer counts are added with atomicAdd (integers: order-independent,
// bit-exact) into the zeroed result - the first, column-serial version took 0.46 ms for 2048 x 2048 (73 GB/s).
__global__ void prdc_cols_kernel(const double* __restrict__ d2, const double* __restrict__ r_real, int* __restrict__ col_hits_real,
                                 int N, int M, int rows_per_slab) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int i0 = blockIdx.y * rows_per_slab, i1 = min(N, i0 + rows_per_slab);
  int c = 0;
  for (int i = i0; i < i1; ++i) c += d2[(size_t)i * M + j] < __ldg(r_real + i);
  if (c) atomicAdd(col_hits_real + j, c);
}

}  // namespace srgan

using namespace srgan;

extern "C" int srgan_cross_entropy_fwd(const float* x, const long long* label, float* loss, float* loss_rows, int N, int J,
                                       void* stream) {
  SRGAN_CHECK_ARG(x && label && loss && loss_rows, "null pointer");
  SRGAN_CHECK_ARG(N > 0 && J > 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  xent_rows_kernel<<<ceil_div(N, 4), 128, 0, st>>>(x, label, loss_rows, N, J);
  mean_fixed_kernel<<<1, 256, 0, st>>>(loss_rows, loss, N);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_cross_entropy_bwd(const float* x, const long long* label, const float* gout, float* dx, int N, int J,
                                       void* stream) {
  SRGAN_CHECK_ARG(x && label && gout && dx, "null pointer");
  SRGAN_CHECK_ARG(N > 0 && J > 0, "bad sizes");
  xent_bwd_kernel<<<ceil_div(N, 4), 128, 0, (cudaStream_t)stream>>>(x, label, gout, dx, N, J);
  SRGAN_RETURN_LAUNCH();
}

extern "C" int srgan_prdc_pairdist2(const float* a, const float* b, double* d2, int N, int M, int D, void* stream) {
  SRGAN_CHECK_ARG(a && b && d2, "null pointer");
  SRGAN_CHECK_ARG(N >= 0 && M >= 0 && D > 0, "bad sizes");
  if (N == 0 || M == 0) return SRGAN_OK;
  pairdist2_kernel<<<dim3(ceil_div(M, kPT), ceil_div(N, kPT)), 256, 0, (cudaStream_t)stream>>>(a, b, d2, N, M, D);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_prdc_kth_radius(const double* d2_self, double* radius, int N, int k, void* stream) {
  SRGAN_CHECK_ARG(d2_self && radius, "null pointer");
  SRGAN_CHECK_ARG(N > 0 && k >= 0 && k < N, "need 0 <= k < N");
  kth_smallest_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(d2_self, radius, N, k);
  SRGAN_RETURN_LAUNCH();
}
extern "C" int srgan_prdc_counts(const double* d2_real_fake, const double* r_real, const double* r_fake, int* col_hits_real,
                                 int* row_hits_fake, int* row_min_in, int N, int M, void* stream) {
  SRGAN_CHECK_ARG(d2_real_fake && r_real && r_fake && col_hits_real && row_hits_fake && row_min_in, "null pointer");
  SRGAN_CHECK_ARG(N > 0 && M > 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  prdc_rows_kernel<<<N, 256, 0, st>>>(d2_real_fake, r_real, r_fake, row_hits_fake, row_min_in, N, M);
  if (cudaMemsetAsync(col_hits_real, 0, (size_t)M * sizeof(int), st) != cudaSuccess) { set_error("prdc_counts: memset failed"); return SRGAN_E_BADARG; }
  const int slabs = N < 64 ? 1 : (N + 63) / 64;
  prdc_cols_kernel<<<dim3(ceil_div(M, 128), slabs), 128, 0, st>>>(d2_real_fake, r_real, col_hits_real, N, M, ceil_div(N, slabs));
  SRGAN_RETURN_LAUNCH();
}
