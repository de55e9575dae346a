// TMA feed-rate probe (bring-up tool, not part of the library): how many bytes per clock can ONE SM pull through
// cp.async.bulk.tensor with 128-byte (32 x fp32, SWIZZLE_128B) box rows, as a function of tensor-map rank, box
// shape, stages in flight and number of issuing threads?  One CTA per SM, a ring of mbarrier stages, a consumer
// thread that only waits and releases.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tools/tma_probe.cu -lcuda
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ int g_waitmode;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const int wm = g_waitmode;
  while (!done) {
    if (wm == 0)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    else if (wm == 1)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(20u) : "memory");
  }
}
__device__ __forceinline__ void tma2(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma4(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma5(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// mode 0: 2D map {C, pixels} box {32, rows}        (GEMM-like)
// mode 1: 4D map {C, W, H, N} box {32, 32, rows/32, 1}
// mode 2: 5D map {C, W, 1, H, N} box {32, 32, 1, rows/32, 1}
// Every stage = `boxes` boxes of `rows` rows of 128 B.  Images are 32x32x256ch; a CTA walks over images.
struct P { int mode, stages, boxes, rows, iters, nimg, rings, pad; };

// `rings` independent producer/consumer pipelines per CTA: ring r = issuer warp r (lane 0) + consumer warp 4+r.
__global__ void __launch_bounds__(288, 1) probe(const __grid_constant__ CUtensorMap map, const __grid_constant__ P p,
                                                 long long* clocks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  const int stage_bytes = p.boxes * p.rows * 128;
  const int ring_bytes = p.stages * stage_bytes;
  uint64_t* bars = (uint64_t*)(smem + p.rings * ring_bytes);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.rings * p.stages * 2; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (lane == 0 && warp < p.rings) {
    uint64_t* full = bars + warp * p.stages * 2;
    uint64_t* empty = full + p.stages;
    uint8_t* base = smem + warp * ring_bytes;
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(empty + stage, phase ^ 1);
      mbar_expect_tx(full + stage, stage_bytes);
      for (int b = 0; b < p.boxes; ++b) {
        uint8_t* dst = base + stage * stage_bytes + b * p.rows * 128;
        // pad == 0: many CTAs touch the same box at the same time; pad == 1: every CTA streams its own boxes
        const int g0 = ((it * p.rings + warp) * p.boxes + b);
        const int rg = 32 / (p.rows / 32);                  // row groups per image
        const int g = p.pad ? (blockIdx.x * 27 + g0) % (p.nimg * rg * 8) : g0;
        const int cc = (g % 8) * 32;
        const int img = p.pad ? (g / 8 / rg) % p.nimg : (blockIdx.x * 7 + g / 8) % p.nimg;
        const int h0 = ((g / 8) % rg) * (p.rows / 32);
        tma5(&map, full + stage, dst, cc, 0, 0, h0, img);
      }
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (lane == 0 && warp >= 4 && warp < 4 + p.rings) {
    uint64_t* full = bars + (warp - 4) * p.stages * 2;
    uint64_t* empty = full + p.stages;
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(full + stage, phase);
      mbar_arrive(empty + stage);
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = clock64() - t0;
}

int main() {
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const int nimg = 64, C = 256, H = 32, W = 32;
  float* x;
  CK(cudaMalloc(&x, (size_t)nimg * H * W * C * 4));
  CK(cudaMemset(x, 0, (size_t)nimg * H * W * C * 4));
  long long* clocks;
  CK(cudaMalloc(&clocks, 148 * 8));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  printf("rings stages boxes rows | KB/stage/ring  clk/iter  B/clk/SM  GB/s(chip)\n");
  { int wm = 0; CK(cudaMemcpyToSymbol(g_waitmode, &wm, sizeof(int))); }
  const int mode = 2, grid = 148;
   for (int distinct = 0; distinct < 2; ++distinct)
    for (int rows : {128})
      for (int boxes : {1})
        for (int stages : {2, 3})
          for (int rings : {1, 2, 4}) {
              if ((size_t)rings * stages * boxes * rows * 128 > 200 * 1024) continue;
              CUtensorMap m;
              uint32_t es[5] = {1, 1, 1, 1, 1};
              CUresult r;
              if (mode == 0) {
                uint64_t d[2] = {(uint64_t)C, (uint64_t)nimg * H * W}; uint64_t s[1] = {(uint64_t)C * 4};
                uint32_t b[2] = {32, (uint32_t)rows};
                r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
              } else if (mode == 1) {
                uint64_t d[4] = {(uint64_t)C, W, H, nimg}; uint64_t s[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
                uint32_t b[4] = {32, 32, (uint32_t)rows / 32, 1};
                r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
              } else {
                uint64_t d[5] = {(uint64_t)C, W, 1, H, nimg};
                uint64_t s[4] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
                uint32_t b[5] = {32, 32, 1, (uint32_t)rows / 32, 1};
                r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, x, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
              }
              if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
              P p = {mode, stages, boxes, rows, 400, nimg, rings, distinct};
              size_t smem = 1024 + (size_t)rings * stages * boxes * rows * 128 + 512;
              probe<<<grid, 288, smem>>>(m, p, clocks);      // warm (L2)
              probe<<<grid, 288, smem>>>(m, p, clocks);
              CK(cudaDeviceSynchronize());
              long long h[148];
              CK(cudaMemcpy(h, clocks, grid * 8, cudaMemcpyDeviceToHost));
              long long mx = 0;
              for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
              double clk_stage = (double)mx / p.iters;
              double bytes = (double)rings * boxes * rows * 128;
              printf("d%d %5d %6d %5d %4d | %12.0f %10.0f %9.1f %10.0f\n", distinct, rings, stages, boxes, rows, bytes / rings / 1024, clk_stage,
                     bytes / clk_stage, bytes / clk_stage * grid * 1.9);
            }
  return 0;
}
