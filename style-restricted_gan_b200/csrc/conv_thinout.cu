// Stride-1 convolutions whose OUTPUT has <= 4 channels (the generator's RGB head, and the input gradient of its
// RGB stem): tcgen05 row-GEMM + in-CTA col2im.  sm_100a only.
//
//   out[h][q][t] = sum_{r,s,f} in[h + r - padH][q + s - padW][f] * Wt[(r,s,t)][f]
//
// A direct implicit GEMM has N = 3 and re-reads the fat input once per tap (49x for the 7x7 head).  Instead each
// input image row h' is multiplied ONCE by the whole packed filter
//   Z_h'[q'][(r, s*4+t)] = sum_f in[h'][q'][f] * Bp[r*32 + s*4 + t][f]          (M = 128 pixels, N = 32*R, K = C)
// with the accumulator in tensor memory (double buffered), and the epilogue warps scatter-free "col2im" it:
//   out[h][q][t] = sum_r sum_s Z_{h+r-padH}[q + s - padW][(r, s*4+t)]
// The column shift crosses lanes, so every 32-column block of Z goes through a transposed shared-memory
// staging buffer; the row shift is a ring of partial output rows in shared memory (a CTA walks down a band
// of rows, so each input row is read from global memory once per band).
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM allocator), warps 2..5 epilogue.
#include <stdlib.h>
#include <cuda_bf16.h>
#include "umma_ptx.cuh"

namespace srgan {

constexpr int kTOThreads = 192;
constexpr int kTOStages = 3;
constexpr int kTOZRow = 144;                 // floats per staging row: 128 pixels + margins (pad <= 8 each side)
constexpr int kTOZBuf = 32 * kTOZRow;        // one staging buffer: 32 packed columns

struct ThinOutP {
  int Hi, Wi, Ho, Wo;
  int tc, R, S, padH, padW;
  int nchunks;                   // input channels / 32
  int bands, BH;                 // row bands per image, rows per band
  int act;
  float slope;
  unsigned int idesc;
};

__global__ void __launch_bounds__(kTOThreads, 1)
conv_thinout_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ ThinOutP p, const float* __restrict__ bias, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ symbol keeps the address space visible: LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_bytes = p.nchunks * p.R * 4096;           // packed filter: per chunk R*32 rows of 128 B
  const int a_bytes = p.nchunks * 16384;                // one input row: per chunk 128 pixels of 128 B
  uint8_t* sb = smem;
  uint8_t* sa = smem + b_bytes;
  float* zs = reinterpret_cast<float*>(sa + kTOStages * a_bytes);
  float* accs = zs + 2 * kTOZBuf;                       // [8 rows][4][128]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(accs + 8 * 4 * 128);
  uint64_t* a_empty = a_full + kTOStages;
  uint64_t* t_full = a_empty + kTOStages;               // [2]
  uint64_t* t_empty = t_full + 2;                       // [2]
  uint64_t* b_full = t_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.bands, band = blockIdx.x % p.bands;
  const int h0 = band * p.BH, h1 = min(p.Ho, h0 + p.BH);
  const int hp_beg = h0 - p.padH, hp_end = (h1 - 1) - p.padH + (p.R - 1);     // inclusive range of input rows

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTOStages; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 4); }
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
  }
  for (int i = threadIdx.x; i < 2 * kTOZBuf + 8 * 4 * 128; i += kTOThreads) zs[i] = 0.f;   // margins stay zero
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(b_full, b_bytes);
      for (int j = 0; j < p.nchunks; ++j) tma_load_2d(&map_b, b_full, sb + j * p.R * 4096, 32 * j, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int hp = hp_beg; hp <= hp_end; ++hp) {
        if (hp < 0 || hp >= p.Hi) continue;
        mbar_wait(a_empty + stage, phase ^ 1);
        uint8_t* dst = sa + stage * a_bytes;
        mbar_expect_tx(a_full + stage, a_bytes);
        for (int j = 0; j < p.nchunks; ++j) tma_load_4d(&map_in, a_full + stage, dst + j * 16384, 32 * j, 0, hp, n);
        if (++stage == kTOStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(b_full, 0);
      int stage = 0, i = 0;
      uint32_t phase = 0;
      for (int hp = hp_beg; hp <= hp_end; ++hp) {
        if (hp < 0 || hp >= p.Hi) continue;
        const int buf = i & 1;
        mbar_wait(a_full + stage, phase);
        mbar_wait(t_empty + buf, ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sa + stage * a_bytes), b0 = smem_u32(sb);
        for (int j = 0; j < p.nchunks; ++j) {
          const uint64_t adesc = smem_desc_sw128(a0 + j * 16384);
          const uint64_t bdesc = smem_desc_sw128(b0 + j * p.R * 4096);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tmem_base + buf * 256, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc, (j | k) != 0);
        }
        umma_commit(a_empty + stage);
        umma_commit(t_full + buf);
        if (++stage == kTOStages) { stage = 0; phase ^= 1; }
        ++i;
      }
    }
  } else {
    const int quad = warp & 3;
    const int qq = quad * 32 + lane;                    // TMEM lane == input pixel q' == output pixel q
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    // ring of partial output rows: accs[h & 7][t][pixel]; a thread only ever touches its own pixel column
    float bv[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) bv[t] = (bias && t < p.tc) ? __ldg(bias + t) : 0.f;
    int i = 0, zi = 0;
    for (int hp = hp_beg; hp <= hp_end; ++hp) {
      if (hp >= 0 && hp < p.Hi) {
        const int buf = i & 1;
        mbar_wait(t_full + buf, (i >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int r = 0; r < p.R; ++r) {                 // (kept rolled: the body is ~150 instructions)
          const int h = hp - r + p.padH;
          if (h < h0 || h >= h1) continue;              // uniform over the CTA
          float v[32];
          tmem_ld32(taddr + buf * 256 + r * 32, v);
          float* zb = zs + (zi & 1) * kTOZBuf;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < p.S * 4) zb[j * kTOZRow + p.padW + qq] = v[j];
          asm volatile("bar.sync 1, 128;" ::: "memory");
          float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int s = 0; s < 8; ++s)
            if (s < p.S) {
#pragma unroll
              for (int t = 0; t < 4; ++t)
                if (t < p.tc) sum[t] += zb[(s * 4 + t) * kTOZRow + qq + s];
            }
          float* ar = accs + (h & 7) * 4 * 128 + qq;
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (t < p.tc) ar[t * 128] += sum[t];
          ++zi;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + buf);
        ++i;
      }
      // output row hp + padH - (R-1) has now received all of its filter rows
      const int hd = hp + p.padH - (p.R - 1);
      if (hd >= h0 && hd < h1) {
        float* ar = accs + (hd & 7) * 4 * 128 + qq;
        float* o = out + (((size_t)n * p.Ho + hd) * p.Wo + qq) * p.tc;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (t < p.tc) {
            if (qq < p.Wo) o[t] = apply_act(ar[t * 128] + bv[t], p.act, p.slope);
            ar[t * 128] = 0.f;
          }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Variant 2: the column shift is done by the tensor core.  For filter column s the A operand is the SAME smem row
// buffer read through a descriptor whose start address is advanced by s pixel rows (s*128 B; the row buffer holds
// pixels -padW .. 127+S-1-padW, TMA zero-fills outside the image), multiplied by the 32-row filter slice
// Bs[s][(r*4+t)][f]:   D[q][(r,t)] = sum_s sum_f in[q + s - padW][f] * W[(r,s,t)][f].
// D has only 32 columns, the epilogue is one tcgen05.ld per image row plus the row ring -- no shared-memory
// transpose.  The row-shifted start is not 1024-byte aligned; measured on B200: the 128-byte swizzle XOR is taken
// from the absolute shared-memory address bits, so the shifted descriptor reads exactly the rows TMA wrote (the
// descriptor's base-offset field must stay 0 -- setting it to (start >> 7) & 7 gives wrong results).
constexpr int kTO2ARows = 144;                              // smem rows per chunk: 128 + S-1 (<= 7) pixels, padded
constexpr int kTO2ABytes = kTO2ARows * 128;                 // 18 KB, multiple of 1024

struct ThinOut2P {
  int Hi, Wi, Ho, Wo;
  int tc, R, S, padH, padW;
  int nchunks, bands, BH;
  int act;
  float slope;
};

// ST: storage type of the FAT input (and of the packed filter): float -> 32 channels per 128-byte row, TF32 MMAs;
// __nv_bfloat16 ("thin16", the bf16 engine's head fprop / stem dgrad) -> 64 channels per row, kind::f16 - half as many
// chunks, i.e. half as many of the N = 32 MMAs whose issue rate bounds this kernel.  The thin output stays fp32.
template <typename ST>
__global__ void __launch_bounds__(kTOThreads, 1)
conv_thinout2_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ ThinOut2P p, const float* __restrict__ bias, float* __restrict__ out) {
  constexpr int kCh = 128 / (int)sizeof(ST);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_bytes = p.S * p.nchunks * 4096;           // Bs[s][chunk][32 rows][32 f]
  const int a_bytes = p.nchunks * kTO2ABytes;
  uint8_t* sb = smem;
  uint8_t* sa = smem + b_bytes;
  float* accs = reinterpret_cast<float*>(sa + kTOStages * a_bytes);      // [8 rows][4][128]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(accs + 8 * 4 * 128);
  uint64_t* a_empty = a_full + kTOStages;
  uint64_t* t_full = a_empty + kTOStages;
  uint64_t* t_empty = t_full + 2;
  uint64_t* b_full = t_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.bands, band = blockIdx.x % p.bands;
  const int h0 = band * p.BH, h1 = min(p.Ho, h0 + p.BH);
  const int hp_beg = h0 - p.padH, hp_end = (h1 - 1) - p.padH + (p.R - 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTOStages; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 4); }
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
  }
  for (int i = threadIdx.x; i < 8 * 4 * 128; i += kTOThreads) accs[i] = 0.f;
  // rows 128+S-1 .. 143 of every A buffer are never written by TMA: zero them once (they are only read by
  // accumulator rows that do not exist, but NaN garbage must not reach the tensor core's exception paths)
  for (int i = threadIdx.x; i < kTOStages * a_bytes / 4; i += kTOThreads) reinterpret_cast<float*>(sa)[i] = 0.f;
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy zero fill before TMA writes
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int box_bytes = (128 + p.S - 1) * 128;

  if (warp == 0) {
    // (measured: one producer warp per stage - three warps - does not change the time; neither does keeping the
    //  partial output rows in registers, below.  What sets the pace is the MMA itself: ~146 clk per N = 32 MMA whose A
    //  start is shifted by s pixel rows inside the 128-byte-swizzled buffer, against 16 clk of math: 28 MMAs = 4100 clk
    //  per image row with bf16 input, 56 x 107 = 6000 clk with fp32 input - the same constant per MMA in both.)
    if (lane == 0) {
      mbar_expect_tx(b_full, b_bytes);
      for (int j = 0; j < p.S * p.nchunks; ++j) tma_load_2d(&map_b, b_full, sb + j * 4096, 0, j * 32);
      int stage = 0;
      uint32_t phase = 0;
      for (int hp = hp_beg; hp <= hp_end; ++hp) {
        if (hp < 0 || hp >= p.Hi) continue;
        mbar_wait(a_empty + stage, phase ^ 1);
        uint8_t* dst = sa + stage * a_bytes;
        mbar_expect_tx(a_full + stage, p.nchunks * box_bytes);
        for (int j = 0; j < p.nchunks; ++j)
          tma_load_4d(&map_in, a_full + stage, dst + j * kTO2ABytes, kCh * j, -p.padW, hp, n);
        if (++stage == kTOStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = f32, A = B = tf32 (format 2) or bf16 (format 1), both K-major, N = 32, M = 128
      constexpr uint32_t kFmt = sizeof(ST) == 4 ? 2u : 1u;
      const uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      mbar_wait(b_full, 0);
      int stage = 0, i = 0;
      uint32_t phase = 0;
      for (int hp = hp_beg; hp <= hp_end; ++hp) {
        if (hp < 0 || hp >= p.Hi) continue;
        const int buf = i & 1;
        mbar_wait(a_full + stage, phase);
        mbar_wait(t_empty + buf, ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sa + stage * a_bytes), b0 = smem_u32(sb);
        // (measured: one accumulator per s, summed in the epilogue, is slower -- 242 vs 208 us -- the 56 N=32 MMAs per
        //  image row are bound by their issue rate, not by the accumulate dependency)
        uint32_t first = 1;
        for (int s = 0; s < p.S; ++s)
          for (int j = 0; j < p.nchunks; ++j) {
            const uint32_t astart = a0 + j * kTO2ABytes + s * 128;        // row-shifted start (see above)
            const uint64_t adesc = smem_desc_sw128(astart);
            const uint64_t bdesc = smem_desc_sw128(b0 + (s * p.nchunks + j) * 4096);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if constexpr (sizeof(ST) == 4)
                umma_tf32(tmem_base + buf * 32, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first ^ 1);
              else
                umma_f16(tmem_base + buf * 32, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first ^ 1);
              first = 0;
            }
          }
        umma_commit(a_empty + stage);
        umma_commit(t_full + buf);
        if (++stage == kTOStages) { stage = 0; phase ^= 1; }
        ++i;
      }
    }
  } else {
    const int quad = warp & 3;
    const int qq = quad * 32 + lane;                    // TMEM lane == output pixel q
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float bv[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) bv[t] = (bias && t < p.tc) ? __ldg(bias + t) : 0.f;
    // Partial output rows live in REGISTERS: part[a] is output row hp + padH - a (a = filter row that input row hp
    // contributes to it); every step the window shifts by one row and row a = R - 1 is complete.  (The first version
    // kept them in a shared-memory ring: 21 dependent load-add-store triples per input row made the four epilogue
    // warps, not the tensor pipe, set the pace - ncu: 14 % of all stall samples on that one line.)
    float part[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int t = 0; t < 4; ++t) part[a][t] = 0.f;
    int i = 0;
    for (int hp = hp_beg; hp <= hp_end; ++hp) {
#pragma unroll
      for (int a = 7; a > 0; --a)
#pragma unroll
        for (int t = 0; t < 4; ++t) part[a][t] = part[a - 1][t];
#pragma unroll
      for (int t = 0; t < 4; ++t) part[0][t] = 0.f;
      if (hp >= 0 && hp < p.Hi) {
        const int buf = i & 1;
        mbar_wait(t_full + buf, (i >> 1) & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(taddr + buf * 32, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + buf);      // accumulator is in registers: release it right away
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int t = 0; t < 4; ++t) part[a][t] += v[a * 4 + t];      // columns of filter rows >= R are zero
        ++i;
      }
      const int hd = hp + p.padH - (p.R - 1);
      if (hd >= h0 && hd < h1 && qq < p.Wo) {
        float done[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 8; ++a)
          if (a == p.R - 1) {
#pragma unroll
            for (int t = 0; t < 4; ++t) done[t] = part[a][t];
          }
        float* o = out + (((size_t)n * p.Ho + hd) * p.Wo + qq) * p.tc;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (t < p.tc) o[t] = apply_act(done[t] + bv[t], p.act, p.slope);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// Bs[s][chunk][r*4 + t][f32]  (rows of 32 floats = one 128-byte swizzle row)
// mode 0: = w[t][r][s][chunk*32 + f] ; mode 1: = w[chunk*32 + f][R-1-r][S-1-s][t]
__device__ __forceinline__ void st_elem(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_elem(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <typename ST>
__global__ void thinout2_pack_filter_kernel(const float* __restrict__ w, ST* __restrict__ bp, int K, int C, int R,
                                            int S, int mode) {
  constexpr int kCh = 128 / (int)sizeof(ST);                 // channels per 128-byte row
  const int F = mode == 0 ? C : K, tc = mode == 0 ? K : C;
  const int nch = F / kCh;
  const int total = S * nch * 32 * kCh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int f = i % kCh, row = (i / kCh) & 31, j = (i / (kCh * 32)) % nch, s = i / (kCh * 32 * nch);
    const int r = row >> 2, t = row & 3, ff = j * kCh + f;
    float v = 0.f;
    if (r < R && t < tc) {
      if (mode == 0) v = w[(((size_t)t * R + r) * S + s) * C + ff];
      else           v = w[(((size_t)ff * R + (R - 1 - r)) * S + (S - 1 - s)) * C + t];
    }
    st_elem(bp + i, v);
  }
}

// mode 0 (fprop, K <= 4):  bp[r*32 + s*4 + k][c] = w[k][r][s][c]                    rows of C floats
// mode 1 (dgrad, C <= 4):  bp[r'*32 + s'*4 + c][k] = w[k][R-1-r'][S-1-s'][c]        rows of K floats
__global__ void thinout_pack_filter_kernel(const float* __restrict__ w, float* __restrict__ bp, int K, int C, int R,
                                           int S, int mode) {
  const int F = mode == 0 ? C : K, tc = mode == 0 ? K : C;
  const int total = R * 32 * F;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int f = i % F, j = (i / F) & 31, r = i / (F * 32);
    const int s = j >> 2, t = j & 3;
    float v = 0.f;
    if (s < S && t < tc) {
      if (mode == 0) v = w[(((size_t)t * R + r) * S + s) * C + f];
      else           v = w[(((size_t)f * R + (R - 1 - r)) * S + (S - 1 - s)) * C + t];
    }
    bp[i] = v;
  }
}

struct ThinOutPlan { int mode, Hi, Wi, Ho, Wo, F, tc, padH, padW, bands, BH; };

static bool thinout_plan(const srgan_conv_desc* d, int pass, ThinOutPlan* t) {
  if (d->stride != 1 || d->S * 4 > 32 || d->R > 8 || d->N < 1) return false;
  ThinOutPlan q = {};
  if (pass == 0) {
    if (d->K > 4) return false;
    q.mode = 0; q.Hi = d->H; q.Wi = d->W; q.Ho = d->P; q.Wo = d->Q; q.F = d->C; q.tc = d->K;
    q.padH = q.padW = d->pad;
  } else if (pass == 1) {
    if (d->C > 4 || d->pad > d->R - 1 || d->pad > d->S - 1) return false;
    q.mode = 1; q.Hi = d->P; q.Wi = d->Q; q.Ho = d->H; q.Wo = d->W; q.F = d->K; q.tc = d->C;
    q.padH = d->R - 1 - d->pad; q.padW = d->S - 1 - d->pad;
  } else {
    return false;
  }
  if (q.F != 32 && q.F != 64) return false;
  if (q.Wi > 128 || q.Wo > 128 || q.padW > 8) return false;
  // rows per band: minimise  waves x (rows + halo) over one-CTA-per-SM waves
  long best = -1;
  for (int b = 1; b <= q.Ho; ++b) {
    int bh = ceil_div(q.Ho, b);
    if (ceil_div(q.Ho, bh) != b) continue;
    long waves = ((long)d->N * b + kNumSMs - 1) / kNumSMs;
    long cost = waves * (bh + d->R - 1 + 4);
    if (best < 0 || cost < best) { best = cost; q.bands = b; q.BH = bh; }
  }
  *t = q;
  return true;
}

bool conv_thinout_supported(const srgan_conv_desc* d, int pass) {
  ThinOutPlan t;
  return thinout_plan(d, pass, &t);
}

size_t conv_thinout_workspace(const srgan_conv_desc* d, int pass) {
  ThinOutPlan t;
  if (!thinout_plan(d, pass, &t)) return 0;
  const int rs = d->R > d->S ? d->R : d->S;
  return (size_t)rs * 32 * t.F * sizeof(float);
}

// the MMA-column-shift kernel for a fat input of storage type ST
template <typename ST>
static int thinout2_launch(const srgan_conv_desc* d, const ThinOutPlan& t, const void* in, const float* w,
                           const float* bias, float* out, int act, float slope, void* ws, cudaStream_t st) {
  constexpr int kCh = 128 / (int)sizeof(ST);
  constexpr uint64_t ES = sizeof(ST);
  const CUtensorMapDataType DT = sizeof(ST) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  ST* bp = (ST*)ws;
  thinout2_pack_filter_kernel<ST><<<ceil_div(d->S * t.F * 32, 256), 256, 0, st>>>(w, bp, d->K, d->C, d->R, d->S, t.mode);
  CUtensorMap min2, mb2;
  {
    uint64_t dims[4] = {(uint64_t)t.F, (uint64_t)t.Wi, (uint64_t)t.Hi, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)t.F * ES, (uint64_t)t.Wi * t.F * ES, (uint64_t)t.Hi * t.Wi * t.F * ES};
    uint32_t box[4] = {(uint32_t)kCh, (uint32_t)(128 + d->S - 1), 1, 1};
    if (int e = encode_map(&min2, in, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, DT)) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)kCh, (uint64_t)d->S * (t.F / kCh) * 32};
    uint64_t str[1] = {128};
    uint32_t box[2] = {(uint32_t)kCh, 32};
    if (int e = encode_map(&mb2, bp, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, DT)) return e;
  }
  ThinOut2P q = {};
  q.Hi = t.Hi; q.Wi = t.Wi; q.Ho = t.Ho; q.Wo = t.Wo; q.tc = t.tc; q.R = d->R; q.S = d->S;
  q.padH = t.padH; q.padW = t.padW; q.nchunks = t.F / kCh; q.bands = t.bands; q.BH = t.BH;
  q.act = act; q.slope = slope;
  const size_t smem2 = 1024 + (size_t)d->S * q.nchunks * 4096 + (size_t)kTOStages * q.nchunks * kTO2ABytes +
                       8 * 4 * 128 * sizeof(float) + 256;
  static unsigned long long attr2 = 0;
  {
    cudaError_t e = ensure_dyn_smem(conv_thinout2_kernel<ST>, 227 * 1024, &attr2);
    if (e != cudaSuccess) { set_error("conv_thinout2 smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  conv_thinout2_kernel<ST><<<d->N * t.bands, kTOThreads, smem2, st>>>(min2, mb2, q, bias, out);
  SRGAN_RETURN_LAUNCH();
}

// "thin16": bf16 fat input (64 channels), fp32 thin output.  pass 0: in = x (bf16), out = y;  pass 1: in = dy (bf16), out = dx
bool conv_thinout16_supported(const srgan_conv_desc* d, int pass) {
  ThinOutPlan t;
  return thinout_plan(d, pass, &t) && t.F == 64;
}
int conv_thinout16_launch(const srgan_conv_desc* d, int pass, const void* in, const float* w, const float* bias,
                          float* out, int act, float slope, void* ws, size_t ws_bytes, cudaStream_t st) {
  ThinOutPlan t;
  if (!thinout_plan(d, pass, &t) || t.F != 64) { set_error("thin16 thin-output conv: unsupported shape"); return SRGAN_E_UNSUPPORTED; }
  const size_t need = conv_thinout_workspace(d, pass);
  if (!ws || ws_bytes < need) { set_error("thin-output conv: workspace %zu < %zu", ws_bytes, need); return SRGAN_E_WORKSPACE; }
  if (((uintptr_t)in | (uintptr_t)ws) % 16) { set_error("thin-output conv: tensors must be 16-byte aligned"); return SRGAN_E_BADARG; }
  return thinout2_launch<__nv_bfloat16>(d, t, in, w, bias, out, act, slope, ws, st);
}

// pass 0: in = x, out = y (bias/activation fused);  pass 1: in = dy, out = dx
int conv_thinout_launch(const srgan_conv_desc* d, int pass, const float* in, const float* w, const float* bias,
                        float* out, int act, float slope, void* ws, size_t ws_bytes, cudaStream_t st) {
  ThinOutPlan t;
  if (!thinout_plan(d, pass, &t)) { set_error("thin-output conv: unsupported shape"); return SRGAN_E_UNSUPPORTED; }
  const size_t need = conv_thinout_workspace(d, pass);
  if (!ws || ws_bytes < need) { set_error("thin-output conv: workspace %zu < %zu", ws_bytes, need); return SRGAN_E_WORKSPACE; }
  if (((uintptr_t)in | (uintptr_t)ws) % 16) { set_error("thin-output conv: tensors must be 16-byte aligned"); return SRGAN_E_BADARG; }
  float* bp = (float*)ws;
  static const char* e_var = getenv("SRGAN_DBG_THINOUT");       // bring-up: 1 = transpose epilogue, 2 = MMA column shift
  const int variant = e_var ? atoi(e_var) : 2;
  if (variant >= 2) return thinout2_launch<float>(d, t, in, w, bias, out, act, slope, ws, st);
  thinout_pack_filter_kernel<<<ceil_div(d->R * 32 * t.F, 256), 256, 0, st>>>(w, bp, d->K, d->C, d->R, d->S, t.mode);
  CUtensorMap min, mb;
  {
    uint64_t dims[4] = {(uint64_t)t.F, (uint64_t)t.Wi, (uint64_t)t.Hi, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)t.F * 4, (uint64_t)t.Wi * t.F * 4, (uint64_t)t.Hi * t.Wi * t.F * 4};
    uint32_t box[4] = {32, 128, 1, 1};
    if (int e = encode_map(&min, in, 4, dims, str, box)) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)t.F, (uint64_t)d->R * 32};
    uint64_t str[1] = {(uint64_t)t.F * 4};
    uint32_t box[2] = {32, (uint32_t)d->R * 32};
    if (int e = encode_map(&mb, bp, 2, dims, str, box)) return e;
  }
  ThinOutP p = {};
  p.Hi = t.Hi; p.Wi = t.Wi; p.Ho = t.Ho; p.Wo = t.Wo; p.tc = t.tc; p.R = d->R; p.S = d->S;
  p.padH = t.padH; p.padW = t.padW; p.nchunks = t.F / 32; p.bands = t.bands; p.BH = t.BH;
  p.act = act; p.slope = slope;
  // D = f32, A = B = tf32, both K-major, N = 32*R, M = 128
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((d->R * 32) >> 3) << 17) | ((128u >> 4) << 24);
  const size_t smem = 1024 + (size_t)p.nchunks * d->R * 4096 + (size_t)kTOStages * p.nchunks * 16384 +
                      (2 * kTOZBuf + 8 * 4 * 128) * sizeof(float) + 256;
  static unsigned long long attr_done = 0;
  {
    cudaError_t e = ensure_dyn_smem(conv_thinout_kernel, 227 * 1024, &attr_done);
    if (e != cudaSuccess) { set_error("conv_thinout smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  conv_thinout_kernel<<<d->N * t.bands, kTOThreads, smem, st>>>(min, mb, p, bias, out);
  SRGAN_RETURN_LAUNCH();
}

}  // namespace srgan
