// tcgen05 / TMEM / TMA implicit-GEMM convolution engine (TF32 operands straight from fp32 NHWC storage,
// fp32 accumulation in tensor memory).  sm_100a only.
//
// One CTA computes a [128 output pixels] x [BN output channels] tile:
//   D[m][n] = sum over taps t, channel chunks c :  A_t[m][c..c+32) . B_t[n][c..c+32)
//   A_t : a TMA box {32 ch, bw, 1, bh, bn} of the activation (viewed 5-D, see below) shifted by the tap
//         offset; out-of-bounds coordinates are zero-filled by TMA  == zero padding for free;
//   B_t : a TMA box {32 ch, 1 tap, BN filters} of the KRSC filter (or its transposed copy);
//   both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes directly.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator), warps 2..5 = epilogue
// (tcgen05.ld -> bias / activation -> global).  smem ring of kStages (full/empty mbarriers),
// tcgen05.commit releases a stage when the MMAs that read it have retired.
//
// The same kernel serves, through a per-launch "tap table" and output-pixel mapping:
//   fprop stride 1        taps (r,s) at offsets (r-pad, s-pad)
//   fprop stride 2        activation viewed as (2C, W/2, 2, H/2, N): a tap selects a row/column parity
//   dgrad stride 1        A = dy, filter transposed to [C][taps][K], taps at offsets (pad-r, pad-s)
//   dgrad stride 2 /      4 output-parity classes (blockIdx.z), each a 2x2-tap stride-1 problem on dy,
//   conv-transpose fwd    written to every second output pixel
#include <stdlib.h>
#include <cuda_bf16.h>
#include "umma_ptx.cuh"

namespace srgan {

// Storage type of the activations / filters of a launch.  Everything in the kernels is laid out in BYTES (a stage
// row is one 128-byte swizzle atom, an MMA K-step is 32 bytes), so the storage type only changes how many reduction
// channels a row holds, the operand format of the instruction descriptor and the MMA kind, and the epilogue's stores.
template <typename ST> struct UmmaElem;
template <> struct UmmaElem<float> {
  static constexpr int kRow = 32;                       // reduction channels per 128-byte row
  static constexpr uint32_t kFmt = 2;                   // TF32
  static constexpr CUtensorMapDataType kTma = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
};
template <> struct UmmaElem<__nv_bfloat16> {
  static constexpr int kRow = 64;
  static constexpr uint32_t kFmt = 1;                   // BF16 (kind::f16)
  static constexpr CUtensorMapDataType kTma = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
};
template <typename ST>
__device__ __forceinline__ void umma_any(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (sizeof(ST) == 4) umma_tf32(d, a, b, idesc, acc);
  else umma_f16(d, a, b, idesc, acc);
}

constexpr int kUmmaThreads = 192;
constexpr int kMaxTaps = 64;
constexpr int kABytes = 128 * 128;          // 128 pixel rows x 128 B (32 fp32 channels)

struct UmmaConvP {
  int c_chunks;                  // reduction channels / 32
  int tiles_w, tiles_h, tiles_n; // M-tile grid over the (N, P, Q) pixel grid of this launch
  int lw, lh;                    // log2 of box width / height (bw * bh * bn == 128)
  int Nn, P, Q;                  // extents of the pixel grid (validity masks)
  int out_H, out_W, out_C;       // output tensor [N][out_H][out_W][out_C]
  int os;                        // output pixel scale (2 for the parity-class launches)
  int K;                         // valid output channels (columns)
  int act;
  float slope;
  int epi_vec;                   // widest aligned vector store of a row: 8, 4 or 0 (scalar) floats
  int cls_oph[4], cls_opw[4];    // output pixel parity of each class
  int gx, gy, gz;                // work items: tile groups (MT tiles each) x filter tiles x parity classes
  int cls_fast;                  // 1: the class is the fastest index of an item (see umma_decode)
  int n_full, n_items;           // the first n_full items are whole tiles; the rest are half-width (BN/2 column)
                                 // tiles, two per remaining tile (balances the last partial wave of CTAs)
  int tap_begin[5];
  int4 taps[kMaxTaps];           // {channel offset, dw, hp | (filter tap << 8), dh}
};

// MT = number of 128-pixel M sub-tiles a CTA accumulates against ONE filter tile per stage: the filter bytes are
// amortised over MT*128 pixels (TF32 operands are 4 bytes, so a 128x256 tile needs 96 B/clk of shared-memory
// fill to keep the tensor pipe busy, a 256x256 tile 64 B/clk).
template <int BN, int MT = 1, typename ST = float>
struct UmmaCfg {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = MT * kABytes + kBBytes;
#ifndef SRGAN_DBG_MAX_STAGES
#define SRGAN_DBG_MAX_STAGES 8          // bring-up: -DSRGAN_DBG_MAX_STAGES=3 shows how the ring depth bounds the kernel
#endif
  static constexpr int kFit = (204 * 1024) / kStageBytes;
  static constexpr int kStages = kFit < SRGAN_DBG_MAX_STAGES ? kFit : SRGAN_DBG_MAX_STAGES;
  static constexpr int kAccBufs = 2 * MT * BN <= 512 ? 2 : 1;   // accumulator double buffering when TMEM allows
  static constexpr int kTmemCols = kAccBufs * MT * BN < 32 ? 32 : kAccBufs * MT * BN;
  // STATS: per-quadrant column sums of one tile [4][BN][2] + a private 32 x 17 transpose pad per epilogue warp
  static constexpr size_t kStatBytes = 4 * BN * 2 * sizeof(float) + 4 * 32 * 17 * sizeof(float);
  static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kStatBytes;
  // TMA issue: measured with tools/tma_probe.cu, the bulk-tensor loads of ONE warp execute back to back (~750-1100
  // clk each, whatever their size) while loads of different warps overlap.  A stage is therefore cut into kBoxes
  // boxes of <= 16 KB (MT activation sub-tiles + filter rows in blocks of <= 128), and 2*kBoxes producer warps each
  // own one box of every second stage: no warp issues more often than once per two stages.
  static constexpr int kBRows = BN < 128 ? BN : 128;
  static constexpr int kBoxes = MT + BN / kBRows;
  static constexpr int kProducers = 2 * kBoxes;
  static constexpr int kThreads = 32 * (5 + kProducers);   // warp 0 MMA, warps 1-4 epilogue, then the producers
  // instruction descriptor: D=f32, A=B=tf32 or bf16, K-major both, N=BN, M=128
  static constexpr uint32_t kIdesc = (1u << 4) | (UmmaElem<ST>::kFmt << 7) | (UmmaElem<ST>::kFmt << 10) |
                                     ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
};

// Persistent: gridDim.x CTAs (one per SM) walk over the work items (tile group, filter tile, parity class) with a
// stride of gridDim.x.  The smem stage ring runs across items without draining; with 2*MT*BN <= 512 TMEM columns
// the accumulator is double buffered, so the epilogue of item i overlaps the MMAs of item i+1 (this is what
// matters for short reductions: a stride-2 dgrad class has 4 taps, the packed RGB stem 7 stages in total).
// ---- epilogue helpers: straight-line code per activation (a per-element switch made the four epilogue warps the
// bottleneck of every short-reduction layer), 32-byte stores so a thread always writes whole sectors
template <int ACT>
__device__ __forceinline__ float act_t(float v, float slope) {
  if (ACT == SRGAN_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == SRGAN_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (ACT == SRGAN_ACT_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ void st_global_v8(float* p, const float* o) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(o[0]), "f"(o[1]), "f"(o[2]),
               "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
}
// act(v[0..N) + bias) -> dst[0..N), N a multiple of 8, dst aligned to VEC floats
template <int ACT, int N, int VEC>
__device__ __forceinline__ void epi_row_chunk(const float (&v)[32], float* __restrict__ dst,
                                              const float* __restrict__ bias, float slope,
                                              const float* __restrict__ add = nullptr) {
#pragma unroll
  for (int j = 0; j < N; j += 8) {
    float o[8];
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    if (add) {                      // same layout as dst: the other gradient that flows into this tensor
      a0 = __ldg(reinterpret_cast<const float4*>(add + j));
      a1 = __ldg(reinterpret_cast<const float4*>(add + j + 4));
    }
    if (bias) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + j));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + j + 4));
      o[0] = v[j] + b0.x; o[1] = v[j + 1] + b0.y; o[2] = v[j + 2] + b0.z; o[3] = v[j + 3] + b0.w;
      o[4] = v[j + 4] + b1.x; o[5] = v[j + 5] + b1.y; o[6] = v[j + 6] + b1.z; o[7] = v[j + 7] + b1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = v[j + e];
    }
    o[0] += a0.x; o[1] += a0.y; o[2] += a0.z; o[3] += a0.w;
    o[4] += a1.x; o[5] += a1.y; o[6] += a1.z; o[7] += a1.w;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = act_t<ACT>(o[e], slope);
    if (VEC == 8) {
      st_global_v8(dst + j, o);
    } else {
      *reinterpret_cast<float4*>(dst + j) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + j + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}
// ---- bf16 storage: the same row chunk packed two outputs per 32-bit word; VEC == 8: 32-byte stores (16 outputs),
// VEC == 4: 16-byte stores (8 outputs)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float* o) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { o[2 * e] = __uint_as_float(w[e] << 16); o[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
}
__device__ __forceinline__ void st_global_v8_b32(void* p, const uint32_t* o) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]),
               "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
}
template <int ACT, int N, int VEC>
__device__ __forceinline__ void epi_row_chunk(const float (&v)[32], __nv_bfloat16* __restrict__ dst,
                                              const float* __restrict__ bias, float slope,
                                              const __nv_bfloat16* __restrict__ add = nullptr) {
#pragma unroll
  for (int j = 0; j < N; j += 16) {
    float o[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) o[e] = v[j + e];
    if (bias) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + j + 4 * q));
        o[4 * q] += b.x; o[4 * q + 1] += b.y; o[4 * q + 2] += b.z; o[4 * q + 3] += b.w;
      }
    }
    if (add) {                      // same layout as dst: the other gradient that flows into this tensor
      float a[16];
      unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(add + j)), a);
      unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(add + j + 8)), a + 8);
#pragma unroll
      for (int e = 0; e < 16; ++e) o[e] += a[e];
    }
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = pack_bf16x2(act_t<ACT>(o[2 * e], slope), act_t<ACT>(o[2 * e + 1], slope));
    if (VEC == 8) {
      st_global_v8_b32(dst + j, w);
    } else {
      *reinterpret_cast<uint4*>(dst + j) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(dst + j + 8) = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
}
__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <int N, typename ST>
__device__ __forceinline__ void epi_row_chunk_any(const float (&v)[32], ST* dst, const float* bias, int act,
                                                  float slope, int vec, const ST* add = nullptr) {
  if (add) {                                     // fused gradient accumulation: no bias, no activation
    if (vec == 8) epi_row_chunk<SRGAN_ACT_NONE, N, 8>(v, dst, nullptr, 0.f, add);
    else epi_row_chunk<SRGAN_ACT_NONE, N, 4>(v, dst, nullptr, 0.f, add);
    return;
  }
#define SRGAN_EPI_CASE(A)                                                         \
  case A:                                                                         \
    if (vec == 8) epi_row_chunk<A, N, 8>(v, dst, bias, slope);                    \
    else epi_row_chunk<A, N, 4>(v, dst, bias, slope);                             \
    break;
  switch (act) {
    SRGAN_EPI_CASE(SRGAN_ACT_RELU)
    SRGAN_EPI_CASE(SRGAN_ACT_LRELU)
    SRGAN_EPI_CASE(SRGAN_ACT_TANH)
    default:
      if (vec == 8) epi_row_chunk<SRGAN_ACT_NONE, N, 8>(v, dst, bias, slope);
      else epi_row_chunk<SRGAN_ACT_NONE, N, 4>(v, dst, bias, slope);
  }
#undef SRGAN_EPI_CASE
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // 4 epilogue warps

// work item -> linear tile index, first column and width of the column range it covers
struct UmmaItem { int lin, n_off, width; };
template <int BN>
__device__ __forceinline__ UmmaItem umma_item(const UmmaConvP& p, int item) {
  UmmaItem u;
  if (item < p.n_full) { u.lin = item; u.n_off = 0; u.width = BN; }
  else { const int h = item - p.n_full; u.lin = p.n_full + (h >> 1); u.n_off = (h & 1) * (BN / 2); u.width = BN / 2; }
  return u;
}

// linear tile index -> (tile group, filter tile, parity class).  With several classes (stride-2 dgrad / transposed
// forward: 4 classes that read the SAME input boxes through different 2 x 2 taps) the class is the fastest index, so the
// CTAs that run at the same time work on the four classes of the same tiles and the input is fetched from HBM once
// instead of once per class (class-major order: 198 MB read for a 67 MB input on G.up1, profiles/r2s_*).
__device__ __forceinline__ void umma_decode(const UmmaConvP& p, int lin, int* bx, int* by, int* cls) {
  if (p.cls_fast) {
    *cls = lin % p.gz;
    const int r = lin / p.gz;
    *bx = r % p.gx;
    *by = r / p.gx;
  } else {
    *bx = lin % p.gx;
    *by = (lin / p.gx) % p.gy;
    *cls = lin / (p.gx * p.gy);
  }
}

// STATS: the epilogue also writes, per 128-pixel tile and output channel, the sum and the sum of squares of the values
// it stores (as stored, i.e. after rounding to ST): stats[(tile_row * K + k) * 2 + {0, 1}], tile_row =
// ((n * classes + cls) * tiles_h + th) * tiles_w + tw - what the instance norm that follows the convolution needs,
// without its own pass over the tensor (srgan_inorm_stats_from_tiles folds the rows of an image in fp64).  Needs tiles
// that lie inside one image (box of 128 pixels of one image), no bias / activation / addend.
// OT: storage type of the OUTPUT (and of the addend); differs from ST only on the thin RGB layers of the bf16 engine
// (TF32 operands from the fp32 image, bf16 result for the bf16 trunk - and the mirrored head dgrad).
template <int BN, int MT, typename ST, bool STATS = false, typename OT = ST>
__global__ void __launch_bounds__((UmmaCfg<BN, MT, ST>::kThreads), 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ UmmaConvP p, const float* __restrict__ bias, OT* __restrict__ y,
                 const OT* __restrict__ addend, float* __restrict__ stats) {
  using Cfg = UmmaCfg<BN, MT, ST>;
  constexpr int NBUF = Cfg::kAccBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tfull = empty + Cfg::kStages;          // [NBUF] accumulator complete
  uint64_t* tempty = tfull + 2;                    // [NBUF] accumulator drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sstat = reinterpret_cast<float*>(smem + (size_t)Cfg::kStages * Cfg::kStageBytes + 256);   // [4][BN][2]
  float* wpad = sstat + 4 * BN * 2;                                                                // [4][32][17]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bw = 1 << p.lw, bh = 1 << p.lh;
  const int bn = 128 >> (p.lw + p.lh);
  const int items = p.n_items;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full + s, Cfg::kBoxes); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 5) {
    // ------------------------------------------------------------------ TMA producers (one box each, every 2nd stage)
    if (lane == 0) {
      const int pw = warp - 5;
      const int box = pw % Cfg::kBoxes, par = pw / Cfg::kBoxes;
      int gi0 = 0;                                   // ring index of the first stage of the current item
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const UmmaItem u = umma_item<BN>(p, item);
        int bx, by, cls;
        umma_decode(p, u.lin, &bx, &by, &cls);
        const int tap0 = p.tap_begin[cls];
        const int iters = (p.tap_begin[cls + 1] - tap0) * p.c_chunks;
        int qb = 0, pb = 0, nb = 0;
        if (box < MT) {
          int t = bx * MT + box;
          const int tw = t % p.tiles_w; t /= p.tiles_w;
          qb = tw * bw; pb = (t % p.tiles_h) * bh; nb = (t / p.tiles_h) * bn;
        }
        for (int it = (gi0 + par) & 1; it < iters; it += 2) {       // stages with (gi0 + it) % 2 == par
          const int gi = gi0 + it;
          const int stage = gi % Cfg::kStages;
          const uint32_t phase = (uint32_t)(gi / Cfg::kStages) & 1u;
          const int4 tp = p.taps[tap0 + it / p.c_chunks];
          const int cc = (it % p.c_chunks) * UmmaElem<ST>::kRow;
          mbar_wait(empty + stage, phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * Cfg::kStageBytes;
          if (box < MT) {
            mbar_expect_tx(full + stage, kABytes);
            tma_load_5d(&map_a, full + stage, sa + box * kABytes, cc + tp.x, qb + tp.y, tp.z & 0xff, pb + tp.w, nb);
          } else {
            const int rb = (box - MT) * Cfg::kBRows;
            if (rb < u.width) {
              mbar_expect_tx(full + stage, Cfg::kBRows * 128);
              tma_load_3d(&map_b, full + stage, sa + MT * kABytes + rb * 128, cc, tp.z >> 8, by * BN + u.n_off + rb);
            } else {
              mbar_arrive(full + stage);         // half-width item: this filter box is not needed
            }
          }
        }
        gi0 += iters;
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      int stage = 0, li = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++li) {
        const UmmaItem u = umma_item<BN>(p, item);
        int bx_, by_, cls;
        umma_decode(p, u.lin, &bx_, &by_, &cls);
        const int iters = (p.tap_begin[cls + 1] - p.tap_begin[cls]) * p.c_chunks;
        const uint32_t idesc = (Cfg::kIdesc & ~(0x3Fu << 17)) | ((uint32_t)(u.width >> 3) << 17);
        const int buf = li % NBUF;
        mbar_wait(tempty + buf, (((uint32_t)(li / NBUF)) & 1u) ^ 1u);       // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * (MT * BN);
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::kStageBytes);
          const uint64_t bdesc = smem_desc_sw128(sa + MT * kABytes);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = smem_desc_sw128(sa + mt * kABytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 4 x (K = 8 tf32 / 16 bf16 = 32 bytes) per 128-byte row; +32 B = +2 in the address field
              umma_any<ST>(acc + mt * BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0);
          }
          umma_commit(empty + stage);     // stage reusable once these MMAs have read it
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull + buf);         // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may read
    const int m = quad * 32 + lane;                    // accumulator row == pixel within the tile
    const int wl = m & (bw - 1), hl = (m >> p.lw) & (bh - 1), nl = m >> (p.lw + p.lh);
    constexpr int kChunk = BN >= 32 ? 32 : 16;
    int li = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++li) {
      const UmmaItem u = umma_item<BN>(p, item);
      int bx, by, cls;
      umma_decode(p, u.lin, &bx, &by, &cls);
      const int col0 = by * BN + u.n_off;
      const int buf = li % NBUF;
      mbar_wait(tfull + buf, ((uint32_t)(li / NBUF)) & 1u);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int t = bx * MT + mt;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int n = (t / p.tiles_h) * bn + nl, pp = (t % p.tiles_h) * bh + hl, qq = tw * bw + wl;
        const bool valid = n < p.Nn && pp < p.P && qq < p.Q;
        OT* yrow = y + (((size_t)n * p.out_H + (size_t)(pp * p.os + p.cls_oph[cls])) * p.out_W +
                        (size_t)(qq * p.os + p.cls_opw[cls])) * p.out_C;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (MT * BN) + mt * BN;
        const float* brow = bias ? bias + col0 : nullptr;
        const OT* arow = addend ? addend + (yrow - y) + col0 : nullptr;         // addend has the layout of y
        if (kChunk == 32) {
          // the TMEM load of chunk c + 1 is in flight while chunk c is converted and stored
          uint32_t ra[32], rb[32];
          tmem_ld32_issue(taddr, ra);
#pragma unroll
          for (int c = 0; c < BN; c += 32) {
            if (c >= u.width) break;
            uint32_t (&cur)[32] = ((c >> 5) & 1) ? rb : ra;
            uint32_t (&nxt)[32] = ((c >> 5) & 1) ? ra : rb;
            tmem_ld_wait();
            if (c + 32 < u.width) tmem_ld32_issue(taddr + c + 32, nxt);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(cur[j]);
            if (STATS) {
              // Column sums over the warp's 32 pixel rows of the values as they are stored (rounded to ST; rows outside
              // the image count as zero), 16 columns at a time through a private 32 x 17 pad: every lane writes its
              // row, then adds 16 rows of one column (lanes 0-15: even rows, lanes 16-31: odd rows), one shuffle joins
              // the two halves.  ~240 instructions per 32 x 32 chunk (a shuffle butterfly costs ~1000: the epilogue
              // warps, not the tensor pipe, then set the pace of the kernel); fixed summation order.
              float* wst = wpad + quad * (32 * 17);
              const int scol = lane & 15, srow = lane >> 4;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float a = valid ? v[h * 16 + j] : 0.f;
                  if (sizeof(OT) == 2) a = __bfloat162float(__float2bfloat16_rn(a));
                  wst[lane * 17 + j] = a;
                }
                __syncwarp();
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int i2 = 0; i2 < 16; ++i2) {
                  const float a = wst[(2 * i2 + srow) * 17 + scol];
                  s1 += a;
                  s2 = fmaf(a, a, s2);
                }
                s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                if (lane < 16)
                  *reinterpret_cast<float2*>(sstat + ((quad * BN) + c + h * 16 + lane) * 2) = make_float2(s1, s2);
                __syncwarp();
              }
            }
            if (valid) {
              if (col0 + c + 32 <= p.K && p.epi_vec) {
                epi_row_chunk_any<32>(v, yrow + col0 + c, brow ? brow + c : nullptr, p.act, p.slope, p.epi_vec,
                                      arow ? arow + c : nullptr);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + c + j < p.K)
                    st_from_float(yrow + col0 + c + j,
                                  apply_act(v[j] + (bias ? __ldg(bias + col0 + c + j) : 0.f) +
                                            (arow ? ld_as_float(arow + c + j) : 0.f), p.act, p.slope));
              }
            }
          }
        } else {
          float v[32];
          tmem_ld16(taddr, v);
          if (valid) {
            if (col0 + 16 <= p.K && p.epi_vec) {
              epi_row_chunk_any<16>(v, yrow + col0, brow, p.act, p.slope, p.epi_vec, arow);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j < p.K)
                  st_from_float(yrow + col0 + j, apply_act(v[j] + (bias ? __ldg(bias + col0 + j) : 0.f) +
                                                           (arow ? ld_as_float(arow + j) : 0.f), p.act, p.slope));
            }
          }
        }
        if (STATS && kChunk == 32) {
          // the four quadrants (32 pixel rows each) of this tile, added in quadrant order; bn == 1: the tile lies in
          // image t / (tiles_w * tiles_h)
          epi_bar_sync();
          int tt = bx * MT + mt;
          const int tw2 = tt % p.tiles_w; tt /= p.tiles_w;
          const int th2 = tt % p.tiles_h, img = tt / p.tiles_h;
          const size_t prow = (((size_t)img * p.gz + cls) * p.tiles_h + th2) * p.tiles_w + tw2;
          if (img < p.Nn) {
            for (int col = quad * 32 + lane; col < u.width; col += 128) {
              float a = 0.f, b = 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) { a += sstat[(q * BN + col) * 2]; b += sstat[(q * BN + col) * 2 + 1]; }
              if (col0 + col < p.K)
                *reinterpret_cast<float2*>(stats + (prow * p.K + col0 + col) * 2) = make_float2(a, b);
            }
          }
          epi_bar_sync();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + buf);
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------ CTA pairs
// 256-wide filter tiles on CTA PAIRS (tcgen05 cta_group::2): a cluster of two CTAs computes a 256-pixel x 256-channel
// tile.  Each CTA loads ITS 128 pixel rows of A and HALF of the filter tile (128 of the 256 output channels); the
// leader's MMAs (M = 256) read both shared memories and leave 128 accumulator rows in each CTA's tensor memory; each
// CTA runs the epilogue of its own rows.  Why: the one-CTA kernel is bound by shared-memory bandwidth - per 128-clock
// MMA it reads 4 KB of A + 8 KB of B while TMA writes the same 12 KB for the next stage: 192 B/clk against a 128 B/clk
// port (tensor pipe active ~70 % of the time, profiles/r1q_res_ncu_raw.csv) - and, one level up, by the L2 -> SM
// traffic of re-reading the whole filter for every 128 pixels (14 TB/s on the residual block).  A pair halves the B
// bytes per CTA on both paths: 64 + 64 B/clk.
//   barriers: full[s]   leader's; its producer warps arrive with expect_tx for BOTH CTAs' boxes, the peer's TMA
//                       (cta_group::2) counts its bytes there;
//             empty[s]  one per CTA, both released by the leader's multicast tcgen05.commit;
//             tfull[b]  one per CTA (multicast commit), tempty[b] leader's (4 epilogue warps x 2 CTAs arrive).
template <typename ST> struct Umma2Cfg {
  static constexpr int kBN = 256, kBHalf = 128;
  static constexpr int kStageBytes = kABytes + kBHalf * 128;                 // 16 KB of A + 16 KB of B per CTA
  static constexpr int kStages = 6;
  static constexpr int kBoxes = 2;                                           // A box, B-half box
  static constexpr int kProducers = 2 * kBoxes;
  // warp 0: MMA issuer; warps 1-8: epilogue (two warps per TMEM lane quadrant, each takes half of the columns: with
  // bf16 operands the MMAs of a tile take ~20 k clocks and ONE warp per quadrant needs longer than that to drain and
  // store 32 rows x 256 columns - the epilogue, not the tensor pipe, set the pace); warps 9-12: TMA producers
  static constexpr int kEpiWarps = 8;
  static constexpr int kThreads = 32 * (1 + kEpiWarps + kProducers);
  static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024 + 256;
  // instruction descriptor: D = f32, A = B = tf32 | bf16, K-major both, N = 256, M = 256 (128 rows per CTA)
  static constexpr uint32_t kIdesc = (1u << 4) | (UmmaElem<ST>::kFmt << 7) | (UmmaElem<ST>::kFmt << 10) |
                                     ((uint32_t)(256 >> 3) << 17) | ((256u >> 4) << 24);
};
template <typename ST>
__device__ __forceinline__ void umma_any_2sm(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (sizeof(ST) == 4) umma_tf32_2sm(d, a, b, idesc, acc);
  else umma_f16_2sm(d, a, b, idesc, acc);
}

// p.gx = PAIRS of 128-pixel tiles per (filter tile, class); work item = (pair, filter tile, class)
template <typename ST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((Umma2Cfg<ST>::kThreads), 1)
conv_umma2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ UmmaConvP p, const float* __restrict__ bias, ST* __restrict__ y,
                  const ST* __restrict__ addend) {
  using Cfg = Umma2Cfg<ST>;
  constexpr int BN = Cfg::kBN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tfull = empty + Cfg::kStages;          // [2]
  uint64_t* tempty = tfull + 2;                    // [2] (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int bw = 1 << p.lw, bh = 1 << p.lh;
  const int bn = 128 >> (p.lw + p.lh);
  const int items = p.n_items;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full + s, Cfg::kBoxes); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 2 * Cfg::kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
  }
  if (warp == 0) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // both CTAs' barriers and tensor memory exist
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp > Cfg::kEpiWarps) {
    // ------------------------------------------------------------------ TMA producers (one box each, every 2nd stage)
    if (lane == 0) {
      const int pw = warp - 1 - Cfg::kEpiWarps;
      const int box = pw % Cfg::kBoxes, par = pw / Cfg::kBoxes;
      int gi0 = 0;
      for (int item = cluster_id; item < items; item += nclusters) {
        const UmmaItem u = umma_item<BN>(p, item);                   // tail items cover half of the filter tile
        const int bx = u.lin % p.gx, by = (u.lin / p.gx) % p.gy, cls = u.lin / (p.gx * p.gy);
        const int brow = by * BN + u.n_off + (int)rank * (u.width >> 1);    // this CTA's half of the item's filters
        const int tap0 = p.tap_begin[cls];
        const int iters = (p.tap_begin[cls + 1] - tap0) * p.c_chunks;
        int t = bx * 2 + (int)rank;                                  // this CTA's 128-pixel tile of the pair
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int qb = tw * bw, pb = (t % p.tiles_h) * bh, nb = (t / p.tiles_h) * bn;
        for (int it = (gi0 + par) & 1; it < iters; it += 2) {
          const int gi = gi0 + it;
          const int stage = gi % Cfg::kStages;
          const uint32_t phase = (uint32_t)(gi / Cfg::kStages) & 1u;
          const int4 tp = p.taps[tap0 + it / p.c_chunks];
          const int cc = (it % p.c_chunks) * UmmaElem<ST>::kRow;
          mbar_wait(empty + stage, phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * Cfg::kStageBytes;
          // the leader's barrier counts this box of BOTH CTAs (same size, OOB rows are zero-filled but counted)
          if (leader) mbar_expect_tx(full + stage, 2 * kABytes);
          if (box == 0)
            tma_load_5d_2sm(&map_a, full + stage, sa, cc + tp.x, qb + tp.y, tp.z & 0xff, pb + tp.w, nb);
          else
            tma_load_3d_2sm(&map_b, full + stage, sa + kABytes, cc, tp.z >> 8, brow);   // tail items use 64 of the 128 rows
        }
        gi0 += iters;
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ MMA issuer (one thread of the leader)
    if (leader && lane == 0) {
      int stage = 0, li = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < items; item += nclusters, ++li) {
        const UmmaItem u = umma_item<BN>(p, item);
        const int cls = u.lin / (p.gx * p.gy);
        const int iters = (p.tap_begin[cls + 1] - p.tap_begin[cls]) * p.c_chunks;
        const uint32_t idesc = (Cfg::kIdesc & ~(0x3Fu << 17)) | ((uint32_t)(u.width >> 3) << 17);
        const int buf = li & 1;
        mbar_wait(tempty + buf, (((uint32_t)(li >> 1)) & 1u) ^ 1u);       // both epilogues drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * BN;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::kStageBytes);
          const uint64_t adesc = smem_desc_sw128(sa), bdesc = smem_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_any_2sm<ST>(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0);
          umma_commit_2sm(empty + stage, 3);       // both CTAs may refill this stage
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(tfull + buf, 3);           // accumulator complete, in both CTAs
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue of this CTA's 128 rows
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may read (warp id % 4)
    const int half = (warp - 1) >> 2;                  // which half of the item's columns this warp drains
    const int m = quad * 32 + lane;
    const int wl = m & (bw - 1), hl = (m >> p.lw) & (bh - 1), nl = m >> (p.lw + p.lh);
    const uint32_t tempty_leader = mapa_shared(smem_u32(tempty), 0);
    int li = 0;
    for (int item = cluster_id; item < items; item += nclusters, ++li) {
      const UmmaItem u = umma_item<BN>(p, item);
      const int bx = u.lin % p.gx, by = (u.lin / p.gx) % p.gy, cls = u.lin / (p.gx * p.gy);
      const int col0 = by * BN + u.n_off;
      const int buf = li & 1;
      mbar_wait(tfull + buf, ((uint32_t)(li >> 1)) & 1u);
      tc_fence_after();
      int t = bx * 2 + (int)rank;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int n = (t / p.tiles_h) * bn + nl, pp = (t % p.tiles_h) * bh + hl, qq = tw * bw + wl;
      const bool valid = n < p.Nn && pp < p.P && qq < p.Q;
      ST* yrow = y + (((size_t)n * p.out_H + (size_t)(pp * p.os + p.cls_oph[cls])) * p.out_W +
                      (size_t)(qq * p.os + p.cls_opw[cls])) * p.out_C;
      const int hw = u.width >> 1, cbeg = half * hw;                  // this warp's columns [cbeg, cbeg + hw)
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN + cbeg;
      const float* brow = bias ? bias + col0 + cbeg : nullptr;
      const ST* arow = addend ? addend + (yrow - y) + col0 + cbeg : nullptr;
      yrow += cbeg;
      uint32_t ra[32], rb[32];
      tmem_ld32_issue(taddr, ra);
#pragma unroll
      for (int c = 0; c < BN / 2; c += 32) {
        if (c >= hw) break;
        uint32_t (&cur)[32] = ((c >> 5) & 1) ? rb : ra;
        uint32_t (&nxt)[32] = ((c >> 5) & 1) ? ra : rb;
        tmem_ld_wait();
        if (c + 32 < hw) tmem_ld32_issue(taddr + c + 32, nxt);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(cur[j]);
        if (valid) {
          if (col0 + cbeg + c + 32 <= p.K && p.epi_vec) {
            epi_row_chunk_any<32>(v, yrow + col0 + c, brow ? brow + c : nullptr, p.act, p.slope, p.epi_vec,
                                  arow ? arow + c : nullptr);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + cbeg + c + j < p.K)
                st_from_float(yrow + col0 + c + j,
                              apply_act(v[j] + (brow ? __ldg(brow + c + j) : 0.f) +
                                        (arow ? ld_as_float(arow + c + j) : 0.f), p.act, p.slope));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + buf * 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // nobody leaves while its partner still needs its memories
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------ wgrad
// dW[k][t][c] = sum over pixels m :  dY[m][k] * X_t[m][c]          (X_t = x shifted by tap t, zero outside)
// Both operands are "MN-major": the reduction index (pixel) is the slow dimension in memory.  A stage holds a
// chunk of 32 pixels: KT*4 dY boxes {32 k, pixel box} (KT sub-tiles of 128 filters) and up to BN/32 X boxes
// {32 c, pixel box}, each a 4 KB swizzled region whose rows are pixels.  UMMA descriptors: MN-block stride (LBO)
// 4096 B, 4-pixel group stride (SBO) 512 B.
// The N dimension of one MMA is a GROUP OF TAPS x channels: for C < 256 a CTA accumulates gt consecutive taps
// (N = gt*C <= 256) against the same dY chunk, so dY is loaded and read once per gt taps; the accumulator columns
// [j*C, (j+1)*C) are exactly dW[k][tap0+j][0..C), contiguous in the KRSC gradient.  For C >= 256, gt = 1 and the
// channels are tiled by 256.  grid = (k tiles * c tiles, tap groups, pixel splits); splits write partial sums that
// a fixed-order reduction kernel adds (deterministic).
struct UmmaWgradP {
  int tiles_c;                   // c tiles per k tile (blockIdx.x = kt * tiles_c + ct)
  int tiles_w, tiles_h, tiles_n; // pixel-chunk grid
  int lw, lh;                    // log2 chunk box width / height (bw*bh*bn == 32)
  int chunks, chunks_per_split;
  int K, C, T;                   // dW[K][T][C]  (row stride of the output = T*C)
  int gt;                        // taps per group
  int cpb;                       // 32-channel blocks per tap inside one CTA's N range
  long long split_stride;        // elements between partial results
  int vec8;                      // every output row segment is 32-byte aligned (set by launch_wgrad)
  unsigned long long desc_hi;    // descriptor bits above the start address (LBO, SBO, version, layout type)
  int4 taps[kMaxTaps];           // {channel offset, dw, hp, dh} of x for each filter tap
};

// MN-major TF32 operands only exist in the "128B swizzle, 32-byte atom" layout (UMMA layout type
// SWIZZLE_128B_BASE32B <-> TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 128 B (32 MN elements), 32-byte
// chunks XOR-ed with (row & 3); canonical K blocks are 4 rows.
__host__ __device__ inline uint64_t mn_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)(lbo_bytes >> 4) << 16;           // leading byte offset: next block of 32 MN elements
  d |= (uint64_t)(sbo_bytes >> 4) << 32;           // stride byte offset: next group of 4 pixel rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;                     // 1 = SWIZZLE_128B_BASE32B
  return d;
}

template <int BN, int KT, int PW = 4>
struct WgradCfg {
  static constexpr int kStageBytes = KT * kABytes + BN * 128;
  static constexpr int kStages = (208 * 1024) / kStageBytes < 8 ? (208 * 1024) / kStageBytes : 8;
  static constexpr int kTmemCols = KT * BN < 32 ? 32 : KT * BN;
  static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024 + 256;
  static constexpr int kProducers = PW;                       // TMA-issuing warps (boxes dealt round robin)
  static constexpr int kThreads = 32 * (5 + kProducers);      // warp 0 MMA, warps 1-4 epilogue, then producers
};

// Storage-type constants of the MN-major operand boxes.  A box is {kCh channels (one 128-byte row), kPix pixels}:
//   fp32 / TF32 : 32 ch x 32 pixels = 4 KB, "128B swizzle with 32-byte atoms" (the only MN-major TF32 layout),
//                 canonical K block = 4 pixel rows (SBO 512), one MMA = 8 pixels = 1024 B further;
//   bf16        : 64 ch x 64 pixels = 8 KB, plain 128B swizzle (16-byte atoms), canonical K block = 8 pixel rows
//                 (SBO 1024), one MMA (K = 16) = 16 pixels = 2048 B further.
// Either way a 128-row operand tile of one pixel chunk is 16 KB (kABytes), a stage runs 4 MMAs per 128-row tile, and
// a 32-column accumulator chunk is 32 output channels.
template <typename ST> struct WgradElem;
template <> struct WgradElem<float> {
  static constexpr int kCh = 32, kPix = 32, kBoxBytes = 4096, kKStep = 64 /* 1024 B >> 4 */;
  static constexpr uint32_t kFmt = 2, kLbo = 4096, kSbo = 512, kLayout = 1;
  static constexpr CUtensorMapSwizzle kSwz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
};
template <> struct WgradElem<__nv_bfloat16> {
  static constexpr int kCh = 64, kPix = 64, kBoxBytes = 8192, kKStep = 128 /* 2048 B >> 4 */;
  static constexpr uint32_t kFmt = 1, kLbo = 8192, kSbo = 1024, kLayout = 2;
  static constexpr CUtensorMapSwizzle kSwz = CU_TENSOR_MAP_SWIZZLE_128B;
};

template <int BN, int KT, int PW, typename ST>
__global__ void __launch_bounds__((WgradCfg<BN, KT, PW>::kThreads), 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ UmmaWgradP p, float* __restrict__ out) {
  using Cfg = WgradCfg<BN, KT, PW>;
  using El = WgradElem<ST>;
  constexpr int kDyBoxes = 128 / El::kCh;             // dY boxes per 128-filter sub-tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tmem_full = empty + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x / p.tiles_c, ct = blockIdx.x % p.tiles_c;
  const int tap0 = blockIdx.y * p.gt;
  const int ntaps = min(p.gt, p.T - tap0);
  const int nblk = ntaps * p.cpb;                     // kCh-column N blocks (one TMA box each) of this CTA
  const int k0 = kt * 128 * KT, c0 = ct * BN;
  const int ch_beg = blockIdx.z * p.chunks_per_split;
  const int ch_end = min(p.chunks, ch_beg + p.chunks_per_split);
  const int iters = ch_end - ch_beg;
  const int bw = 1 << p.lw, bh = 1 << p.lh;
  const int bn = El::kPix >> (p.lw + p.lh);

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full + s, Cfg::kProducers); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_x) : "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 5) {
    if (lane == 0) {
      const int pw = warp - 5;
      const int nboxes = KT * kDyBoxes + nblk;
      int mine = 0;
      for (int b = pw; b < nboxes; b += Cfg::kProducers) ++mine;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        int ch = ch_beg + it;
        const int tw = ch % p.tiles_w; ch /= p.tiles_w;
        const int th = ch % p.tiles_h;
        const int tn = ch / p.tiles_h;
        const int q0 = tw * bw, p0 = th * bh, n0 = tn * bn;
        mbar_wait(empty + stage, phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * Cfg::kStageBytes;
        mbar_expect_tx(full + stage, mine * El::kBoxBytes);
        for (int b = pw; b < nboxes; b += Cfg::kProducers) {
          if (b < KT * kDyBoxes) {
            tma_load_5d(&map_dy, full + stage, sa + b * El::kBoxBytes, k0 + El::kCh * b, q0, 0, p0, n0);
          } else {
            const int j = b - KT * kDyBoxes;
            const int4 tj = p.taps[tap0 + j / p.cpb];
            tma_load_5d(&map_x, full + stage, sa + KT * kABytes + j * El::kBoxBytes, c0 + El::kCh * (j % p.cpb) + tj.x,
                        q0 + tj.y, tj.z, p0 + tj.w, n0);
          }
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      // D = f32, A = B = tf32 or bf16, both MN-major, N = kCh * nblk, M = 128
      const uint32_t idesc = (1u << 4) | (El::kFmt << 7) | (El::kFmt << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)((nblk * El::kCh) >> 3) << 17) | ((128u >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full + stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::kStageBytes);
        const uint64_t bdesc = p.desc_hi | (uint64_t)(((sa + KT * kABytes) >> 4) & 0x3FFF);
#pragma unroll
        for (int t = 0; t < KT; ++t) {
          const uint64_t adesc = p.desc_hi | (uint64_t)(((sa + t * kABytes) >> 4) & 0x3FFF);
#pragma unroll
          for (int k = 0; k < 4; ++k)     // 4 x (8 tf32 | 16 bf16) pixels, each 1024 | 2048 B further
            umma_any<ST>(tmem_base + t * BN, adesc + (uint64_t)(El::kKStep * k), bdesc + (uint64_t)(El::kKStep * k), idesc,
                         (it | k) != 0);
        }
        umma_commit(empty + stage);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);
    }
  } else {
    const int quad = warp & 3;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      const int k = k0 + t * 128 + quad * 32 + lane;     // accumulator row == filter index
      const bool valid = k < p.K && iters > 0;
      float* orow = out + (size_t)blockIdx.z * p.split_stride + ((size_t)k * p.T + tap0) * p.C + c0;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + t * BN;
      // two 32-column chunks per trip: the TMEM load of the next chunk is in flight while this one is stored
      const int ncols = nblk * El::kCh;
      auto put = [&](const uint32_t (&r)[32], int c) {
        if (!(valid && c0 + (c % (p.cpb * El::kCh)) < p.C)) return;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.vec8) epi_row_chunk<SRGAN_ACT_NONE, 32, 8>(v, orow + c, nullptr, 0.f);
        else epi_row_chunk<SRGAN_ACT_NONE, 32, 4>(v, orow + c, nullptr, 0.f);
      };
      uint32_t ra[32], rb[32];
      tmem_ld32_issue(taddr, ra);
#pragma unroll 1
      for (int c = 0; c < ncols; c += 64) {
        tmem_ld_wait();
        if (c + 32 < ncols) tmem_ld32_issue(taddr + c + 32, rb);
        put(ra, c);
        if (c + 32 < ncols) {
          tmem_ld_wait();
          if (c + 64 < ncols) tmem_ld32_issue(taddr + c + 64, ra);
          put(rb, c + 32);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// w[K][T][C] -> wt[C][T][K]  (filter transpose for the dgrad-shaped problems)
template <typename ST>
__global__ void filter_transpose_kernel(const ST* __restrict__ w, ST* __restrict__ wt, int K, int T, int C) {
  __shared__ ST tile[32][33 + (sizeof(ST) == 2 ? 1 : 0)];
  const int tp = blockIdx.z;
  const int k0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int k = k0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (k < K && c < C) ? w[((size_t)k * T + tp) * C + c] : ST(0.f);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, k = k0 + threadIdx.x;
    if (k < K && c < C) wt[((size_t)c * T + tp) * K + k] = tile[threadIdx.x][j];
  }
}

// ------------------------------------------------------------------------------------------ host side
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// choose the pixel box (bw x bh x bn = 128, powers of two) covering a P x Q grid
static void pick_box(int P, int Q, int* lw, int* lh) {
  int w = 1 << ilog2(Q);
  if (w > 128) w = 128;
  int h = 1 << ilog2(P);
  if (h > 128 / w) h = 128 / w;
  *lw = ilog2(w);
  *lh = ilog2(h);
}

static int pick_bn(int K) {
  if (K > 128) return 256;
  if (K > 64) return 128;
  if (K > 32) return 64;
  if (K > 16) return 32;
  return 16;
}

// CTA pairs for 256-wide tiles: SRGAN_CONV_PAIRS = 0 (off, default) | 1 (bf16 storage) | 2 (bf16 and TF32).
// Measured on B200 (tools/conv_k_probe.py, profiles/r2q_*): per 64-channel stage (4 MMAs = 512 tensor clocks) both
// kernels need ~390 ns in steady state - one CTA 734 "1.9 GHz clocks", pairs 770 - whatever the ring depth (3 stages:
// 779, 4: 734, pairs with 6: 770) and whatever the epilogue width (4 or 8 warps).  390 ns for 512 clocks is 1.32 GHz:
// the clock the driver's own cuBLAS peak measurement records under load (MEASURED_PEAKS.json sm_mhz_median 1320, 1000 W
// power limit).  In steady state the MMAs are therefore back to back in both kernels and the pair's halved operand
// traffic (ncu: shared-memory wavefronts 43 % -> 30 %, tensor pipe active 59 % -> 73 % of active cycles at the 1.8 GHz
// of a cold profiling run) buys nothing under the power cap, while its fixed cost per launch (cluster launch, two
// cluster barriers) is 4 us higher: residual block 62.4 vs 58.8 us back to back.  Off by default, kept selectable.
template <typename ST>
static bool use_cta_pairs(long tiles) {
  static const int mode = getenv("SRGAN_CONV_PAIRS") ? atoi(getenv("SRGAN_CONV_PAIRS")) : 0;
  if (mode <= 0 || (sizeof(ST) == 4 && mode < 2)) return false;
  return tiles >= 2 * kNumSMs;                      // at least two full waves of single tiles: every pair has work
}

template <typename ST>
static int launch_pairs(const CUtensorMap& ma, const CUtensorMap& mb_full, const UmmaConvP& p, const float* bias, ST* y,
                        dim3 grid, cudaStream_t st, const ST* addend) {
  using Cfg = Umma2Cfg<ST>;
  static unsigned long long attr_done = 0;
  static int max_clusters[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  {
    cudaError_t e = ensure_dyn_smem(conv_umma2_kernel<ST>, (int)Cfg::kSmem, &attr_done);
    if (e != cudaSuccess) { set_error("conv_umma2 smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  if (dev < 64 && max_clusters[dev] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * kNumSMs); cfg.blockDim = dim3(Cfg::kThreads); cfg.dynamicSmemBytes = Cfg::kSmem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv_umma2_kernel<ST>, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = kNumSMs / 2; }
    max_clusters[dev] = n;
  }
  const int nmax = dev < 64 ? max_clusters[dev] : kNumSMs / 2;
  UmmaConvP q = p;
  q.gx = (grid.x + 1) / 2; q.gy = grid.y; q.gz = grid.z;          // pairs of 128-pixel tiles
  long items = (long)q.gx * q.gy * q.gz;
  q.n_full = (int)items;
  // last, partial wave of clusters: when it fills at most half of them, its items run as twice as many half-width
  // items (128 of the 256 filters; both CTAs then use 64 rows of their filter box) - see launch_bn
  const long rem = items % nmax;
  if (rem > 0 && 2 * rem <= nmax) { q.n_full = (int)(items - rem); items += rem; }
  q.n_items = (int)items;
  const unsigned clusters = (unsigned)(items < nmax ? items : nmax);
  // the filter map of the pair kernel: boxes of 128 filters (half a tile) - same geometry as mb_full's kBRows box
  conv_umma2_kernel<ST><<<2 * clusters, Cfg::kThreads, Cfg::kSmem, st>>>(ma, mb_full, q, bias, y, addend);
  SRGAN_RETURN_LAUNCH();
}

template <int BN, int MT, typename ST, typename OT = ST>
static int launch_bn(const CUtensorMap& ma, const CUtensorMap& mb, const UmmaConvP& p, const float* bias, OT* y,
                     dim3 grid, cudaStream_t st, const OT* addend, float* stats = nullptr) {
  using Cfg = UmmaCfg<BN, MT, ST>;
  constexpr bool kSame = sizeof(ST) == sizeof(OT);
  constexpr bool kCanStat = kSame && sizeof(ST) == 2 && BN >= 32;       // tile statistics: bf16 trunk only
  static unsigned long long attr_done = 0, attr_done_s = 0;
  {
    cudaError_t e = ensure_dyn_smem(conv_umma_kernel<BN, MT, ST, false, OT>, (int)Cfg::kSmem, &attr_done);
    if (e == cudaSuccess && kCanStat && stats)
      e = ensure_dyn_smem(conv_umma_kernel<BN, MT, ST, kCanStat, OT>, (int)Cfg::kSmem, &attr_done_s);
    if (e != cudaSuccess) { set_error("conv_umma smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  if (stats && !kCanStat) { set_error("conv_umma: tile statistics need bf16 storage and >= 32 output channels"); return SRGAN_E_UNSUPPORTED; }
  if constexpr (BN == 256 && MT == 1 && kSame) {
    if (!stats && use_cta_pairs<ST>((long)grid.x * grid.y * grid.z))
      return launch_pairs<ST>(ma, mb, p, bias, y, grid, st, addend);
  }
  UmmaConvP q = p;
  q.gx = (grid.x + MT - 1) / MT; q.gy = grid.y; q.gz = grid.z;
  // A/B switch: SRGAN_DBG_CLASS_MAJOR=1 restores the class-major item order
  static const bool class_major = getenv("SRGAN_DBG_CLASS_MAJOR") && atoi(getenv("SRGAN_DBG_CLASS_MAJOR")) != 0;
  q.cls_fast = (q.gz > 1 && !class_major) ? 1 : 0;
  long items = (long)q.gx * q.gy * q.gz;
  q.n_full = (int)items;
  // The CTAs of the last, partial wave would each run a whole tile while the other SMs idle.  With 256-wide tiles
  // (two filter boxes) a tile splits into two half-width items at no extra operand traffic per FLOP on the filter
  // side: when the tail fills at most half of the SMs, run it as twice as many half-width items.
  // Bring-up override: SRGAN_DBG_CONV_TAIL=0.
  static const bool tail_split = !(getenv("SRGAN_DBG_CONV_TAIL") && atoi(getenv("SRGAN_DBG_CONV_TAIL")) == 0);
  const long rem = items % kNumSMs;
  if (BN == 256 && MT == 1 && tail_split && rem > 0 && 2 * rem <= kNumSMs) {
    q.n_full = (int)(items - rem);
    items += rem;
  }
  q.n_items = (int)items;
  const unsigned ctas = (unsigned)(items < kNumSMs ? items : kNumSMs);     // persistent: one CTA per SM
  if (kCanStat && stats)
    conv_umma_kernel<BN, MT, ST, kCanStat, OT><<<ctas, Cfg::kThreads, Cfg::kSmem, st>>>(ma, mb, q, bias, y, addend, stats);
  else
    conv_umma_kernel<BN, MT, ST, false, OT><<<ctas, Cfg::kThreads, Cfg::kSmem, st>>>(ma, mb, q, bias, y, addend, nullptr);
  SRGAN_RETURN_LAUNCH();
}

// Generic launcher.  act_*: the tensor providing the A operand, [aN][aH][aW][aC] NHWC; `a_stride` 1 or 2
// (2 => parity view).  filt: [fK][T][fC] with fC == aC the reduction channels, fK = output channels.
// Activation, filter, output and addend share one storage type ST (float: TF32 MMAs; __nv_bfloat16: kind::f16).
struct Problem {
  const void* act; int aN, aH, aW, aC; int a_stride;
  const void* filt; int fK, T;
  int Nn, P, Q;              // pixel grid per class
  int out_H, out_W, os;
  int ncls;
  UmmaConvP p;               // taps / classes prefilled
};

template <typename T> struct NoDeduce { using type = T; };
template <typename ST, typename OT = ST>
static int run_problem_t(Problem& pr, const float* bias, OT* y, int act, float slope, cudaStream_t st,
                         const CUtensorMap* ma_prebuilt, const typename NoDeduce<OT>::type* addend,
                         float* stats = nullptr) {
  constexpr uint64_t ES = sizeof(ST);
  constexpr uint32_t ROW = UmmaElem<ST>::kRow;
  constexpr CUtensorMapDataType DT = UmmaElem<ST>::kTma;
  if (((uintptr_t)pr.act | (uintptr_t)pr.filt | (uintptr_t)y) % 16) {
    set_error("tcgen05 conv: tensors must be 16-byte aligned");
    return SRGAN_E_BADARG;
  }
  CUtensorMap ma, mb;
  const int C = pr.aC;
  if (C % (int)ROW) { set_error("tcgen05 conv: reduction channels must be a multiple of %d", (int)ROW); return SRGAN_E_UNSUPPORTED; }
  if (ma_prebuilt) {
    ma = *ma_prebuilt;            // caller picked the box with pick_box(pr.P, pr.Q) and filled pr.p.lw / lh
  } else if (pr.a_stride == 1) {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)pr.aW, 1, (uint64_t)pr.aH, (uint64_t)pr.aN};
    uint64_t str[4] = {(uint64_t)C * ES, (uint64_t)pr.aW * C * ES, (uint64_t)pr.aW * C * ES,
                       (uint64_t)pr.aH * pr.aW * C * ES};
    pick_box(pr.P, pr.Q, &pr.p.lw, &pr.p.lh);
    uint32_t box[5] = {ROW, 1u << pr.p.lw, 1, 1u << pr.p.lh, 128u >> (pr.p.lw + pr.p.lh)};
    if (int e = encode_map(&ma, pr.act, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, DT)) return e;
  } else {
    uint64_t dims[5] = {(uint64_t)2 * C, (uint64_t)pr.aW / 2, 2, (uint64_t)pr.aH / 2, (uint64_t)pr.aN};
    uint64_t str[4] = {(uint64_t)2 * C * ES, (uint64_t)pr.aW * C * ES, (uint64_t)2 * pr.aW * C * ES,
                       (uint64_t)pr.aH * pr.aW * C * ES};
    pick_box(pr.P, pr.Q, &pr.p.lw, &pr.p.lh);
    uint32_t box[5] = {ROW, 1u << pr.p.lw, 1, 1u << pr.p.lh, 128u >> (pr.p.lw + pr.p.lh)};
    if (int e = encode_map(&ma, pr.act, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, DT)) return e;
  }
  const int BN = pick_bn(pr.fK);
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)pr.T, (uint64_t)pr.fK};
    uint64_t str[2] = {(uint64_t)C * ES, (uint64_t)pr.T * C * ES};
    uint32_t box[3] = {ROW, 1, (uint32_t)(BN < 128 ? BN : 128)};      // == UmmaCfg::kBRows
    if (int e = encode_map(&mb, pr.filt, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, DT)) return e;
  }
  UmmaConvP& p = pr.p;
  p.c_chunks = C / (int)ROW;
  const int bw = 1 << p.lw, bh = 1 << p.lh, bn = 128 >> (p.lw + p.lh);
  p.tiles_w = ceil_div(pr.Q, bw); p.tiles_h = ceil_div(pr.P, bh); p.tiles_n = ceil_div(pr.Nn, bn);
  p.Nn = pr.Nn; p.P = pr.P; p.Q = pr.Q;
  p.out_H = pr.out_H; p.out_W = pr.out_W; p.out_C = pr.fK; p.os = pr.os; p.K = pr.fK;
  p.act = act; p.slope = slope;
  // epi_vec 8: a row segment of 32 accumulator columns is written with 32-byte stores, 4: 16-byte stores, 0: scalar
  constexpr int V8 = 32 / (int)sizeof(OT), V4 = 16 / (int)sizeof(OT);      // outputs per 32-byte / 16-byte store
  p.epi_vec = (pr.fK % V8 == 0 && (uintptr_t)y % 32 == 0) ? 8 : (pr.fK % V4 == 0 ? 4 : 0);
  if (bias && (uintptr_t)bias % 16) p.epi_vec = 0;            // vector bias loads need an aligned bias
  if (addend && ((uintptr_t)addend % 16 || pr.fK % V4)) { set_error("conv: addend must be 16-byte aligned"); return SRGAN_E_BADARG; }
  if (stats && (bn != 1 || bias || addend || act != SRGAN_ACT_NONE)) {
    set_error("conv: tile statistics need 128-pixel tiles inside one image and a plain epilogue");
    return SRGAN_E_UNSUPPORTED;
  }
  dim3 grid(p.tiles_w * p.tiles_h * p.tiles_n, ceil_div(pr.fK, BN), pr.ncls);
  // Two M sub-tiles per CTA (one filter tile feeds 256 pixels) when TMEM can still double-buffer the accumulator
  // (2 x 2 x 128 columns) and every SM keeps work; 256-wide tiles stay at MT = 1: overlapping the epilogue with the
  // next item's MMAs is worth more than sharing the filter tile (res conv 139 us vs 165 us).
  // Bring-up override: SRGAN_DBG_CONV_MT=1|2.
  static const char* e_mt = getenv("SRGAN_DBG_CONV_MT");
  const long ctas = (long)grid.x * grid.y * grid.z;
  const int mt = e_mt ? atoi(e_mt) : ((BN == 128 || BN == 64) && ctas >= 2 * kNumSMs ? 2 : 1);
  if constexpr (sizeof(ST) != sizeof(OT)) {
    // mixed storage exists for the thin RGB layers only: 64 / 32 output channels (stems, the head's input gradient)
    if (mt == 2 && BN == 64) return launch_bn<64, 2, ST, OT>(ma, mb, p, bias, y, grid, st, addend, stats);
    if (BN == 64) return launch_bn<64, 1, ST, OT>(ma, mb, p, bias, y, grid, st, addend, stats);
    if (BN == 32) return launch_bn<32, 1, ST, OT>(ma, mb, p, bias, y, grid, st, addend, stats);
    set_error("tcgen05 conv: mixed storage needs 32 or 64 output channels (got %d)", pr.fK);
    return SRGAN_E_UNSUPPORTED;
  } else {
    if (mt == 2 && BN == 256) return launch_bn<256, 2, ST>(ma, mb, p, bias, y, grid, st, addend, stats);
    if (mt == 2 && BN == 128) return launch_bn<128, 2, ST>(ma, mb, p, bias, y, grid, st, addend, stats);
    if (mt == 2 && BN == 64) return launch_bn<64, 2, ST>(ma, mb, p, bias, y, grid, st, addend, stats);
    switch (BN) {
      case 256: return launch_bn<256, 1, ST>(ma, mb, p, bias, y, grid, st, addend, stats);
      case 128: return launch_bn<128, 1, ST>(ma, mb, p, bias, y, grid, st, addend, stats);
      case 64:  return launch_bn<64, 1, ST>(ma, mb, p, bias, y, grid, st, addend, stats);
      case 32:  return launch_bn<32, 1, ST>(ma, mb, p, bias, y, grid, st, addend, stats);
      default:  return launch_bn<16, 1, ST>(ma, mb, p, bias, y, grid, st, addend, stats);
    }
  }
}

// fp32 storage (TF32 MMAs): the entry point of the thin-tensor and plain fp32 paths
static int run_problem(Problem& pr, const float* bias, float* y, int act, float slope, cudaStream_t st,
                       const CUtensorMap* ma_prebuilt = nullptr, const float* addend = nullptr) {
  return run_problem_t<float>(pr, bias, y, act, slope, st, ma_prebuilt, addend);
}


// ------------------------------------------------------------------------------------------ thin tensors
// Convolutions with <= 4 channels on one side (RGB stems and heads) would waste the tensor core on a
// 3-wide GEMM dimension.  They are run "row-packed" instead: the thin tensor is copied once into a zero-
// padded NHWC4 buffer TP[N][Hp][Wp][4]; for a filter row r the S taps x 4 channels of one output pixel are
// then 4*S CONTIGUOUS floats, so a TMA map whose pixel stride (16 B * conv stride) is smaller than its
// 32-float inner box delivers im2col rows  A_r[pixel][j = s*4 + c]  straight into the swizzled smem layout.
// The GEMM reduction per filter row is one 32-wide chunk (S <= 8), i.e. R "taps" of 32 "channels".
//   fprop, thin input  (C <= 4)             y  = sum_r A_r(xpad)  . Bp[k][r][s*4+c]
//   dgrad, thin dy     (K <= 4, stride 1)   dx = sum_r' A_r'(dypad) . Bp[c][r'][s'*4+k]   (flipped filter)
//   wgrad, thin x      (C <= 4)             out[k][r][s*4+c]   = sum_pix dy[pix][k] A_r(xpad)[pix][.]
//   wgrad, thin dy     (K <= 4, stride 1)   out[c][r'][s'*4+k] = sum_pix x[pix][c]  A_r'(dypad)[pix][.]
// The wgrad forms run wgrad_umma_kernel in "taps as N" mode: all R filter rows are N-blocks of ONE MMA, so
// the fat tensor is read exactly once.
struct ThinPlan {
  int mode;                 // 0: the thin tensor is the conv input x ; 1: it is dy (flipped taps)
  int st;                   // pixel stride of the packed view (conv stride, 1 for mode 1)
  int tc, tH, tW;           // thin tensor [N][tH][tW][tc]
  int padH, padW;           // physical zero padding of TP (top/left)
  int Hp, Wp;
  int fH, fW, fC;           // fat tensor [N][fH][fW][fC] == the pixel grid of the GEMM
  int R, S, N;
};

static bool thin_fits(const srgan_conv_desc* d) { return d->S * 4 <= 32 && d->R <= 8 && (d->stride == 1 || d->stride == 2); }

static bool thin_plan(const srgan_conv_desc* d, int pass, ThinPlan* t) {
  if (!thin_fits(d)) return false;
  const bool thin_in = d->C <= 4, thin_out = d->K <= 4;
  ThinPlan q = {};
  q.R = d->R; q.S = d->S; q.N = d->N;
  if (pass == 0 || (pass == 2 && thin_in)) {
    if (!thin_in) return false;
    q.mode = 0; q.st = d->stride; q.tc = d->C; q.tH = d->H; q.tW = d->W; q.padH = q.padW = d->pad;
    q.fH = d->P; q.fW = d->Q; q.fC = d->K;
    if (pass == 2 && d->K % 4) return false;
  } else {
    if (!thin_out || d->stride != 1 || d->pad > d->R - 1 || d->pad > d->S - 1) return false;
    q.mode = 1; q.st = 1; q.tc = d->K; q.tH = d->P; q.tW = d->Q; q.padH = d->R - 1 - d->pad; q.padW = d->S - 1 - d->pad;
    q.fH = d->H; q.fW = d->W; q.fC = d->C;
    if (d->C % 4) return false;
  }
  if (pass == 2 && q.fC > 128) return false;            // one 128-row accumulator tile
  int hp = q.tH + 2 * q.padH, need_h = (q.fH - 1) * q.st + q.R;
  if (need_h > hp) hp = need_h;
  q.Hp = (hp + q.st - 1) / q.st * q.st;
  int wp = q.tW + 2 * q.padW, need_w = (q.fW - 1) * q.st + 8;
  q.Wp = need_w > wp ? need_w : wp;
  *t = q;
  return true;
}

// TP[n][hp][wp][0..3] = thin[n][hp - padH][wp - padW][0..tc) or 0
__global__ void thin_pad_kernel(const float* __restrict__ src, float4* __restrict__ dst, int N, int tH, int tW, int tc,
                                int Hp, int Wp, int padH, int padW) {
  const size_t total = (size_t)N * Hp * Wp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wp = (int)(i % Wp);
    const size_t t = i / Wp;
    const int hp = (int)(t % Hp);
    const int n = (int)(t / Hp);
    const int h = hp - padH, w = wp - padW;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (h >= 0 && h < tH && w >= 0 && w < tW) {
      const float* s = src + (((size_t)n * tH + h) * tW + w) * tc;
      for (int c = 0; c < tc; ++c) v[c] = __ldg(s + c);
    }
    dst[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// mode 0: bp[k][r][s*4+c]  = w[k][r][s][c]                 (rows = K)
// mode 1: bp[c][r'][s'*4+k] = w[k][R-1-r'][S-1-s'][c]       (rows = C)
__global__ void thin_pack_filter_kernel(const float* __restrict__ w, float* __restrict__ bp, int K, int C, int R, int S,
                                        int mode) {
  const int rows = mode == 0 ? K : C;
  const int total = rows * R * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 31, r = (i >> 5) % R, row = i / (32 * R);
    const int s = j >> 2, t = j & 3;
    float v = 0.f;
    if (s < S) {
      if (mode == 0) { if (t < C) v = w[(((size_t)row * R + r) * S + s) * C + t]; }
      else           { if (t < K) v = w[(((size_t)t * R + (R - 1 - r)) * S + (S - 1 - s)) * C + row]; }
    }
    bp[i] = v;
  }
}

// dw[k][r][s][c] = sum over splits of  part[.][k][r][s*4+c]  (mode 0)  or  part[.][c][R-1-r][(S-1-s)*4+k]  (mode 1)
__global__ void thin_unpack_wgrad_kernel(const float* __restrict__ part, float* __restrict__ dw, int K, int C, int R, int S,
                                         int mode, int splits, long long split_stride) {
  const int total = K * R * S * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, s = (i / C) % S, r = (i / (C * S)) % R, k = i / (C * S * R);
    const size_t src = mode == 0 ? (((size_t)k * R + r) * 32 + s * 4 + c)
                                 : (((size_t)c * R + (R - 1 - r)) * 32 + (S - 1 - s) * 4 + k);
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += part[(size_t)z * split_stride + src];
    dw[i] = acc;
  }
}

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

static void thin_box(int fH, int fW, int pixels, int* lw, int* lh) {   // pow2 box bw x bh x bn = pixels
  int w = 1 << ilog2(fW);
  if (w > pixels) w = pixels;
  int h = 1 << ilog2(fH);
  if (h > pixels / w) h = pixels / w;
  *lw = ilog2(w); *lh = ilog2(h);
}

static int thin_map(CUtensorMap* m, const float* tp, const ThinPlan& t, int lw, int lh, int pixels, CUtensorMapSwizzle swz) {
  // {32 packed floats, fat column (stride st pixels), row parity, fat row, image}; strides overlap on purpose
  uint64_t dims[5] = {32, (uint64_t)t.fW, (uint64_t)t.st, (uint64_t)(t.Hp / t.st), (uint64_t)t.N};
  uint64_t str[4] = {(uint64_t)t.st * 16, (uint64_t)t.Wp * 16, (uint64_t)t.st * t.Wp * 16, (uint64_t)t.Hp * t.Wp * 16};
  uint32_t box[5] = {32, 1u << lw, 1, 1u << lh, (uint32_t)pixels >> (lw + lh)};
  return encode_map(m, tp, 5, dims, str, box, swz);
}

static size_t thin_tp_bytes(const ThinPlan& t) { return align256((size_t)t.N * t.Hp * t.Wp * 16); }

static int thin_pad_launch(const ThinPlan& t, const float* thin, float* tp, cudaStream_t st) {
  const size_t total = (size_t)t.N * t.Hp * t.Wp;
  unsigned blocks = (unsigned)((total + 255) / 256 < (size_t)kNumSMs * 16 ? (total + 255) / 256 : (size_t)kNumSMs * 16);
  thin_pad_kernel<<<blocks, 256, 0, st>>>(thin, (float4*)tp, t.N, t.tH, t.tW, t.tc, t.Hp, t.Wp, t.padH, t.padW);
  SRGAN_RETURN_LAUNCH();
}

// direct form (below): the tensor core reads the im2col rows from one copy of the image row
static bool thin_direct_plan(const srgan_conv_desc* d, int pass, ThinPlan* t);
template <typename OT>
static int conv_thin_direct_launch(const srgan_conv_desc* d, const ThinPlan& t, const float* thin, const float* w,
                                   const float* bias, OT* out, int act, float slope, void* ws, size_t ws_bytes,
                                   cudaStream_t st);

// y = act(conv(x) + bias) with thin x (pass 0)  /  dx = conv_transpose(dy) with thin dy (pass 1)
template <typename OT = float>
static int conv_thin_fwdlike_launch(const srgan_conv_desc* d, int pass, const float* thin, const float* w,
                                    const float* bias, OT* out, int act, float slope, void* ws, size_t ws_bytes,
                                    cudaStream_t st) {
  ThinPlan t;
  if (thin_direct_plan(d, pass, &t)) return conv_thin_direct_launch<OT>(d, t, thin, w, bias, out, act, slope, ws, ws_bytes, st);
  if (!thin_plan(d, pass, &t)) { set_error("thin conv: unsupported shape"); return SRGAN_E_UNSUPPORTED; }
  const size_t tpb = thin_tp_bytes(t), bpb = align256((size_t)t.fC * t.R * 32 * sizeof(float));
  if (!ws || ws_bytes < tpb + bpb) { set_error("thin conv: workspace %zu < %zu", ws_bytes, tpb + bpb); return SRGAN_E_WORKSPACE; }
  float* tp = (float*)ws;
  float* bp = (float*)((uint8_t*)ws + tpb);
  if (int e = thin_pad_launch(t, thin, tp, st)) return e;
  thin_pack_filter_kernel<<<ceil_div(t.fC * t.R * 32, 256), 256, 0, st>>>(w, bp, d->K, d->C, d->R, d->S, t.mode);
  Problem pr = {};
  pr.act = tp; pr.aN = t.N; pr.aH = t.Hp; pr.aW = t.Wp; pr.aC = 32; pr.a_stride = 1;
  pr.filt = bp; pr.fK = t.fC; pr.T = t.R;
  pr.Nn = t.N; pr.P = t.fH; pr.Q = t.fW; pr.out_H = t.fH; pr.out_W = t.fW; pr.os = 1; pr.ncls = 1;
  UmmaConvP& p = pr.p;
  p.tap_begin[0] = 0; p.tap_begin[1] = t.R;
  for (int r = 0; r < t.R; ++r) p.taps[r] = make_int4(0, 0, (r % t.st) | (r << 8), r / t.st);
  p.cls_oph[0] = 0; p.cls_opw[0] = 0;
  pick_box(pr.P, pr.Q, &p.lw, &p.lh);
  CUtensorMap ma;
  if (int e = thin_map(&ma, tp, t, p.lw, p.lh, 128, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  return run_problem_t<float, OT>(pr, bias, out, act, slope, st, &ma, nullptr);
}

static inline int floordiv2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }

// conv_thinout.cu
bool conv_thinout_supported(const srgan_conv_desc* d, int pass);
size_t conv_thinout_workspace(const srgan_conv_desc* d, int pass);
int conv_thinout_launch(const srgan_conv_desc* d, int pass, const float* in, const float* w, const float* bias,
                        float* out, int act, float slope, void* ws, size_t ws_bytes, cudaStream_t st);

bool conv_umma_supported(const srgan_conv_desc* d, int pass) {
  if (d->N < 1) return false;
  if (conv_thinout_supported(d, pass)) return true;
  { ThinPlan t; if (thin_plan(d, pass, &t)) return true; }
  if (pass == 0) {
    if (d->C % 32) return false;
    if (d->R * d->S > kMaxTaps) return false;
    if (d->stride == 1) return true;
    if (d->stride == 2) return d->H % 2 == 0 && d->W % 2 == 0;
    return false;
  }
  if (pass == 1) {
    if (d->K % 32) return false;             // reduction runs over the output channels of the conv
    if (d->stride == 1) return d->R * d->S <= kMaxTaps && d->pad < d->R && d->pad < d->S;
    if (d->stride == 2)   // four output-parity classes over an (H/2, W/2) grid; dy rows outside [0,P) are TMA zero-fill
      return d->H % 2 == 0 && d->W % 2 == 0 && d->R * d->S <= kMaxTaps && d->R >= 2 && d->S >= 2;
    return false;
  }
  // wgrad (dy is re-packed to a multiple of 4 channels when K % 4 != 0: TMA needs 16-byte pixel rows)
  if (d->C % 32 || d->R * d->S > kMaxTaps) return false;
  if (d->stride == 1) return true;
  if (d->stride == 2) return d->H % 2 == 0 && d->W % 2 == 0;
  return false;
}

constexpr int kColsumBlocks = 64;

// dst[pix][0..Kp) = src[pix][0..K) followed by zeros
__global__ void pad_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t pixels, int K,
                                    int Kp) {
  const size_t total = pixels * Kp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t pix = i / Kp;
    int k = (int)(i - pix * Kp);
    dst[i] = k < K ? __ldg(src + pix * K + k) : 0.f;
  }
}

struct WgradPlan { int BN, KT, gt, cpb, ngroups, lw, lh, tiles_w, tiles_h, tiles_n, chunks, cps, splits, tiles_k, tiles_c; };

static int pow2_at_least(int v, int lo) { int b = lo; while (b < v) b *= 2; return b; }

// ch / pix: channels and pixels of one operand box (WgradElem: 32 x 32 for fp32 storage, 64 x 64 for bf16)
static WgradPlan plan_wgrad(const srgan_conv_desc* d, int ch = 32, int pix = 32) {
  WgradPlan w;
  const int T = d->R * d->S;
  if (d->C >= 256) {
    w.BN = 256; w.cpb = 256 / ch; w.gt = 1; w.ngroups = T; w.tiles_c = ceil_div(d->C, 256);
  } else {
    w.cpb = d->C / ch;
    int gt_max = 256 / d->C;
    if (gt_max < 1) gt_max = 1;
    w.ngroups = ceil_div(T, gt_max);
    w.gt = ceil_div(T, w.ngroups);
    w.tiles_c = 1;
    w.BN = pow2_at_least(w.gt * d->C, ch);
  }
  w.KT = d->K > 128 ? 2 : 1;
  w.tiles_k = ceil_div(d->K, 128 * w.KT);
  int bw = 1 << ilog2(d->Q);
  if (bw > pix) bw = pix;
  int bh = 1 << ilog2(d->P);
  if (bh > pix / bw) bh = pix / bw;
  w.lw = ilog2(bw); w.lh = ilog2(bh);
  int bn = pix / (bw * bh);
  w.tiles_w = ceil_div(d->Q, bw); w.tiles_h = ceil_div(d->P, bh); w.tiles_n = ceil_div(d->N, bn);
  w.chunks = w.tiles_w * w.tiles_h * w.tiles_n;
  // Pixel splits (split-K): one CTA per SM is resident, so the launch runs in ceil(CTAs / 148) rounds of
  // (chunks per split x time per chunk + prologue / epilogue), and every split adds one pass over dW to the
  // fixed-order reduction that follows.  Pick the split count that minimises that estimate (a count that spills one
  // CTA into a third round costs 50 %: 9 taps x 33 splits = 297 CTAs was the old choice for the residual blocks).
  // Constants from tools/conv_bench.py: 0.29 us per 32-pixel chunk of a 128 x 256 MMA tile, ~5 us to drain a
  // 256 x 256 accumulator, ~4 TB/s for the reduction.  Bring-up override: SRGAN_DBG_WGRAD_SPLITS.
  const int tiles = w.tiles_k * w.tiles_c * w.ngroups;
  const double chunk_us = 0.29 * w.KT * (w.BN / 256.0);
  const double epi_us = 3.0 + 2.5 * w.KT * (w.BN / 256.0);
  const double dw_mb = (double)d->K * T * d->C * 4e-6;
  int max_splits = ceil_div(w.chunks, 8);
  if (max_splits > 128) max_splits = 128;
  if (max_splits < 1) max_splits = 1;
  int best = 1;
  double best_t = 1e30;
  for (int sp = 1; sp <= max_splits; ++sp) {
    const int cps = ceil_div(w.chunks, sp);
    const int eff = ceil_div(w.chunks, cps);
    if (eff != sp) continue;
    const int rounds = ceil_div(tiles * eff, kNumSMs);
    double t = rounds * (cps * chunk_us + epi_us);
    if (eff > 1) t += 3.0 + eff * dw_mb / 4.0;
    if (t < best_t) { best_t = t; best = eff; }
  }
  static const char* e_sp = getenv("SRGAN_DBG_WGRAD_SPLITS");
  if (e_sp && atoi(e_sp) >= 1 && atoi(e_sp) <= max_splits) best = atoi(e_sp);
  w.cps = ceil_div(w.chunks, best);
  w.splits = ceil_div(w.chunks, w.cps);
  return w;
}

// introspection for tests / tools: how the wgrad of this layer is decomposed
void conv_umma_wgrad_plan(const srgan_conv_desc* d, int* splits, int* ctas) {
  const WgradPlan w = plan_wgrad(d);
  *splits = w.splits;
  *ctas = w.tiles_k * w.tiles_c * w.ngroups * w.splits;
}

static size_t thin_workspace(const srgan_conv_desc* d, int pass, const ThinPlan& t);

size_t conv_umma_workspace(const srgan_conv_desc* d, int pass) {
  if (conv_thinout_supported(d, pass)) return conv_thinout_workspace(d, pass);
  { ThinPlan t; if (thin_plan(d, pass, &t)) return thin_workspace(d, pass, t); }
  if (pass == 1) return (size_t)d->K * d->R * d->S * d->C * sizeof(float);   // transposed filter
  if (pass == 2) {
    WgradPlan w = plan_wgrad(d);
    size_t b = w.splits > 1 ? (size_t)w.splits * d->K * d->R * d->S * d->C * sizeof(float) : 0;
    if (d->K % 4) b += (size_t)d->N * d->P * d->Q * ((d->K + 3) / 4 * 4) * sizeof(float);   // re-packed dy
    b += (size_t)kColsumBlocks * d->K * sizeof(float);                                      // bias-gradient partials
    return b;
  }
  return 0;
}

// plain (non-thin) forward problem: taps (r, s) at offsets (r - pad, s - pad); stride 2 through the parity view
static void fprop_problem(const srgan_conv_desc* d, const void* x, const void* w, Problem& pr) {
  pr.act = x; pr.aN = d->N; pr.aH = d->H; pr.aW = d->W; pr.aC = d->C; pr.a_stride = d->stride;
  pr.filt = w; pr.fK = d->K; pr.T = d->R * d->S;
  pr.Nn = d->N; pr.P = d->P; pr.Q = d->Q; pr.out_H = d->P; pr.out_W = d->Q; pr.os = 1; pr.ncls = 1;
  UmmaConvP& p = pr.p;
  p.tap_begin[0] = 0;
  int nt = 0;
  for (int r = 0; r < d->R; ++r)
    for (int s = 0; s < d->S; ++s) {
      int a = r - d->pad, b = s - d->pad;
      int4 tp;
      if (d->stride == 1) {
        tp = make_int4(0, b, 0 | ((r * d->S + s) << 8), a);
      } else {
        int hp = ((a % 2) + 2) % 2, wp = ((b % 2) + 2) % 2;
        tp = make_int4(wp * d->C, floordiv2(b), hp | ((r * d->S + s) << 8), floordiv2(a));
      }
      p.taps[nt++] = tp;
    }
  p.tap_begin[1] = nt;
  p.cls_oph[0] = 0; p.cls_opw[0] = 0;
}

// plain input-gradient problem on dy and the transposed filter wt[C][T][K]
static void dgrad_problem(const srgan_conv_desc* d, const void* dy, const void* wt, Problem& pr) {
  const int T = d->R * d->S;
  pr.act = dy; pr.aN = d->N; pr.aH = d->P; pr.aW = d->Q; pr.aC = d->K; pr.a_stride = 1;
  pr.filt = wt; pr.fK = d->C; pr.T = T;
  UmmaConvP& p = pr.p;
  if (d->stride == 1) {
    pr.Nn = d->N; pr.P = d->H; pr.Q = d->W; pr.out_H = d->H; pr.out_W = d->W; pr.os = 1; pr.ncls = 1;
    int nt = 0;
    p.tap_begin[0] = 0;
    for (int r = 0; r < d->R; ++r)
      for (int s = 0; s < d->S; ++s) p.taps[nt++] = make_int4(0, d->pad - s, 0 | ((r * d->S + s) << 8), d->pad - r);
    p.tap_begin[1] = nt;
    p.cls_oph[0] = 0; p.cls_opw[0] = 0;
  } else {
    // output pixel (2*i + ph, 2*j + pw): taps r with (ph + pad - r) even, source row i + (ph + pad - r)/2
    pr.Nn = d->N; pr.P = d->H / 2; pr.Q = d->W / 2; pr.out_H = d->H; pr.out_W = d->W; pr.os = 2; pr.ncls = 4;
    int nt = 0;
    for (int cls = 0; cls < 4; ++cls) {
      const int ph = cls >> 1, pw = cls & 1;
      p.tap_begin[cls] = nt;
      p.cls_oph[cls] = ph; p.cls_opw[cls] = pw;
      for (int r = 0; r < d->R; ++r) {
        if ((ph + d->pad - r) % 2) continue;
        for (int s = 0; s < d->S; ++s) {
          if ((pw + d->pad - s) % 2) continue;
          p.taps[nt++] = make_int4(0, floordiv2(pw + d->pad - s), 0 | ((r * d->S + s) << 8), floordiv2(ph + d->pad - r));
        }
      }
    }
    p.tap_begin[4] = nt;
  }
}

int conv_fprop_umma_launch(const srgan_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                           int act, float slope, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (conv_thinout_supported(d, 0)) return conv_thinout_launch(d, 0, x, w, bias, y, act, slope, ws, ws_bytes, st);
  { ThinPlan t; if (thin_plan(d, 0, &t)) return conv_thin_fwdlike_launch(d, 0, x, w, bias, y, act, slope, ws, ws_bytes, st); }
  Problem pr = {};
  fprop_problem(d, x, w, pr);
  return run_problem(pr, bias, y, act, slope, st);
}

// addend (optional, layout of dx): dx = dgrad(dy) + addend in the epilogue; only on the generic stride-1 path
bool conv_dgrad_umma_add_supported(const srgan_conv_desc* d) {
  ThinPlan t;
  return d->stride == 1 && d->C % 4 == 0 && !conv_thinout_supported(d, 1) && !thin_plan(d, 1, &t);
}

int conv_dgrad_umma_launch(const srgan_conv_desc* d, const float* dy, const float* w, float* dx, void* ws,
                           size_t ws_bytes, cudaStream_t st, const float* addend) {
  if (addend && !conv_dgrad_umma_add_supported(d)) { set_error("conv dgrad: fused addend not available for this shape"); return SRGAN_E_UNSUPPORTED; }
  if (conv_thinout_supported(d, 1)) return conv_thinout_launch(d, 1, dy, w, nullptr, dx, SRGAN_ACT_NONE, 0.f, ws, ws_bytes, st);
  { ThinPlan t; if (thin_plan(d, 1, &t)) return conv_thin_fwdlike_launch(d, 1, dy, w, nullptr, dx, SRGAN_ACT_NONE, 0.f, ws, ws_bytes, st); }
  const int T = d->R * d->S;
  size_t need = (size_t)d->K * T * d->C * sizeof(float);
  if (ws_bytes < need || !ws) { set_error("conv dgrad: workspace %zu < %zu", ws_bytes, need); return SRGAN_E_WORKSPACE; }
  float* wt = (float*)ws;       // [C][T][K]
  {
    dim3 g(ceil_div(d->C, 32), ceil_div(d->K, 32), T);
    filter_transpose_kernel<float><<<g, dim3(32, 8), 0, st>>>(w, wt, d->K, T, d->C);
  }
  Problem pr = {};
  dgrad_problem(d, dy, wt, pr);
  return run_problem(pr, nullptr, dx, SRGAN_ACT_NONE, 0.f, st, nullptr, addend);
}

// ------------------------------------------------------------------------------------------ thin layers, direct form
// Stride-1 convolutions with a <= 4-channel INPUT side and 64 output channels (RGB stem fprop; the head's input gradient
// with the roles of dy / flipped filter) without materialising im2col rows.  The row-packed form above makes TMA
// deliver, per filter row, 128 overlapping rows of 128 B (16 KB for 2 KB of distinct image data): 7.4 x expansion on
// the L2 -> shared-memory path, which bounds that kernel (ncu: 1.1 GB of TMA reads for a 134 MB result).
// Here ONE copy of every padded NHWC4 image row (Wp pixels x 16 B, contiguous) is bulk-copied into a ring of row slots
// and the tensor core reads the overlapping im2col rows itself: without swizzle a K-major core matrix is 8 rows of 16
// bytes that are 16 bytes apart - eight consecutive pixels - so
//     A_r[m][j]  (pixel m of the output row, j = s*4 + c)  = slot(h + r)[(m + j/4) * 16 + (j%4) * 4]
// is the canonical no-swizzle layout with LBO (next K core matrix) = 16 B and SBO (next 8 rows) = 128 B: overlapping
// core matrices over the same bytes.  One output row (<= 128 pixels) = R x 4 MMAs (M 128, N 64, K 8 tf32) into a
// four-deep ring of TMEM accumulators; consecutive output rows share R - 1 input rows, so a CTA walking down a band of rows
// loads one new 2 KB row per 16 KB (bf16) of output.  The packed filter (R x 8 KB, no-swizzle K-major: chunk stride 1 KB,
// 8-filter group stride 128 B) is loaded once per CTA.  Warp 0: bulk-copy producer, warp 1: MMA issuer (+ TMEM),
// warps 2-5: epilogue (tcgen05.ld -> bias / activation -> 32-byte stores, fp32 or bf16 output).
constexpr int kTDThreads = 192;
constexpr int kTDRing = 16;                    // image-row slots
constexpr int kTDSlotBytes = 2304;             // >= (128 + 8) pixels x 16 B, multiple of 128
constexpr int kTDAcc = 4;                      // TMEM accumulators of 64 columns (one per output row in flight)
struct ThinDirectP {
  int N, Ho, Wo, K;          // output [N][Ho][Wo][K], Wo <= 128, K == 64
  int Hp, Wp;                // padded thin tensor TP[N][Hp][Wp][4] (fp32)
  int R;                     // filter rows (<= 8)
  int bands, BH;             // row bands per image, rows per band
  int row_bytes;             // bytes of a TP row that are copied (multiple of 16, <= kTDSlotBytes)
  int act;
  float slope;
};
__device__ __forceinline__ uint64_t smem_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo >> 4) << 16;                 // K-direction stride between core matrices
  d |= (uint64_t)(sbo >> 4) << 32;                 // M / N-direction stride between 8-row groups
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ void bulk_g2s_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <typename OT>
__global__ void __launch_bounds__(kTDThreads, 1)
conv_thin_direct_kernel(const __grid_constant__ ThinDirectP p, const float* __restrict__ tp, const float* __restrict__ bp,
                        const float* __restrict__ bias, OT* __restrict__ y) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_bytes = p.R * 8192;
  uint8_t* sb = smem;
  uint8_t* ring = smem + b_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + kTDRing * kTDSlotBytes);
  uint64_t* empty = full + kTDRing;
  uint64_t* t_full = empty + kTDRing;              // [kTDAcc]
  uint64_t* t_empty = t_full + kTDAcc;             // [kTDAcc]
  uint64_t* b_full = t_empty + kTDAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.N * p.bands;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTDRing; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < kTDAcc; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 4); }
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // slots start as zeros: the MMAs read up to 16 bytes past a copied row (taps beyond S carry zero filter weights, but
  // 0 x NaN garbage would still poison the accumulator)
  for (int i = threadIdx.x; i < kTDRing * kTDSlotBytes / 16; i += kTDThreads)
    reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 1) tmem_alloc(tmem_slot, kTDAcc * 64);
  tc_fence_before();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(b_full, b_bytes);
      bulk_g2s_1d(sb, bp, b_bytes, b_full);
      int li = 0;                                  // image rows loaded so far by this CTA
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / p.bands, band = item - n * p.bands;
        const int h0 = band * p.BH, h1 = min(p.Ho, h0 + p.BH);
        for (int hp = h0; hp < h1 + p.R - 1; ++hp, ++li) {       // padded rows h0 .. h1 - 1 + R - 1
          const int slot = li % kTDRing;
          mbar_wait(empty + slot, (((uint32_t)(li / kTDRing)) & 1u) ^ 1u);
          mbar_expect_tx(full + slot, p.row_bytes);
          bulk_g2s_1d(ring + slot * kTDSlotBytes, tp + ((size_t)n * p.Hp + hp) * p.Wp * 4, p.row_bytes, full + slot);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = f32, A = B = tf32, both K-major, N = 64, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      mbar_wait(b_full, 0);
      const uint64_t adesc0 = smem_desc_nosw(smem_u32(ring), 16, 128);        // A: next K chunk + 16 B, next 8 pixels + 128 B
      const uint64_t bdesc0 = smem_desc_nosw(smem_u32(sb), 1024, 128);        // B: next K chunk + 1 KB, next 8 filters + 128 B
      int li = 0, waited = 0, ti = 0;              // first row of the current item, rows waited for, output rows done
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / p.bands, band = item - n * p.bands;
        const int h0 = band * p.BH, h1 = min(p.Ho, h0 + p.BH);
        (void)n;
        for (int h = h0; h < h1; ++h, ++ti) {
          const int first = li + (h - h0);          // load index of padded row h
          for (; waited < first + p.R; ++waited)    // rows h .. h + R - 1 have landed
            mbar_wait(full + waited % kTDRing, ((uint32_t)(waited / kTDRing)) & 1u);
          const int buf = ti % kTDAcc;
          mbar_wait(t_empty + buf, (((uint32_t)(ti / kTDAcc)) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t acc = tmem_base + buf * 64;
          // one thread issues all MMAs: keep its instruction count per MMA minimal (descriptors advance by adds: the
          // first version rebuilt both descriptors per MMA and was bound by this thread, 3000 clk per output row).
          // Measured and not kept: four output rows interleaved on four accumulators (so that consecutive MMAs are
          // independent): 130 instead of 105 us - the pace is set by the operand fetch (80 shared-memory wavefronts per
          // MMA: 128 pixel rows x 32 B through core matrices that straddle 128-byte lines), not by the dependent chain.
          int slot = first % kTDRing;
          uint64_t bdesc = bdesc0;
          for (int r = 0; r < p.R; ++r) {
            const uint64_t adesc = adesc0 + (uint64_t)(slot * (kTDSlotBytes >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k)             // K step = 8 floats = 2 pixels of A (+32 B), 2 chunks of B (+2 KB)
              umma_tf32(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(128 * k), idesc, (r | k) != 0);
            bdesc += 8192 >> 4;
            if (++slot == kTDRing) slot = 0;
          }
          umma_commit(empty + first % kTDRing);     // padded row h is not needed by later output rows
          umma_commit(t_full + buf);
        }
        // the last R - 1 rows of the band are released once its last output row has read them
        for (int r = 1; r < p.R; ++r) umma_commit(empty + (li + (h1 - h0) - 1 + r) % kTDRing);
        li += (h1 - h0) + p.R - 1;
      }
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;                 // TMEM lane == output pixel of the row
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    constexpr int V8 = 32 / (int)sizeof(OT);
    const int vec = ((uintptr_t)y % 32 == 0 && p.K % V8 == 0 && (!bias || (uintptr_t)bias % 16 == 0)) ? 8 : 0;
    int ti = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int n = item / p.bands, band = item - n * p.bands;
      const int h0 = band * p.BH, h1 = min(p.Ho, h0 + p.BH);
      for (int h = h0; h < h1; ++h, ++ti) {
        const int buf = ti % kTDAcc;
        mbar_wait(t_full + buf, ((uint32_t)(ti / kTDAcc)) & 1u);
        tc_fence_after();
        uint32_t ra[32], rb[32];
        tmem_ld32_issue(taddr + buf * 64, ra);
        tmem_ld32_issue(taddr + buf * 64 + 32, rb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + buf);  // the accumulator is in registers
        if (m < p.Wo) {
          OT* yrow = y + (((size_t)n * p.Ho + h) * p.Wo + m) * p.K;
          float v[32];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(half ? rb[j] : ra[j]);
            if (vec) {
              epi_row_chunk_any<32>(v, yrow + half * 32, bias ? bias + half * 32 : nullptr, p.act, p.slope, 8);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                st_from_float(yrow + half * 32 + j,
                              apply_act(v[j] + (bias ? __ldg(bias + half * 32 + j) : 0.f), p.act, p.slope));
            }
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTDAcc * 64);
  }
}

// packed filter of the direct form, per filter row r: [chunk c = s (4 floats)][8-filter group g][filter f % 8][4 floats]
// mode 0: value = w[f][r][s][e] ; mode 1 (input gradient of a thin-output layer): = w[e][R-1-r][S-1-s][f]   (e < tc)
__global__ void thin_direct_pack_filter_kernel(const float* __restrict__ w, float* __restrict__ bp, int K, int C, int R,
                                               int S, int mode) {
  const int total = R * 8 * 64 * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 3, f8 = (i >> 2) & 7, g = (i >> 5) & 7, c = (i >> 8) & 7, r = i >> 11;
    const int f = g * 8 + f8, s = c;
    float v = 0.f;
    if (s < S) {
      if (mode == 0) { if (e < C && f < K) v = w[(((size_t)f * R + r) * S + s) * C + e]; }
      else           { if (e < K && f < C) v = w[(((size_t)e * R + (R - 1 - r)) * S + (S - 1 - s)) * C + f]; }
    }
    bp[i] = v;
  }
}

static bool thin_direct_plan(const srgan_conv_desc* d, int pass, ThinPlan* t) {
  static const bool off = getenv("SRGAN_THIN_DIRECT") && atoi(getenv("SRGAN_THIN_DIRECT")) == 0;
  if (off || (pass != 0 && pass != 1) || !thin_plan(d, pass, t)) return false;
  return t->st == 1 && t->fC == 64 && t->fW <= 128 && t->S * 4 <= 32 && t->R <= 8 && t->Wp * 16 <= kTDSlotBytes &&
         t->Wp >= t->fW + t->S - 1;
}
static size_t thin_direct_workspace(const ThinPlan& t) { return thin_tp_bytes(t) + align256((size_t)t.R * 8192); }

template <typename OT>
static int conv_thin_direct_launch(const srgan_conv_desc* d, const ThinPlan& t, const float* thin, const float* w,
                                   const float* bias, OT* out, int act, float slope, void* ws, size_t ws_bytes,
                                   cudaStream_t st) {
  const size_t tpb = thin_tp_bytes(t), need = thin_direct_workspace(t);
  if (!ws || ws_bytes < need) { set_error("thin direct conv: workspace %zu < %zu", ws_bytes, need); return SRGAN_E_WORKSPACE; }
  if (((uintptr_t)thin | (uintptr_t)ws | (uintptr_t)out) % 16) { set_error("thin direct conv: tensors must be 16-byte aligned"); return SRGAN_E_BADARG; }
  float* tp = (float*)ws;
  float* bp = (float*)((uint8_t*)ws + tpb);
  if (int e = thin_pad_launch(t, thin, tp, st)) return e;
  thin_direct_pack_filter_kernel<<<ceil_div(t.R * 8192 / 4, 256), 256, 0, st>>>(w, bp, d->K, d->C, d->R, d->S, t.mode);
  ThinDirectP p = {};
  p.N = t.N; p.Ho = t.fH; p.Wo = t.fW; p.K = t.fC; p.Hp = t.Hp; p.Wp = t.Wp; p.R = t.R;
  p.act = act; p.slope = slope;
  p.row_bytes = t.Wp * 16;
  // rows per band: whole waves of one CTA per SM; every band re-reads R - 1 halo rows (cheap: 2 KB each)
  long best = -1;
  for (int b = 1; b <= p.Ho; ++b) {
    const int bh = ceil_div(p.Ho, b);
    if (ceil_div(p.Ho, bh) != b) continue;
    const long waves = ((long)p.N * b + kNumSMs - 1) / kNumSMs;
    const long cost = waves * (4 * bh + t.R - 1 + 8);
    if (best < 0 || cost < best) { best = cost; p.bands = b; p.BH = bh; }
  }
  const size_t smem = 1024 + (size_t)t.R * 8192 + (size_t)kTDRing * kTDSlotBytes + 512;
  static unsigned long long attr = 0;
  {
    cudaError_t e = ensure_dyn_smem(conv_thin_direct_kernel<OT>, 227 * 1024, &attr);
    if (e != cudaSuccess) { set_error("conv_thin_direct smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  const long items = (long)p.N * p.bands;
  const unsigned ctas = (unsigned)(items < kNumSMs ? items : kNumSMs);
  conv_thin_direct_kernel<OT><<<ctas, kTDThreads, smem, st>>>(p, tp, bp, bias, out);
  SRGAN_RETURN_LAUNCH();
}

// ------------------------------------------------------------------------------------------ thin layers, bf16 fat side
// The RGB layers of the bf16 engine: the 3-channel side (image, image gradient) and the filter stay fp32, the fat
// 64-channel side is bf16 (it belongs to the bf16 trunk).  Same kernels as above with another storage type on one side:
//   pass 0, thin input  (stem fprop)          fp32 image -> bf16 activation     conv_umma_kernel<.., float, .., bf16>
//   pass 1, thin output (head input gradient) fp32 dy    -> bf16 dx             same
// (pass 0 thin output / pass 1 thin input: conv_thinout.cu; pass 2: thin wgrad below)
static bool thin16_wgrad_supported(const srgan_conv_desc* d, ThinPlan* t);
static size_t thin16_wgrad_workspace(const srgan_conv_desc* d, const ThinPlan& t);
static int conv_wgrad_thin16_launch(const srgan_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                                    void* ws, size_t ws_bytes, cudaStream_t st);
bool conv_thinout16_supported(const srgan_conv_desc* d, int pass);
int conv_thinout16_launch(const srgan_conv_desc* d, int pass, const void* in, const float* w, const float* bias,
                          float* out, int act, float slope, void* ws, size_t ws_bytes, cudaStream_t st);
bool conv_thin16_supported(const srgan_conv_desc* d, int pass) {
  ThinPlan t;
  if (d->N < 1) return false;
  if (pass == 2) return thin16_wgrad_supported(d, &t);
  if ((pass == 0 && d->K <= 4) || (pass == 1 && d->C <= 4)) return conv_thinout16_supported(d, pass);
  if (pass == 0 && d->C <= 4) return thin_plan(d, 0, &t) && d->K % 8 == 0 && d->K <= 64 && d->K > 16;
  if (pass == 1 && d->K <= 4) return thin_plan(d, 1, &t) && d->C % 8 == 0 && d->C <= 64 && d->C > 16;
  return false;
}
size_t conv_thin16_workspace(const srgan_conv_desc* d, int pass) {
  ThinPlan t;
  if (!conv_thin16_supported(d, pass)) return 0;
  if ((pass == 0 && d->K <= 4) || (pass == 1 && d->C <= 4)) return conv_thinout_workspace(d, pass);
  if (!thin_plan(d, pass, &t)) return 0;
  if (pass == 2) return thin16_wgrad_workspace(d, t);
  return thin_workspace(d, pass, t);
}
int conv_thin16_wgrad_launch(const srgan_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* ws,
                             size_t ws_bytes, cudaStream_t st) {
  return conv_wgrad_thin16_launch(d, x, dy, dw, dbias, ws, ws_bytes, st);
}
int conv_thin16_launch(const srgan_conv_desc* d, int pass, const void* in, const float* w, const float* bias, void* out,
                       int act, float slope, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!conv_thin16_supported(d, pass)) { set_error("thin16 conv: unsupported shape / pass %d", pass); return SRGAN_E_UNSUPPORTED; }
  if ((pass == 0 && d->K <= 4) || (pass == 1 && d->C <= 4))      // bf16 fat input -> fp32 thin output
    return conv_thinout16_launch(d, pass, in, w, bias, (float*)out, act, slope, ws, ws_bytes, st);
  return conv_thin_fwdlike_launch<__nv_bfloat16>(d, pass, (const float*)in, w, bias, (__nv_bfloat16*)out, act, slope, ws,
                                                 ws_bytes, st);
}

// ------------------------------------------------------------------------------------------ bf16 storage
// Activations, filters (a bf16 shadow of the fp32 master weights) and outputs in bf16, fp32 accumulation and bias.
// Plain layers only (reduction channels a multiple of 64); thin RGB layers stay on the fp32 kernels.
bool conv_umma_bf16_supported(const srgan_conv_desc* d, int pass) {
  if (d->N < 1 || d->R * d->S > kMaxTaps) return false;
  const bool even = d->H % 2 == 0 && d->W % 2 == 0;
  if (pass == 0) return d->C % 64 == 0 && d->K % 8 == 0 && (d->stride == 1 || (d->stride == 2 && even));
  if (pass == 1) {
    if (d->K % 64 || d->C % 8) return false;
    if (d->stride == 1) return d->pad < d->R && d->pad < d->S;
    return d->stride == 2 && even && d->R >= 2 && d->S >= 2;
  }
  return false;
}

size_t conv_umma_bf16_workspace(const srgan_conv_desc* d, int pass) {
  return pass == 1 ? (size_t)d->K * d->R * d->S * d->C * sizeof(__nv_bfloat16) : 0;     // transposed filter
}

// Rows of per-tile statistics one image contributes when the pass is run with `stats` (0: not available): the
// pixel grid of a class must be cut into 128-pixel boxes that lie inside one image.
int conv_umma_bf16_stat_rows(const srgan_conv_desc* d, int pass) {
  if (pass > 1 || !conv_umma_bf16_supported(d, pass)) return 0;
  if ((pass == 0 ? d->K : d->C) < 32) return 0;
  int P, Q, ncls = 1;
  if (pass == 0) { P = d->P; Q = d->Q; }
  else if (d->stride == 1) { P = d->H; Q = d->W; }
  else { P = d->H / 2; Q = d->W / 2; ncls = 4; }
  int lw, lh;
  pick_box(P, Q, &lw, &lh);
  if ((128 >> (lw + lh)) != 1) return 0;
  return ncls * ceil_div(P, 1 << lh) * ceil_div(Q, 1 << lw);
}

int conv_fprop_umma_bf16_launch(const srgan_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                                int act, float slope, cudaStream_t st, float* stats) {
  if (!conv_umma_bf16_supported(d, 0)) { set_error("bf16 conv fprop: unsupported shape"); return SRGAN_E_UNSUPPORTED; }
  if (stats && !conv_umma_bf16_stat_rows(d, 0)) { set_error("bf16 conv fprop: tile statistics not available for this shape"); return SRGAN_E_UNSUPPORTED; }
  Problem pr = {};
  fprop_problem(d, x, w, pr);
  return run_problem_t<__nv_bfloat16>(pr, bias, (__nv_bfloat16*)y, act, slope, st, nullptr, nullptr, stats);
}

int conv_dgrad_umma_bf16_launch(const srgan_conv_desc* d, const void* dy, const void* w, void* dx, void* ws,
                                size_t ws_bytes, cudaStream_t st, const void* addend, float* stats) {
  if (!conv_umma_bf16_supported(d, 1)) { set_error("bf16 conv dgrad: unsupported shape"); return SRGAN_E_UNSUPPORTED; }
  if (stats && !conv_umma_bf16_stat_rows(d, 1)) { set_error("bf16 conv dgrad: tile statistics not available for this shape"); return SRGAN_E_UNSUPPORTED; }
  if (addend && d->stride != 1) { set_error("bf16 conv dgrad: fused addend only for stride 1"); return SRGAN_E_UNSUPPORTED; }
  const int T = d->R * d->S;
  const size_t need = conv_umma_bf16_workspace(d, 1);
  if (ws_bytes < need || !ws) { set_error("bf16 conv dgrad: workspace %zu < %zu", ws_bytes, need); return SRGAN_E_WORKSPACE; }
  __nv_bfloat16* wt = (__nv_bfloat16*)ws;       // [C][T][K]
  {
    dim3 g(ceil_div(d->C, 32), ceil_div(d->K, 32), T);
    filter_transpose_kernel<__nv_bfloat16><<<g, dim3(32, 8), 0, st>>>((const __nv_bfloat16*)w, wt, d->K, T, d->C);
  }
  Problem pr = {};
  dgrad_problem(d, dy, wt, pr);
  return run_problem_t<__nv_bfloat16>(pr, nullptr, (__nv_bfloat16*)dx, SRGAN_ACT_NONE, 0.f, st, nullptr,
                                      (const __nv_bfloat16*)addend, stats);
}

void splitk_reduce_launch(const float* part, float* out, long long n, int splits, cudaStream_t st);
int colsum_launch(const float* x, float* out, long long rows, int C, float* scratch, int scratch_blocks,
                  cudaStream_t st);

template <int BN, int KT, int PW, typename ST>
static int launch_wgrad_pw(const CUtensorMap& mdy, const CUtensorMap& mx, const UmmaWgradP& p, float* out, dim3 grid,
                           cudaStream_t st) {
  using Cfg = WgradCfg<BN, KT, PW>;
  static unsigned long long attr_done = 0;
  {
    cudaError_t e = ensure_dyn_smem(wgrad_umma_kernel<BN, KT, PW, ST>, (int)Cfg::kSmem, &attr_done);
    if (e != cudaSuccess) { set_error("wgrad_umma smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  wgrad_umma_kernel<BN, KT, PW, ST><<<grid, Cfg::kThreads, Cfg::kSmem, st>>>(mdy, mx, p, out);
  SRGAN_RETURN_LAUNCH();
}

// Producer warps: a stage of the widest tiles is 16 boxes of 4 KB; boxes issued by one warp complete one after the
// other, so the wide tiles deal them to 8 warps.  Bring-up override: SRGAN_DBG_WGRAD_PW=4|8.
template <int BN, int KT, typename ST>
static int launch_wgrad_cfg(const CUtensorMap& mdy, const CUtensorMap& mx, const UmmaWgradP& p, float* out, dim3 grid,
                            cudaStream_t st) {
  static const char* e_pw = getenv("SRGAN_DBG_WGRAD_PW");
  constexpr int boxes = (KT * 128 + BN) / WgradElem<ST>::kCh;      // TMA boxes of one full stage
  const int pw = e_pw ? atoi(e_pw) : (boxes * WgradElem<ST>::kBoxBytes >= 12 * 4096 ? 8 : 4);
  if constexpr (sizeof(ST) == 4) {
    if (pw == 16) return launch_wgrad_pw<BN, KT, 16, ST>(mdy, mx, p, out, grid, st);
  }
  if (pw == 8) return launch_wgrad_pw<BN, KT, 8, ST>(mdy, mx, p, out, grid, st);
  return launch_wgrad_pw<BN, KT, 4, ST>(mdy, mx, p, out, grid, st);
}

static int launch_wgrad(int BN, int KT, const CUtensorMap& mdy, const CUtensorMap& mx, const UmmaWgradP& p_in,
                        float* out, dim3 grid, cudaStream_t st) {
  UmmaWgradP p = p_in;
  p.vec8 = (uintptr_t)out % 32 == 0 && p.C % 8 == 0 && p.split_stride % 8 == 0;
  if (KT == 2) {
    switch (BN) {
      case 256: return launch_wgrad_cfg<256, 2, float>(mdy, mx, p, out, grid, st);
      case 128: return launch_wgrad_cfg<128, 2, float>(mdy, mx, p, out, grid, st);
      case 64:  return launch_wgrad_cfg<64, 2, float>(mdy, mx, p, out, grid, st);
      default:  return launch_wgrad_cfg<32, 2, float>(mdy, mx, p, out, grid, st);
    }
  }
  switch (BN) {
    case 256: return launch_wgrad_cfg<256, 1, float>(mdy, mx, p, out, grid, st);
    case 128: return launch_wgrad_cfg<128, 1, float>(mdy, mx, p, out, grid, st);
    case 64:  return launch_wgrad_cfg<64, 1, float>(mdy, mx, p, out, grid, st);
    default:  return launch_wgrad_cfg<32, 1, float>(mdy, mx, p, out, grid, st);
  }
}

// bf16 operands: N blocks are 64 channels wide, so the narrowest tile is 64 columns
static int launch_wgrad_bf16(int BN, int KT, const CUtensorMap& mdy, const CUtensorMap& mx, const UmmaWgradP& p_in,
                             float* out, dim3 grid, cudaStream_t st) {
  using B = __nv_bfloat16;
  UmmaWgradP p = p_in;
  p.vec8 = (uintptr_t)out % 32 == 0 && p.C % 8 == 0 && p.split_stride % 8 == 0;
  if (KT == 2) {
    switch (BN) {
      case 256: return launch_wgrad_cfg<256, 2, B>(mdy, mx, p, out, grid, st);
      case 128: return launch_wgrad_cfg<128, 2, B>(mdy, mx, p, out, grid, st);
      default:  return launch_wgrad_cfg<64, 2, B>(mdy, mx, p, out, grid, st);
    }
  }
  switch (BN) {
    case 256: return launch_wgrad_cfg<256, 1, B>(mdy, mx, p, out, grid, st);
    case 128: return launch_wgrad_cfg<128, 1, B>(mdy, mx, p, out, grid, st);
    default:  return launch_wgrad_cfg<64, 1, B>(mdy, mx, p, out, grid, st);
  }
}


// ---- thin wgrad: out[row][r][32] = sum_pix fat[pix][row] * A_r(TP)[pix][.]  (rows = fat channels <= 128)
struct ThinWgradPlan { int BN, lw, lh, tiles_w, tiles_h, tiles_n, chunks, cps, splits; };

static ThinWgradPlan plan_thin_wgrad(const ThinPlan& t) {
  ThinWgradPlan w;
  w.BN = t.R * 32 > 128 ? 256 : (t.R * 32 > 64 ? 128 : (t.R * 32 > 32 ? 64 : 32));
  thin_box(t.fH, t.fW, 32, &w.lw, &w.lh);
  const int bw = 1 << w.lw, bh = 1 << w.lh, bn = 32 / (bw * bh);
  w.tiles_w = ceil_div(t.fW, bw); w.tiles_h = ceil_div(t.fH, bh); w.tiles_n = ceil_div(t.N, bn);
  w.chunks = w.tiles_w * w.tiles_h * w.tiles_n;
  int splits = w.chunks < kNumSMs ? w.chunks : kNumSMs;     // one resident CTA per SM, every CTA the full tile
  w.cps = ceil_div(w.chunks, splits);
  w.splits = ceil_div(w.chunks, w.cps);
  return w;
}

static size_t thin_workspace(const srgan_conv_desc* d, int pass, const ThinPlan& t) {
  size_t b = thin_tp_bytes(t);
  if (pass != 2) return b + align256((size_t)t.fC * t.R * 32 * sizeof(float));
  ThinWgradPlan w = plan_thin_wgrad(t);
  b += align256((size_t)w.splits * t.fC * t.R * 32 * sizeof(float));
  b += (size_t)kColsumBlocks * d->K * sizeof(float);
  return b;
}

static int conv_wgrad_thin_launch(const srgan_conv_desc* d, const ThinPlan& t, const float* x, const float* dy,
                                  float* dw, float* dbias, void* ws, size_t ws_bytes, cudaStream_t st) {
  const ThinWgradPlan w = plan_thin_wgrad(t);
  const size_t need = thin_workspace(d, 2, t);
  if (need > ws_bytes || !ws) { set_error("thin wgrad: workspace %zu < %zu", ws_bytes, need); return SRGAN_E_WORKSPACE; }
  if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)ws) % 16) {
    set_error("thin wgrad: tensors must be 16-byte aligned");
    return SRGAN_E_BADARG;
  }
  const size_t row_elems = (size_t)t.R * 32, part_elems = (size_t)t.fC * row_elems;
  float* tp = (float*)ws;
  float* part = (float*)((uint8_t*)ws + thin_tp_bytes(t));
  float* csum = (float*)((uint8_t*)part + align256((size_t)w.splits * part_elems * sizeof(float)));
  if (dbias) {
    if (int e = colsum_launch(dy, dbias, (long long)d->N * d->P * d->Q, d->K, csum, kColsumBlocks, st)) return e;
  }
  if (!dw) return SRGAN_OK;
  const float* thin = t.mode == 0 ? x : dy;
  const float* fat = t.mode == 0 ? dy : x;
  if (int e = thin_pad_launch(t, thin, tp, st)) return e;
  CUtensorMap mfat, mthin;
  const uint32_t bw = 1u << w.lw, bh = 1u << w.lh, bn = 32u / (bw * bh);
  {
    uint64_t dims[5] = {(uint64_t)t.fC, (uint64_t)t.fW, 1, (uint64_t)t.fH, (uint64_t)t.N};
    uint64_t str[4] = {(uint64_t)t.fC * 4, (uint64_t)t.fW * t.fC * 4, (uint64_t)t.fW * t.fC * 4,
                       (uint64_t)t.fH * t.fW * t.fC * 4};
    uint32_t box[5] = {32, bw, 1, bh, bn};
    if (int e = encode_map(&mfat, fat, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  }
  if (int e = thin_map(&mthin, tp, t, w.lw, w.lh, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  UmmaWgradP p = {};
  p.tiles_c = 1; p.tiles_w = w.tiles_w; p.tiles_h = w.tiles_h; p.tiles_n = w.tiles_n;
  p.lw = w.lw; p.lh = w.lh; p.chunks = w.chunks; p.chunks_per_split = w.cps;
  p.K = t.fC; p.C = 32; p.T = t.R;                // out[row][r][32]: the R filter rows are one tap group
  p.gt = t.R; p.cpb = 1;
  p.split_stride = (long long)part_elems;
  p.desc_hi = mn_desc_hi(4096, 512, 1);
  for (int r = 0; r < t.R; ++r) p.taps[r] = make_int4(0, 0, r % t.st, r / t.st);
  dim3 grid(1, 1, w.splits);
  if (int e = launch_wgrad(w.BN, 1, mfat, mthin, p, part, grid, st)) return e;
  thin_unpack_wgrad_kernel<<<ceil_div(d->K * d->R * d->S * d->C, 256), 256, 0, st>>>(
      part, dw, d->K, d->C, d->R, d->S, t.mode, w.splits, (long long)part_elems);
  SRGAN_RETURN_LAUNCH();
}

// ---- thin wgrad with a bf16 fat tensor ("thin16"): the thin tensor is packed as bf16 NHWC8 (16 B per pixel, the same
// byte geometry as fp32 NHWC4: an im2col row of a filter row is S x 8 <= 64 contiguous bf16 = one 128-byte row), both
// operands are MN-major bf16 boxes of 64 x 64 (WgradElem<bf16>).  A filter row is a 64-column N block, so the R rows
// are run as tap groups of 4 (N = 256) on adjacent CTAs: the fat tensor comes from HBM once, the second group's read
// is an L2 hit.  out[row][r][s*8 + c].
constexpr int kThin16Gt = 4;
struct Thin16WgradPlan { int lw, lh, tiles_w, tiles_h, tiles_n, chunks, cps, splits, groups; };
static Thin16WgradPlan plan_thin16_wgrad(const ThinPlan& t) {
  Thin16WgradPlan w;
  thin_box(t.fH, t.fW, 64, &w.lw, &w.lh);
  const int bw = 1 << w.lw, bh = 1 << w.lh, bn = 64 / (bw * bh);
  w.tiles_w = ceil_div(t.fW, bw); w.tiles_h = ceil_div(t.fH, bh); w.tiles_n = ceil_div(t.N, bn);
  w.chunks = w.tiles_w * w.tiles_h * w.tiles_n;
  w.groups = ceil_div(t.R, kThin16Gt);
  const int slots = kNumSMs / w.groups > 0 ? kNumSMs / w.groups : 1;      // one resident CTA per SM
  int splits = w.chunks < slots ? w.chunks : slots;
  w.cps = ceil_div(w.chunks, splits);
  w.splits = ceil_div(w.chunks, w.cps);
  return w;
}
// TP16[n][hp][wp][0..7] (bf16) = thin[n][hp - padH][wp - padW][0..tc) or 0
__global__ void thin_pad16_kernel(const float* __restrict__ src, uint4* __restrict__ dst, int N, int tH, int tW, int tc,
                                  int Hp, int Wp, int padH, int padW) {
  const size_t total = (size_t)N * Hp * Wp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wp = (int)(i % Wp);
    const size_t t = i / Wp;
    const int hp = (int)(t % Hp);
    const int n = (int)(t / Hp);
    const int h = hp - padH, w = wp - padW;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (h >= 0 && h < tH && w >= 0 && w < tW) {
      const float* s = src + (((size_t)n * tH + h) * tW + w) * tc;
      for (int c = 0; c < tc; ++c) v[c] = __ldg(s + c);
    }
    dst[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), 0u, 0u);
  }
}
// dw[k][r][s][c] = sum over splits of  part[.][k][r][s*8+c]  (mode 0)  or  part[.][c][R-1-r][(S-1-s)*8+k]  (mode 1)
__global__ void thin16_unpack_wgrad_kernel(const float* __restrict__ part, float* __restrict__ dw, int K, int C, int R, int S,
                                           int mode, int splits, long long split_stride) {
  const int total = K * R * S * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, s = (i / C) % S, r = (i / (C * S)) % R, k = i / (C * S * R);
    const size_t src = mode == 0 ? (((size_t)k * R + r) * 64 + s * 8 + c)
                                 : (((size_t)c * R + (R - 1 - r)) * 64 + (S - 1 - s) * 8 + k);
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += part[(size_t)z * split_stride + src];
    dw[i] = acc;
  }
}
// column sums of a bf16 [rows][C] tensor (bias gradient of the stem: dy is the bf16 fat tensor), two fixed-order stages
__global__ void colsum16_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ part, long long rows, int C,
                                        long long rows_per_block) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f;
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += __bfloat162float(x[r * C + c]);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    part[(long long)blockIdx.y * C + c] = t;
  }
}
constexpr int kColsum16Blocks = 592;

static bool thin16_wgrad_supported(const srgan_conv_desc* d, ThinPlan* t) {
  if (d->N < 1 || !thin_plan(d, 2, t)) return false;
  return d->S * 8 <= 64 && (t->fC == 64 || t->fC == 128);       // whole 64-channel boxes of the fat tensor
}
static size_t thin16_wgrad_workspace(const srgan_conv_desc* d, const ThinPlan& t) {
  const Thin16WgradPlan w = plan_thin16_wgrad(t);
  size_t b = thin_tp_bytes(t);                                                // bf16 NHWC8: 16 B per pixel as well
  b += align256((size_t)w.splits * t.fC * t.R * 64 * sizeof(float));
  b += align256((size_t)kColsum16Blocks * d->K * sizeof(float));
  return b;
}
static int conv_wgrad_thin16_launch(const srgan_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                                    void* ws, size_t ws_bytes, cudaStream_t st) {
  using El = WgradElem<__nv_bfloat16>;
  ThinPlan t;
  if (!thin16_wgrad_supported(d, &t)) { set_error("thin16 wgrad: unsupported shape"); return SRGAN_E_UNSUPPORTED; }
  const Thin16WgradPlan w = plan_thin16_wgrad(t);
  const size_t need = thin16_wgrad_workspace(d, t);
  if (need > ws_bytes || !ws) { set_error("thin16 wgrad: workspace %zu < %zu", ws_bytes, need); return SRGAN_E_WORKSPACE; }
  if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)ws) % 16) { set_error("thin16 wgrad: tensors must be 16-byte aligned"); return SRGAN_E_BADARG; }
  const size_t part_elems = (size_t)t.fC * t.R * 64;
  uint8_t* tp = (uint8_t*)ws;
  float* part = (float*)((uint8_t*)ws + thin_tp_bytes(t));
  float* csum = (float*)((uint8_t*)part + align256((size_t)w.splits * part_elems * sizeof(float)));
  const long long pixels = (long long)d->N * d->P * d->Q;
  if (dbias) {
    if (t.mode == 0) {                       // dy is the bf16 fat tensor
      const long long rpb = ceil_div64(pixels, kColsum16Blocks);
      const int nb = (int)ceil_div64(pixels, rpb);
      colsum16_partial_kernel<<<dim3(ceil_div(d->K, 32), nb), dim3(32, 8), 0, st>>>((const __nv_bfloat16*)dy, csum, pixels,
                                                                                  d->K, rpb);
      if (int e = colsum_launch(csum, dbias, nb, d->K, nullptr, 0, st)) return e;
    } else {                                 // dy is the thin fp32 tensor
      if (int e = colsum_launch((const float*)dy, dbias, pixels, d->K, csum, 64, st)) return e;
    }
  }
  if (!dw) return SRGAN_OK;
  const float* thin = (const float*)(t.mode == 0 ? x : dy);
  const void* fat = t.mode == 0 ? dy : x;
  {
    const size_t total = (size_t)t.N * t.Hp * t.Wp;
    unsigned blocks = (unsigned)((total + 255) / 256 < (size_t)kNumSMs * 16 ? (total + 255) / 256 : (size_t)kNumSMs * 16);
    thin_pad16_kernel<<<blocks, 256, 0, st>>>(thin, (uint4*)tp, t.N, t.tH, t.tW, t.tc, t.Hp, t.Wp, t.padH, t.padW);
  }
  CUtensorMap mfat, mthin;
  const uint32_t bw = 1u << w.lw, bh = 1u << w.lh, bn = 64u / (bw * bh);
  const CUtensorMapDataType DT = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    uint64_t dims[5] = {(uint64_t)t.fC, (uint64_t)t.fW, 1, (uint64_t)t.fH, (uint64_t)t.N};
    uint64_t str[4] = {(uint64_t)t.fC * 2, (uint64_t)t.fW * t.fC * 2, (uint64_t)t.fW * t.fC * 2,
                       (uint64_t)t.fH * t.fW * t.fC * 2};
    uint32_t box[5] = {(uint32_t)El::kCh, bw, 1, bh, bn};
    if (int e = encode_map(&mfat, fat, 5, dims, str, box, El::kSwz, DT)) return e;
  }
  {
    // {64 packed bf16, fat column (stride st pixels), row parity, fat row, image}; strides overlap on purpose
    uint64_t dims[5] = {64, (uint64_t)t.fW, (uint64_t)t.st, (uint64_t)(t.Hp / t.st), (uint64_t)t.N};
    uint64_t str[4] = {(uint64_t)t.st * 16, (uint64_t)t.Wp * 16, (uint64_t)t.st * t.Wp * 16, (uint64_t)t.Hp * t.Wp * 16};
    uint32_t box[5] = {64, bw, 1, bh, bn};
    if (int e = encode_map(&mthin, tp, 5, dims, str, box, El::kSwz, DT)) return e;
  }
  UmmaWgradP p = {};
  p.tiles_c = 1; p.tiles_w = w.tiles_w; p.tiles_h = w.tiles_h; p.tiles_n = w.tiles_n;
  p.lw = w.lw; p.lh = w.lh; p.chunks = w.chunks; p.chunks_per_split = w.cps;
  p.K = t.fC; p.C = 64; p.T = t.R;                // out[row][r][64]
  p.gt = kThin16Gt; p.cpb = 1;
  p.split_stride = (long long)part_elems;
  p.desc_hi = mn_desc_hi(El::kLbo, El::kSbo, El::kLayout);
  for (int r = 0; r < t.R; ++r) p.taps[r] = make_int4(0, 0, r % t.st, r / t.st);
  dim3 grid(1, w.groups, w.splits);
  if (int e = launch_wgrad_bf16(256, 1, mfat, mthin, p, part, grid, st)) return e;
  thin16_unpack_wgrad_kernel<<<ceil_div(d->K * d->R * d->S * d->C, 256), 256, 0, st>>>(
      part, dw, d->K, d->C, d->R, d->S, t.mode, w.splits, (long long)part_elems);
  SRGAN_RETURN_LAUNCH();
}

int conv_wgrad_umma_launch(const srgan_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                           void* ws, size_t ws_bytes, cudaStream_t st) {
  { ThinPlan t; if (thin_plan(d, 2, &t)) return conv_wgrad_thin_launch(d, t, x, dy, dw, dbias, ws, ws_bytes, st); }
  const WgradPlan w = plan_wgrad(d);
  const int T = d->R * d->S;
  const size_t need = conv_umma_workspace(d, 2);
  if (need > ws_bytes || !ws) { set_error("conv wgrad: workspace %zu < %zu", ws_bytes, need); return SRGAN_E_WORKSPACE; }
  if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw | (uintptr_t)ws) % 16) {
    set_error("tcgen05 wgrad: tensors must be 16-byte aligned");
    return SRGAN_E_BADARG;
  }
  // workspace layout: [split partials][re-packed dy][colsum partials]
  float* part = (float*)ws;
  float* dyp = part + (w.splits > 1 ? (size_t)w.splits * d->K * T * d->C : 0);
  const int Kp = (d->K + 3) / 4 * 4;
  const size_t pixels = (size_t)d->N * d->P * d->Q;
  float* csum = dyp + (d->K % 4 ? pixels * Kp : 0);
  if (dbias) {
    if (int e = colsum_launch(dy, dbias, (long long)pixels, d->K, csum, kColsumBlocks, st)) return e;
  }
  if (!dw) return SRGAN_OK;
  if (d->K % 4) {
    size_t total = pixels * Kp;
    unsigned blocks = (unsigned)((total + 255) / 256 < (size_t)kNumSMs * 8 ? (total + 255) / 256 : (size_t)kNumSMs * 8);
    pad_channels_kernel<<<blocks, 256, 0, st>>>(dy, dyp, pixels, d->K, Kp);
    dy = dyp;
  }
  CUtensorMap mdy, mx;
  const uint32_t bw = 1u << w.lw, bh = 1u << w.lh, bn = 32u / (bw * bh);
  const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  {
    uint64_t dims[5] = {(uint64_t)Kp, (uint64_t)d->Q, 1, (uint64_t)d->P, (uint64_t)d->N};
    uint64_t str[4] = {(uint64_t)Kp * 4, (uint64_t)d->Q * Kp * 4, (uint64_t)d->Q * Kp * 4,
                       (uint64_t)d->P * d->Q * Kp * 4};
    uint32_t box[5] = {32, bw, 1, bh, bn};
    if (int e = encode_map(&mdy, dy, 5, dims, str, box, swz)) return e;
  }
  const int C = d->C;
  if (d->stride == 1) {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)d->W, 1, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t str[4] = {(uint64_t)C * 4, (uint64_t)d->W * C * 4, (uint64_t)d->W * C * 4, (uint64_t)d->H * d->W * C * 4};
    uint32_t box[5] = {32, bw, 1, bh, bn};
    if (int e = encode_map(&mx, x, 5, dims, str, box, swz)) return e;
  } else {
    uint64_t dims[5] = {(uint64_t)2 * C, (uint64_t)d->W / 2, 2, (uint64_t)d->H / 2, (uint64_t)d->N};
    uint64_t str[4] = {(uint64_t)2 * C * 4, (uint64_t)d->W * C * 4, (uint64_t)2 * d->W * C * 4,
                       (uint64_t)d->H * d->W * C * 4};
    uint32_t box[5] = {32, bw, 1, bh, bn};
    if (int e = encode_map(&mx, x, 5, dims, str, box, swz)) return e;
  }
  UmmaWgradP p = {};
  p.tiles_c = w.tiles_c; p.tiles_w = w.tiles_w; p.tiles_h = w.tiles_h; p.tiles_n = w.tiles_n;
  p.lw = w.lw; p.lh = w.lh; p.chunks = w.chunks; p.chunks_per_split = w.cps;
  p.K = d->K; p.C = C; p.T = T;
  p.gt = w.gt; p.cpb = w.cpb;
  p.split_stride = (long long)d->K * T * C;
  p.desc_hi = mn_desc_hi(4096, 512, 1);
  for (int r = 0; r < d->R; ++r)
    for (int s = 0; s < d->S; ++s) {
      int a = r - d->pad, b = s - d->pad;
      if (d->stride == 1) p.taps[r * d->S + s] = make_int4(0, b, 0, a);
      else p.taps[r * d->S + s] = make_int4((((b % 2) + 2) % 2) * C, floordiv2(b), ((a % 2) + 2) % 2, floordiv2(a));
    }
  float* out = w.splits > 1 ? part : dw;
  dim3 grid(w.tiles_k * w.tiles_c, w.ngroups, w.splits);
  if (int e = launch_wgrad(w.BN, w.KT, mdy, mx, p, out, grid, st)) return e;
  if (w.splits > 1) splitk_reduce_launch(part, dw, (long long)d->K * T * C, w.splits, st);
  SRGAN_RETURN_LAUNCH();
}

// ---- bf16 storage: x and dy are bf16 NHWC, dW (and the split partials) fp32 KRSC, dbias fp32
bool conv_umma_bf16_wgrad_supported(const srgan_conv_desc* d) {
  if (d->N < 1 || d->R * d->S > kMaxTaps) return false;
  if (d->C % 64 || d->K % 64) return false;
  if (d->C < 256 && 256 % d->C) return false;                 // tap groups: N = gt * C must tile 64-column blocks
  return d->stride == 1 || (d->stride == 2 && d->H % 2 == 0 && d->W % 2 == 0);
}

size_t conv_umma_bf16_wgrad_workspace(const srgan_conv_desc* d) {
  const WgradPlan w = plan_wgrad(d, 64, 64);
  return w.splits > 1 ? (size_t)w.splits * d->K * d->R * d->S * d->C * sizeof(float) : 0;
}

void conv_umma_bf16_wgrad_plan(const srgan_conv_desc* d, int* splits, int* ctas) {
  const WgradPlan w = plan_wgrad(d, 64, 64);
  *splits = w.splits;
  *ctas = w.tiles_k * w.tiles_c * w.ngroups * w.splits;
}

int conv_wgrad_umma_bf16_launch(const srgan_conv_desc* d, const void* x, const void* dy, float* dw, void* ws,
                                size_t ws_bytes, cudaStream_t st) {
  using El = WgradElem<__nv_bfloat16>;
  if (!conv_umma_bf16_wgrad_supported(d)) { set_error("bf16 conv wgrad: unsupported shape"); return SRGAN_E_UNSUPPORTED; }
  const WgradPlan w = plan_wgrad(d, El::kCh, El::kPix);
  const int T = d->R * d->S;
  const size_t need = conv_umma_bf16_wgrad_workspace(d);
  if (need > ws_bytes || (need && !ws)) { set_error("bf16 conv wgrad: workspace %zu < %zu", ws_bytes, need); return SRGAN_E_WORKSPACE; }
  if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw | (uintptr_t)ws) % 16) {
    set_error("tcgen05 wgrad: tensors must be 16-byte aligned");
    return SRGAN_E_BADARG;
  }
  float* part = (float*)ws;
  CUtensorMap mdy, mx;
  const uint32_t bw = 1u << w.lw, bh = 1u << w.lh, bn = (uint32_t)El::kPix / (bw * bh);
  const CUtensorMapDataType DT = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const uint64_t K = d->K, C = d->C;
  {
    uint64_t dims[5] = {K, (uint64_t)d->Q, 1, (uint64_t)d->P, (uint64_t)d->N};
    uint64_t str[4] = {K * 2, (uint64_t)d->Q * K * 2, (uint64_t)d->Q * K * 2, (uint64_t)d->P * d->Q * K * 2};
    uint32_t box[5] = {(uint32_t)El::kCh, bw, 1, bh, bn};
    if (int e = encode_map(&mdy, dy, 5, dims, str, box, El::kSwz, DT)) return e;
  }
  if (d->stride == 1) {
    uint64_t dims[5] = {C, (uint64_t)d->W, 1, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t str[4] = {C * 2, (uint64_t)d->W * C * 2, (uint64_t)d->W * C * 2, (uint64_t)d->H * d->W * C * 2};
    uint32_t box[5] = {(uint32_t)El::kCh, bw, 1, bh, bn};
    if (int e = encode_map(&mx, x, 5, dims, str, box, El::kSwz, DT)) return e;
  } else {
    uint64_t dims[5] = {2 * C, (uint64_t)d->W / 2, 2, (uint64_t)d->H / 2, (uint64_t)d->N};
    uint64_t str[4] = {2 * C * 2, (uint64_t)d->W * C * 2, (uint64_t)2 * d->W * C * 2, (uint64_t)d->H * d->W * C * 2};
    uint32_t box[5] = {(uint32_t)El::kCh, bw, 1, bh, bn};
    if (int e = encode_map(&mx, x, 5, dims, str, box, El::kSwz, DT)) return e;
  }
  UmmaWgradP p = {};
  p.tiles_c = w.tiles_c; p.tiles_w = w.tiles_w; p.tiles_h = w.tiles_h; p.tiles_n = w.tiles_n;
  p.lw = w.lw; p.lh = w.lh; p.chunks = w.chunks; p.chunks_per_split = w.cps;
  p.K = d->K; p.C = d->C; p.T = T;
  p.gt = w.gt; p.cpb = w.cpb;
  p.split_stride = (long long)d->K * T * d->C;
  p.desc_hi = mn_desc_hi(El::kLbo, El::kSbo, El::kLayout);
  for (int r = 0; r < d->R; ++r)
    for (int s2 = 0; s2 < d->S; ++s2) {
      int a = r - d->pad, b = s2 - d->pad;
      if (d->stride == 1) p.taps[r * d->S + s2] = make_int4(0, b, 0, a);
      else p.taps[r * d->S + s2] = make_int4((((b % 2) + 2) % 2) * d->C, floordiv2(b), ((a % 2) + 2) % 2, floordiv2(a));
    }
  float* out = w.splits > 1 ? part : dw;
  dim3 grid(w.tiles_k * w.tiles_c, w.ngroups, w.splits);
  if (int e = launch_wgrad_bf16(w.BN, w.KT, mdy, mx, p, out, grid, st)) return e;
  if (w.splits > 1) splitk_reduce_launch(part, dw, (long long)d->K * T * d->C, w.splits, st);
  SRGAN_RETURN_LAUNCH();
}

}  // namespace srgan

extern "C" int srgan_has_tcgen05(void) { return 1; }
