// GPU side of the notebooks' image pre-processing: decoded RGB uint8 images ->
// CenterCrop(crop) -> Pillow-exact bilinear Resize(out) -> optional horizontal flip -> ToTensor -> MinMax(True),
// written as the NHWC fp32 batch the generator's stem reads.
//   ref: notebook/01-train_Conventional_SingleGAN.ipynb cell 9 (transform["train"] / ["test"]),
//        pyfiles/dataset.py:127-141 (FaceDataset.__getitem__), pyfiles/util.py:108-116,148-153 (min_max, MinMax).
// The resize is Pillow's 8-bit two-pass resampling restated in integer arithmetic: horizontal pass into a uint8
// intermediate (kept in shared memory), vertical pass, both with the 22-bit fixed-point triangle-filter coefficients
// the host computes exactly like Pillow's precompute_coeffs / normalize_coeffs_8bpc; ToTensor and MinMax use
// correctly rounded fp32 division / subtraction in the reference's order.  Every output bit equals the CPU pipeline's.
//
// One CTA per image: the cropped image is read once from HBM (3 * crop^2 bytes), the result written once
// (12 * out^2 bytes); everything between lives in shared memory.  Byte work, HBM bound, no tensor cores.
#include "common.cuh"

namespace srgan {

constexpr int kFaceThreads = 512;
constexpr int kFacePrecision = 22;

struct FaceP {
  int H, W, crop, out, top, left, ksize_h, ksize_v;
};

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kFacePrecision;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// smem: [coef_h out*ksize_h][bounds_h out*2][coef_v out*ksize_v][bounds_v out*2] int32, then tmp[crop][out][3] u8,
// then res[out][out][3] u8
__global__ void __launch_bounds__(kFaceThreads, 1)
face_transform_kernel(FaceP p, const uint8_t* __restrict__ img, const int* __restrict__ coef_h,
                      const int* __restrict__ bounds_h, const int* __restrict__ coef_v,
                      const int* __restrict__ bounds_v, const uint8_t* __restrict__ flip, float* __restrict__ y) {
  extern __shared__ __align__(16) uint8_t face_smem[];
  __shared__ int s_min, s_max;
  int* sch = reinterpret_cast<int*>(face_smem);
  int* sbh = sch + p.out * p.ksize_h;
  int* scv = sbh + p.out * 2;
  int* sbv = scv + p.out * p.ksize_v;
  uint8_t* tmp = reinterpret_cast<uint8_t*>(sbv + p.out * 2);
  uint8_t* res = tmp + (size_t)p.crop * p.out * 3;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < p.out * p.ksize_h; i += blockDim.x) sch[i] = __ldg(coef_h + i);
  for (int i = threadIdx.x; i < p.out * p.ksize_v; i += blockDim.x) scv[i] = __ldg(coef_v + i);
  for (int i = threadIdx.x; i < p.out * 2; i += blockDim.x) { sbh[i] = __ldg(bounds_h + i); sbv[i] = __ldg(bounds_v + i); }
  if (threadIdx.x == 0) { s_min = 255; s_max = 0; }
  __syncthreads();

  // horizontal pass: tmp[r][xx][c] from the cropped row r (consecutive threads: consecutive output columns)
  const uint8_t* src = img + ((size_t)b * p.H + p.top) * p.W * 3 + (size_t)p.left * 3;
  for (int i = threadIdx.x; i < p.crop * p.out; i += blockDim.x) {
    const int r = i / p.out, xx = i - r * p.out;
    const int x0 = sbh[2 * xx], n = sbh[2 * xx + 1];
    const int* k = sch + xx * p.ksize_h;
    const uint8_t* row = src + (size_t)r * p.W * 3 + x0 * 3;
    int a0 = 1 << (kFacePrecision - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < n; ++j) {
      const int w = k[j];
      a0 += (int)__ldg(row + 3 * j) * w; a1 += (int)__ldg(row + 3 * j + 1) * w; a2 += (int)__ldg(row + 3 * j + 2) * w;
    }
    uint8_t* t = tmp + (size_t)i * 3;
    t[0] = clip8(a0); t[1] = clip8(a1); t[2] = clip8(a2);
  }
  __syncthreads();

  // vertical pass: res[yy][xx][c]; track the image's min / max (ToTensor and MinMax are monotone in the byte value)
  int lo = 255, hi = 0;
  for (int i = threadIdx.x; i < p.out * p.out * 3; i += blockDim.x) {
    const int yy = i / (p.out * 3), rem = i - yy * (p.out * 3);
    const int y0 = sbv[2 * yy], n = sbv[2 * yy + 1];
    const int* k = scv + yy * p.ksize_v;
    int a = 1 << (kFacePrecision - 1);
    for (int j = 0; j < n; ++j) a += (int)tmp[(size_t)(y0 + j) * p.out * 3 + rem] * k[j];
    const uint8_t v = clip8(a);
    res[i] = v;
    lo = min(lo, (int)v); hi = max(hi, (int)v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&s_min, lo); atomicMax(&s_max, hi); }
  __syncthreads();

  // ToTensor: t = v / 255 ; MinMax(True): ((t - min) / (max - min + 1e-8)) * 2 - 1, every operation rounded to fp32
  // as numpy does (ref pyfiles/util.py:108-116)
  const float flo = __fdiv_rn((float)s_min, 255.f), fhi = __fdiv_rn((float)s_max, 255.f);
  const float den = __fadd_rn(__fsub_rn(fhi, flo), 1e-8f);
  const bool fl = flip != nullptr && flip[b] != 0;
  float* yb = y + (size_t)b * p.out * p.out * 3;
  for (int i = threadIdx.x; i < p.out * p.out * 3; i += blockDim.x) {
    const int yy = i / (p.out * 3), rem = i - yy * (p.out * 3);
    const int xx = rem / 3, c = rem - xx * 3;
    const int sx = fl ? p.out - 1 - xx : xx;
    const float t = __fdiv_rn((float)res[(yy * p.out + sx) * 3 + c], 255.f);
    const float q = __fdiv_rn(__fsub_rn(t, flo), den);
    yb[i] = __fsub_rn(__fmul_rn(q, 2.f), 1.f);
  }
}

}  // namespace srgan

using namespace srgan;

extern "C" size_t srgan_face_transform_smem(int crop, int out, int ksize_h, int ksize_v) {
  return (size_t)out * (ksize_h + ksize_v + 4) * sizeof(int) + (size_t)crop * out * 3 + (size_t)out * out * 3;
}

extern "C" int srgan_face_transform(const uint8_t* img, int B, int H, int W, int crop, int out, const int* coef_h,
                                    const int* bounds_h, int ksize_h, const int* coef_v, const int* bounds_v,
                                    int ksize_v, const uint8_t* flip, float* y, void* stream) {
  SRGAN_CHECK_ARG(img && coef_h && bounds_h && coef_v && bounds_v && y, "null pointer");
  SRGAN_CHECK_ARG(B >= 0 && H > 0 && W > 0 && crop > 0 && out > 0 && ksize_h > 0 && ksize_v > 0, "bad sizes");
  SRGAN_CHECK_ARG(crop <= H && crop <= W, "crop window larger than the image (torchvision would pad)");
  if (B == 0) return SRGAN_OK;
  FaceP p;
  p.H = H; p.W = W; p.crop = crop; p.out = out; p.ksize_h = ksize_h; p.ksize_v = ksize_v;
  // torchvision center_crop: int(round((size - crop) / 2.0)) with Python's round-half-to-even
  auto origin = [](int size, int c) { const int d = size - c; return (d % 2 == 0) ? d / 2 : ((d / 2) % 2 == 0 ? d / 2 : d / 2 + 1); };
  p.top = origin(H, crop); p.left = origin(W, crop);
  const size_t smem = srgan_face_transform_smem(crop, out, ksize_h, ksize_v);
  SRGAN_CHECK_ARG(smem <= 220 * 1024, "crop / output size too large for the shared-memory staging");
  static unsigned long long attr_done = 0;
  {
    cudaError_t e = ensure_dyn_smem(face_transform_kernel, 220 * 1024, &attr_done);
    if (e != cudaSuccess) { set_error("face_transform smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  face_transform_kernel<<<B, kFaceThreads, smem, (cudaStream_t)stream>>>(p, img, coef_h, bounds_h, coef_v, bounds_v,
                                                                         flip, y);
  SRGAN_RETURN_LAUNCH();
}
