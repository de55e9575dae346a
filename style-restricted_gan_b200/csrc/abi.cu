// C-ABI front door: error reporting, convolution engine dispatch, small wrappers.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace srgan {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// conv_ffma.cu
size_t conv_ffma_workspace(const srgan_conv_desc* d, int pass);
int conv_fprop_ffma_launch(const srgan_conv_desc*, const float*, const float*, const float*, float*, int, float,
                           cudaStream_t);
int conv_dgrad_ffma_launch(const srgan_conv_desc*, const float*, const float*, float*, cudaStream_t);
int conv_wgrad_ffma_launch(const srgan_conv_desc*, const float*, const float*, float*, float*, void*, size_t,
                           cudaStream_t);
int colsum_launch(const float*, float*, long long, int, float*, int, cudaStream_t);
bool conv_head_supported(const srgan_conv_desc* d);
int conv_fprop_head_launch(const srgan_conv_desc*, const float*, const float*, const float*, float*, int, float,
                           cudaStream_t);

// conv_umma.cu (tcgen05 engine)
bool conv_umma_supported(const srgan_conv_desc* d, int pass);
size_t conv_umma_workspace(const srgan_conv_desc* d, int pass);
int conv_fprop_umma_launch(const srgan_conv_desc*, const float*, const float*, const float*, float*, int, float,
                           void*, size_t, cudaStream_t);
int conv_dgrad_umma_launch(const srgan_conv_desc*, const float*, const float*, float*, void*, size_t,
                           cudaStream_t, const float* addend);
bool conv_dgrad_umma_add_supported(const srgan_conv_desc* d);
void conv_umma_wgrad_plan(const srgan_conv_desc* d, int* splits, int* ctas);
int conv_wgrad_umma_launch(const srgan_conv_desc*, const float*, const float*, float*, float*, void*, size_t,
                           cudaStream_t);

// conv_umma.cu, bf16 storage (void*: __nv_bfloat16 tensors)
bool conv_umma_bf16_supported(const srgan_conv_desc* d, int pass);
size_t conv_umma_bf16_workspace(const srgan_conv_desc* d, int pass);
int conv_umma_bf16_stat_rows(const srgan_conv_desc* d, int pass);
int conv_fprop_umma_bf16_launch(const srgan_conv_desc*, const void*, const void*, const float*, void*, int, float,
                                cudaStream_t, float* stats);
int conv_dgrad_umma_bf16_launch(const srgan_conv_desc*, const void*, const void*, void*, void*, size_t, cudaStream_t,
                                const void* addend, float* stats);
bool conv_umma_bf16_wgrad_supported(const srgan_conv_desc* d);
size_t conv_umma_bf16_wgrad_workspace(const srgan_conv_desc* d);
void conv_umma_bf16_wgrad_plan(const srgan_conv_desc* d, int* splits, int* ctas);
int conv_wgrad_umma_bf16_launch(const srgan_conv_desc*, const void*, const void*, float*, void*, size_t, cudaStream_t);

// conv_umma.cu / conv_thinout.cu: thin RGB layers with a bf16 fat side
bool conv_thin16_supported(const srgan_conv_desc* d, int pass);
size_t conv_thin16_workspace(const srgan_conv_desc* d, int pass);
int conv_thin16_launch(const srgan_conv_desc* d, int pass, const void* in, const float* w, const float* bias, void* out,
                       int act, float slope, void* ws, size_t ws_bytes, cudaStream_t st);

int conv_thin16_wgrad_launch(const srgan_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* ws,
                             size_t ws_bytes, cudaStream_t st);

static int check_desc(const srgan_conv_desc* d) {
  if (!d) { set_error("conv: null descriptor"); return SRGAN_E_BADARG; }
  if (d->N < 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || d->K <= 0 || d->R <= 0 || d->S <= 0 || d->stride <= 0 ||
      d->pad < 0) {
    set_error("conv: non-positive dimension"); return SRGAN_E_BADARG;
  }
  int P = (d->H + 2 * d->pad - d->R) / d->stride + 1, Q = (d->W + 2 * d->pad - d->S) / d->stride + 1;
  if (d->H + 2 * d->pad < d->R || d->W + 2 * d->pad < d->S || P != d->P || Q != d->Q) {
    set_error("conv: output size mismatch (expected %dx%d, got %dx%d)", P, Q, d->P, d->Q);
    return SRGAN_E_BADARG;
  }
  return SRGAN_OK;
}

static bool dense_x(const srgan_conv_desc* d) {
  if (d->xs_n == 0 && d->xs_h == 0 && d->xs_w == 0 && d->xs_c == 0) return true;
  return d->xs_c == 1 && d->xs_w == d->C && d->xs_h == (int64_t)d->W * d->C &&
         d->xs_n == (int64_t)d->H * d->W * d->C;
}

static int resolve_engine(const srgan_conv_desc* d, int pass, int engine) {
  bool ok = (pass == 1 || dense_x(d)) && conv_umma_supported(d, pass);
  if (engine == SRGAN_CONV_FP32) return SRGAN_CONV_FP32;
  if (engine == SRGAN_CONV_TF32) return ok ? SRGAN_CONV_TF32 : SRGAN_E_UNSUPPORTED;
  return ok ? SRGAN_CONV_TF32 : SRGAN_CONV_FP32;
}

}  // namespace srgan

using namespace srgan;

extern "C" const char* srgan_last_error(void) { return g_err; }
extern "C" int srgan_abi_version(void) { return SRGAN_ABI_VERSION; }

extern "C" int srgan_conv2d_engine(const srgan_conv_desc* d, int pass) {
  if (int e = check_desc(d)) return e;
  return resolve_engine(d, pass, SRGAN_CONV_AUTO);
}

extern "C" size_t srgan_conv2d_workspace(const srgan_conv_desc* d, int pass, int engine) {
  if (check_desc(d)) return 0;
  int e = resolve_engine(d, pass, engine);
  if (e == SRGAN_CONV_TF32) return conv_umma_workspace(d, pass);
  if (e == SRGAN_CONV_FP32) return conv_ffma_workspace(d, pass);
  return 0;
}

extern "C" int srgan_conv2d_fprop(const srgan_conv_desc* d, const float* x, const float* w, const float* bias,
                                  float* y, int act, float slope, int engine, void* ws, size_t ws_bytes,
                                  void* stream) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(x && w && y, "null pointer");
  if (engine != SRGAN_CONV_TF32 && dense_x(d) && conv_head_supported(d))      // long-reduction 1..4-logit heads
    return conv_fprop_head_launch(d, x, w, bias, y, act, slope, (cudaStream_t)stream);
  int e = resolve_engine(d, 0, engine);
  if (e < 0) { set_error("conv fprop: shape not supported by the tcgen05 engine"); return e; }
  if (e == SRGAN_CONV_TF32)
    return conv_fprop_umma_launch(d, x, w, bias, y, act, slope, ws, ws_bytes, (cudaStream_t)stream);
  return conv_fprop_ffma_launch(d, x, w, bias, y, act, slope, (cudaStream_t)stream);
}

extern "C" int srgan_conv2d_dgrad(const srgan_conv_desc* d, const float* dy, const float* w, float* dx,
                                  int engine, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(dy && w && dx, "null pointer");
  int e = resolve_engine(d, 1, engine);
  if (e < 0) { set_error("conv dgrad: shape not supported by the tcgen05 engine"); return e; }
  if (e == SRGAN_CONV_TF32) return conv_dgrad_umma_launch(d, dy, w, dx, ws, ws_bytes, (cudaStream_t)stream, nullptr);
  return conv_dgrad_ffma_launch(d, dy, w, dx, (cudaStream_t)stream);
}

// dx = dgrad(dy) + addend: the second gradient that flows into the convolution's input (residual skip) is added in
// the epilogue instead of by a separate pass over the tensor
extern "C" int srgan_conv2d_dgrad_add_supported(const srgan_conv_desc* d, int engine) {
  if (check_desc(d)) return 0;
  if (resolve_engine(d, 1, engine) != SRGAN_CONV_TF32) return 0;
  return conv_dgrad_umma_add_supported(d) ? 1 : 0;
}
extern "C" int srgan_conv2d_dgrad_add(const srgan_conv_desc* d, const float* dy, const float* w, const float* addend,
                                      float* dx, int engine, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(dy && w && dx && addend, "null pointer");
  if (resolve_engine(d, 1, engine) != SRGAN_CONV_TF32) { set_error("conv dgrad_add: tcgen05 engine only"); return SRGAN_E_UNSUPPORTED; }
  return conv_dgrad_umma_launch(d, dy, w, dx, ws, ws_bytes, (cudaStream_t)stream, addend);
}

extern "C" int srgan_conv2d_wgrad(const srgan_conv_desc* d, const float* x, const float* dy, float* dw,
                                  float* dbias, int engine, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(x && dy && (dw || dbias), "null pointer");
  int e = resolve_engine(d, 2, engine);
  if (e < 0) { set_error("conv wgrad: shape not supported by the tcgen05 engine"); return e; }
  if (e == SRGAN_CONV_TF32)
    return conv_wgrad_umma_launch(d, x, dy, dw, dbias, ws, ws_bytes, (cudaStream_t)stream);
  return conv_wgrad_ffma_launch(d, x, dy, dw, dbias, ws, ws_bytes, (cudaStream_t)stream);
}

// ---- bf16 storage (experimental): NHWC bf16 activations, KRSC bf16 filters, fp32 bias and accumulation
extern "C" int srgan_conv2d_bf16_supported(const srgan_conv_desc* d, int pass) {
  if (check_desc(d)) return 0;
  if (pass == 2) return dense_x(d) && conv_umma_bf16_wgrad_supported(d) ? 1 : 0;
  return (pass == 1 || dense_x(d)) && conv_umma_bf16_supported(d, pass) ? 1 : 0;
}
extern "C" size_t srgan_conv2d_bf16_workspace(const srgan_conv_desc* d, int pass) {
  if (check_desc(d)) return 0;
  if (pass == 2) return conv_umma_bf16_wgrad_supported(d) ? conv_umma_bf16_wgrad_workspace(d) : 0;
  return conv_umma_bf16_workspace(d, pass);
}
extern "C" int srgan_conv2d_wgrad_bf16(const srgan_conv_desc* d, const void* x, const void* dy, float* dw, void* ws,
                                       size_t ws_bytes, void* stream) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(x && dy && dw, "null pointer");
  SRGAN_CHECK_ARG(dense_x(d), "bf16 conv: dense NHWC input only");
  return conv_wgrad_umma_bf16_launch(d, x, dy, dw, ws, ws_bytes, (cudaStream_t)stream);
}
extern "C" int srgan_conv2d_wgrad_bf16_plan(const srgan_conv_desc* d, int* splits, int* ctas) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(splits && ctas, "null pointer");
  SRGAN_CHECK_ARG(conv_umma_bf16_wgrad_supported(d), "shape does not qualify for the bf16 wgrad");
  conv_umma_bf16_wgrad_plan(d, splits, ctas);
  return SRGAN_OK;
}
extern "C" int srgan_conv2d_fprop_bf16(const srgan_conv_desc* d, const void* x, const void* w, const float* bias,
                                       void* y, int act, float slope, float* tile_stats, void* stream) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(x && w && y, "null pointer");
  SRGAN_CHECK_ARG(dense_x(d), "bf16 conv: dense NHWC input only");
  return conv_fprop_umma_bf16_launch(d, x, w, bias, y, act, slope, (cudaStream_t)stream, tile_stats);
}
extern "C" int srgan_conv2d_dgrad_bf16(const srgan_conv_desc* d, const void* dy, const void* w, const void* addend,
                                       void* dx, float* tile_stats, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(dy && w && dx, "null pointer");
  return conv_dgrad_umma_bf16_launch(d, dy, w, dx, ws, ws_bytes, (cudaStream_t)stream, addend, tile_stats);
}
extern "C" int srgan_conv2d_thin16_supported(const srgan_conv_desc* d, int pass) {
  if (check_desc(d) || !dense_x(d)) return 0;
  return conv_thin16_supported(d, pass) ? 1 : 0;
}
extern "C" size_t srgan_conv2d_thin16_workspace(const srgan_conv_desc* d, int pass) {
  if (check_desc(d)) return 0;
  return conv_thin16_workspace(d, pass);
}
extern "C" int srgan_conv2d_fprop_thin16(const srgan_conv_desc* d, const void* x, const float* w, const float* bias,
                                         void* y, int act, float slope, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(x && w && y, "null pointer");
  SRGAN_CHECK_ARG(dense_x(d), "thin16 conv: dense NHWC input only");
  return conv_thin16_launch(d, 0, x, w, bias, y, act, slope, ws, ws_bytes, (cudaStream_t)stream);
}
extern "C" int srgan_conv2d_dgrad_thin16(const srgan_conv_desc* d, const void* dy, const float* w, void* dx, void* ws,
                                         size_t ws_bytes, void* stream) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(dy && w && dx, "null pointer");
  return conv_thin16_launch(d, 1, dy, w, nullptr, dx, SRGAN_ACT_NONE, 0.f, ws, ws_bytes, (cudaStream_t)stream);
}
extern "C" int srgan_conv2d_wgrad_thin16(const srgan_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                                         void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(x && dy && (dw || dbias), "null pointer");
  SRGAN_CHECK_ARG(dense_x(d), "thin16 conv: dense NHWC input only");
  return conv_thin16_wgrad_launch(d, x, dy, dw, dbias, ws, ws_bytes, (cudaStream_t)stream);
}
extern "C" int srgan_conv2d_bf16_stat_rows(const srgan_conv_desc* d, int pass) {
  if (check_desc(d)) return 0;
  return conv_umma_bf16_stat_rows(d, pass);
}

extern "C" int srgan_conv2d_wgrad_plan(const srgan_conv_desc* d, int* splits, int* ctas) {
  if (int e = check_desc(d)) return e;
  SRGAN_CHECK_ARG(splits && ctas, "null pointer");
  conv_umma_wgrad_plan(d, splits, ctas);
  return SRGAN_OK;
}

extern "C" int srgan_colsum(const float* x, float* out, size_t rows, int C, void* stream) {
  SRGAN_CHECK_ARG(x && out && C >= 0, "bad argument");
  return colsum_launch(x, out, (long long)rows, C, nullptr, 0, (cudaStream_t)stream);
}
