/*
 * srgan_b200.h -- C ABI of the B200-native SRGAN training-step kernels.
 *
 * This is the drop-in boundary underneath the Python surface of the reference's
 * pyfiles/ (model.py, util.py, util_notebook.py).  The reference has no FFI of its
 * own: every entry point below replaces a PyTorch library call that the reference
 * issues on the hot path; the call site it replaces is cited as
 * "ref: <file>:<line>" (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     name starts with h_;
 *   - activations are fp32, NHWC ("channels last"): x[n][h][w][c];
 *   - convolution filters are fp32, KRSC: w[k][r][s][c] (== torch channels_last of a
 *     [K,C,R,S] parameter);
 *   - every function enqueues on `stream` (a cudaStream_t passed as void*), never
 *     synchronises, never allocates; scratch is passed in (query the size first);
 *   - return value: 0 = ok, <0 = invalid argument (SRGAN_E_*), >0 = cudaError_t.
 *     Nothing throws across this boundary.  srgan_last_error() gives a message.
 */
#ifndef SRGAN_B200_H_
#define SRGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRGAN_ABI_VERSION 1

/* status codes */
#define SRGAN_OK            0
#define SRGAN_E_BADARG     -1
#define SRGAN_E_UNSUPPORTED -2
#define SRGAN_E_WORKSPACE  -3

/* activation ids (epilogues / fused norm) */
#define SRGAN_ACT_NONE  0
#define SRGAN_ACT_RELU  1
#define SRGAN_ACT_LRELU 2
#define SRGAN_ACT_TANH  3

/* convolution engine selection */
#define SRGAN_CONV_AUTO 0   /* tcgen05 (TF32 in, fp32 accumulate) where the shape qualifies, else FFMA */
#define SRGAN_CONV_FP32 1   /* fp32 FFMA implicit GEMM: exact-fp32 parity path and odd shapes */
#define SRGAN_CONV_TF32 2   /* force the tcgen05 path; SRGAN_E_UNSUPPORTED when the shape does not qualify */

typedef struct srgan_conv_desc {
  int32_t N, H, W, C;   /* input activation x[N][H][W][C] */
  int32_t K, R, S;      /* filter w[K][R][S][C] */
  int32_t P, Q;         /* output y[N][P][Q][K]; P = (H + 2*pad - R)/stride + 1 */
  int32_t stride, pad;  /* symmetric, zero padding (reflect padding = srgan_reflect_pad_* + pad 0) */
  /* element strides of x for fprop/wgrad; all 0 means dense NHWC.  Lets the stem read an
     NCHW image batch in place. */
  int64_t xs_n, xs_h, xs_w, xs_c;
} srgan_conv_desc;

const char* srgan_last_error(void);
int  srgan_abi_version(void);
/* 1 when the library was built with the tcgen05 convolution kernels */
int  srgan_has_tcgen05(void);

/* ---------------------------------------------------------------- convolutions
 * ref: nn.Conv2d forward  pyfiles/model.py:191,193,212,215,232,262,269,274,302,309,328-331,
 *      358,364,369,385,419,425,430,445 ; nn.ConvTranspose2d pyfiles/model.py:227,230
 *      (a transposed convolution's forward is srgan_conv2d_dgrad of the mirrored conv,
 *      its input gradient is srgan_conv2d_fprop, its filter gradient srgan_conv2d_wgrad
 *      with x and dy swapped).
 * fprop : y = act(conv(x, w) + bias)
 * dgrad : dx = conv_transpose(dy, w)            (overwrites dx)
 * wgrad : dw = sum_pixels x (*) dy ; dbias = sum_pixels dy (either may be NULL)
 */
size_t srgan_conv2d_workspace(const srgan_conv_desc* d, int pass /*0 fprop,1 dgrad,2 wgrad*/, int engine);
int srgan_conv2d_fprop(const srgan_conv_desc* d, const float* x, const float* w, const float* bias,
                       float* y, int act, float slope, int engine,
                       void* workspace, size_t workspace_bytes, void* stream);
int srgan_conv2d_dgrad(const srgan_conv_desc* d, const float* dy, const float* w, float* dx,
                       int engine, void* workspace, size_t workspace_bytes, void* stream);
int srgan_conv2d_wgrad(const srgan_conv_desc* d, const float* x, const float* dy, float* dw,
                       float* dbias, int engine, void* workspace, size_t workspace_bytes, void* stream);
/* dx = dgrad(dy) + addend (addend: layout of dx).  The residual blocks (ref SingleResidualBlock, pyfiles/model.py:196-201)
 * send two gradients into their input - through c1 and through the skip connection; autograd would add them with a
 * separate pass over the tensor, here the skip gradient is added in the dgrad epilogue.  tcgen05 engine, stride 1. */
int srgan_conv2d_dgrad_add_supported(const srgan_conv_desc* d, int engine);
int srgan_conv2d_dgrad_add(const srgan_conv_desc* d, const float* dy, const float* w, const float* addend,
                           float* dx, int engine, void* workspace, size_t workspace_bytes, void* stream);
/* ---- bf16-storage engine (the generator's trunk; ref layers pyfiles/model.py:188-249).  x, w, y, dy, dx, addend are
 * NHWC / KRSC tensors of bfloat16 (passed as void*), bias / dw are fp32, accumulation is fp32 in TMEM (tcgen05
 * kind::f16).  Plain layers only: fprop needs C % 64 == 0, dgrad K % 64 == 0, wgrad both (pass: 0 fprop, 1 dgrad,
 * 2 wgrad).  w is the bf16 shadow of the fp32 master filter (srgan_cast_f32_bf16).  addend may be NULL (stride 1 only
 * otherwise).  wgrad writes dw[K][R][S][C] in fp32 (split partials in the workspace, fixed-order reduction). */
#define SRGAN_DT_F32  0
#define SRGAN_DT_BF16 1
int srgan_conv2d_bf16_supported(const srgan_conv_desc* d, int pass);
size_t srgan_conv2d_bf16_workspace(const srgan_conv_desc* d, int pass);
int srgan_conv2d_fprop_bf16(const srgan_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                            int act, float slope, float* tile_stats, void* stream);
int srgan_conv2d_dgrad_bf16(const srgan_conv_desc* d, const void* dy, const void* w, const void* addend, void* dx,
                            float* tile_stats, void* workspace, size_t workspace_bytes, void* stream);
/* tile_stats (optional, forward use of fprop / of dgrad as the transposed convolution): the epilogue also writes the
 * per-tile statistics the instance norm behind the convolution needs (ref CBINorm2d pyfiles/model.py:54-67 follows
 * every generator convolution :236-249), sparing that norm its own statistics pass:
 *   tile_stats[((n * rows + r) * K_out + k) * 2 + {0,1}] = sum, sum of squares over the 128 pixels of tile r of image
 *   n of the values AS STORED (after rounding to bf16), rows = srgan_conv2d_bf16_stat_rows(d, pass) (0: shape cannot
 *   provide them: tiles must lie inside one image; needs bias == NULL, act none, no addend).
 * srgan_inorm_stats_from_tiles folds the rows of every image in row order in fp64 (independent of the batch). */
int srgan_conv2d_bf16_stat_rows(const srgan_conv_desc* d, int pass);
/* Thin RGB layers of the bf16 engine ("thin16"): the 3-channel side (image / image gradient) and the filter are fp32,
 * the fat side (the 64-channel activation of the bf16 trunk or its gradient) is bfloat16 - the stem writes what the
 * trunk reads and the head reads what the trunk wrote, without an fp32 copy of the largest tensors of the step.
 * ref: nn.Conv2d(nch_in, nch, 7, 1, 3) / nn.Conv2d(nch, nch_in, 7, 1, 3) + Tanh, SingleGenerator pyfiles/model.py:
 * 280-318 (first and last layer), and their backward passes.
 *   fprop : C <= 4: x fp32 -> y bf16 (bias, activation)           K <= 4: x bf16 -> y fp32 (bias, activation)
 *   dgrad : K <= 4: dy fp32 -> dx bf16                            C <= 4: dy bf16 -> dx fp32
 *   wgrad : C <= 4: x fp32, dy bf16 -> dw, dbias fp32             K <= 4: x bf16, dy fp32 -> dw, dbias fp32
 * TF32 tensor-core arithmetic on the fp32 side's packed rows where the fp32 tensor is the streamed operand, kind::f16
 * (thin side rounded to bf16) where the bf16 tensor is.  pass: 0 fprop, 1 dgrad, 2 wgrad. */
int srgan_conv2d_thin16_supported(const srgan_conv_desc* d, int pass);
size_t srgan_conv2d_thin16_workspace(const srgan_conv_desc* d, int pass);
int srgan_conv2d_fprop_thin16(const srgan_conv_desc* d, const void* x, const float* w, const float* bias, void* y,
                              int act, float slope, void* workspace, size_t workspace_bytes, void* stream);
int srgan_conv2d_dgrad_thin16(const srgan_conv_desc* d, const void* dy, const float* w, void* dx, void* workspace,
                              size_t workspace_bytes, void* stream);
int srgan_conv2d_wgrad_thin16(const srgan_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                              void* workspace, size_t workspace_bytes, void* stream);   /* dw or dbias may be NULL */
int srgan_inorm_stats_from_tiles(const float* tile_stats, int rows, int N, int HW, int C, float eps, float* mean,
                                 float* rstd, void* stream);
int srgan_conv2d_wgrad_bf16(const srgan_conv_desc* d, const void* x, const void* dy, float* dw, void* workspace,
                            size_t workspace_bytes, void* stream);
int srgan_conv2d_wgrad_bf16_plan(const srgan_conv_desc* d, int* splits, int* ctas);
/* g[i] += p1[i] (+ p2[i], may be NULL), then p1[i] = p2[i] = 0; n % 4 == 0, 16-byte aligned.  Folds the later
 * gradient contributions of one backward pass (written by the wgrad / norm kernels into zeroed copies of the flat
 * gradient buffer instead of being added tensor by tensor; ref: autograd accumulation of `.grad` across the several
 * generator / encoder passes of one loss.backward(), pyfiles/util_notebook.py:664-665, 689) in arrival order. */
int srgan_grad_fold(float* g, float* p1, float* p2, size_t n, void* stream);
/* bf16 discriminator tower (engine bf16; ref: the LeakyReLU(0.2) after every tower convolution, pyfiles/model.py:
 * 255-346): dz = dy * act'(y) on bf16 tensors (y = the activation OUTPUT, as in srgan_act_bwd), and the widening cast
 * at the tower's fp32 boundary (the 1- / 4-logit heads read fp32). */
int srgan_act_bwd_bf16(const void* dy, const void* y, void* dx, size_t n, int act, float slope, void* stream);
int srgan_cast_bf16_f32(const void* src, float* dst, size_t n, void* stream);
/* dst[i] = bf16(src[i]) (round to nearest even), n elements; both 16-byte aligned */
int srgan_cast_f32_bf16(const float* src, void* dst, size_t n, void* stream);
/* introspection (tests, tools): pixel splits and CTAs of the tcgen05 wgrad launch for this layer (host only) */
int srgan_conv2d_wgrad_plan(const srgan_conv_desc* d, int* splits, int* ctas);
/* which engine AUTO resolves to for this shape/pass: SRGAN_CONV_FP32 or SRGAN_CONV_TF32 */
int srgan_conv2d_engine(const srgan_conv_desc* d, int pass);

/* ---------------------------------------------------------------- layout / padding
 * ref: images arrive NCHW from the DataLoader (notebook/03-train... cell 24:16-18);
 *      padding_mode="reflect" pyfiles/model.py:358,364,419,425 */
int srgan_nchw_to_nhwc(const float* x, float* y, int N, int C, int H, int W, void* stream);
int srgan_nhwc_to_nchw(const float* x, float* y, int N, int C, int H, int W, void* stream);
int srgan_reflect_pad_fwd(const float* x, float* y, int N, int H, int W, int C, int pad, void* stream);
int srgan_reflect_pad_bwd(const float* dy, float* dx, int N, int H, int W, int C, int pad, void* stream);

/* ---------------------------------------------------------------- fused instance norm
 * ref: CBINorm2d.forward pyfiles/model.py:54-67 (F.instance_norm + ConBias + affine),
 *      nn.InstanceNorm2d(affine=False) pyfiles/model.py:178, followed by ReLU
 *      (:199,240,246) / LeakyReLU(0.2) (:357,361,417,422) / residual add (:201).
 *   xh = (x - mean_hw) * rstd_hw ; v = (xh + cbias[n][c]) * gamma[c] + beta[c]
 *   y  = act(v) (+ residual)
 * gamma/beta/cbias/residual may be NULL.  mean/rstd [N*C] are written by fwd and read by bwd.
 * bwd writes dx and the per-(n,c) sums s1 = sum dv, s2 = sum dv*xh  (dv = dy*act'(v));
 * srgan_inorm_param_grads turns (s1,s2) into dgamma, dbeta (accumulated over n in a fixed
 * order, overwritten) and dcbias[n][c].
 */
size_t srgan_inorm_workspace(int N, int HW, int C);   /* bytes of slice-partial scratch for fwd and bwd */
/* 1 when the slice partials are accumulated in fp64 over fixed 4-row atoms: the statistics of an image are then
 * independent of N (of how the grid planner slices the plane), which data-parallel reproducibility relies on. */
int srgan_norm_partials_fp64(void);
int srgan_inorm_fwd(const float* x, float* y, float* mean, float* rstd,
                    const float* gamma, const float* beta, const float* cbias, const float* residual,
                    int N, int HW, int C, float eps, int act, float slope,
                    void* workspace, size_t workspace_bytes, void* stream);
int srgan_inorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, const float* cbias,
                    float* dx, float* s1, float* s2,
                    int N, int HW, int C, int act, float slope,
                    void* workspace, size_t workspace_bytes, void* stream);
int srgan_inorm_param_grads(const float* s1, const float* s2, const float* gamma, const float* cbias,
                            float* dgamma, float* dbeta, float* dcbias, int N, int C, void* stream);
/* Mixed storage (norm8.cu): x / dx have dtype x_dtype, y / dy / residual have dtype y_dtype (SRGAN_DT_F32 or
 * SRGAN_DT_BF16, independently); statistics, parameters and arithmetic stay fp32 (partials fp64 over fixed atoms, as
 * above).  Otherwise as srgan_inorm_fwd / srgan_inorm_bwd.  The generator uses f32 -> bf16 behind its RGB stem,
 * bf16 -> bf16 in the trunk and bf16 -> f32 in front of the RGB head.
 * workspace: srgan_inorm_mixed_workspace(N, HW, C) bytes of scratch (slice partials).
 * counters: srgan_inorm_mixed_counters(N, C) bytes of int32 ticket counters that are ZERO on entry; the kernels leave
 * them zero (the CTA that completes an image folds its partials and resets the counter), so one zero-initialised
 * buffer serves every launch issued on the same stream. */
size_t srgan_inorm_mixed_workspace(int N, int HW, int C);
size_t srgan_inorm_mixed_counters(int N, int C);
int srgan_inorm_fwd_mixed(const void* x, int x_dtype, void* y, int y_dtype, float* mean, float* rstd,
                          const float* gamma, const float* beta, const float* cbias, const void* residual,
                          int N, int HW, int C, float eps, int act, float slope, int stats_given,
                          void* workspace, size_t workspace_bytes, int* counters, void* stream);
/* stats_given != 0: mean / rstd are inputs (srgan_inorm_stats_from_tiles); only the apply kernel runs. */
/* One-pass form (norm8c.cu): when a thread-block cluster of <= 8 CTAs can hold the resident tensor of one image (x
 * forward, dy backward; <= 64 KB per CTA, C <= 256) the two entry points above run ONE kernel that reads x once and
 * writes y once (backward: dy resident, x streamed twice, dx written once).  The choice depends on (HW, C, storage
 * type) only - never on N - and the statistics keep the atoms of the two-kernel path.  Returns 1 and the cluster size
 * / pixels per CTA when the plane is eligible, 0 otherwise (or while the path is switched off, the default: measured
 * slower than two kernels at batch 64, see srgan_inorm_onepass_enable). */
int srgan_inorm_onepass_plan(int HW, int C, int resident_dtype, int* cluster, int* slice_px);
/* Process-wide switch of the one-pass form (default: OFF, or SRGAN_NORM_ONEPASS from the environment); on < 0 only
 * queries.  Returns the previous setting.  Do not flip it while a captured CUDA graph of a step is alive. */
int srgan_inorm_onepass_enable(int on);
/* Introspection: clusters of the one-pass forward kernel the current device holds at once for this plane
 * (cudaOccupancyMaxActiveClusters); 0 = plane not eligible, -1 = query failed. */
int srgan_inorm_onepass_max_clusters(int HW, int C, int resident_dtype);
int srgan_inorm_bwd_mixed(const void* dy, int y_dtype, const void* x, int x_dtype, const float* mean,
                          const float* rstd, const float* gamma, const float* beta, const float* cbias,
                          void* dx, float* s1, float* s2, int N, int HW, int C, int act, float slope,
                          void* workspace, size_t workspace_bytes, int* counters, void* stream);

/* ---------------------------------------------------------------- batch-statistics norms
 * ref: CBBNorm2d / _CBBNorm.forward pyfiles/model.py:75-171 and nn.BatchNorm2d(affine=True) chosen by
 *      get_norm_layer("batch", ...) :173-177.
 *   CBBNorm2d (cond = 1):  y = act(((x - m_nc) * r_c + cbias[n][c]) * gamma[c] + beta[c])
 *   BatchNorm2d (cond = 0): y = act((x - mu_c) * r_c * gamma[c] + beta[c])
 * m_nc = mean over the pixels of image n, mu_c / r_c = batch mean / rsqrt(biased batch variance + eps) in training,
 * the running statistics otherwise (training updates them: momentum, unbiased variance).
 * Staged so that data-parallel ranks can all-gather the [N][C] tables between stages (N_all rows, this rank owns
 * rows [n0, n0 + N_loc)):
 *   forward : image_stats(x) -> mean_nc, m2_nc ; batch_stats(tables) -> mean, rstd, batch_mean ; apply
 *   backward: bwd_sums(dy, x) -> s1 = sum dv, s2 = sum dv*xh ; bwd_coeffs(tables) -> m1, m2 ;
 *             bwd_apply: dx = rstd*gamma*(dv - m1 - xh*m2) ; parameter gradients: srgan_inorm_param_grads(s1, s2). */
int srgan_bnorm_image_stats(const float* x, float* mean_nc, float* m2_nc, int N, int HW, int C,
                            void* workspace, size_t workspace_bytes, void* stream);   /* srgan_inorm_workspace */
int srgan_bnorm_batch_stats(const float* mean_nc_all, const float* m2_nc_all, int N_all, int n0, int N_loc,
                            int HW, int C, float eps, int cond, int training, float* running_mean,
                            float* running_var, float momentum, float* mean, float* rstd, float* batch_mean,
                            void* stream);
int srgan_bnorm_apply(const float* x, float* y, const float* mean, const float* rstd, const float* gamma,
                      const float* beta, const float* cbias, const float* residual, int N, int HW, int C,
                      int act, float slope, void* stream);
int srgan_bnorm_bwd_sums(const float* dy, const float* x, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, const float* cbias, float* s1, float* s2,
                         int N, int HW, int C, int act, float slope, void* workspace, size_t workspace_bytes,
                         void* stream);
int srgan_bnorm_bwd_coeffs(const float* s1_all, const float* s2_all, const float* mean, const float* rstd,
                           const float* batch_mean, int N_all, int n0, int N_loc, int HW, int C, int cond,
                           int training, float* m1, float* m2, void* stream);
int srgan_bnorm_bwd_apply(const float* dy, const float* x, const float* mean, const float* rstd,
                          const float* gamma, const float* beta, const float* cbias, const float* m1,
                          const float* m2, float* dx, int N, int HW, int C, int act, float slope, void* stream);

/* ---------------------------------------------------------------- conditional bias
 * ref: ConBias = nn.Sequential(nn.Linear(num_con, C), nn.Tanh()) pyfiles/model.py:16-19,57
 *   t[n][c] = tanh(sum_j con[n][j] * w[c][j] + b[c])
 * bwd: dpre = dt * (1 - t^2); dw[c][j] = sum_n dpre*con ; db[c] = sum_n dpre ;
 *      dcon[n][j] = sum_c dpre * w[c][j]   (any output may be NULL) */
int srgan_condbias_fwd(const float* con, const float* w, const float* b, float* t,
                       int N, int J, int C, void* stream);
int srgan_condbias_bwd(const float* dt, const float* t, const float* con, const float* w,
                       float* dw, float* db, float* dcon, int N, int J, int C, void* stream);

/* ---------------------------------------------------------------- pooling / elementwise
 * ref: nn.AvgPool2d(2,2) pyfiles/model.py:365,368,426,429 ; nn.AvgPool2d(3, stride=2, padding=1,
 *      count_include_pad=False) :286,324 ; LeakyReLU(0.2)+AdaptiveAvgPool2d(1) :394,454 ;
 *      LeakyReLU(0.01) :263,270,303,310 ; Tanh :248 ; residual/shortcut adds :201,375,436 */
/* bf16 storage variants for the encoder's trunk under the bf16 engine (ref BasicBlock_classification / BasicBlock
 * pyfiles/model.py:412-450, 355-380: reflect-padded 3x3 convolutions, average pooling joined with the shortcut):
 * reflect padding of bf16 NHWC tensors; avgpool2(a: bf16) + b (fp32) -> y (fp32) and its backward dy (fp32) -> da (bf16).
 * C % 8 == 0. */
int srgan_reflect_pad_fwd_bf16(const void* x, void* y, int N, int H, int W, int C, int pad, void* stream);
int srgan_reflect_pad_bwd_bf16(const void* dy, void* dx, int N, int H, int W, int C, int pad, void* stream);
int srgan_avgpool2_add_fwd_mixed(const void* a_bf16, const float* b, float* y, int N, int H, int W, int C, void* stream);
int srgan_avgpool2_bwd_mixed(const float* dy, void* dx_bf16, int N, int H, int W, int C, void* stream);
int srgan_avgpool2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream);
int srgan_avgpool2_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream);
/* y = avgpool2(a) + b  (encoder block tail: cmp-conv pooled + shortcut) */
int srgan_avgpool2_add_fwd(const float* a, const float* b, float* y, int N, int H, int W, int C, void* stream);
int srgan_avgpool3s2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream);
int srgan_avgpool3s2_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream);
/* f[n][c] = mean_hw lrelu(x[n][hw][c]) */
int srgan_lrelu_gap_fwd(const float* x, float* f, int N, int HW, int C, float slope, void* stream);
int srgan_lrelu_gap_bwd(const float* df, const float* x, float* dx, int N, int HW, int C, float slope, void* stream);
/* dx = dy * act'(.) evaluated from the saved OUTPUT y of the activation */
int srgan_act_bwd(const float* dy, const float* y, float* dx, size_t n, int act, float slope, void* stream);
int srgan_add(const float* a, const float* b, float* y, size_t n, void* stream);
/* column sums of a [rows][C] matrix (bias gradients): out[c] = sum_r x[r][c] */
int srgan_colsum(const float* x, float* out, size_t rows, int C, void* stream);

/* ---------------------------------------------------------------- encoder head
 * ref: softmax over the class dimension, nn.Softmax() pyfiles/model.py:333-334 ;
 *      reparametrize pyfiles/model.py:398-402,459-463 : z = eps*exp(0.5*logvar) + mu */
int srgan_softmax_fwd(const float* x, float* y, int N, int J, void* stream);
int srgan_softmax_bwd(const float* dy, const float* y, float* dx, int N, int J, void* stream);
int srgan_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z, size_t n, void* stream);
int srgan_reparam_bwd(const float* dz, const float* logvar, const float* eps, float* dmu, float* dlogvar,
                      size_t n, void* stream);

/* ---------------------------------------------------------------- image / patch losses
 * ref: torch.mean(torch.abs(a-b)) pyfiles/util_notebook.py:295,309,348,359,625,639,676,686 ;
 *      get_loss_D (MSE against a constant) pyfiles/util.py:457-462 ;
 *      get_domainloss_D (MSE between two tensors) pyfiles/util.py:464-468
 * fwd kernels write ONE fp32 scalar to out (overwrite); reductions use a fixed order
 * (bit-reproducible run to run).  scratch: srgan_reduce_scratch_bytes(n).          */
size_t srgan_reduce_scratch_bytes(size_t n);
int srgan_l1_mean_fwd(const float* a, const float* b, size_t n, float* out, void* scratch, void* stream);
/* da = g[0]*sign(a-b)/n ; db = -da  (either may be NULL); g is a device scalar */
int srgan_l1_mean_bwd(const float* a, const float* b, const float* g, float* da, float* db, size_t n, void* stream);
int srgan_mse_const_fwd(const float* x, float target, size_t n, float* out, void* scratch, void* stream);
int srgan_mse_const_bwd(const float* x, float target, const float* g, float* dx, size_t n, void* stream);
int srgan_mse_fwd(const float* a, const float* b, size_t n, float* out, void* scratch, void* stream);
int srgan_mse_bwd(const float* a, const float* b, const float* g, float* da, float* db, size_t n, void* stream);

/* ---------------------------------------------------------------- latent batch losses
 * ref: conventional KL pyfiles/util_notebook.py:300-304,630-634 ; batch-KL :314-320,644-650 ;
 *      corrcoef / corrcoef_loss pyfiles/util.py:470-517 ; GaussianHistogram /
 *      histogram_imitation pyfiles/util.py:521-553.
 * mu is [n][D] (row = sample), D <= 32, bins <= 64.
 * out (fp32) layout, SRGAN_LATENT_OUT_FLOATS(D,bins) floats:
 *   [0] batch-KL  [1] corr loss  [2] hist loss  [3] conventional KL
 *   [4 .. 4+D)        mean_d
 *   [.. +D)           var_d   (unbiased * n_cfg/(n_cfg-1))
 *   [.. +D)           unbiased variance c_dd (diagonal of the covariance)
 *   [.. +D*D)         corrcoef matrix (clamped)
 *   [.. +D*bins)      soft histogram h[d][b]
 *   [.. +D)           histogram mass S_d = sum_b h[d][b]
 * flags: bit0 batch-KL, bit1 corr, bit2 hist, bit3 conventional KL (needs logvar).
 * One CTA, fixed summation order: results are bit-identical for a given (mu, n).
 */
#define SRGAN_LATENT_BKL  1
#define SRGAN_LATENT_CORR 2
#define SRGAN_LATENT_HIST 4
#define SRGAN_LATENT_KL   8
#define SRGAN_LATENT_OUT_FLOATS(D, bins) (4 + 4 * (D) + (D) * (D) + (D) * (bins))
int srgan_latent_losses_fwd(const float* mu, const float* logvar, int n, int D, float n_cfg,
                            const float* hist_target, int bins, float hist_min, float hist_max, float sigma,
                            int flags, float* out, void* stream);
/* dmu[n][D] = sum_k g4[k] * dLoss_k/dmu, g4 = DEVICE pointer to the 4 upstream gradients (same
 * order as out[0..4)); dlogvar likewise for the KL term (may be NULL).  Reads the statistics fwd
 * left in `out`.  Only rows [row0,row0+rows) of the batch are differentiated and written to
 * dmu[rows][D] / dlogvar[rows][D] (a data-parallel rank's slice of an all-gathered batch).   */
int srgan_latent_losses_bwd(const float* mu, const float* logvar, int n, int D, float n_cfg,
                            const float* hist_target, int bins, float hist_min, float hist_max, float sigma,
                            int flags, const float* out, const float* g4,
                            float* dmu, float* dlogvar, int row0, int rows, void* stream);
/* stats = the `out` blob of srgan_latent_losses_fwd(flags with SRGAN_LATENT_CORR) */
int srgan_corrcoef_bwd(const float* mu, int n, int D, const float* stats, const float* dcorr,
                       float* dmu, void* stream);
int srgan_softhist_fwd(const float* x, int n, int bins, float hist_min, float hist_max, float sigma,
                       float* h, void* stream);
int srgan_softhist_bwd(const float* x, const float* dh, int n, int bins, float hist_min, float hist_max,
                       float sigma, float* dx, void* stream);

/* ---------------------------------------------------------------- optimizer
 * ref: optim.Adam(..., betas=(0.5,0.999)) pyfiles/util_notebook.py:117-131,500-507 (torch.optim.Adam
 *      semantics: bias-corrected, eps added to sqrt(v_hat), no weight decay, no amsgrad).
 * One launch updates a whole flat parameter buffer.                                  */
int srgan_adam_step(float* p, const float* g, float* m, float* v, size_t n,
                    float lr, float beta1, float beta2, float eps, int step, void* stream);
/* Same update, step-dependent scalars in DEVICE memory: hyper = [lr, beta1, beta2, eps, 1 - beta1^t, sqrt(1 - beta2^t)]
 * (the caller uploads them before each step), so the launch can be part of a replayed CUDA graph. */
int srgan_adam_step_dev(float* p, const float* g, float* m, float* v, size_t n, const float* hyper, void* stream);

/* ---------------------------------------------------------------- input pipeline
 * ref: notebook/01-train_Conventional_SingleGAN.ipynb cell 9 (transform["train"]: CenterCrop(178) -> Resize(128) ->
 *      RandomHorizontalFlip -> ToTensor -> MinMax(True)), pyfiles/dataset.py:127-141, pyfiles/util.py:108-116,148-153.
 * img: B decoded RGB images [B][H][W][3] uint8 (device); y: [B][out][out][3] fp32 NHWC (device) = the channels-last
 * storage of the logical [B,3,out,out] batch.  coef_* / bounds_*: Pillow's fixed-point (22 fractional bits)
 * triangle-filter tables for crop -> out along each axis ([out][ksize] int32 and [out][2] = first tap, tap count), as
 * computed by dataset.resample_coeffs; flip: B bytes (non-zero = mirror) or NULL.  Bit-exact with the CPU pipeline. */
size_t srgan_face_transform_smem(int crop, int out, int ksize_h, int ksize_v);
int srgan_face_transform(const uint8_t* img, int B, int H, int W, int crop, int out, const int* coef_h,
                         const int* bounds_h, int ksize_h, const int* coef_v, const int* bounds_v, int ksize_v,
                         const uint8_t* flip, float* y, void* stream);

/* ---------------------------------------------------------------- notebook-04 classifier loss, PRDC evaluation (f4)
 * ref: criterion = nn.CrossEntropyLoss() ; loss = criterion(net(x), label)   notebook 04_Facial_Recognition-Encoder
 *      cells 18, 22 (the net already ends in a softmax, pyfiles/model.py:484-508: the loss sees probabilities).
 *   loss = mean_n (logsumexp_j x[n][j] - x[n][label[n]]) ; loss_rows [N] scratch holds the per-row terms (fixed-order
 *   mean) ; bwd: dx[n][j] = (softmax(x[n])[j] - [j == label[n]]) * gout[0] / N.   label: int64. */
int srgan_cross_entropy_fwd(const float* x, const long long* label, float* loss, float* loss_rows, int N, int J,
                            void* stream);
int srgan_cross_entropy_bwd(const float* x, const long long* label, const float* gout, float* dx, int N, int J,
                            void* stream);
/* ref: GAN_evaluation.get_prdc pyfiles/evaluation.py:98-110 -> prdc.compute_prdc(real_features, fake_features,
 *      nearest_k) of the un-vendored dependency prdc==0.2 (Docker/requirements.txt:13; published algorithm restated in
 *      oracle/eval_oracle.py).
 *   srgan_prdc_pairdist2 : d2[i][j] = sum_k (a[i][k] - b[j][k])^2, fp64, k ascending  (a [N][D], b [M][D] fp32)
 *   srgan_prdc_kth_radius: radius[i] = (k+1)-th smallest entry of row i of the SELF distance matrix [N][N]
 *                          (= squared distance to the k-th nearest other sample; duplicates counted one by one)
 *   srgan_prdc_counts    : on d2 of (real x fake) [N][M]:
 *                            col_hits_real[j] = #{i : d2[i][j] < r_real[i]}   (precision = mean_j [hits > 0],
 *                                                                             density = sum_j hits / (k M))
 *                            row_hits_fake[i] = #{j : d2[i][j] < r_fake[j]}   (recall = mean_i [hits > 0])
 *                            row_min_in[i]    = [min_j d2[i][j] < r_real[i]]  (coverage = mean_i)
 *   Squared distances order like the reference's Euclidean distances; the integer counts are the result. */
int srgan_prdc_pairdist2(const float* a, const float* b, double* d2, int N, int M, int D, void* stream);
int srgan_prdc_kth_radius(const double* d2_self, double* radius, int N, int k, void* stream);
int srgan_prdc_counts(const double* d2_real_fake, const double* r_real, const double* r_fake, int* col_hits_real,
                      int* row_hits_fake, int* row_min_in, int N, int M, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRGAN_B200_H_ */
