// tcgen05 / TMEM / TMA implicit-GEMM convolution engine (TF32 operands, fp32 accumulate).
// Placeholder until the kernels land: reports "unsupported" so AUTO resolves to the FFMA engine.
#include "common.cuh"

namespace srgan {
bool conv_umma_supported(const srgan_conv_desc*, int) { return false; }
size_t conv_umma_workspace(const srgan_conv_desc*, int) { return 0; }
int conv_fprop_umma_launch(const srgan_conv_desc*, const float*, const float*, const float*, float*, int, float,
                           void*, size_t, cudaStream_t) { return SRGAN_E_UNSUPPORTED; }
int conv_dgrad_umma_launch(const srgan_conv_desc*, const float*, const float*, float*, void*, size_t,
                           cudaStream_t) { return SRGAN_E_UNSUPPORTED; }
int conv_wgrad_umma_launch(const srgan_conv_desc*, const float*, const float*, float*, float*, void*, size_t,
                           cudaStream_t) { return SRGAN_E_UNSUPPORTED; }
}  // namespace srgan
extern "C" int srgan_has_tcgen05(void) { return 0; }
