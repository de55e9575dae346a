"""ctypes binding of libsrgan_b200.so (the C ABI declared in include/srgan_b200.h).

There is no CPU or PyTorch fallback behind this module: if the shared library is missing or a
kernel reports an error, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SRGAN_LIB (bring-up only): load another build of the library for A/B timing
LIB_PATH = os.environ.get("SRGAN_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libsrgan_b200.so")

c_int, c_float, c_size_t, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p
P = c_void_p


class ConvDesc(ctypes.Structure):
    """Mirror of `srgan_conv_desc`."""
    _fields_ = [(n, ctypes.c_int32) for n in ("N", "H", "W", "C", "K", "R", "S", "P", "Q", "stride", "pad")] + \
               [(n, ctypes.c_int64) for n in ("xs_n", "xs_h", "xs_w", "xs_c")]


DP = ctypes.POINTER(ConvDesc)

# name -> (restype, argtypes); every symbol of include/srgan_b200.h
SIGNATURES = {
    "srgan_last_error": (ctypes.c_char_p, []),
    "srgan_abi_version": (c_int, []),
    "srgan_has_tcgen05": (c_int, []),
    "srgan_conv2d_workspace": (c_size_t, [DP, c_int, c_int]),
    "srgan_conv2d_fprop": (c_int, [DP, P, P, P, P, c_int, c_float, c_int, P, c_size_t, P]),
    "srgan_conv2d_dgrad": (c_int, [DP, P, P, P, c_int, P, c_size_t, P]),
    "srgan_conv2d_wgrad_plan": (c_int, [DP, P, P]),
    "srgan_conv2d_bf16_supported": (c_int, [DP, c_int]),
    "srgan_conv2d_bf16_workspace": (c_size_t, [DP, c_int]),
    "srgan_conv2d_fprop_bf16": (c_int, [DP, P, P, P, P, c_int, c_float, P, P]),
    "srgan_conv2d_dgrad_bf16": (c_int, [DP, P, P, P, P, P, P, c_size_t, P]),
    "srgan_conv2d_bf16_stat_rows": (c_int, [DP, c_int]),
    "srgan_conv2d_thin16_supported": (c_int, [DP, c_int]),
    "srgan_conv2d_thin16_workspace": (c_size_t, [DP, c_int]),
    "srgan_conv2d_fprop_thin16": (c_int, [DP, P, P, P, P, c_int, c_float, P, c_size_t, P]),
    "srgan_conv2d_dgrad_thin16": (c_int, [DP, P, P, P, P, c_size_t, P]),
    "srgan_conv2d_wgrad_thin16": (c_int, [DP, P, P, P, P, P, c_size_t, P]),
    "srgan_inorm_stats_from_tiles": (c_int, [P, c_int, c_int, c_int, c_int, c_float, P, P, P]),
    "srgan_conv2d_wgrad_bf16": (c_int, [DP, P, P, P, P, c_size_t, P]),
    "srgan_conv2d_wgrad_bf16_plan": (c_int, [DP, P, P]),
    "srgan_cast_f32_bf16": (c_int, [P, P, c_size_t, P]),
    "srgan_grad_fold": (c_int, [P, P, P, c_size_t, P]),
    "srgan_act_bwd_bf16": (c_int, [P, P, P, c_size_t, c_int, c_float, P]),
    "srgan_cast_bf16_f32": (c_int, [P, P, c_size_t, P]),
    "srgan_inorm_mixed_workspace": (c_size_t, [c_int, c_int, c_int]),
    "srgan_inorm_mixed_counters": (c_size_t, [c_int, c_int]),
    "srgan_inorm_fwd_mixed": (c_int, [P, c_int, P, c_int, P, P, P, P, P, P, c_int, c_int, c_int, c_float, c_int,
                                      c_float, c_int, P, c_size_t, P, P]),
    "srgan_inorm_bwd_mixed": (c_int, [P, c_int, P, c_int, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float,
                                      P, c_size_t, P, P]),
    "srgan_inorm_onepass_plan": (c_int, [c_int, c_int, c_int, P, P]),
    "srgan_inorm_onepass_enable": (c_int, [c_int]),
    "srgan_inorm_onepass_max_clusters": (c_int, [c_int, c_int, c_int]),
    "srgan_conv2d_dgrad_add_supported": (c_int, [DP, c_int]),
    "srgan_conv2d_dgrad_add": (c_int, [DP, P, P, P, P, c_int, P, c_size_t, P]),
    "srgan_conv2d_wgrad": (c_int, [DP, P, P, P, P, c_int, P, c_size_t, P]),
    "srgan_conv2d_engine": (c_int, [DP, c_int]),
    "srgan_nchw_to_nhwc": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "srgan_nhwc_to_nchw": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "srgan_reflect_pad_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "srgan_reflect_pad_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "srgan_inorm_workspace": (c_size_t, [c_int, c_int, c_int]),
    "srgan_norm_partials_fp64": (c_int, []),
    "srgan_inorm_fwd": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_float, c_int, c_float, P, c_size_t, P]),
    "srgan_inorm_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, P, c_size_t, P]),
    "srgan_bnorm_image_stats": (c_int, [P, P, P, c_int, c_int, c_int, P, c_size_t, P]),
    "srgan_bnorm_batch_stats": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, P, P, c_float,
                                        P, P, P, P]),
    "srgan_bnorm_apply": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, P]),
    "srgan_bnorm_bwd_sums": (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, P, c_size_t, P]),
    "srgan_bnorm_bwd_coeffs": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "srgan_bnorm_bwd_apply": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, P]),
    "srgan_inorm_param_grads": (c_int, [P, P, P, P, P, P, P, c_int, c_int, P]),
    "srgan_condbias_fwd": (c_int, [P, P, P, P, c_int, c_int, c_int, P]),
    "srgan_condbias_bwd": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, P]),
    "srgan_reflect_pad_fwd_bf16": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "srgan_reflect_pad_bwd_bf16": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "srgan_avgpool2_add_fwd_mixed": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    "srgan_avgpool2_bwd_mixed": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "srgan_avgpool2_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "srgan_avgpool2_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "srgan_avgpool2_add_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    "srgan_avgpool3s2_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "srgan_avgpool3s2_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "srgan_lrelu_gap_fwd": (c_int, [P, P, c_int, c_int, c_int, c_float, P]),
    "srgan_lrelu_gap_bwd": (c_int, [P, P, P, c_int, c_int, c_int, c_float, P]),
    "srgan_act_bwd": (c_int, [P, P, P, c_size_t, c_int, c_float, P]),
    "srgan_add": (c_int, [P, P, P, c_size_t, P]),
    "srgan_colsum": (c_int, [P, P, c_size_t, c_int, P]),
    "srgan_softmax_fwd": (c_int, [P, P, c_int, c_int, P]),
    "srgan_softmax_bwd": (c_int, [P, P, P, c_int, c_int, P]),
    "srgan_cross_entropy_fwd": (c_int, [P, P, P, P, c_int, c_int, P]),
    "srgan_cross_entropy_bwd": (c_int, [P, P, P, P, c_int, c_int, P]),
    "srgan_prdc_pairdist2": (c_int, [P, P, P, c_int, c_int, c_int, P]),
    "srgan_prdc_kth_radius": (c_int, [P, P, c_int, c_int, P]),
    "srgan_prdc_counts": (c_int, [P, P, P, P, P, P, c_int, c_int, P]),
    "srgan_reparam_fwd": (c_int, [P, P, P, P, c_size_t, P]),
    "srgan_reparam_bwd": (c_int, [P, P, P, P, P, c_size_t, P]),
    "srgan_reduce_scratch_bytes": (c_size_t, [c_size_t]),
    "srgan_l1_mean_fwd": (c_int, [P, P, c_size_t, P, P, P]),
    "srgan_l1_mean_bwd": (c_int, [P, P, P, P, P, c_size_t, P]),
    "srgan_mse_const_fwd": (c_int, [P, c_float, c_size_t, P, P, P]),
    "srgan_mse_const_bwd": (c_int, [P, c_float, P, P, c_size_t, P]),
    "srgan_mse_fwd": (c_int, [P, P, c_size_t, P, P, P]),
    "srgan_mse_bwd": (c_int, [P, P, P, P, P, c_size_t, P]),
    "srgan_latent_losses_fwd": (c_int, [P, P, c_int, c_int, c_float, P, c_int, c_float, c_float, c_float, c_int,
                                        P, P]),
    "srgan_latent_losses_bwd": (c_int, [P, P, c_int, c_int, c_float, P, c_int, c_float, c_float, c_float, c_int,
                                        P, P, P, P, c_int, c_int, P]),
    "srgan_corrcoef_bwd": (c_int, [P, c_int, c_int, P, P, P, P]),
    "srgan_softhist_fwd": (c_int, [P, c_int, c_int, c_float, c_float, c_float, P, P]),
    "srgan_softhist_bwd": (c_int, [P, P, c_int, c_int, c_float, c_float, c_float, P, P]),
    "srgan_face_transform_smem": (c_size_t, [c_int, c_int, c_int, c_int]),
    "srgan_face_transform": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P, c_int, P, P, P]),
    "srgan_adam_step": (c_int, [P, P, P, P, c_size_t, c_float, c_float, c_float, c_float, c_int, P]),
    "srgan_adam_step_dev": (c_int, [P, P, P, P, c_size_t, P, P]),
}

_lib = None


class SrganKernelError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SrganKernelError(
            f"{LIB_PATH} not found: build it with `python style-restricted_gan_b200/csrc/build.py` "
            "(the SRGAN B200 kernels have no CPU / PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("SRGAN_LIB") and not hasattr(lib, name):
            continue              # an older build under A/B test may lack the newest entry points
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.srgan_abi_version() != 1:
        raise SrganKernelError("libsrgan_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().srgan_last_error().decode("utf-8", "replace")
        raise SrganKernelError(f"{what} failed with status {status}: {msg}")
