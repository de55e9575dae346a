"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/srgan_b200.h declares
(no kernel is launched here)."""
import ctypes
import os
import re
import subprocess

import pytest

import _srgan_lib


def _declared(repo_root):
    text = open(os.path.join(repo_root, "include", "srgan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srgan_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(repo_root):
    assert os.path.exists(_srgan_lib.LIB_PATH), "run python style-restricted_gan_b200/csrc/build.py"
    lib = ctypes.CDLL(_srgan_lib.LIB_PATH)
    names = _declared(repo_root)
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes signature table covers exactly the header
    assert sorted(_srgan_lib.SIGNATURES) == names


def test_binding_loads_and_reports_version():
    lib = _srgan_lib.load()
    assert lib.srgan_abi_version() == 1
    assert lib.srgan_reduce_scratch_bytes(0) > 0


def test_sass_is_sm100a(repo_root):
    out = subprocess.run(["cuobjdump", "-lelf", _srgan_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_bad_arguments_fail_loudly_without_a_gpu():
    """Argument validation happens before any CUDA call: status < 0 and a message, never a crash."""
    lib = _srgan_lib.load()
    d = _srgan_lib.ConvDesc()
    d.N, d.H, d.W, d.C, d.K, d.R, d.S, d.P, d.Q, d.stride, d.pad = 1, 8, 8, 4, 4, 3, 3, 5, 5, 1, 1   # P,Q wrong
    st = lib.srgan_conv2d_fprop(d, 16, 16, None, 16, 0, 0.0, 1, None, 0, None)
    assert st == -1 and b"output size mismatch" in lib.srgan_last_error()
    with pytest.raises(_srgan_lib.SrganKernelError):
        _srgan_lib.check(st, "srgan_conv2d_fprop")
    assert lib.srgan_inorm_fwd(16, 16, 16, 16, None, None, None, None, 1, 16, 12, 1e-5, 0, 0.0, None, 0, None) == -1


def test_product_ops_refuse_cpu_tensors():
    import torch
    import srgan_ops
    x = torch.zeros(1, 8, 4, 4)
    with pytest.raises(_srgan_lib.SrganKernelError):
        srgan_ops.instance_norm_act(x)
    with pytest.raises(_srgan_lib.SrganKernelError):
        srgan_ops.l1_mean(x, x)


def test_wgrad_split_plan_fills_whole_waves():
    """The tcgen05 wgrad launches one CTA per SM and round: the split count must not spill a handful of CTAs into an
    extra round (the first heuristic gave the residual blocks 9 taps x 33 splits = 297 CTAs = 2 rounds + 1 CTA).
    Host-side planning only: runs without a GPU."""
    import ctypes
    lib = _srgan_lib.load()
    SM = 148

    def plan(N, H, W, C, K, R, stride, pad):
        P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
        d = _srgan_lib.ConvDesc(N, H, W, C, K, R, R, P, Q, stride, pad, 0, 0, 0, 0)
        s, c = ctypes.c_int(0), ctypes.c_int(0)
        assert lib.srgan_conv2d_wgrad_plan(ctypes.byref(d), ctypes.addressof(s), ctypes.addressof(c)) == 0
        return s.value, c.value
    splits, ctas = plan(64, 32, 32, 256, 256, 3, 1, 1)          # residual block at batch 64
    assert ctas <= SM and ctas >= 0.9 * SM, (splits, ctas)
    for shape in ((64, 64, 64, 128, 256, 4, 2, 1), (64, 16, 16, 256, 512, 4, 2, 1), (64, 17, 17, 256, 512, 3, 1, 0),
                  (8, 32, 32, 256, 256, 3, 1, 1), (256, 32, 32, 256, 256, 3, 1, 1)):
        splits, ctas = plan(*shape)
        rounds = -(-ctas // SM)
        assert splits >= 1 and ctas > (rounds - 1) * SM + 0.5 * SM or rounds == 1, (shape, splits, ctas)
