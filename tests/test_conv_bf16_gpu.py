"""GPU: bf16-storage convolutions (tcgen05 kind::f16) through the C ABI against torch on the same bf16-rounded
operands (products exact in fp32, fp32 accumulation), i.e. the only differences are the accumulation order and the
final rounding of the output to bf16: rel-L2 <= 3e-3 (bf16 has 8 mantissa bits: 2^-9 = 2e-3 per element)."""
import pytest
import torch
import torch.nn.functional as F

import srgan_ops as ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
CL = torch.channels_last
TOL = 3e-3


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _nhwc_bf16(t):
    return t.to(DEV).to(torch.bfloat16).contiguous(memory_format=CL)


GEOMS = [  # N, C, H, W, K, R, stride, pad
    (4, 256, 32, 32, 256, 3, 1, 1),     # residual block
    (2, 64, 64, 64, 128, 4, 2, 1),      # down path
    (3, 128, 17, 23, 64, 3, 1, 1),      # ragged plane, 64-wide tile
    (2, 64, 16, 16, 8, 1, 1, 0),        # 1x1, narrow output
    (2, 512, 8, 8, 512, 4, 2, 1),       # deep discriminator layer
]


@pytest.mark.parametrize("g", GEOMS)
def test_fprop_bf16(g):
    N, C, H, W, K, R, stride, pad = g
    torch.manual_seed(0)
    x = _nhwc_bf16(torch.randn(N, C, H, W))
    w = _nhwc_bf16(torch.randn(K, C, R, R) * 0.05)
    b = torch.randn(K, device=DEV)
    d = ops._desc(N, H, W, C, K, R, R, stride, pad)
    assert ops._lib().srgan_conv2d_bf16_supported(d, 0) == 1
    y = torch.empty((N, K, d.P, d.Q), dtype=torch.bfloat16, device=DEV).contiguous(memory_format=CL)
    for act, slope, fn in ((ops.ACT_NONE, 0.0, lambda t: t), (ops.ACT_LRELU, 0.2, lambda t: F.leaky_relu(t, 0.2))):
        ops._call("srgan_conv2d_fprop_bf16", d, ops._p(x), ops._p(w), ops._p(b), ops._p(y), act, slope, ops._stream())
        ref = fn(F.conv2d(x.float(), w.float(), b, stride, pad))
        assert _rel(y.float(), ref) < TOL, (g, act)


@pytest.mark.parametrize("g", GEOMS)
def test_dgrad_bf16(g):
    N, C, H, W, K, R, stride, pad = g
    if K % 64:
        pytest.skip("dgrad reduces over K: needs K % 64 == 0")
    torch.manual_seed(1)
    d = ops._desc(N, H, W, C, K, R, R, stride, pad)
    if (H + 2 * pad - R) % stride:
        pytest.skip("input rows not covered by the output grid")
    dy = _nhwc_bf16(torch.randn(N, K, d.P, d.Q))
    w = _nhwc_bf16(torch.randn(K, C, R, R) * 0.05)
    assert ops._lib().srgan_conv2d_bf16_supported(d, 1) == 1
    nb = ops._lib().srgan_conv2d_bf16_workspace(d, 1)
    ws = ops._workspace(torch.device(DEV, torch.cuda.current_device()), nb)
    dx = torch.empty((N, C, H, W), dtype=torch.bfloat16, device=DEV).contiguous(memory_format=CL)
    ops._call("srgan_conv2d_dgrad_bf16", d, ops._p(dy), ops._p(w), None, ops._p(dx), ops._p(ws), nb, ops._stream())
    ref = torch.nn.grad.conv2d_input((N, C, H, W), w.float(), dy.float(), stride, pad)
    assert _rel(dx.float(), ref) < TOL, g
    if stride == 1:
        add = _nhwc_bf16(torch.randn(N, C, H, W))
        ops._call("srgan_conv2d_dgrad_bf16", d, ops._p(dy), ops._p(w), ops._p(add), ops._p(dx), ops._p(ws), nb,
                  ops._stream())
        assert _rel(dx.float(), ref + add.float()) < TOL, g


def test_bf16_res_conv_rate():
    """Not a pass/fail timing: prints the rate of the residual-block convolution at batch 64 next to the TF32 engine."""
    N, C, H, W, K = 64, 256, 32, 32, 256
    x = _nhwc_bf16(torch.randn(N, C, H, W))
    w = _nhwc_bf16(torch.randn(K, C, 3, 3) * 0.05)
    d = ops._desc(N, H, W, C, K, 3, 3, 1, 1)
    y = torch.empty((N, K, H, W), dtype=torch.bfloat16, device=DEV).contiguous(memory_format=CL)
    xf, wf = x.float().contiguous(memory_format=CL), w.float().contiguous(memory_format=CL)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

    def timed(fn, n=20):
        ts = []
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]
    t16 = timed(lambda: ops._call("srgan_conv2d_fprop_bf16", d, ops._p(x), ops._p(w), None, ops._p(y), 0, 0.0,
                                  ops._stream()))
    t32 = timed(lambda: ops.conv2d(xf, wf, None, 1, 1))
    fl = 2.0 * N * H * W * K * C * 9
    print("res conv fprop: bf16 %.1f us (%.0f TF/s)   tf32 %.1f us (%.0f TF/s)" %
          (t16 * 1e3, fl / t16 / 1e9, t32 * 1e3, fl / t32 / 1e9))
