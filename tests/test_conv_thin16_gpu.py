"""GPU: the thin RGB layers of the bf16 engine ("thin16", include/srgan_b200.h): the 3-channel side stays fp32, the fat
64-channel side is bf16.  Against fp64 torch convolutions on the SAME inputs (the bf16 side rounded as stored); the
error left is the tensor-core operand format (TF32 where the fp32 tensor streams, bf16 where the bf16 tensor does) and
the rounding of a bf16 result: rel-L2 <= 4e-3.
ref: first / last layer of SingleGenerator pyfiles/model.py:280-318, Encoder first_layer :519, discriminator stems."""
import pytest
import torch
import torch.nn.functional as F

import srgan_ops as ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
CL = torch.channels_last
BF = torch.bfloat16
TOL = 4e-3


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _ws(d, p):
    nb = ops._lib().srgan_conv2d_thin16_workspace(d, p)
    return ops._workspace(torch.device(DEV, torch.cuda.current_device()), nb), nb


# N, H, W, fat channels, R, stride, pad
STEMS = [
    (3, 128, 128, 64, 7, 1, 3),      # generator stem
    (64, 128, 128, 64, 7, 1, 3),     # ... at the production batch (multi-wave persistent loop)
    (4, 128, 128, 64, 7, 2, 1),      # encoder first layer
    (5, 128, 128, 64, 4, 2, 1),      # discriminator 1 stem
    (5, 64, 64, 32, 4, 2, 1),        # discriminator 2 stem (32-wide tile)
    (2, 37, 53, 64, 7, 1, 3),        # ragged plane
]


@pytest.mark.parametrize("g", STEMS)
def test_fprop_thin_input_bf16_output(g):
    N, H, W, K, R, stride, pad = g
    torch.manual_seed(0)
    x = (torch.rand(N, 3, H, W, device=DEV) * 2 - 1).contiguous(memory_format=CL)
    w = (torch.randn(K, 3, R, R, device=DEV) * 0.1).contiguous(memory_format=CL)
    b = torch.randn(K, device=DEV) * 0.1
    d = ops._desc(N, H, W, 3, K, R, R, stride, pad)
    assert ops._lib().srgan_conv2d_thin16_supported(d, 0) == 1
    y = torch.empty((N, K, d.P, d.Q), dtype=BF, device=DEV).contiguous(memory_format=CL)
    ws, nb = _ws(d, 0)
    for act, slope, fn in ((ops.ACT_NONE, 0.0, lambda t: t), (ops.ACT_LRELU, 0.2, lambda t: F.leaky_relu(t, 0.2))):
        ops._call("srgan_conv2d_fprop_thin16", d, ops._p(x), ops._p(w), ops._p(b), ops._p(y), act, slope, ops._p(ws), nb,
                  ops._stream())
        ref = fn(F.conv2d(x.double(), w.double(), b.double(), stride, pad))
        assert _rel(y, ref) < TOL, (g, act, _rel(y, ref))


HEADS = [(3, 128, 128, 64, 7, 1, 3), (64, 128, 128, 64, 7, 1, 3), (2, 40, 24, 64, 7, 1, 3), (2, 32, 32, 32, 3, 1, 1)]


@pytest.mark.parametrize("g", HEADS)
def test_dgrad_thin_output_bf16_dx(g):
    """input gradient of the RGB head: dy [N, 3, P, Q] fp32 -> dx [N, C, H, W] bf16"""
    N, H, W, C, R, stride, pad = g
    torch.manual_seed(1)
    d = ops._desc(N, H, W, C, 3, R, R, stride, pad)
    dy = torch.randn(N, 3, d.P, d.Q, device=DEV).contiguous(memory_format=CL)
    w = (torch.randn(3, C, R, R, device=DEV) * 0.1).contiguous(memory_format=CL)
    assert ops._lib().srgan_conv2d_thin16_supported(d, 1) == 1
    dx = torch.empty((N, C, H, W), dtype=BF, device=DEV).contiguous(memory_format=CL)
    ws, nb = _ws(d, 1)
    ops._call("srgan_conv2d_dgrad_thin16", d, ops._p(dy), ops._p(w), ops._p(dx), ops._p(ws), nb, ops._stream())
    ref = torch.nn.grad.conv2d_input((N, C, H, W), w.double(), dy.double(), stride, pad)
    assert _rel(dx, ref) < TOL, (g, _rel(dx, ref))


WGRADS = [  # N, H, W, fat channels, R, stride, pad
    (3, 128, 128, 64, 7, 1, 3),
    (64, 128, 128, 64, 7, 1, 3),     # production batch: 1024 pixel chunks over 74 splits x 2 tap groups
    (4, 128, 128, 64, 7, 2, 1),      # encoder first layer (stride-2 packed view)
    (5, 128, 128, 64, 4, 2, 1),      # discriminator stem: one tap group
    (2, 37, 53, 64, 7, 1, 3),        # ragged plane
]


@pytest.mark.parametrize("g", WGRADS)
def test_wgrad_thin_input_bf16_dy(g):
    """stem: x [N, 3, H, W] fp32, dy [N, K, P, Q] bf16 -> dw [K, 3, R, R], dbias [K] fp32"""
    N, H, W, K, R, stride, pad = g
    torch.manual_seed(2)
    d = ops._desc(N, H, W, 3, K, R, R, stride, pad)
    x = (torch.rand(N, 3, H, W, device=DEV) * 2 - 1).contiguous(memory_format=CL)
    dy = torch.randn(N, K, d.P, d.Q, device=DEV).to(BF).contiguous(memory_format=CL)
    assert ops._lib().srgan_conv2d_thin16_supported(d, 2) == 1
    dw = torch.empty((K, 3, R, R), device=DEV).contiguous(memory_format=CL)
    db = torch.empty((K,), device=DEV)
    ws, nb = _ws(d, 2)
    ops._call("srgan_conv2d_wgrad_thin16", d, ops._p(x), ops._p(dy), ops._p(dw), ops._p(db), ops._p(ws), nb, ops._stream())
    ref = torch.nn.grad.conv2d_weight(x.double(), (K, 3, R, R), dy.double(), stride, pad)
    assert _rel(dw, ref) < TOL, (g, _rel(dw, ref))
    assert _rel(db, dy.double().sum(dim=(0, 2, 3))) < 1e-5
    dw2 = torch.empty_like(dw)
    ops._call("srgan_conv2d_wgrad_thin16", d, ops._p(x), ops._p(dy), ops._p(dw2), None, ops._p(ws), nb, ops._stream())
    assert torch.equal(dw, dw2)                                       # fixed-order reduction


@pytest.mark.parametrize("g", [(3, 128, 128, 64, 7, 1, 3), (64, 128, 128, 64, 7, 1, 3), (2, 40, 24, 64, 7, 1, 3),
                               (2, 32, 32, 128, 3, 1, 1)])
def test_wgrad_thin_output_bf16_x(g):
    """head: x [N, C, H, W] bf16, dy [N, 3, P, Q] fp32 -> dw [3, C, R, R], dbias [3] fp32"""
    N, H, W, C, R, stride, pad = g
    torch.manual_seed(3)
    d = ops._desc(N, H, W, C, 3, R, R, stride, pad)
    x = torch.randn(N, C, H, W, device=DEV).to(BF).contiguous(memory_format=CL)
    dy = torch.randn(N, 3, d.P, d.Q, device=DEV).contiguous(memory_format=CL)
    assert ops._lib().srgan_conv2d_thin16_supported(d, 2) == 1
    dw = torch.empty((3, C, R, R), device=DEV).contiguous(memory_format=CL)
    db = torch.empty((3,), device=DEV)
    ws, nb = _ws(d, 2)
    ops._call("srgan_conv2d_wgrad_thin16", d, ops._p(x), ops._p(dy), ops._p(dw), ops._p(db), ops._p(ws), nb, ops._stream())
    ref = torch.nn.grad.conv2d_weight(x.double(), (3, C, R, R), dy.double(), stride, pad)
    assert _rel(dw, ref) < TOL, (g, _rel(dw, ref))
    assert _rel(db, dy.double().sum(dim=(0, 2, 3))) < 1e-5


@pytest.mark.parametrize("g", [(3, 128, 128, 7, 3), (64, 128, 128, 7, 3), (2, 40, 24, 7, 3), (2, 32, 32, 3, 1)])
def test_fprop_thin_output_bf16_input(g):
    """RGB head: x [N, 64, H, W] bf16 -> y [N, 3, H, W] fp32 with bias and tanh (row GEMM + col2im kernel, kind::f16)"""
    N, H, W, R, pad = g
    torch.manual_seed(4)
    d = ops._desc(N, H, W, 64, 3, R, R, 1, pad)
    x = torch.randn(N, 64, H, W, device=DEV).to(BF).contiguous(memory_format=CL)
    w = (torch.randn(3, 64, R, R, device=DEV) * 0.05).contiguous(memory_format=CL)
    b = torch.randn(3, device=DEV) * 0.1
    assert ops._lib().srgan_conv2d_thin16_supported(d, 0) == 1
    y = torch.empty((N, 3, d.P, d.Q), device=DEV).contiguous(memory_format=CL)
    ws, nb = _ws(d, 0)
    for act, fn in ((ops.ACT_NONE, lambda t: t), (ops.ACT_TANH, torch.tanh)):
        ops._call("srgan_conv2d_fprop_thin16", d, ops._p(x), ops._p(w), ops._p(b), ops._p(y), act, 0.0, ops._p(ws), nb,
                  ops._stream())
        ref = fn(F.conv2d(x.double(), w.double(), b.double(), 1, pad))
        assert _rel(y, ref) < TOL, (g, act, _rel(y, ref))


@pytest.mark.parametrize("g", [(3, 128, 128, 7, 3), (64, 128, 128, 7, 3), (2, 40, 24, 7, 3)])
def test_dgrad_thin_input_bf16_dy(g):
    """input gradient of the RGB stem: dy [N, 64, P, Q] bf16 -> dx [N, 3, H, W] fp32"""
    N, H, W, R, pad = g
    torch.manual_seed(5)
    d = ops._desc(N, H, W, 3, 64, R, R, 1, pad)
    dy = torch.randn(N, 64, d.P, d.Q, device=DEV).to(BF).contiguous(memory_format=CL)
    w = (torch.randn(64, 3, R, R, device=DEV) * 0.05).contiguous(memory_format=CL)
    assert ops._lib().srgan_conv2d_thin16_supported(d, 1) == 1
    dx = torch.empty((N, 3, H, W), device=DEV).contiguous(memory_format=CL)
    ws, nb = _ws(d, 1)
    ops._call("srgan_conv2d_dgrad_thin16", d, ops._p(dy), ops._p(w), ops._p(dx), ops._p(ws), nb, ops._stream())
    ref = torch.nn.grad.conv2d_input((N, 3, H, W), w.double(), dy.double(), 1, pad)
    assert _rel(dx, ref) < TOL, (g, _rel(dx, ref))


def test_thin16_rates():
    """Not pass / fail: the six thin16 launches of the generator's RGB stem and head at batch 64 next to their fp32 forms."""
    N, H = 64, 128
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

    def timed(fn, n=7):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(e) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]
    img = (torch.rand(N, 3, H, H, device=DEV) * 2 - 1).contiguous(memory_format=CL)
    fat32 = torch.randn(N, 64, H, H, device=DEV).contiguous(memory_format=CL)
    fat16 = fat32.to(BF).contiguous(memory_format=CL)
    ws_ = torch.randn(64, 3, 7, 7, device=DEV).contiguous(memory_format=CL)
    wh_ = torch.randn(3, 64, 7, 7, device=DEV).contiguous(memory_format=CL)
    ds, dh = ops._desc(N, H, H, 3, 64, 7, 7, 1, 3), ops._desc(N, H, H, 64, 3, 7, 7, 1, 3)
    lib = ops._lib()
    dev = torch.device(DEV, torch.cuda.current_device())
    out16 = torch.empty_like(fat16)
    out32 = torch.empty_like(fat32)
    thin = torch.empty_like(img)
    dws, dwh = torch.empty_like(ws_), torch.empty_like(wh_)
    rows = []
    for name, d, p, call16, call32 in (
        ("stem fprop", ds, 0,
         lambda ws, nb: ops._call("srgan_conv2d_fprop_thin16", ds, ops._p(img), ops._p(ws_), None, ops._p(out16), 0, 0.0, ops._p(ws), nb, ops._stream()),
         lambda ws, nb: ops._call("srgan_conv2d_fprop", ds, ops._p(img), ops._p(ws_), None, ops._p(out32), 0, 0.0, 0, ops._p(ws), nb, ops._stream())),
        ("stem dgrad", ds, 1,
         lambda ws, nb: ops._call("srgan_conv2d_dgrad_thin16", ds, ops._p(fat16), ops._p(ws_), ops._p(thin), ops._p(ws), nb, ops._stream()),
         lambda ws, nb: ops._call("srgan_conv2d_dgrad", ds, ops._p(fat32), ops._p(ws_), ops._p(thin), 0, ops._p(ws), nb, ops._stream())),
        ("stem wgrad", ds, 2,
         lambda ws, nb: ops._call("srgan_conv2d_wgrad_thin16", ds, ops._p(img), ops._p(fat16), ops._p(dws), None, ops._p(ws), nb, ops._stream()),
         lambda ws, nb: ops._call("srgan_conv2d_wgrad", ds, ops._p(img), ops._p(fat32), ops._p(dws), None, 0, ops._p(ws), nb, ops._stream())),
        ("head fprop", dh, 0,
         lambda ws, nb: ops._call("srgan_conv2d_fprop_thin16", dh, ops._p(fat16), ops._p(wh_), None, ops._p(thin), 0, 0.0, ops._p(ws), nb, ops._stream()),
         lambda ws, nb: ops._call("srgan_conv2d_fprop", dh, ops._p(fat32), ops._p(wh_), None, ops._p(thin), 0, 0.0, 0, ops._p(ws), nb, ops._stream())),
        ("head dgrad", dh, 1,
         lambda ws, nb: ops._call("srgan_conv2d_dgrad_thin16", dh, ops._p(img), ops._p(wh_), ops._p(out16), ops._p(ws), nb, ops._stream()),
         lambda ws, nb: ops._call("srgan_conv2d_dgrad", dh, ops._p(img), ops._p(wh_), ops._p(out32), 0, ops._p(ws), nb, ops._stream())),
        ("head wgrad", dh, 2,
         lambda ws, nb: ops._call("srgan_conv2d_wgrad_thin16", dh, ops._p(fat16), ops._p(img), ops._p(dwh), None, ops._p(ws), nb, ops._stream()),
         lambda ws, nb: ops._call("srgan_conv2d_wgrad", dh, ops._p(fat32), ops._p(img), ops._p(dwh), None, 0, ops._p(ws), nb, ops._stream())),
    ):
        nb16 = lib.srgan_conv2d_thin16_workspace(d, p)
        nb32 = lib.srgan_conv2d_workspace(d, p, 0)
        w16, w32 = ops._workspace(dev, max(nb16, nb32)), None
        t16 = timed(lambda: call16(w16, max(nb16, nb32)))
        t32 = timed(lambda: call32(w16, max(nb16, nb32)))
        rows.append("%s: thin16 %.0f us, fp32 %.0f us" % (name, t16, t32))
    print("RGB layers at batch 64 (L2 flushed): " + "; ".join(rows))
