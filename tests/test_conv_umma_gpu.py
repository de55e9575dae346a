"""GPU: the tcgen05/TMEM/TMA convolution engine (TF32 operands, fp32 accumulate) against fp64 PyTorch and
against the exact-fp32 FFMA engine, through the C ABI.  Tolerance: TF32 rounds both operands to 11 significant
bits (rel 4.9e-4 each); for the reduction lengths here the result differs from fp64 by < 2e-3 relative L2."""
import pytest
import torch
import torch.nn.functional as F

import srgan_ops as ops

pytestmark = pytest.mark.gpu
CL = torch.channels_last
TOL = 2e-3


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.fixture(autouse=True)
def _engine():
    ops.set_conv_engine("auto")       # tcgen05 where the shape/pass qualifies, FFMA otherwise
    yield
    ops.set_conv_engine("auto")


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).cuda()


# name, N, C, H, W, K, R, stride, pad, bias
GEOMS = [
    ("G.down1", 2, 64, 128, 128, 128, 4, 2, 1, False),
    ("G.down2", 2, 128, 64, 64, 256, 4, 2, 1, False),
    ("G.res", 3, 256, 32, 32, 256, 3, 1, 1, False),
    ("G.out", 2, 64, 128, 128, 3, 7, 1, 3, False),
    ("E.conv1.62", 2, 64, 64, 64, 64, 3, 1, 0, False),        # reflect-padded input, valid conv
    ("E.cmp.31", 2, 128, 33, 33, 256, 3, 1, 0, False),
    ("E.cmp.15", 2, 256, 17, 17, 512, 3, 1, 0, False),
    ("E.cmp.7", 5, 512, 9, 9, 1024, 3, 1, 0, False),
    ("E.short", 2, 64, 31, 31, 128, 1, 1, 0, True),
    ("E.short.3", 2, 512, 3, 3, 1024, 1, 1, 0, True),
    ("D.c1", 2, 64, 64, 64, 128, 4, 2, 1, False),
    ("D.c3", 2, 256, 16, 16, 512, 4, 2, 1, False),
    ("D2.c3", 2, 128, 8, 8, 256, 4, 2, 1, False),
    ("D.patch", 2, 512, 8, 8, 1, 4, 1, 1, True),
    ("D.class", 2, 512, 8, 8, 4, 8, 1, 0, True),
    # thin (<= 4 channel) tensors: row-packed im2col through an overlapping-stride TMA map
    ("G.stem", 2, 3, 128, 128, 64, 7, 1, 3, False),
    ("E.stem", 2, 3, 128, 128, 64, 7, 2, 1, True),
    ("D.stem", 3, 3, 128, 128, 64, 4, 2, 1, False),
    ("D2.stem", 2, 3, 64, 64, 32, 4, 2, 1, False),
    ("thin.odd", 3, 2, 37, 29, 32, 5, 1, 2, True),
    ("thin.odd.s2", 2, 4, 30, 22, 40, 3, 2, 1, False),
    ("thin.out.odd", 2, 32, 19, 27, 2, 5, 1, 2, True),
]


@pytest.mark.parametrize("geom", GEOMS, ids=[g[0] for g in GEOMS])
def test_fprop_and_dgrad_tcgen05(geom):
    _, N, C, H, W, K, R, stride, pad, bias = geom
    x = _rand(N, C, H, W, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    w = _rand(K, C, R, R, seed=2, scale=(C * R * R) ** -0.5).contiguous(memory_format=CL).requires_grad_(True)
    b = _rand(K, seed=3).requires_grad_(True) if bias else None
    d = ops._desc(N, H, W, C, K, R, R, stride, pad)
    lib = ops._lib()
    assert lib.srgan_conv2d_engine(d, 0) == ops.ENGINE_TF32, "shape should qualify for the tcgen05 engine"
    y = ops.conv2d(x, w, b, stride, pad)
    xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yr = F.conv2d(xr, wr, None if b is None else b.detach().double(), stride, pad)
    assert _rel(y, yr) < TOL
    gy = _rand(*yr.shape, seed=4).contiguous(memory_format=CL)
    y.backward(gy)
    yr.backward(gy.double())
    if lib.srgan_conv2d_engine(d, 1) == ops.ENGINE_TF32:
        assert _rel(x.grad, xr.grad) < TOL
    else:
        assert _rel(x.grad, xr.grad) < 3e-5
    assert _rel(w.grad, wr.grad) < TOL
    # engine-vs-engine on the device
    ops.set_conv_engine("fp32")
    y32 = ops.conv2d(x.detach(), w.detach(), None if b is None else b.detach(), stride, pad)
    ops.set_conv_engine("auto")
    assert _rel(y, y32) < TOL


@pytest.mark.parametrize("shape", [(2, 256, 32, 32, 128), (2, 128, 64, 64, 64)])
def test_conv_transpose_tcgen05(shape):
    N, Cin, H, W, Cout = shape
    x = _rand(N, Cin, H, W, seed=8).contiguous(memory_format=CL).requires_grad_(True)
    w = _rand(Cin, Cout, 4, 4, seed=9, scale=(Cin * 4) ** -0.5).contiguous(memory_format=CL).requires_grad_(True)
    y = ops.conv_transpose2d(x, w, 2, 1)
    xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yr = F.conv_transpose2d(xr, wr, None, 2, 1)
    assert _rel(y, yr) < TOL
    gy = _rand(*yr.shape, seed=5).contiguous(memory_format=CL)
    y.backward(gy)
    yr.backward(gy.double())
    assert _rel(x.grad, xr.grad) < TOL and _rel(w.grad, wr.grad) < TOL


def test_fused_epilogue_tcgen05():
    x = _rand(2, 64, 32, 32, seed=11).contiguous(memory_format=CL)
    w = _rand(128, 64, 4, 4, seed=12, scale=0.03).contiguous(memory_format=CL)
    b = _rand(128, seed=13)
    y = ops.conv2d(x, w, b, 2, 1, "zeros", ops.ACT_LRELU, 0.01)
    ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), 2, 1), 0.01)
    assert _rel(y, ref) < TOL


def test_thin_output_row_gemm_tanh_and_bands():
    """RGB head (64 -> 3, 7x7) through the row-GEMM + col2im kernel: fused tanh, bias, several row bands,
    width below the 128-pixel tile, batch large enough for more than one wave."""
    for (N, C, H, W, K, R, pad) in ((5, 64, 40, 128, 3, 7, 3), (3, 32, 23, 50, 4, 3, 1), (160, 64, 8, 16, 1, 7, 3)):
        x = _rand(N, C, H, W, seed=21).contiguous(memory_format=CL)
        w = _rand(K, C, R, R, seed=22, scale=(C * R * R) ** -0.5).contiguous(memory_format=CL)
        b = _rand(K, seed=23)
        y = ops.conv2d(x, w, b, 1, pad, "zeros", ops.ACT_TANH, 0.0)
        ref = torch.tanh(F.conv2d(x.double(), w.double(), b.double(), 1, pad))
        assert _rel(y, ref) < TOL, (N, C, H, W, K, R, pad)


def test_residual_skip_gradient_fused_into_dgrad_epilogue():
    """conv2d_skip returns (conv(x), x) through one autograd node: d x = dgrad(d conv) + d skip with the add in the
    dgrad epilogue (srgan_conv2d_dgrad_add), also for shapes where the fused form is not available (FFMA engine,
    odd channel counts) - same values as the two separate ops."""
    for (N, C, H, W, K, engine) in ((3, 256, 32, 32, 256, "auto"), (2, 64, 16, 16, 64, "auto"), (70, 32, 8, 8, 32, "auto"),
                                    (2, 64, 16, 16, 64, "fp32")):
        ops.set_conv_engine(engine)
        try:
            x = _rand(N, C, H, W, seed=41).contiguous(memory_format=CL).requires_grad_(True)
            w = _rand(K, C, 3, 3, seed=42, scale=(C * 9) ** -0.5).contiguous(memory_format=CL).requires_grad_(True)
            g1, g2 = _rand(N, K, H, W, seed=43).contiguous(memory_format=CL), _rand(N, C, H, W, seed=44).contiguous(memory_format=CL)
            y, skip = ops.conv2d_skip(x, w, 1, 1)
            dx, dw = torch.autograd.grad([y, skip], [x, w], [g1, g2])
            xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
            yr = F.conv2d(xr, wr, None, 1, 1)
            dxr, dwr = torch.autograd.grad([yr, xr * 1.0], [xr, wr], [g1.double(), g2.double()])
            tol = TOL if engine == "auto" else 3e-5
            assert _rel(y, yr) < tol and _rel(dx, dxr) < tol and _rel(dw, dwr) < tol, (N, C, engine, _rel(dx, dxr))
            # skip gradient absent (the skip output unused): plain dgrad
            y2, _ = ops.conv2d_skip(x, w, 1, 1)
            dx2, = torch.autograd.grad(y2, x, g1)
            dx2r, = torch.autograd.grad(F.conv2d(xr, wr, None, 1, 1), xr, g1.double())
            assert _rel(dx2, dx2r) < tol
        finally:
            ops.set_conv_engine("auto")
