"""GPU: the convolution engine at the PRODUCTION shape of bench.py (batch 64, full-width layers) against fp64
PyTorch, through the C ABI.  The small-batch cases of test_conv_umma_gpu.py never reach three code paths that only
exist at this size:
  * the persistent multi-wave loop of conv_umma_kernel (512 tiles on 148 CTAs, accumulator double buffering across
    work items),
  * the half-width tail items that balance the last partial wave (conv_umma.cu launch_bn),
  * the cost-model pixel split of wgrad (16 splits for the residual blocks) and its fixed-order reduction.
Reference: F.conv2d / F.conv_transpose2d in fp64 on the same device (ref layers: pyfiles/model.py:188-249 generator,
:385,445 encoder stems, :504-563 encoder blocks).  Tolerance: TF32 operands (11 significant bits each), fp32
accumulation: relative L2 error < 2e-3, stated per assert; the measured values are printed."""
import pytest
import torch
import torch.nn.functional as F

import srgan_ops as ops

pytestmark = pytest.mark.gpu
CL = torch.channels_last
TOL = 2e-3
N = 64


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g, device="cuda") * scale


@pytest.fixture(autouse=True)
def _engine():
    ops.set_conv_engine("auto")
    yield
    ops.set_conv_engine("auto")


# name, C, H, K, R, stride, pad, bias, act
CONVS = [
    ("G.res", 256, 32, 256, 3, 1, 1, False, ops.ACT_NONE),          # 512 tiles: 3.46 waves, tail items, 16-split wgrad
    ("G.down1", 64, 128, 128, 4, 2, 1, False, ops.ACT_NONE),        # MT=2 tiles; stride-2 dgrad in 4 parity classes
    ("G.down2", 128, 64, 256, 4, 2, 1, False, ops.ACT_NONE),
    ("E.l3.cmp", 512, 9, 1024, 3, 1, 0, False, ops.ACT_NONE),       # 7x7 output planes, K = 1024 in 4 filter tiles
    ("E.l0.conv1", 64, 64, 64, 3, 1, 0, False, ops.ACT_LRELU),      # <64,2> tiles, fused LeakyReLU
    ("D1.1", 64, 64, 128, 4, 2, 1, True, ops.ACT_LRELU),
    ("G.stem", 3, 128, 64, 7, 1, 3, False, ops.ACT_NONE),           # row-packed thin input
    ("E.first", 3, 128, 64, 7, 2, 1, True, ops.ACT_NONE),           # thin input, stride 2
    ("G.head", 64, 128, 3, 7, 1, 3, True, ops.ACT_TANH),            # thin output: row GEMM + col2im
]


@pytest.mark.parametrize("geom", CONVS, ids=[g[0] for g in CONVS])
def test_production_batch_conv_fprop_dgrad_wgrad(geom):
    name, C, H, K, R, stride, pad, bias, act = geom
    x = _rand(N, C, H, H, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    w = _rand(K, C, R, R, seed=2, scale=(C * R * R) ** -0.5).contiguous(memory_format=CL).requires_grad_(True)
    b = _rand(K, seed=3).requires_grad_(True) if bias else None
    d = ops._desc(N, H, H, C, K, R, R, stride, pad)
    lib = ops._lib()
    for p in (0, 1, 2):
        assert lib.srgan_conv2d_engine(d, p) == ops.ENGINE_TF32, "pass %d should run on the tcgen05 engine" % p
    y = ops.conv2d(x, w, b, stride, pad, "zeros", act, 0.2)
    gy = _rand(*y.shape, seed=4).contiguous(memory_format=CL)
    dx, dw = torch.autograd.grad(y, [x, w], gy, retain_graph=bias)
    db = torch.autograd.grad(y, b, gy)[0] if bias else None

    xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    br = b.detach().double().requires_grad_(True) if bias else None
    zr = F.conv2d(xr, wr, br, stride, pad)
    # The backward of a fused activation is evaluated from the activation OUTPUT the product saved (autograd
    # semantics).  A pre-activation within TF32 rounding of zero may land on the other side of the LeakyReLU kink than
    # in fp64 (measured: 2e-4 of the elements, i.e. 1e-2 relative L2 on dx if the masks were compared too), so the
    # reference back-propagates the PRODUCT's mask: what is checked is dgrad / wgrad, not the sign of rounding noise.
    if act == ops.ACT_LRELU:
        yr = F.leaky_relu(zr, 0.2)
        dz = gy.double() * torch.where(y.detach() > 0, 1.0, 0.2).double()
    elif act == ops.ACT_TANH:
        yr = torch.tanh(zr)
        dz = gy.double() * (1.0 - y.detach().double() ** 2)
    else:
        yr, dz = zr, gy.double()
    grads = torch.autograd.grad(zr, [xr, wr] + ([br] if bias else []), dz)
    e = dict(y=_rel(y, yr), dx=_rel(dx, grads[0]), dw=_rel(dw, grads[1]))
    if bias:
        e["db"] = _rel(db, grads[2])
    print("%s N=%d rel-L2 vs fp64: %s" % (name, N, {k: "%.2e" % v for k, v in e.items()}))
    for k, v in e.items():
        assert v < TOL, (name, k, v)
    if name == "G.res":
        splits, ctas = ops.wgrad_plan(d)
        assert splits > 1, "the production residual wgrad is expected to run split over pixels"


# name, Cin, H, Cout : ConvTranspose2d(4, stride 2, pad 1) of the generator's up path (ref pyfiles/model.py:227,230)
UPS = [("G.up0", 256, 32, 128), ("G.up1", 128, 64, 64)]


@pytest.mark.parametrize("geom", UPS, ids=[g[0] for g in UPS])
def test_production_batch_conv_transpose(geom):
    name, Cin, H, Cout = geom
    x = _rand(N, Cin, H, H, seed=8).contiguous(memory_format=CL).requires_grad_(True)
    w = _rand(Cin, Cout, 4, 4, seed=9, scale=(Cin * 4) ** -0.5).contiguous(memory_format=CL).requires_grad_(True)
    y = ops.conv_transpose2d(x, w, 2, 1)
    gy = _rand(*y.shape, seed=5).contiguous(memory_format=CL)
    dx, dw = torch.autograd.grad(y, [x, w], gy)
    xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yr = F.conv_transpose2d(xr, wr, None, 2, 1)
    dxr, dwr = torch.autograd.grad(yr, [xr, wr], gy.double())
    e = dict(y=_rel(y, yr), dx=_rel(dx, dxr), dw=_rel(dw, dwr))
    print("%s N=%d rel-L2 vs fp64: %s" % (name, N, {k: "%.2e" % v for k, v in e.items()}))
    for k, v in e.items():
        assert v < TOL, (name, k, v)


def test_production_batch_residual_block_skip_dgrad():
    """The dgrad of the residual blocks' first convolution with the skip gradient added in the epilogue, at N = 64
    (persistent loop + tail items + addend loads)."""
    x = _rand(N, 256, 32, 32, seed=41).contiguous(memory_format=CL).requires_grad_(True)
    w = _rand(256, 256, 3, 3, seed=42, scale=2304 ** -0.5).contiguous(memory_format=CL).requires_grad_(True)
    g1 = _rand(N, 256, 32, 32, seed=43).contiguous(memory_format=CL)
    g2 = _rand(N, 256, 32, 32, seed=44).contiguous(memory_format=CL)
    y, skip = ops.conv2d_skip(x, w, 1, 1)
    dx, dw = torch.autograd.grad([y, skip], [x, w], [g1, g2])
    xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, 1, 1)
    dxr, dwr = torch.autograd.grad([yr, xr * 1.0], [xr, wr], [g1.double(), g2.double()])
    e = dict(y=_rel(y, yr), dx=_rel(dx, dxr), dw=_rel(dw, dwr))
    print("G.res skip N=%d rel-L2 vs fp64: %s" % (N, {k: "%.2e" % v for k, v in e.items()}))
    for k, v in e.items():
        assert v < TOL, (k, v)


def test_production_batch_instance_norm():
    """CBINorm + ReLU of the residual trunk at batch 64 (one-wave grid, fp64 partial sums) against fp64."""
    C, H = 256, 32
    x = (_rand(N, C, H, H, seed=51) * 1.7 + 0.4).contiguous(memory_format=CL).requires_grad_(True)
    g = _rand(C, seed=52).requires_grad_(True)
    b = _rand(C, seed=53).requires_grad_(True)
    cb = _rand(N, C, seed=54).requires_grad_(True)
    y = ops.instance_norm_act(x, g, b, cb, None, 1e-5, ops.ACT_RELU, 0.0)
    gy = _rand(*y.shape, seed=55).contiguous(memory_format=CL)
    got = torch.autograd.grad(y, [x, g, b, cb], gy)
    xr, gr, br, cr = (t.detach().double().requires_grad_(True) for t in (x, g, b, cb))
    yr = torch.relu((F.instance_norm(xr, eps=1e-5) + cr[:, :, None, None]) * gr[None, :, None, None]
                    + br[None, :, None, None])
    ref = torch.autograd.grad(yr, [xr, gr, br, cr], gy.double())
    e = dict(y=_rel(y, yr), dx=_rel(got[0], ref[0]), dgamma=_rel(got[1], ref[1]), dbeta=_rel(got[2], ref[2]),
             dcbias=_rel(got[3], ref[3]))
    print("IN 256@32 N=%d rel-L2 vs fp64: %s" % (N, {k: "%.2e" % v for k, v in e.items()}))
    for k, v in e.items():
        assert v < 1e-4, (k, v)          # fp32 arithmetic, fp64 partial sums
