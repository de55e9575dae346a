"""GPU: the one-pass instance norm (norm8c.cu: a thread-block cluster holds the image in shared memory, x read once
forward) behind srgan_inorm_{fwd,bwd}_mixed, through the C ABI:
  * against fp64 on the same (rounded) inputs, with the tolerances of the two-kernel path (test_conv_bf16_gpu.py);
  * against the two-kernel path of norm8.cu on the same inputs: the atoms of the summation are the same, so mean / rstd
    and the backward sums agree to the last bit up to an fp64 reassociation (1 ulp allowed), outputs within one
    rounding of the storage type;
  * batch-split invariance (an image normalised alone or with others: identical bits), which N-GPU = 1-GPU relies on;
  * cluster sizes 1 .. 8 including odd ones, ragged last chunks, every storage pair, ReLU / LeakyReLU / residual.
ref: CBINorm2d.forward pyfiles/model.py:54-67, nn.InstanceNorm2d(affine=False) :178."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

import srgan_ops as ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
CL = torch.channels_last
BF, F32 = torch.bfloat16, torch.float32


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _plan(hw, c, dtype):
    cl, sl = ctypes.c_int(0), ctypes.c_int(0)
    before = ops._lib().srgan_inorm_onepass_enable(1)        # the plan reports 0 while the path is switched off
    try:
        ok = ops._lib().srgan_inorm_onepass_plan(hw, c, ops._dt(torch.empty(0, dtype=dtype)), ctypes.byref(cl),
                                                 ctypes.byref(sl))
    finally:
        ops._lib().srgan_inorm_onepass_enable(before)
    return ok, cl.value, sl.value


class _onepass:
    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.before = ops._lib().srgan_inorm_onepass_enable(int(self.on))

    def __exit__(self, *a):
        ops._lib().srgan_inorm_onepass_enable(self.before)


def _run(x, g, b, cb, res, gy, act, slope, yd):
    xs = x.clone().contiguous(memory_format=CL).requires_grad_(True)
    y = ops.instance_norm_act(xs, g, b, cb, res, 1e-5, act, slope, out_dtype=yd)
    ins = [xs, g, b, cb] + ([res] if res is not None else [])
    grads = torch.autograd.grad(y, ins, gy)
    return y.detach(), grads


# N, C, H, W, x dtype, y dtype, expected forward cluster size (resident x), expected backward cluster size (resident dy)
PLANES = [
    (6, 256, 32, 32, BF, BF, 8, 8),       # the residual trunk: 8 CTAs x 128 pixels x 512 B
    (3, 256, 15, 15, BF, BF, 2, 2),       # encoder: 225 pixels, ragged last chunk
    (2, 128, 31, 31, F32, BF, 8, 4),      # encoder block entry (f32 -> bf16): 961 pixels
    (5, 128, 24, 24, BF, BF, 3, 3),       # odd cluster size
    (5, 128, 24, 24, BF, F32, 3, 5),
    (4, 64, 16, 16, BF, BF, 1, 1),        # one CTA per image
    (3, 192, 20, 20, BF, BF, 3, 3),       # 24 threads per pixel row, 16 idle threads
    (2, 64, 62, 62, BF, BF, 8, 8),        # encoder, first block: 3844 pixels, 31 chunks of 128
]


@pytest.mark.parametrize("pl", PLANES)
def test_onepass_plan_and_parity(pl):
    N, C, H, W, xd, yd, cl_f, cl_b = pl
    okf, clf, _ = _plan(H * W, C, xd)
    okb, clb, _ = _plan(H * W, C, yd)
    assert (okf, clf) == (1, cl_f) and (okb, clb) == (1, cl_b), (pl, clf, clb)
    torch.manual_seed(5)
    x = (torch.randn(N, C, H, W, device=DEV) * 1.3 + 0.7).to(xd).contiguous(memory_format=CL)
    g = torch.randn(C, device=DEV).requires_grad_(True)
    b = torch.randn(C, device=DEV).requires_grad_(True)
    cb = torch.randn(N, C, device=DEV).requires_grad_(True)
    res = torch.randn(N, C, H, W, device=DEV).to(yd).contiguous(memory_format=CL).requires_grad_(True)
    gy = torch.randn(N, C, H, W, device=DEV).to(yd).contiguous(memory_format=CL)
    for act, slope, residual in ((ops.ACT_RELU, 0.0, None), (ops.ACT_LRELU, 0.2, None), (ops.ACT_NONE, 0.0, res)):
        with _onepass(True):
            y1, g1 = _run(x, g, b, cb, residual, gy, act, slope, yd)
        with _onepass(False):
            y2, g2 = _run(x, g, b, cb, residual, gy, act, slope, yd)
        # fp64 reference on the same inputs
        xr, gr, br, cr = (t.detach().double().requires_grad_(True) for t in (x, g, b, cb))
        rr = res.detach().double().requires_grad_(True)
        v = (F.instance_norm(xr, eps=1e-5) + cr[:, :, None, None]) * gr[None, :, None, None] + br[None, :, None, None]
        yr = torch.relu(v) if act == ops.ACT_RELU else (F.leaky_relu(v, slope) if act == ops.ACT_LRELU else v + rr)
        ref = torch.autograd.grad(yr, [xr, gr, br, cr] + ([rr] if residual is not None else []), gy.double())
        tol_y = 4e-3 if yd == BF else 1e-5
        tol_x = 4e-3 if xd == BF else 1e-5
        assert _rel(y1, yr) < tol_y, (pl, act)
        assert _rel(g1[0], ref[0]) < tol_x, (pl, act)
        for a_, r_ in zip(g1[1:4], ref[1:4]):
            assert _rel(a_, r_) < 1e-4, (pl, act)
        if residual is not None:
            assert torch.equal(g1[4], gy)
        # the two-kernel path: same atoms
        assert g1[0].dtype == xd and y1.dtype == yd
        ulp_y = 2.0 ** -7 if yd == BF else 2.0 ** -21
        ulp_x = 2.0 ** -7 if xd == BF else 2.0 ** -21
        dy_ = (y1.float() - y2.float()).abs()
        assert float((dy_ / y2.float().abs().clamp_min(1.0)).max()) <= ulp_y, (pl, act)
        assert float((dy_ > 0).float().mean()) < 1e-3, (pl, act)
        dx_ = (g1[0].float() - g2[0].float()).abs()
        assert float((dx_ / g2[0].float().abs().clamp_min(1.0)).max()) <= 4 * ulp_x, (pl, act)
        for a_, b_ in zip(g1[1:4], g2[1:4]):
            assert _rel(a_, b_) < 1e-6, (pl, act)


def test_onepass_statistics_equal_two_kernel_path():
    """mean / rstd written by the one-pass kernel against the statistics kernel of norm8.cu, straight through the ABI."""
    lib = ops._lib()
    for (N, C, H, W, dt) in ((7, 256, 32, 32, BF), (3, 128, 31, 31, F32), (2, 64, 62, 62, BF), (3, 256, 15, 15, BF)):
        torch.manual_seed(6)
        x = (torch.randn(N, C, H, W, device=DEV) * 2 + 3).to(dt).contiguous(memory_format=CL)
        outs = []
        for on in (True, False):
            with _onepass(on):
                y = torch.empty_like(x, dtype=BF)
                mean = torch.empty((N, C), device=DEV)
                rstd = torch.empty((N, C), device=DEV)
                nb = lib.srgan_inorm_mixed_workspace(N, H * W, C)
                ws = ops._workspace(x.device, nb)
                ops._call("srgan_inorm_fwd_mixed", ops._p(x), ops._dt(x), ops._p(y), ops._dt(y), ops._p(mean),
                          ops._p(rstd), None, None, None, None, N, H * W, C, 1e-5, ops.ACT_NONE, 0.0, 0, ops._p(ws), nb,
                          ops._p(ops._norm_counters(x.device, N, C)), ops._stream())
                outs.append((mean.clone(), rstd.clone(), y.clone()))
        (m1, r1, y1), (m2, r2, y2) = outs
        ref = x.double().mean(dim=(2, 3))
        assert _rel(m1, ref) < 1e-6
        # identical atoms, fp64 above them: equal up to one fp32 ulp in (rare) reassociation cases
        assert float(((m1 - m2).abs() / m2.abs().clamp_min(1e-3)).max()) <= 2.0 ** -22
        assert float(((r1 - r2).abs() / r2.abs()).max()) <= 2.0 ** -22
        assert float((m1 != m2).float().mean()) < 1e-3 and float((r1 != r2).float().mean()) < 1e-3


def test_onepass_batch_split_invariance():
    """Image 0 normalised alone, with 3 and with 15 others: bit-identical y and dx (the cluster layout of an image does
    not depend on the batch)."""
    torch.manual_seed(11)
    for (c, h, w, xd, yd) in ((256, 32, 32, BF, BF), (128, 31, 31, F32, BF), (64, 62, 62, BF, BF)):
        assert _plan(h * w, c, xd)[0] == 1
        x = (torch.randn(16, c, h, w, device=DEV) * 2 + 3).to(xd).contiguous(memory_format=CL)
        dy = torch.randn(16, c, h, w, device=DEV).to(yd).contiguous(memory_format=CL)
        g = torch.randn(c, device=DEV)
        b = torch.randn(c, device=DEV)
        outs = []
        for n in (16, 4, 1):
            xs = x[:n].clone().contiguous(memory_format=CL).requires_grad_(True)
            with _onepass(True):
                y = ops.instance_norm_act(xs, g, b, None, None, 1e-5, ops.ACT_LRELU, 0.2, out_dtype=yd)
                y.backward(dy[:n].clone().contiguous(memory_format=CL))
            outs.append((y.detach()[:1].clone(), xs.grad[:1].clone()))
        for y, dx in outs[1:]:
            assert torch.equal(y, outs[0][0]), (c, h, w)
            assert torch.equal(dx, outs[0][1]), (c, h, w)


def test_onepass_not_eligible_falls_back():
    """Planes a cluster cannot hold (or channel counts the mapping cannot serve) report 0 and run the two-kernel path."""
    assert _plan(128 * 128, 64, BF)[0] == 0          # 2 MB per image
    assert _plan(64 * 64, 128, BF)[0] == 0           # 1 MB per image
    assert _plan(49, 512, BF)[0] == 0                # C > 256
    torch.manual_seed(2)
    x = torch.randn(2, 512, 7, 7, device=DEV).to(BF).contiguous(memory_format=CL)
    with _onepass(True):
        y = ops.instance_norm_act(x, None, None, None, None, 1e-5, ops.ACT_RELU, 0.0, out_dtype=BF)
    assert _rel(y, torch.relu(F.instance_norm(x.double(), eps=1e-5))) < 4e-3


def test_onepass_rate_production_plane():
    """Not pass / fail: the 256-channel 32x32 bf16 plane at batch 64, one-pass against two-kernel, L2 flushed."""
    N, C, H = 64, 256, 32
    x = torch.randn(N, C, H, H, device=DEV).to(BF).contiguous(memory_format=CL)
    dy = torch.randn(N, C, H, H, device=DEV).to(BF).contiguous(memory_format=CL)
    g = torch.randn(C, device=DEV)
    b = torch.randn(C, device=DEV)
    cb = torch.randn(N, C, device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    lib = ops._lib()
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    mean = torch.empty((N, C), device=DEV)
    rstd = torch.empty((N, C), device=DEV)
    s1 = torch.empty((N, C), device=DEV)
    s2 = torch.empty((N, C), device=DEV)
    nb = lib.srgan_inorm_mixed_workspace(N, H * H, C)
    ws = ops._workspace(x.device, nb)
    ctr = ops._norm_counters(x.device, N, C)

    def fwd():
        ops._call("srgan_inorm_fwd_mixed", ops._p(x), ops._dt(x), ops._p(y), ops._dt(y), ops._p(mean), ops._p(rstd),
                  ops._p(g), ops._p(b), ops._p(cb), None, N, H * H, C, 1e-5, ops.ACT_RELU, 0.0, 0, ops._p(ws), nb,
                  ops._p(ctr), ops._stream())

    def bwd():
        ops._call("srgan_inorm_bwd_mixed", ops._p(dy), ops._dt(dy), ops._p(x), ops._dt(x), ops._p(mean), ops._p(rstd),
                  ops._p(g), ops._p(b), ops._p(cb), ops._p(dx), ops._p(s1), ops._p(s2), N, H * H, C, ops.ACT_RELU, 0.0,
                  ops._p(ws), nb, ops._p(ctr), ops._stream())

    def timed(fn, n=20, cold=True):
        ts = []
        for _ in range(n):
            if cold:
                flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        ts.sort()
        return ts[len(ts) // 2] * 1e3
    mb = x.numel() * 2 / 1e6
    for on in (True, False):
        with _onepass(on):
            fwd(); bwd(); torch.cuda.synchronize()
            tf, tb = timed(fwd), timed(bwd)
            tfw, tbw = timed(fwd, cold=False), timed(bwd, cold=False)
            print("norm 256ch@32x32 bf16 batch 64, %s: fwd %.1f us (%.2f TB/s algorithmic), bwd %.1f us (%.2f TB/s); "
                  "warm L2: fwd %.1f us, bwd %.1f us"
                  % ("one pass" if on else "two kernels", tf, 2 * mb / tf, tb, 3 * mb / tb, tfw, tbw))
