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
        ops._call("srgan_conv2d_fprop_bf16", d, ops._p(x), ops._p(w), ops._p(b), ops._p(y), act, slope, None,
                  ops._stream())
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
    ops._call("srgan_conv2d_dgrad_bf16", d, ops._p(dy), ops._p(w), None, ops._p(dx), None, ops._p(ws), nb, ops._stream())
    ref = torch.nn.grad.conv2d_input((N, C, H, W), w.float(), dy.float(), stride, pad)
    assert _rel(dx.float(), ref) < TOL, g
    if stride == 1:
        add = _nhwc_bf16(torch.randn(N, C, H, W))
        ops._call("srgan_conv2d_dgrad_bf16", d, ops._p(dy), ops._p(w), ops._p(add), ops._p(dx), None, ops._p(ws), nb,
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
    t16 = timed(lambda: ops._call("srgan_conv2d_fprop_bf16", d, ops._p(x), ops._p(w), None, ops._p(y), 0, 0.0, None,
                                  ops._stream()))
    rows = ops._lib().srgan_conv2d_bf16_stat_rows(d, 0)
    assert rows == 8
    ts = torch.empty((N, rows, K, 2), device=DEV)
    t16s = timed(lambda: ops._call("srgan_conv2d_fprop_bf16", d, ops._p(x), ops._p(w), None, ops._p(y), 0, 0.0,
                                   ops._p(ts), ops._stream()))
    print("res conv fprop with tile statistics in the epilogue: %.1f us" % (t16s * 1e3))
    t32 = timed(lambda: ops.conv2d(xf, wf, None, 1, 1))
    fl = 2.0 * N * H * W * K * C * 9
    print("res conv fprop: bf16 %.1f us (%.0f TF/s)   tf32 %.1f us (%.0f TF/s)" %
          (t16 * 1e3, fl / t16 / 1e9, t32 * 1e3, fl / t32 / 1e9))


WGRAD_GEOMS = [  # N, C, H, W, K, R, stride, pad
    (4, 256, 32, 32, 256, 3, 1, 1),     # residual block: KT = 2, 256-wide channel tile, 9 tap groups
    (3, 64, 64, 64, 128, 4, 2, 1),      # down1: 4 taps x 64 channels per MMA, stride-2 parity view of x
    (2, 128, 32, 32, 256, 4, 2, 1),     # down2: 2 taps x 128 channels
    (2, 64, 40, 24, 128, 4, 2, 1),      # ragged plane (pixel chunks cross image rows), mirrored up1
    (5, 128, 9, 13, 64, 3, 1, 1),       # odd plane, 64 filters (half an accumulator tile)
    (64, 256, 32, 32, 256, 3, 1, 1),    # production batch: cost-model pixel split + fixed-order reduction
]


@pytest.mark.parametrize("g", WGRAD_GEOMS)
def test_wgrad_bf16(g):
    """dW from bf16 x / dy (MN-major operands, plain 128B swizzle, 64-pixel chunks) against torch on the same
    bf16-rounded operands: products exact in fp32, so only the accumulation order differs (rel-L2 <= 1e-4)."""
    N, C, H, W, K, R, stride, pad = g
    torch.manual_seed(3)
    d = ops._desc(N, H, W, C, K, R, R, stride, pad)
    x = _nhwc_bf16(torch.randn(N, C, H, W))
    dy = _nhwc_bf16(torch.randn(N, K, d.P, d.Q))
    assert ops._lib().srgan_conv2d_bf16_supported(d, 2) == 1
    dw, _ = ops._wgrad(d, x, dy, True, False)
    assert dw.dtype == torch.float32 and dw.is_contiguous(memory_format=CL)
    ref = torch.nn.grad.conv2d_weight(x.float(), (K, C, R, R), dy.float(), stride, pad)
    assert _rel(dw, ref) < 1e-4, (g, _rel(dw, ref))


@pytest.mark.parametrize("xd,yd", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16),
                                   (torch.bfloat16, torch.float32)])
def test_instance_norm_mixed_storage(xd, yd):
    """srgan_inorm_{fwd,bwd}_mixed: x / dx and y / dy / residual in fp32 or bf16 independently; arithmetic fp32.
    Reference: fp64 on the SAME (rounded) inputs; the only error left is the rounding of the bf16 outputs (2^-9)."""
    N, C, H = 5, 128, 24
    torch.manual_seed(4)
    x = (torch.randn(N, C, H, H, device=DEV) * 1.3 + 0.2).to(xd).contiguous(memory_format=CL).requires_grad_(True)
    g = torch.randn(C, device=DEV).requires_grad_(True)
    b = torch.randn(C, device=DEV).requires_grad_(True)
    cb = torch.randn(N, C, device=DEV).requires_grad_(True)
    res = torch.randn(N, C, H, H, device=DEV).to(yd).contiguous(memory_format=CL).requires_grad_(True)
    for act, residual in ((ops.ACT_RELU, None), (ops.ACT_NONE, res)):
        y = ops.instance_norm_act(x, g, b, cb, residual, 1e-5, act, 0.0, out_dtype=yd)
        assert y.dtype == yd and y.is_contiguous(memory_format=CL)
        gy = torch.randn(N, C, H, H, device=DEV).to(yd).contiguous(memory_format=CL)
        ins = [x, g, b, cb] + ([res] if residual is not None else [])
        got = torch.autograd.grad(y, ins, gy)
        assert got[0].dtype == xd
        xr, gr, br, cr = (t.detach().double().requires_grad_(True) for t in (x, g, b, cb))
        rr = res.detach().double().requires_grad_(True)
        v = (F.instance_norm(xr, eps=1e-5) + cr[:, :, None, None]) * gr[None, :, None, None] + br[None, :, None, None]
        yr = torch.relu(v) if act == ops.ACT_RELU else v + rr
        ref = torch.autograd.grad(yr, [xr, gr, br, cr] + ([rr] if residual is not None else []), gy.double())
        tol_y = 4e-3 if yd == torch.bfloat16 else 1e-5
        tol_x = 4e-3 if xd == torch.bfloat16 else 1e-5
        assert _rel(y, yr) < tol_y
        assert _rel(got[0], ref[0]) < tol_x
        for a_, r_ in zip(got[1:4], ref[1:4]):
            assert _rel(a_, r_) < 1e-4                      # parameter gradients: fp32 sums of exact products
        if residual is not None:
            assert torch.equal(got[4], gy)


def _generator_pair(batch=3):
    """Full-width generator and a forward/backward on it, once per engine, same weights and inputs."""
    import cases
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(0)
    G = model.SingleGenerator(3, 64, 2, 2, 6, "instance", num_con=12).to(DEV)
    x = (torch.rand(batch, 3, 128, 128, device=DEV) * 2 - 1).requires_grad_(True)
    c = torch.randn(batch, 12, device=DEV)
    gy = torch.randn(batch, 3, 128, 128, device=DEV)
    out = {}
    for eng in ("fp32", "auto", "bf16"):
        ops.set_conv_engine(eng)
        try:
            for p in G.parameters():
                p.grad = None
            y = G(x, c)
            dx, = torch.autograd.grad(y, x, gy, retain_graph=True)
            y.backward(gy)
            out[eng] = (y.detach().clone(), dx.clone(), {n: p.grad.detach().clone() for n, p in G.named_parameters()})
        finally:
            ops.set_conv_engine("auto")
    return out


def _rel_dict(a, b):
    num = sum(float((a[k].double() - b[k].double()).pow(2).sum()) for k in b)
    den = sum(float(b[k].double().pow(2).sum()) for k in b)
    return (num / den) ** 0.5


def _cos_dict(a, b):
    dot = sum(float((a[k].double() * b[k].double()).sum()) for k in b)
    na = sum(float(a[k].double().pow(2).sum()) for k in b) ** 0.5
    nb_ = sum(float(b[k].double().pow(2).sum()) for k in b) ** 0.5
    return dot / (na * nb_)


def test_generator_bf16_trunk_against_fp32_engine():
    """SingleGenerator forward / backward with the bf16 trunk against the exact-fp32 engine on the same weights
    (ref pyfiles/model.py:236-249), the TF32 engine's distance printed next to it.
    Measured on B200 (batch 3, random-init weights, white-noise output gradient):
        TF32  : y 1.7e-3, dx 6.6e-2, parameter gradients 5.4e-2
        bf16  : y 1.5e-2, dx 1.9e-1, parameter gradients 1.6e-1   (cosine 0.987)
    The building blocks are exact to 1e-4 .. 4e-3 on identical operands (tests above); what is measured here is how 17
    convolutions and 17 norms amplify operand rounding (2^-9 per stored bf16 value, 2^-11 per TF32 operand): the
    network's gradient is discontinuous in its pre-activations (ReLU masks), so both engines sit ~40 x above their
    per-element rounding and bf16 sits 3 x above TF32, the ratio of the roundings.  Stated bf16 tolerance for this
    worst-case probe: y 3e-2, gradients 3.5e-1 relative L2 and cosine >= 0.95 (2 x measured); the training-step losses
    are held to 1e-2 against the oracle below and in bench.py."""
    out = _generator_pair()
    y0, dx0, g0 = out["fp32"]
    for eng in ("auto", "bf16"):
        y, dx, g = out[eng]
        e = (_rel(y, y0), _rel(dx, dx0), _rel_dict(g, g0), _cos_dict(g, g0))
        print("generator %s vs fp32 engine: y %.2e  dx %.2e  param grads %.2e (cos %.4f)" % ((eng,) + e))
        if eng == "bf16":
            assert e[0] < 3e-2 and e[1] < 3.5e-1 and e[2] < 3.5e-1 and e[3] > 0.95, e
        else:
            assert e[0] < 5e-3 and e[1] < 2e-1 and e[2] < 2e-1, e
    assert out["bf16"][0].dtype == torch.float32


def test_bf16_step_matches_oracle_within_bf16_tolerance():
    """One full-width SRGAN step (nb03 recipe, batch 2, k = 1) with the bf16 trunk against the CPU oracle.
    Stated bf16 tolerance: losses 1e-2 relative; discriminator gradients (fp32 D on bf16-generated fakes) 5e-2."""
    import numpy as np
    import cases
    import srgan_oracle as so
    name = "srgan_full"
    c = cases.CASES[name]
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(c["seed"])
    np.random.seed(c["seed"])
    nets = cases.build_nets(model, name, DEV)
    sds = cases.state_dicts(nets)
    torch.manual_seed(c["seed"] + 500)
    oracle = cases.build_oracle(name, sds, so)
    torch.manual_seed(c["seed"] + 500)
    sg = cases.build_trainer(nb, name, tuple(n.to(DEV) for n in nets), DEV)
    x, label = cases.synthetic_batch(c["batch"], util.get_target)
    torch.manual_seed(c["seed"] + 1000)
    ref = [float(e) for e in oracle.train(x, label)]
    ops.set_conv_engine("bf16")
    try:
        torch.manual_seed(c["seed"] + 1000)
        got = [float(e) for e in sg.train(x.to(DEV), {"source": label["source"].to(DEV), "target": label["target"]})]
        torch.cuda.synchronize()
    finally:
        ops.set_conv_engine("auto")
    err = [abs(a - b) / max(1.0, abs(b)) for a, b in zip(got, ref)]
    print("bf16 step losses %s vs oracle %s: rel %s" % (got, ref, ["%.2e" % e for e in err]))
    assert max(err) < 1e-2, (got, ref)


def test_bf16_cuda_graph_replay_is_bit_identical_to_eager():
    """The bf16 trunk under CUDA-graph replay (shadow refresh after each Adam step is part of the captured step)."""
    import numpy as np
    import cases
    c = dict(cases.CASES["srgan_full"], batch=2, k=2, res_num=2)
    model, util, nb = cases.use_product_modules()

    def run(graph):
        torch.manual_seed(0)
        np.random.seed(0)
        nets = tuple(n.to(DEV) for n in cases.build_nets(model, c, DEV))
        torch.manual_seed(1)
        sg = cases.build_trainer(nb, c, nets, DEV)
        if graph:
            sg.enable_cuda_graph(warmup=1)
        losses = []
        torch.manual_seed(2)
        for step in range(4):
            x, label = cases.synthetic_batch(c["batch"], util.get_target, seed=100 + step)
            errs = sg.train(x.to(DEV), {"source": label["source"].to(DEV), "target": label["target"]})
            losses.append([float(e) for e in errs])
        torch.cuda.synchronize()
        return losses, torch.cat([p.detach().reshape(-1) for n in nets for p in n.parameters()]).cpu()
    ops.set_conv_engine("bf16")
    try:
        le, pe = run(False)
        lg, pg = run(True)
    finally:
        ops.set_conv_engine("auto")
    assert le == lg, (le, lg)
    assert torch.equal(pe, pg)


STAT_GEOMS = [  # N, C, H, W, K, R, stride, pad, transposed
    (5, 256, 32, 32, 256, 3, 1, 1, False),     # residual block: 8 tiles per image, 256-wide tiles (+ half-width tail)
    (3, 64, 128, 128, 128, 4, 2, 1, False),    # down1: MT = 2 sub-tiles
    (70, 256, 32, 32, 256, 3, 1, 1, False),    # more than three waves of tiles
    (3, 256, 32, 32, 128, 4, 2, 1, True),      # up0: transposed conv = dgrad with 4 output-parity classes
    (2, 128, 64, 64, 64, 4, 2, 1, True),       # up1
]


@pytest.mark.parametrize("g", STAT_GEOMS)
def test_tile_statistics_from_the_conv_epilogue(g):
    """The bf16 forward convolutions write per-tile sum / sum of squares of their (stored) outputs; the instance norm
    behind them folds those instead of reading the tensor (srgan_inorm_stats_from_tiles).  Checks the raw sums against
    torch on the stored output and the fused conv -> norm against the same norm run stand-alone."""
    N, C, H, W, K, R, stride, pad, transposed = g
    prev = ops.set_tile_stats(True)           # opt-in path (srgan_ops: the conv kernel has no smem bandwidth to spare)
    try:
        _tile_stats_case(N, C, H, W, K, R, stride, pad, transposed)
    finally:
        ops.set_tile_stats(prev)


def _tile_stats_case(N, C, H, W, K, R, stride, pad, transposed):
    torch.manual_seed(5)
    x = _nhwc_bf16(torch.randn(N, C, H, W)).requires_grad_(True)
    gam, bet, cb = torch.randn(K, device=DEV), torch.randn(K, device=DEV), torch.randn(N, K, device=DEV)
    if transposed:
        w = (torch.randn(C, K, R, R, device=DEV) * 0.05).contiguous(memory_format=CL)
        conv = lambda: ops.conv_transpose2d(x, w, stride, pad)
    else:
        w = (torch.randn(K, C, R, R, device=DEV) * 0.05).contiguous(memory_format=CL)
        conv = lambda: ops.conv2d(x, w, None, stride, pad)
    y = conv()
    pending = ops._tile_stats
    assert pending is not None and pending[0] == y.data_ptr(), "the forward convolution should leave tile statistics"
    tiles, rows = pending[1], pending[2]
    HW = y.shape[2] * y.shape[3]
    assert rows * 128 == HW
    yf = y.detach().float()
    s1 = tiles[..., 0].sum(1)
    s2 = tiles[..., 1].sum(1)
    assert _rel(s1, yf.sum((2, 3))) < 1e-4 and _rel(s2, (yf * yf).sum((2, 3))) < 1e-5
    fused = ops.instance_norm_act(y, gam, bet, cb, None, 1e-5, ops.ACT_RELU, 0.0)
    assert ops._tile_stats is None                       # consumed
    alone = ops.instance_norm_act(y.detach().clone(), gam, bet, cb, None, 1e-5, ops.ACT_RELU, 0.0)
    assert _rel(fused, alone) < 2e-3                     # same statistics up to fp32 rounding, outputs rounded to bf16
    assert float((fused.float() - alone.float()).abs().max()) < 0.1
    # backward goes through the saved statistics
    gy = _nhwc_bf16(torch.randn(*fused.shape))
    dx_f, = torch.autograd.grad(fused, x, gy)
    y2 = conv()
    ops._tile_stats = None
    dx_a, = torch.autograd.grad(ops.instance_norm_act(y2, gam, bet, cb, None, 1e-5, ops.ACT_RELU, 0.0), x, gy)
    assert _rel(dx_f, dx_a) < 1e-2


PAIR_GEOMS = [  # N, C, H, W, K, R, stride, pad  -- >= 296 tiles of 128 pixels and 256-wide filter tiles: CTA pairs
    (64, 256, 32, 32, 256, 3, 1, 1),    # the residual block at the production batch (512 tiles, 256 pairs, 3.5 waves)
    (99, 64, 24, 16, 256, 3, 1, 1),     # 3 tiles per image x 99 images: an odd tile count, the last pair is half empty
    (40, 128, 64, 64, 512, 4, 2, 1),    # stride 2 (parity view of x), two filter tiles
]


def _run_in_subprocess_with_pairs(test_id):
    """The kernel choice is read from the environment once per process: run the test body in a child with
    SRGAN_CONV_PAIRS=1."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, SRGAN_CONV_PAIRS="1", SRGAN_PAIR_CHILD="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", __file__ + "::" + test_id],
                       capture_output=True, text=True, env=env, cwd=os.path.dirname(os.path.dirname(__file__)))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("g", PAIR_GEOMS)
def test_cta_pair_kernel_fprop_dgrad(g, request):
    """conv_umma2_kernel (tcgen05 cta_group::2: two CTAs per 256 x 256 tile, leader issues the MMAs; opt-in with
    SRGAN_CONV_PAIRS=1) against torch on the same bf16-rounded operands, forward with bias + LeakyReLU and input
    gradient with the fused skip addend."""
    import os
    if os.environ.get("SRGAN_PAIR_CHILD") != "1":
        return _run_in_subprocess_with_pairs(request.node.name)
    N, C, H, W, K, R, stride, pad = g
    torch.manual_seed(7)
    d = ops._desc(N, H, W, C, K, R, R, stride, pad)
    x = _nhwc_bf16(torch.randn(N, C, H, W))
    w = _nhwc_bf16(torch.randn(K, C, R, R) * 0.05)
    b = torch.randn(K, device=DEV)
    y = torch.empty((N, K, d.P, d.Q), dtype=torch.bfloat16, device=DEV).contiguous(memory_format=CL)
    ops._call("srgan_conv2d_fprop_bf16", d, ops._p(x), ops._p(w), ops._p(b), ops._p(y), ops.ACT_LRELU, 0.2, None,
              ops._stream())
    ref = F.leaky_relu(F.conv2d(x.float(), w.float(), b, stride, pad), 0.2)
    assert _rel(y.float(), ref) < TOL, (g, _rel(y.float(), ref))
    if C % 64 == 0 and K % 64 == 0 and C >= 129:         # the dgrad's filter tile (C output channels) is 256 wide
        dy = _nhwc_bf16(torch.randn(N, K, d.P, d.Q))
        nb = ops._lib().srgan_conv2d_bf16_workspace(d, 1)
        ws = ops._workspace(torch.device(DEV, torch.cuda.current_device()), nb)
        dx = torch.empty((N, C, H, W), dtype=torch.bfloat16, device=DEV).contiguous(memory_format=CL)
        add = _nhwc_bf16(torch.randn(N, C, H, W)) if stride == 1 else None
        ops._call("srgan_conv2d_dgrad_bf16", d, ops._p(dy), ops._p(w), ops._p(add), ops._p(dx), None, ops._p(ws), nb,
                  ops._stream())
        refx = torch.nn.grad.conv2d_input((N, C, H, W), w.float(), dy.float(), stride, pad)
        if add is not None:
            refx = refx + add.float()
        assert _rel(dx.float(), refx) < TOL, (g, _rel(dx.float(), refx))


def test_encoder_bf16_trunk_against_fp32_engine():
    """Encoder (ref pyfiles/model.py:452-482) forward / backward with the bf16 block trunk against the exact-fp32
    engine on the same weights; the TF32 engine's distance is printed next to it.  Stated bf16 tolerance: mu / logvar
    2e-2 relative L2, parameter gradients 2e-1 and cosine >= 0.97 (the generator probe explains why whole-network
    gradients sit well above the per-element rounding)."""
    import cases
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(0)
    E = model.Encoder(3, 8, 64, 4, "instance", 4, DEV).to(DEV)
    x = (torch.rand(6, 3, 128, 128, device=DEV) * 2 - 1).requires_grad_(True)
    gmu, glv = torch.randn(6, 8, device=DEV), torch.randn(6, 8, device=DEV)
    out = {}
    for eng in ("fp32", "auto", "bf16"):
        ops.set_conv_engine(eng)
        try:
            for p in E.parameters():
                p.grad = None
            torch.manual_seed(1)
            _, mu, logvar, cls, _ = E(x)
            dx, = torch.autograd.grad([mu, logvar], x, [gmu, glv], retain_graph=True)
            torch.autograd.backward([mu, logvar], [gmu, glv])
            out[eng] = (torch.cat([mu, logvar], 1).detach().clone(), dx.clone(),
                        {n: p.grad.detach().clone() for n, p in E.named_parameters() if p.grad is not None})
        finally:
            ops.set_conv_engine("auto")
    y0, dx0, g0 = out["fp32"]
    for eng in ("auto", "bf16"):
        y, dx, g = out[eng]
        e = (_rel(y, y0), _rel(dx, dx0), _rel_dict(g, g0), _cos_dict(g, g0))
        print("encoder %s vs fp32 engine: mu/logvar %.2e  dx %.2e  param grads %.2e (cos %.4f)" % ((eng,) + e))
        if eng == "bf16":
            assert e[0] < 2e-2 and e[2] < 2e-1 and e[3] > 0.97, e


@pytest.mark.parametrize("shape", [(3, 64, 62, 62, 1), (2, 128, 9, 7, 1), (2, 64, 8, 8, 2)])
def test_reflect_pad_and_pool_bf16(shape):
    """bf16 reflect padding (pure copy forward; backward adds the mirrored contributions in fp32, one rounding) and the
    mixed pool-add tail of the encoder block, against torch."""
    N, C, H, W, pad = shape
    torch.manual_seed(2)
    x = _nhwc_bf16(torch.randn(N, C, H, W)).requires_grad_(True)
    y = ops._ReflectPadFn.apply(x, pad)
    ref = F.pad(x.detach().float(), (pad,) * 4, mode="reflect")
    assert y.dtype == torch.bfloat16 and torch.equal(y.float(), ref)
    gy = _nhwc_bf16(torch.randn(*y.shape))
    dx, = torch.autograd.grad(y, x, gy)
    xr = x.detach().float().requires_grad_(True)
    dxr, = torch.autograd.grad(F.pad(xr, (pad,) * 4, mode="reflect"), xr, gy.float())
    assert dx.dtype == torch.bfloat16 and _rel(dx, dxr) < 4e-3
    a = _nhwc_bf16(torch.randn(N, C, H, W)).requires_grad_(True)
    b = torch.randn(N, C, H // 2, W // 2, device=DEV).contiguous(memory_format=CL).requires_grad_(True)
    z = ops.avg_pool2_add(a, b)
    assert z.dtype == torch.float32 and _rel(z, F.avg_pool2d(a.detach().float(), 2) + b.detach()) < 1e-6
    gz = torch.randn_like(z)
    da, db = torch.autograd.grad(z, [a, b], gz)
    ar = a.detach().float().requires_grad_(True)
    dar, = torch.autograd.grad(F.avg_pool2d(ar, 2), ar, gz)
    assert da.dtype == torch.bfloat16 and _rel(da, dar) < 4e-3 and torch.equal(db, gz)


def test_act_bwd_bf16_and_cast_f32():
    """LeakyReLU backward on bf16 tensors and the widening cast at a bf16 tower's fp32 boundary."""
    torch.manual_seed(7)
    y = torch.randn(3, 64, 9, 11, device=DEV).to(torch.bfloat16).contiguous(memory_format=CL)
    dy = torch.randn(3, 64, 9, 11, device=DEV).to(torch.bfloat16).contiguous(memory_format=CL)
    dz = ops._act_bwd(dy, y, ops.ACT_LRELU, 0.01)
    ref = (dy.float() * torch.where(y.float() > 0, 1.0, 0.01)).to(torch.bfloat16)
    assert dz.dtype == torch.bfloat16 and torch.equal(dz, ref)
    x = y.clone().requires_grad_(True)
    f = ops.cast_f32(x)
    assert f.dtype == torch.float32 and torch.equal(f, y.float())
    g, = torch.autograd.grad(f, x, dy.float())
    assert g.dtype == torch.bfloat16 and torch.equal(g, dy)
    z = torch.randn(5, 7, device=DEV)
    assert ops.cast_f32(z) is z


def test_discriminator_bf16_tower_against_fp32_engine():
    """SingleDiscriminator_solo_multi (nb02 / 03 / 05) forward / backward with the wide tower in bf16 (thin16 stem with
    fused LeakyReLU, three kind::f16 convolutions, fp32 heads; the narrow tower stays TF32) against the exact-fp32 engine
    on the same weights; the TF32 engine's distance printed next to it.  ref pyfiles/model.py:294-346."""
    import cases
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(0)
    D = model.SingleDiscriminator_solo_multi(3, 64, 2, 4, "instance", 4).to(DEV)
    x = (torch.rand(6, 3, 128, 128, device=DEV) * 2 - 1).requires_grad_(True)
    out = {}
    for eng in ("fp32", "auto", "bf16"):
        ops.set_conv_engine(eng)
        try:
            for p in D.parameters():
                p.grad = None
            o, c = D(x)
            assert all(t.dtype == torch.float32 for t in o + c)
            loss = sum((t ** 2).mean() for t in o) + sum((t[:, 0]).mean() for t in c)
            dx, = torch.autograd.grad(loss, x, retain_graph=True)
            loss.backward()
            out[eng] = ([t.detach().clone() for t in o + c], dx.clone(),
                        {n: p.grad.detach().clone() for n, p in D.named_parameters()})
        finally:
            ops.set_conv_engine("auto")
    y0, dx0, g0 = out["fp32"]
    for eng in ("auto", "bf16"):
        y, dx, g = out[eng]
        ey = max(_rel(a, b) for a, b in zip(y, y0))
        e = (ey, _rel(dx, dx0), _rel_dict(g, g0), _cos_dict(g, g0))
        print("discriminator %s vs fp32 engine: outputs %.2e  dx %.2e  param grads %.2e (cos %.4f)" % ((eng,) + e))
        # measured on B200: TF32 2.1e-3 / 3.3e-2 / 5.3e-3, bf16 tower 2.4e-3 / 7.4e-2 / 1.1e-2 (cos 0.9999); bounds = 3 x
        if eng == "bf16":
            assert e[0] < 8e-3 and e[1] < 2.5e-1 and e[2] < 4e-2 and e[3] > 0.999, e
        else:
            assert e[0] < 7e-3 and e[1] < 1e-1 and e[2] < 2e-2, e


def test_discriminator_original_multi_bf16_tower_against_fp32_engine():
    """SingleDiscriminator_original_multi (notebook 01: the patch head is the last convolution of the tower's stack):
    the wide tower in bf16 up to the head, which reads an fp32 copy; against the exact-fp32 engine on the same weights.
    ref pyfiles/model.py:255-292."""
    import cases
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(1)
    D = model.SingleDiscriminator_original_multi(3, 64, 2, 4, "instance").to(DEV)
    x = (torch.rand(4, 3, 128, 128, device=DEV) * 2 - 1).requires_grad_(True)
    out = {}
    for eng in ("fp32", "bf16"):
        ops.set_conv_engine(eng)
        try:
            for p in D.parameters():
                p.grad = None
            o = D(x)
            assert all(t.dtype == torch.float32 for t in o) and o[0].shape == (4, 1, 7, 7)
            loss = sum((t ** 2).mean() for t in o)
            dx, = torch.autograd.grad(loss, x, retain_graph=True)
            loss.backward()
            out[eng] = ([t.detach().clone() for t in o], dx.clone(),
                        {n: p.grad.detach().clone() for n, p in D.named_parameters()})
        finally:
            ops.set_conv_engine("auto")
    y0, dx0, g0 = out["fp32"]
    y, dx, g = out["bf16"]
    e = (max(_rel(a, b) for a, b in zip(y, y0)), _rel(dx, dx0), _rel_dict(g, g0), _cos_dict(g, g0))
    print("discriminator (original, multi) bf16 vs fp32 engine: outputs %.2e  dx %.2e  param grads %.2e (cos %.4f)" % e)
    assert e[0] < 1e-2 and e[1] < 2.5e-1 and e[2] < 5e-2 and e[3] > 0.998, e


def test_discriminator_two_streams_is_bit_identical():
    """The two towers of the multi-scale discriminators on two streams (srgan_ops.tower_streams) against the
    single-stream schedule: same kernels on the same data - outputs, input gradient and every parameter gradient must be
    bit-identical (per-stream scratch, autograd's cross-stream synchronisation)."""
    import cases
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(2)
    D = model.SingleDiscriminator_solo_multi(3, 64, 2, 4, "instance", 4).to(DEV)
    x = (torch.rand(8, 3, 128, 128, device=DEV) * 2 - 1).requires_grad_(True)
    res = {}
    prev = ops.TOWER_STREAMS
    try:
        for eng in ("auto", "bf16"):
            ops.set_conv_engine(eng)
            for on in (True, False, True):
                ops.TOWER_STREAMS = on
                for p in D.parameters():
                    p.grad = None
                o, c = D(x)
                loss = sum((t ** 2).mean() for t in o) + sum((t[:, 1]).mean() for t in c)
                dx, = torch.autograd.grad(loss, x, retain_graph=True)
                loss.backward()
                torch.cuda.synchronize()
                cur = ([t.detach().clone() for t in o + c], dx.clone(), [p.grad.detach().clone() for p in D.parameters()])
                key = eng
                if key in res:
                    for a, b in zip(cur[0], res[key][0]):
                        assert torch.equal(a, b), (eng, on)
                    assert torch.equal(cur[1], res[key][1]), (eng, on)
                    for a, b in zip(cur[2], res[key][2]):
                        assert torch.equal(a, b), (eng, on)
                else:
                    res[key] = cur
    finally:
        ops.TOWER_STREAMS = prev
        ops.set_conv_engine("auto")
