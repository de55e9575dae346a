"""GPU: every kernel entry point against a plain PyTorch reference of the same op (fp64 on the device),
forward and backward, over the layer geometries of SURVEY Appendix A.  Calls go through the C ABI
(srgan_ops -> ctypes -> libsrgan_b200.so)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import srgan_ops as ops
import srgan_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda"
CL = torch.channels_last


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _check(prod_fn, ref_fn, inputs, tol, grad_tol=None, names=None):
    """inputs: list of fp32 cuda tensors (those with requires_grad get gradient checks)."""
    grad_tol = grad_tol or tol
    p_in = [t.detach().clone().requires_grad_(t.requires_grad) for t in inputs]
    r_in = [t.detach().double().clone().requires_grad_(t.requires_grad) for t in inputs]
    y = prod_fn(*p_in)
    yr = ref_fn(*r_in)
    assert y.shape == yr.shape, (y.shape, yr.shape)
    assert _rel(y, yr) < tol, ("forward", _rel(y, yr))
    w = torch.randn(yr.shape, generator=torch.Generator().manual_seed(7), dtype=torch.float64).to(DEV)
    (y * w.float()).sum().backward()
    (yr * w).sum().backward()
    for i, (a, b) in enumerate(zip(p_in, r_in)):
        if b.requires_grad:
            assert a.grad is not None, i
            e = _rel(a.grad, b.grad)
            assert e < grad_tol, ("grad", names[i] if names else i, e)
    return y


def _clear_of_kink(x, pre_fn, margin=1e-4):
    """Nudge the few elements of x whose fp64 pre-activation pre_fn(x) lies within `margin` of 0.  With a
    piecewise-linear activation the gradient mask of such an element is decided by fp32 rounding, and one
    flipped element out of 2M is a 4e-4 relative-L2 difference in dx (5e-4 in dbeta) -- a property of the
    comparison, not of the kernel.  ~2M * 0.8 * margin elements are touched."""
    x = x.detach().clone()
    for _ in range(8):
        near = pre_fn(x.double()).abs() < margin
        if not bool(near.any()):
            return x
        x = torch.where(near, x + 0.01, x)
    raise AssertionError("could not move the inputs away from the activation kink")


@pytest.fixture(autouse=True)
def _fp32_engine():
    ops.set_conv_engine("fp32")
    yield
    ops.set_conv_engine("auto")


# (name, N, C, H, W, K, R, stride, pad, bias)  -- Appendix A, batch 2
CONV_GEOMS = [
    ("G.down0", 2, 3, 128, 128, 64, 7, 1, 3, False),
    ("G.down1", 2, 64, 128, 128, 128, 4, 2, 1, False),
    ("G.down2", 2, 128, 64, 64, 256, 4, 2, 1, False),
    ("G.res", 2, 256, 32, 32, 256, 3, 1, 1, False),
    ("G.out", 2, 64, 128, 128, 3, 7, 1, 3, False),
    ("E.first", 2, 3, 128, 128, 64, 7, 2, 1, True),
    ("E.short", 2, 64, 31, 31, 128, 1, 1, 0, True),
    ("E.tail", 3, 512, 7, 7, 1024, 3, 1, 0, False),
    ("D.c0", 2, 3, 128, 128, 64, 4, 2, 1, False),
    ("D.c3", 2, 256, 16, 16, 512, 4, 2, 1, False),
    ("D.patch", 2, 512, 8, 8, 1, 4, 1, 1, True),
    ("D.class", 2, 512, 8, 8, 4, 8, 1, 0, True),
    ("D2.class", 2, 256, 4, 4, 4, 4, 1, 0, True),
    ("odd", 1, 24, 15, 13, 40, 3, 2, 1, True),
]


@pytest.mark.parametrize("geom", CONV_GEOMS, ids=[g[0] for g in CONV_GEOMS])
def test_conv2d_fp32_engine(geom):
    _, N, C, H, W, K, R, stride, pad, bias = geom
    x = _rand(N, C, H, W, seed=1).contiguous(memory_format=CL).requires_grad_(True)
    w = _rand(K, C, R, R, seed=2, scale=(C * R * R) ** -0.5).contiguous(memory_format=CL).requires_grad_(True)
    b = _rand(K, seed=3).requires_grad_(True) if bias else None
    ins = [x, w] + ([b] if bias else [])
    _check(lambda x, w, b=None: ops.conv2d(x, w, b, stride, pad),
           lambda x, w, b=None: F.conv2d(x, w, b, stride, pad), ins, 1e-5, 3e-5, ["x", "w", "b"])  # fp32 sums of up to 32768 terms


def test_conv2d_nchw_input_and_fused_activation():
    x = _rand(2, 16, 20, 20, seed=4).requires_grad_(True)                      # NCHW-contiguous
    w = _rand(24, 16, 4, 4, seed=5, scale=0.1).contiguous(memory_format=CL).requires_grad_(True)
    for act, slope, ref in ((ops.ACT_LRELU, 0.01, lambda t: F.leaky_relu(t, 0.01)), (ops.ACT_TANH, 0.0, torch.tanh),
                            (ops.ACT_RELU, 0.0, F.relu)):
        _check(lambda x, w: ops.conv2d(x, w, None, 2, 1, "zeros", act, slope),
               lambda x, w: ref(F.conv2d(x, w, None, 2, 1)), [x, w], 2e-6, 3e-5)


def test_conv2d_reflect():
    for (C, H) in ((64, 62), (128, 31), (256, 15), (512, 7)):
        x = _rand(2, C, H, H, seed=6).contiguous(memory_format=CL).requires_grad_(True)
        w = _rand(C, C, 3, 3, seed=7, scale=(9 * C) ** -0.5).contiguous(memory_format=CL).requires_grad_(True)
        _check(lambda x, w: ops.conv2d(x, w, None, 1, 1, "reflect"),
               lambda x, w: F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w), [x, w], 2e-6, 3e-5)


@pytest.mark.parametrize("shape", [(2, 256, 32, 32, 128), (2, 128, 64, 64, 64), (1, 16, 5, 7, 8)])
def test_conv_transpose2d(shape):
    N, Cin, H, W, Cout = shape
    x = _rand(N, Cin, H, W, seed=8).contiguous(memory_format=CL).requires_grad_(True)
    w = _rand(Cin, Cout, 4, 4, seed=9, scale=(Cin * 4) ** -0.5).contiguous(memory_format=CL).requires_grad_(True)
    _check(lambda x, w: ops.conv_transpose2d(x, w, 2, 1),
           lambda x, w: F.conv_transpose2d(x, w, None, 2, 1), [x, w], 2e-6, 3e-5)


def test_linear():
    x = _rand(5, 1024, seed=10).requires_grad_(True)
    w = _rand(8, 1024, seed=11, scale=0.03).requires_grad_(True)
    b = _rand(8, seed=12).requires_grad_(True)
    _check(ops.linear, F.linear, [x, w, b], 2e-6, 2e-5)


NORM_SHAPES = [(2, 64, 128, 128), (2, 128, 64, 64), (3, 256, 32, 32), (2, 64, 62, 62), (2, 128, 31, 31),
               (2, 256, 15, 15), (5, 512, 7, 7), (1, 8, 3, 5), (2, 16, 9, 9), (40, 32, 16, 16),
               (2, 24, 10, 10), (150, 64, 32, 32), (1, 64, 200, 200), (3, 1024, 3, 3)]


@pytest.mark.parametrize("shape", NORM_SHAPES, ids=[str(s) for s in NORM_SHAPES])
def test_instance_norm_cond_affine_act(shape):
    N, C, H, W = shape
    x = (_rand(N, C, H, W, seed=13) * 1.7 + 0.3).contiguous(memory_format=CL).requires_grad_(True)
    gamma = (_rand(C, seed=14) * 0.2 + 1).requires_grad_(True)
    beta = (_rand(C, seed=15) * 0.2).requires_grad_(True)
    cb = (_rand(N, C, seed=16) * 0.5).requires_grad_(True)

    def ref(act):
        def f(x, gamma, beta, cb):
            h = (F.instance_norm(x, eps=1e-5) + cb[:, :, None, None]) * gamma[None, :, None, None] \
                + beta[None, :, None, None]
            return act(h)
        return f
    x = _clear_of_kink(x, lambda t: ref(lambda h: h)(t, gamma.detach().double(), beta.detach().double(),
                                                     cb.detach().double()))
    x = x.contiguous(memory_format=CL).requires_grad_(True)
    _check(lambda x, g, b, c: ops.instance_norm_act(x, g, b, c, None, 1e-5, ops.ACT_RELU),
           ref(F.relu), [x, gamma, beta, cb], 3e-6, 2e-4, ["x", "gamma", "beta", "cbias"])
    _check(lambda x, g, b, c: ops.instance_norm_act(x, g, b, c, None, 1e-5, ops.ACT_LRELU, 0.2),
           ref(lambda t: F.leaky_relu(t, 0.2)), [x, gamma, beta, cb], 3e-6, 2e-4)


def test_instance_norm_plain_and_residual():
    x = _clear_of_kink(_rand(2, 128, 64, 64, seed=17), lambda t: F.instance_norm(t, eps=1e-5))
    x = x.contiguous(memory_format=CL).requires_grad_(True)
    _check(lambda x: ops.instance_norm_act(x, act=ops.ACT_RELU), lambda x: F.relu(F.instance_norm(x, eps=1e-5)),
           [x], 3e-6, 2e-4)
    x = _rand(2, 256, 32, 32, seed=18).contiguous(memory_format=CL).requires_grad_(True)
    r = _rand(2, 256, 32, 32, seed=19).contiguous(memory_format=CL).requires_grad_(True)
    g = (_rand(256, seed=20) * 0.1 + 1).requires_grad_(True)
    b = (_rand(256, seed=21) * 0.1).requires_grad_(True)
    cb = _rand(2, 256, seed=22).requires_grad_(True)
    _check(lambda x, r, g, b, cb: ops.instance_norm_act(x, g, b, cb, r),
           lambda x, r, g, b, cb: (F.instance_norm(x, eps=1e-5) + cb[:, :, None, None]) * g[None, :, None, None]
           + b[None, :, None, None] + r, [x, r, g, b, cb], 3e-6, 2e-4)


def test_cond_bias():
    con = _rand(6, 12, seed=23).requires_grad_(True)
    w = _rand(256, 12, seed=24, scale=0.3).requires_grad_(True)
    b = _rand(256, seed=25, scale=0.3).requires_grad_(True)
    _check(ops.cond_bias, lambda c, w, b: torch.tanh(F.linear(c, w, b)), [con, w, b], 2e-6, 2e-5)


def test_pools():
    for H, W in ((62, 62), (31, 31), (15, 15), (7, 7), (8, 6)):
        x = _rand(2, 16, H, W, seed=26).contiguous(memory_format=CL).requires_grad_(True)
        _check(ops.avg_pool2, lambda x: F.avg_pool2d(x, 2, 2), [x], 1e-6)
        b = _rand(2, 16, H // 2, W // 2, seed=27).contiguous(memory_format=CL).requires_grad_(True)
        _check(ops.avg_pool2_add, lambda a, b: F.avg_pool2d(a, 2, 2) + b, [x, b], 1e-6)
    for H, W in ((128, 128), (9, 7), (4, 4)):
        x = _rand(2, 3, H, W, seed=28).contiguous(memory_format=CL).requires_grad_(True)
        # reference on the CPU: torch's CUDA avg_pool2d backward disagrees with its own CPU kernel for
        # channels-last inputs with count_include_pad=False (seen on torch 2.11 / B200)
        y = ops.avg_pool3s2(x)
        xr = x.detach().cpu().double().contiguous().requires_grad_(True)
        yr = F.avg_pool2d(xr, 3, 2, 1, count_include_pad=False)
        assert _rel(y, yr) < 1e-6
        gy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
        y.backward(gy.float().to(DEV))
        yr.backward(gy)
        assert _rel(x.grad, xr.grad) < 1e-6
    x = _rand(3, 1024, 3, 3, seed=29).contiguous(memory_format=CL).requires_grad_(True)
    _check(lambda x: ops.lrelu_gap(x, 0.2),
           lambda x: F.adaptive_avg_pool2d(F.leaky_relu(x, 0.2), 1).flatten(1), [x], 1e-6)


def test_softmax_reparam_add_layout():
    x = _rand(7, 4, seed=30).requires_grad_(True)
    _check(ops.softmax_rows, lambda x: F.softmax(x, dim=1), [x], 1e-6, 1e-5)
    mu, lv = _rand(5, 8, seed=31).requires_grad_(True), _rand(5, 8, seed=32).requires_grad_(True)
    eps = _rand(5, 8, seed=33)
    _check(lambda m, l: ops.reparametrize(m, l, eps), lambda m, l: eps.double() * torch.exp(0.5 * l) + m, [mu, lv],
           1e-6, 1e-5)
    a = _rand(2, 8, 5, 5, seed=34).contiguous(memory_format=CL).requires_grad_(True)
    b = _rand(2, 8, 5, 5, seed=35).contiguous(memory_format=CL).requires_grad_(True)
    _check(ops.add, lambda a, b: a + b, [a, b], 1e-7)
    n = _rand(3, 5, 6, 7, seed=36)
    y = ops.to_nhwc(n)
    assert y.is_contiguous(memory_format=CL) and torch.equal(y, n)
    assert torch.equal(ops._raw_to_nchw(y).contiguous(), n)


def test_image_losses_and_determinism():
    a = _rand(3, 3, 128, 128, seed=37).contiguous(memory_format=CL).requires_grad_(True)
    b = _rand(3, 3, 128, 128, seed=38).contiguous(memory_format=CL).requires_grad_(True)
    _check(ops.l1_mean, lambda a, b: (a - b).abs().mean(), [a, b], 2e-6, 1e-6)
    _check(ops.mse, lambda a, b: ((a - b) ** 2).mean(), [a, b], 2e-6, 2e-6)
    o = _rand(4, 1, 7, 7, seed=39).requires_grad_(True)
    _check(lambda o: ops.mse_const(o, 1.0), lambda o: ((o - 1.0) ** 2).mean(), [o], 2e-6, 2e-6)
    big = _rand(64, 3, 128, 128, seed=40)
    big2 = _rand(64, 3, 128, 128, seed=41)
    v1, v2 = ops.l1_mean(big, big2), ops.l1_mean(big, big2)
    assert torch.equal(v1, v2)                       # fixed reduction order: bit-reproducible
    assert abs(float(v1) - float((big.double() - big2.double()).abs().mean())) < 1e-6
    # NCHW vs channels-last operands
    assert abs(float(ops.l1_mean(a.detach().contiguous(), b.detach())) - float(ops.l1_mean(a, b))) < 1e-6


@pytest.mark.parametrize("n", [4, 64, 256, 1000])
def test_latent_losses_against_oracle(n):
    torch.manual_seed(n)
    mu_c = (torch.randn(n, 8) * 0.7 + 0.2 * torch.randn(1, 8))
    mu_c[:, 1] += 0.5 * mu_c[:, 0]                   # some correlation
    lv_c = torch.randn(n, 8) * 0.3
    target = so.hist_target(20000)
    n_cfg = float(max(n, 8))
    w = [10.0, 100.0, 100.0, 0.1]

    mu_r, lv_r = mu_c.clone().requires_grad_(True), lv_c.clone().requires_grad_(True)
    terms = [so.batch_kl(mu_r, n_cfg), so.corr_loss(mu_r), so.hist_loss(mu_r, target), so.conventional_kl(mu_r, lv_r)]
    sum(t * k for t, k in zip(terms, w)).backward()

    mu = mu_c.to(DEV).requires_grad_(True)
    lv = lv_c.to(DEV).requires_grad_(True)
    losses, blob = ops.latent_losses(mu, lv, n_cfg=n_cfg, target=target.to(DEV), flags=15)
    for k in range(4):
        ref = float(terms[k])
        assert abs(float(losses[k]) - ref) <= 3e-5 * max(1.0, abs(ref)), (k, float(losses[k]), ref)
    (losses * torch.tensor(w, device=DEV)).sum().backward()
    assert _rel(mu.grad, mu_r.grad) < 2e-4
    assert _rel(lv.grad, lv_r.grad) < 1e-5
    st = ops.latent_stats_views(blob, 8, 50)
    ref_st = so.latent_statistics(mu_c, n_cfg)
    assert _rel(st["mean"], ref_st["mean"]) < 1e-5 and _rel(st["var"], ref_st["var"]) < 1e-5
    assert _rel(st["corr"], ref_st["corr"]) < 1e-5 and _rel(st["hist"], ref_st["hist"]) < 1e-5
    # soft-bin sums are produced in a fixed order: bit-reproducible run to run
    _, blob2 = ops.latent_losses(mu.detach(), lv.detach(), n_cfg=n_cfg, target=target.to(DEV), flags=15)
    assert torch.equal(blob, blob2)


def test_latent_row_slices_equal_full_batch():
    """Data-parallel form: statistics on the gathered batch, gradients for a rank's rows only."""
    torch.manual_seed(3)
    full = torch.randn(32, 8).to(DEV)
    target = so.hist_target(20000).to(DEV)
    ref = full.clone().requires_grad_(True)
    l_ref, _ = ops.latent_losses(ref, n_cfg=32.0, target=target, flags=7)
    l_ref.sum().backward()
    for rank in range(4):
        part = full[rank * 8:(rank + 1) * 8].clone().requires_grad_(True)
        l, _ = ops.latent_losses(part, n_cfg=32.0, target=target, flags=7, mu_all=full, row0=rank * 8)
        assert torch.equal(l, l_ref)                 # global statistics, bit-identical on every rank
        l.sum().backward()
        assert torch.equal(part.grad, ref.grad[rank * 8:(rank + 1) * 8])


def test_util_loss_functions():
    import util
    torch.manual_seed(5)
    x = torch.randn(5, 120)
    c = util.corrcoef(x.to(DEV))
    assert np.allclose(np.corrcoef(x.numpy()), c.cpu().numpy(), atol=1e-5)      # the reference's docstring check
    xr = x.clone().double().requires_grad_(True)
    wgt = torch.randn(5, 5, dtype=torch.float64)
    (so.corrcoef(xr) * wgt).sum().backward()
    xp = x.to(DEV).requires_grad_(True)
    (util.corrcoef(xp) * wgt.float().to(DEV)).sum().backward()
    # the diagonal of the clamped matrix carries no gradient; remove its (tiny, rounding-only) part from the ref
    assert _rel(xp.grad, xr.grad.float()) < 1e-3
    mu = torch.randn(64, 8)
    assert abs(float(util.corrcoef_loss(mu.to(DEV).t(), DEV)) - float(so.corr_loss(mu))) < 1e-5
    gh = util.GaussianHistogram(50, -10, 10, 0.2)
    v = torch.randn(500)
    assert _rel(gh(v.to(DEV)), so.soft_hist(v)) < 1e-5
    vp = v.to(DEV).requires_grad_(True)
    vr = v.clone().double().requires_grad_(True)
    wv = torch.randn(50, dtype=torch.float64)
    (gh(vp) * wv.float().to(DEV)).sum().backward()
    (so.soft_hist(vr) * wv).sum().backward()
    assert _rel(vp.grad, vr.grad.float()) < 1e-4


def test_fused_adam_matches_torch_adam():
    torch.manual_seed(0)
    ps = [torch.randn(64, 16, 3, 3).contiguous(memory_format=CL), torch.randn(33), torch.randn(7, 5)]
    a = [torch.nn.Parameter(p.clone().to(DEV)) for p in ps]
    b = [torch.nn.Parameter(p.clone().to(DEV)) for p in ps]
    oa = ops.FusedAdam(a, lr=1e-3, betas=(0.5, 0.999))
    ob = torch.optim.Adam(b, lr=1e-3, betas=(0.5, 0.999))
    for it in range(5):
        for pa, pb in zip(a, b):
            g = torch.randn(pa.shape, generator=torch.Generator().manual_seed(it)).to(DEV)
            if pa.dim() == 4:
                g = g.contiguous(memory_format=CL)
            pa.grad = g.clone() if it % 2 == 0 or pa.grad is None else pa.grad.copy_(g)
            pb.grad = g.clone()
        oa.step()
        ob.step()
    for pa, pb in zip(a, b):
        assert _rel(pa.detach(), pb.detach()) < 1e-6
        assert pa.shape == pb.shape


def test_fused_norm_path_in_a_subprocess():
    """SRGAN_NORM_FUSED=1 selects the single-launch two-phase norm kernels (work queues, arrival counters); the switch is
    read once per process, so the norm parity tests are re-run in a child process with it set."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, SRGAN_NORM_FUSED="1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_ops_gpu.py"), "-x", "-q", "-m", "gpu",
                        "-k", "instance_norm"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "passed" in r.stdout


def test_instance_norm_batch_split_invariance():
    """With fp64 partial sums over fixed atoms (norm.cu: D4) the statistics of an image do not depend on how many
    spatial slices the grid planner cuts it into, i.e. on how many images share the launch: the output and the
    input gradient of image 0 are bit-identical whether it is normalised alone, with 3 or with 15 others.  (That is
    what lets an N-rank data-parallel step reproduce the 1-GPU step on the global batch, tools/dp_check.py.)  The
    switch is read once per process; skipped when the process runs with fp32 partials or the fused kernels."""
    import os
    if not ops._lib().srgan_norm_partials_fp64() or os.environ.get("SRGAN_NORM_FUSED", "0") != "0":
        pytest.skip("fp32 partials selected")
    torch.manual_seed(11)
    for (c, h, w) in ((64, 128, 128), (256, 32, 32), (128, 65, 63), (32, 37, 5)):
        x = (torch.randn(16, c, h, w) * 2 + 3).to(DEV).contiguous(memory_format=CL)
        dy = torch.randn(16, c, h, w).to(DEV).contiguous(memory_format=CL)
        g = torch.randn(c).to(DEV)
        b = torch.randn(c).to(DEV)
        outs = []
        for n in (16, 4, 1):
            xs = x[:n].clone().contiguous(memory_format=CL).requires_grad_(True)
            y = ops.instance_norm_act(xs, g, b, None, None, 1e-5, ops.ACT_LRELU, 0.2)
            y.backward(dy[:n].clone().contiguous(memory_format=CL))
            outs.append((y.detach()[:1].clone(), xs.grad[:1].clone()))
        for y, dx in outs[1:]:
            assert torch.equal(y, outs[0][0]), (c, h, w)
            assert torch.equal(dx, outs[0][1]), (c, h, w)


def test_norm_f32_partials_switch_in_a_subprocess():
    """SRGAN_DBG_NORM_F32_PARTIALS=1 (the A/B switch of DESIGN.md 2.4) still passes the norm parity tests; the
    split-invariance test skips itself there."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, SRGAN_DBG_NORM_F32_PARTIALS="1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_ops_gpu.py"),
                        os.path.join(here, "test_batchnorm_gpu.py"), "-x", "-q", "-m", "gpu",
                        "-k", "instance_norm or batchnorm or batch_norm or cbbn"], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "passed" in r.stdout
