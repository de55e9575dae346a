"""GPU: the batch-statistics norms (CBBNorm2d, BatchNorm2d of norm_type="batch") through the C ABI against the CPU
oracle restatement of the reference (oracle.cbbn, pinned by tests/golden/cbbnorm.npz) and fp64 PyTorch."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases
import srgan_ops as ops
import srgan_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda"
CL = torch.channels_last


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_cbbnorm_module_matches_reference_golden():
    """Product CBBNorm2d on the inputs recorded from the reference: outputs, all gradients, running statistics over
    two training calls and one evaluation call (tolerance 1e-4: fp32, different summation order)."""
    model, _, _ = cases.use_product_modules()
    g = np.load(os.path.join(cases.GOLDEN, "cbbnorm.npz"))
    m = model.CBBNorm2d(g["weight"].shape[0], num_con=g["lin_w"].shape[1]).to(DEV)
    with torch.no_grad():
        m.weight.copy_(torch.tensor(g["weight"])); m.bias.copy_(torch.tensor(g["bias"]))
        m.ConBias[0].weight.copy_(torch.tensor(g["lin_w"])); m.ConBias[0].bias.copy_(torch.tensor(g["lin_b"]))
    for call in range(3):
        pre = "call%d." % call
        m.train(bool(g[pre + "training"]))
        x = torch.tensor(g[pre + "x"]).to(DEV).requires_grad_(True)
        con = torch.tensor(g[pre + "con"]).to(DEV).requires_grad_(True)
        y = m(x, con)
        m.zero_grad()
        (y * torch.tensor(g[pre + "probe"]).to(DEV)).sum().backward()
        assert _rel(y, torch.tensor(g[pre + "y"])) < 1e-5, (call, _rel(y, torch.tensor(g[pre + "y"])))
        got = {"dx": x.grad, "dcon": con.grad, "dweight": m.weight.grad, "dbias": m.bias.grad,
               "dlin_w": m.ConBias[0].weight.grad, "dlin_b": m.ConBias[0].bias.grad}
        for k, v in got.items():
            assert _rel(v, torch.tensor(g[pre + k])) < 1e-4, (call, k, _rel(v, torch.tensor(g[pre + k])))
        assert _rel(m.running_mean, torch.tensor(g[pre + "running_mean"])) < 1e-5
        assert _rel(m.running_var, torch.tensor(g[pre + "running_var"])) < 1e-5
    assert int(m.num_batches_tracked) == 2


@pytest.mark.parametrize("shape", [(4, 64, 32, 32), (3, 256, 15, 15), (70, 32, 8, 8), (2, 24, 10, 10)])
def test_cbbnorm_op_against_oracle(shape):
    N, C, H, W = shape
    gen = torch.Generator().manual_seed(5 + N + C)
    x = (torch.randn(N, C, H, W, generator=gen) * 1.5 + 0.4)
    con = torch.randn(N, 12, generator=gen)
    w, b = torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen) * 0.2
    lw, lb = torch.randn(C, 12, generator=gen) * 0.3, torch.randn(C, generator=gen) * 0.1
    probe = torch.randn(N, C, H, W, generator=gen)
    ref_in = [t.double().requires_grad_(True) for t in (x, con, w, b, lw, lb)]
    yr, rm_r, rv_r = so.cbbn(*ref_in, torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64))
    gr = torch.autograd.grad((F.leaky_relu(yr, 0.2) * probe.double()).sum(), ref_in)
    prod_in = [t.to(DEV).requires_grad_(True) for t in (x, con, w, b, lw, lb)]
    px, pcon, pw, pb, plw, plb = prod_in
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    t = ops.cond_bias(pcon, plw, plb)
    y = ops.batch_norm_act(px.contiguous(memory_format=CL), pw, pb, t, None, rm, rv, True, 0.1, 1e-5, True,
                           ops.ACT_LRELU, 0.2)
    gp = torch.autograd.grad((y * probe.to(DEV)).sum(), prod_in)
    assert _rel(y, F.leaky_relu(yr, 0.2)) < 3e-6
    for name, a, r in zip(("dx", "dcon", "dw", "db", "dlin_w", "dlin_b"), gp, gr):
        assert _rel(a, r) < 3e-4, (name, _rel(a, r))
    assert _rel(rm, rm_r) < 1e-5 and _rel(rv, rv_r) < 1e-5


def test_plain_batchnorm_training_and_eval_against_torch():
    model, _, _ = cases.use_product_modules()
    N, C, H, W = 5, 64, 16, 16
    gen = torch.Generator().manual_seed(11)
    ref = torch.nn.BatchNorm2d(C).double()
    prod = model._KernelBatchNorm2d(C).to(DEV)
    with torch.no_grad():
        ref.weight.copy_(torch.rand(C, generator=gen) + 0.5); ref.bias.copy_(torch.randn(C, generator=gen) * 0.1)
        prod.weight.copy_(ref.weight.float()); prod.bias.copy_(ref.bias.float())
    for training in (True, True, False):
        ref.train(training); prod.train(training)
        x = torch.randn(N, C, H, W, generator=gen) * 2 + 1
        probe = torch.randn(N, C, H, W, generator=gen)
        xr = x.double().requires_grad_(True)
        xp = x.to(DEV).contiguous(memory_format=CL).requires_grad_(True)
        yr, yp = F.relu(ref(xr)), prod(xp, act=ops.ACT_RELU)
        ref.zero_grad(); prod.zero_grad()
        (yr * probe.double()).sum().backward()
        (yp * probe.to(DEV)).sum().backward()
        assert _rel(yp, yr) < 3e-6
        assert _rel(xp.grad, xr.grad) < 3e-4, _rel(xp.grad, xr.grad)
        assert _rel(prod.weight.grad, ref.weight.grad) < 3e-4 and _rel(prod.bias.grad, ref.bias.grad) < 3e-4
        assert _rel(prod.running_mean, ref.running_mean) < 1e-5 and _rel(prod.running_var, ref.running_var) < 1e-5
    assert list(prod.state_dict().keys()) == list(ref.state_dict().keys())


def test_generator_with_batch_norm_type_runs_and_matches_torch_composition():
    """SingleGenerator(norm_type="batch") (CBBNorm2d + BatchNorm2d everywhere): forward + backward against the same
    network evaluated with plain PyTorch ops on the product's parameters (fp32 conv engine)."""
    model, _, _ = cases.use_product_modules()
    ops.set_conv_engine("fp32")
    try:
        torch.manual_seed(3)
        G = model.SingleGenerator(3, 8, 2, 2, 2, "batch", num_con=12).to(DEV)
        sd = {k: v.detach().double() for k, v in G.state_dict().items()}
        x = torch.rand(3, 3, 32, 32, device=DEV) * 2 - 1
        c = torch.randn(3, 12, device=DEV)
        y = G(x, c)
        y.square().mean().backward()

        def cb(pre, h):
            z = torch.zeros(h.shape[1], dtype=torch.float64, device=DEV)
            return so.cbbn(h, c.double(), sd[pre + "weight"], sd[pre + "bias"], sd[pre + "ConBias.0.weight"],
                           sd[pre + "ConBias.0.bias"], z, z + 1)[0]
        h = x.double()
        for i in range(3):
            h = F.conv2d(h, sd["down_convs.%d.weight" % i], stride=1 if i == 0 else 2, padding=3 if i == 0 else 1)
            h = F.relu(cb("down_cnorms.%d." % i, h))
        for bidx in range(2):
            p = "resBlocks.%d." % bidx
            r = F.relu(cb(p + "cn1.", F.conv2d(h, sd[p + "c1.weight"], padding=1)))
            h = cb(p + "cn2.", F.conv2d(r, sd[p + "c2.weight"], padding=1)) + h
        for i in range(2):
            h = F.conv_transpose2d(h, sd["up_convs.%d.weight" % i], stride=2, padding=1)
            h = F.relu(F.batch_norm(h, None, None, sd["up_norms.%d.weight" % i], sd["up_norms.%d.bias" % i], True))
        ref = torch.tanh(F.conv2d(h, sd["up_convs.2.weight"], padding=3))
        assert _rel(y, ref) < 2e-5, _rel(y, ref)
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in G.parameters())
    finally:
        ops.set_conv_engine("auto")
