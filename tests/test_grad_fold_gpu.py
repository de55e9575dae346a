"""GPU: later gradient contributions of one backward pass are parked in copies of the optimizer's flat gradient buffer
and folded by srgan_grad_fold (srgan_ops._grad_sink / fold_pending_grads).  A small generator applied THREE times
inside one loss (like phase 1 of the SRGAN step, ref pyfiles/util_notebook.py:619-665): the folded gradients must equal
autograd's own accumulation (same kernels, only the association of the sum differs), the parking copies must be left
zero, and a fourth contribution must fall back to autograd's add."""
import pytest
import torch

import cases
import srgan_ops as ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _grads(passes, pending):
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(0)
    G = model.SingleGenerator(3, 8, 2, 2, 2, "instance", num_con=12).to(DEV)
    opt = ops.FusedAdam(G.parameters(), lr=1e-4, betas=(0.5, 0.999))
    x = torch.rand(2, 3, 32, 32, device=DEV) * 2 - 1
    c = torch.randn(2, 12, device=DEV)
    prev = ops._MAX_PENDING
    ops._MAX_PENDING = pending
    try:
        opt.zero_grad()
        y, loss = x, 0
        for i in range(passes):
            y = G(y, c)
            loss = loss + (y * (i + 1)).abs().mean()
        with ops.direct_param_grads():
            loss.backward()
        torch.cuda.synchronize()
        st = opt._flat[0]
        flat = st["g"].clone()
        for p in st.get("pend", []):
            assert float(p.abs().max()) == 0.0                    # folded and zeroed
        assert not ops._pending_states
    finally:
        ops._MAX_PENDING = prev
    return flat, {n: p.grad.detach().clone() for n, p in G.named_parameters()}


@pytest.mark.parametrize("passes", [2, 3, 4])
def test_folded_gradients_equal_autograd_accumulation(passes):
    f0, g0 = _grads(passes, 0)          # every later contribution through autograd's add
    f2, g2 = _grads(passes, 2)          # up to two parked contributions, the rest through autograd
    assert _rel(f2, f0) < 2e-6, (passes, _rel(f2, f0))
    for k in g0:
        assert g2[k].shape == g0[k].shape and _rel(g2[k], g0[k]) < 5e-6, k
    f2b, _ = _grads(passes, 2)
    assert torch.equal(f2, f2b)         # reproducible


def test_grad_fold_kernel():
    n = 4096 + 64
    g, p1, p2 = (torch.randn(n, device=DEV) for _ in range(3))
    want = (g + p1) + p2
    ops._call("srgan_grad_fold", ops._p(g), ops._p(p1), ops._p(p2), n, ops._stream())
    assert torch.equal(g, want) and float(p1.abs().max()) == 0 and float(p2.abs().max()) == 0
    g, p1 = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
    want = g + p1
    ops._call("srgan_grad_fold", ops._p(g), ops._p(p1), None, n, ops._stream())
    assert torch.equal(g, want) and float(p1.abs().max()) == 0
