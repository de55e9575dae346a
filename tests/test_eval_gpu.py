"""GPU: notebook-04 classifier job and PRDC evaluation through the C ABI against the CPU oracle.
  * cross entropy: against torch in fp64;
  * PRDC: the integer counts of the kernels equal the oracle's exactly (golden features, random features with exact
    duplicates, ragged sizes), metrics equal the literal prdc restatement;
  * one training iteration of notebook 04 (`Classifier_training.train_step`) against tests/golden/classifier.npz
    recorded from the unmodified reference: fp32 engine tight, TF32 / bf16 engines at their stated tolerances."""
import os
import time

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases
import eval_oracle as eo
import srgan_ops as ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _golden(name):
    return dict(np.load(os.path.join(cases.GOLDEN, name + ".npz")))


def _rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("nj", [(6, 4), (512, 4), (37, 10), (3, 33)])
def test_cross_entropy_matches_torch(nj):
    N, J = nj
    torch.manual_seed(N)
    x = torch.randn(N, J, device=DEV).requires_grad_(True)
    label = torch.randint(0, J, (N,), device=DEV)
    loss = ops.cross_entropy(x, label)
    gx, = torch.autograd.grad(loss * 1.7, x)
    xr = x.detach().double().requires_grad_(True)
    ref = F.cross_entropy(xr, label)
    gr, = torch.autograd.grad(ref * 1.7, xr)
    assert abs(float(loss.detach()) - float(ref.detach())) < 2e-6 * max(1.0, abs(float(ref)))
    assert _rel(gx, gr) < 2e-6
    assert torch.equal(ops.cross_entropy(x, label), loss)            # fixed-order mean: bit-reproducible
    with pytest.raises(ops.SrganKernelError):
        ops.cross_entropy(x, label.int())


def _check_counts(real, fake, k):
    want = eo.prdc_counts(real, fake, k)
    got = ops.prdc_counts(torch.from_numpy(real).to(DEV), torch.from_numpy(fake).to(DEV), k)
    for q in ("col_hits_real", "row_hits_fake", "row_min_in"):
        assert np.array_equal(got[q].cpu().numpy().astype(np.int64), want[q]), q
    for q in ("r2_real", "r2_fake"):
        assert np.allclose(got[q].cpu().numpy(), want[q], rtol=1e-12, atol=0), q
    met = ops.compute_prdc(real, fake, k)
    ref = eo.metrics_from_counts(want, k)
    assert met == ref, (met, ref)
    return met


def test_prdc_counts_golden_bit_exact():
    g = _golden("prdc")
    for tag in ("a", "b"):
        met = _check_counts(g[tag + "/real"], g[tag + "/fake"], int(g[tag + "/k"]))
        for i, q in enumerate(("precision", "recall", "density", "coverage")):
            assert abs(met[q] - g[tag + "/metrics"][i]) < 1e-12, (tag, q)


@pytest.mark.parametrize("shape", [(300, 257, 100, 5, 21), (65, 64, 4096, 5, 22), (64, 129, 17, 1, 23),
                                   (130, 70, 3, 9, 24)])
def test_prdc_counts_random_bit_exact(shape):
    n, m, d, k, seed = shape
    real, fake = eo.synthetic_features(n, m, d, seed)
    met = _check_counts(real, fake, k)
    lit = eo.compute_prdc_literal(real, fake, k)
    for q in lit:
        assert abs(met[q] - lit[q]) < 1e-12, (shape, q)


def test_prdc_through_gan_evaluation_identity():
    model, util, nb = cases.use_product_modules()
    import evaluation as ev
    torch.manual_seed(0)
    true = torch.rand(40, 3, 8, 8) * 2 - 1
    pred = (true + 0.3 * torch.randn_like(true)).clamp(-1, 1)
    e = ev.GAN_evaluation("identity", device=DEV)
    met = e.get_prdc(true, pred, nearest_k=5, preprocess=True)
    ref = eo.compute_prdc(true.reshape(40, -1).numpy(), pred.reshape(40, -1).numpy(), 5)
    assert met == ref
    with pytest.raises(ValueError):
        ops.compute_prdc(true.reshape(40, -1), pred.reshape(40, -1), 40)


def _job(engine):
    model, util, nb = cases.use_product_modules()
    g = _golden("classifier")
    net = model.Encoder_classifier(3, 8, 8, 4, "instance", 4)
    net.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("init/")})
    net = net.to(DEV)
    ops.set_conv_engine(engine)
    try:
        job = nb.Classifier_training(net, lr=1e-4)
        loss, acc = job.train_step(torch.from_numpy(g["x"]), torch.from_numpy(g["label"]))
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().float().cpu() for n, p in net.named_parameters()}
        after = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        with torch.no_grad():
            net.eval()
    finally:
        ops.set_conv_engine("auto")
    return g, float(loss), float(acc), grads, after


# loss (relative), gradients (rel-L2 over all parameters): stated tolerance per engine; measured values are printed
# measured on B200: fp32 0 / 6.9e-7, TF32 2.3e-5 / 2.4e-2, bf16 trunk 2.3e-5 / 2.7e-2 -> bounds = 3 x measured
CLS_TOL = {"fp32": (2e-6, 3e-4), "auto": (1e-4, 7e-2), "bf16": (1e-4, 8e-2)}


@pytest.mark.parametrize("engine", ["fp32", "auto", "bf16"])
def test_classifier_training_step_matches_reference_golden(engine):
    g, loss, acc, grads, after = _job(engine)
    tol_l, tol_g = CLS_TOL[engine]
    ref_loss = float(g["loss"])
    num = sum(float((grads[k].double() - torch.from_numpy(g["grad/" + k]).double()).pow(2).sum()) for k in grads)
    den = sum(float(torch.from_numpy(g["grad/" + k]).double().pow(2).sum()) for k in grads)
    rel_g = (num / den) ** 0.5
    print("classifier step [%s]: loss rel err %.2e (tol %.0e), gradient rel-L2 %.2e (tol %.0e)"
          % (engine, abs(loss - ref_loss) / ref_loss, tol_l, rel_g, tol_g))
    assert abs(loss - ref_loss) <= tol_l * ref_loss
    assert rel_g < tol_g
    ref_acc = float((np.argmax(g["y"], axis=1) == g["label"]).mean())
    assert abs(acc - ref_acc) < 1e-6 or engine != "fp32"
    if engine == "fp32":
        for k, v in after.items():
            du = v.numpy() - g["init/" + k]
            dr = g["after/" + k] - g["init/" + k]
            assert np.abs(du - dr).max() <= 2.5e-5, k            # first Adam step: +-lr per weight
            assert np.mean(np.abs(du - dr) > 1e-6) < 0.02, k


def test_classifier_job_fit_and_rate():
    """`fit` over a tiny synthetic loader (loss decreases), do_test, and - not pass / fail - the rate of the
    notebook's configuration: Encoder_classifier(3, 8, 64, 4, "instance", 4), batch 512."""
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(0)
    net = model.Encoder_classifier(3, 8, 8, 4, "instance", 4).to(DEV)
    x = torch.rand(16, 3, 128, 128) * 2 - 1
    lab = torch.arange(16) % 4
    x += lab.view(-1, 1, 1, 1).float() * 0.2                      # separable classes
    loader = [(x[:8], lab[:8]), (x[8:], lab[8:])]
    job = nb.Classifier_training(net, lr=1e-3)
    losses, accs, val = job.fit(loader, 6, valloader=loader, test_interval=3)
    assert len(losses) == 6 and len(val) == 2 and losses[-1] < losses[0]
    labels, outputs = nb.do_test(net, loader, DEV, "eval")
    assert outputs.shape == (16, 4) and np.allclose(outputs.sum(axis=1), 1.0, atol=1e-5)
    assert abs(job.optimizer.param_groups[0]["lr"] - 1e-3 * 0.99 ** 6) < 1e-12

    ops.set_conv_engine("bf16")
    try:
        torch.manual_seed(1)
        net = model.Encoder_classifier(3, 8, 64, 4, "instance", 4).to(DEV)
        job = nb.Classifier_training(net, lr=1e-4)
        xb = (torch.rand(512, 3, 128, 128, device=DEV) * 2 - 1)
        lb = torch.randint(0, 4, (512,), device=DEV)
        for _ in range(3):
            job.train_step(xb, lb)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            job.train_step(xb, lb)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print("notebook-04 job, batch 512, bf16 trunk: %.1f ms/iteration = %.0f images/s" % (dt * 1e3, 512 / dt))
    finally:
        ops.set_conv_engine("auto")
