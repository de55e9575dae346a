"""CPU: the oracle restatement (oracle/srgan_oracle.py) against the golden vectors that
oracle/make_golden.py recorded from the UNMODIFIED reference.  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

import cases
import srgan_oracle as so

SMALL = ["srgan_small", "single_solo_small", "single_multi_small", "srgan_frozen_small"]


def _golden(name):
    return dict(np.load(os.path.join(cases.GOLDEN, name + ".npz")))


def _run_oracle(name):
    """Replay a case with the oracle, starting from the PRODUCT modules' seeded default init."""
    c = cases.CASES[name]
    model, util, _ = cases.use_product_modules()
    torch.set_num_threads(8)
    torch.manual_seed(c["seed"])
    np.random.seed(c["seed"])
    nets = cases.build_nets(model, name)
    sds = cases.state_dicts(nets)
    tr = cases.build_oracle(name, sds, so)          # draws the histogram target (CPU RNG), like the trainer
    x, label = cases.synthetic_batch(c["batch"], util.get_target)
    tr.record = {}
    torch.manual_seed(c["seed"] + 1000)
    errs = tr.train(x, label)
    return c, sds, tr, x, label, errs


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference_golden(name):
    g = _golden(name)
    c, sds, tr, x, label, errs = _run_oracle(name)

    # same synthetic inputs
    assert np.array_equal(g["label.source"], label["source"].numpy())
    assert np.array_equal(g["label.target"], label["target"].numpy())
    assert _rel(so.digest(x).numpy(), g["x.digest"]) == 0.0

    # product modules: same seed -> bit-identical initial weights as the reference's modules
    g_sd, d_sd, e_sd = sds
    nets = {"G": g_sd, "E": e_sd}
    if isinstance(d_sd, list):
        nets.update({"Dc%d" % i: d for i, d in enumerate(d_sd)})
    else:
        nets["D"] = d_sd
    n_init = 0
    for net, sd in nets.items():
        for k, v in sd.items():
            assert np.array_equal(so.digest(v).numpy(), g["init.%s.%s" % (net, k)]), (net, k)
            n_init += 1
    assert n_init == sum(1 for k in g if k.startswith("init."))

    # losses returned by train(): fp32, same ATen kernels in the same order -> tight
    for got, ref in zip(errs, g["errs"]):
        assert abs(float(got) - ref) <= 2e-5 * max(1.0, abs(ref)), (float(got), ref)

    # first encoder pass of update_GandE
    assert _rel(tr.stats["mu"].numpy(), g["enc.mu"]) < 1e-5
    assert _rel(tr.stats["logvar"].numpy(), g["enc.logvar"]) < 1e-5

    # gradients at every optimizer step and weights after it (digests: norm, probe dot, 32 samples).
    # G gradients are ill-conditioned (L1 sign flips): looser bound for phase 2.
    checked = 0
    agg = {}
    for key, ref in g.items():
        if "." not in key or key.split(".")[0] in ("init", "enc", "stats", "ref", "label", "x") \
                or key in ("errs", "hist_target"):
            continue
        step, kind, pname = key.split(".", 2)
        assert step + "." + kind in tr.record, key
        got = tr.record[step + "." + kind][pname]
        if ref.size == 0:
            assert got is None, key
            continue
        d = so.digest(got).numpy()
        scale = max(abs(ref[0]), 1e-12)          # digest[0] = L2 norm of the tensor
        if step in ("G1", "E_final"):
            # phase 2 is chaotic under Adam (two runs of the reference itself differ by 3.6e-2 here, SURVEY
            # F12); it is pinned tightly, with teacher forcing, in test_oracle_vs_reference_inprocess.py
            if kind == "grad":
                assert abs(d[0] - ref[0]) <= 0.15 * scale, (key, d[0], ref[0])
            continue
        # G (and E, which is fed through G's L1 losses) gradients carry the L1-sign noise floor of SURVEY F12:
        # loose per parameter, tight on the whole step (vector of per-parameter norms + all samples)
        tol = 2e-2 if step in ("G0", "E0") else 1e-3
        assert abs(d[0] - ref[0]) <= tol * scale, (key, d[0], ref[0])
        assert np.max(np.abs(d[2:] - ref[2:])) <= 2 * tol * max(np.max(np.abs(ref[2:])), scale * 1e-2), key
        a = agg.setdefault(step + "." + kind, [[], []])
        a[0].append(d)
        a[1].append(ref)
        checked += 1
    for stepkind, (got, ref) in agg.items():
        assert _rel(np.concatenate(got), np.concatenate(ref)) < 3e-3, stepkind
    assert checked > 50


@pytest.mark.parametrize("name", ["srgan_small", "single_solo_small"])
def test_oracle_latent_functions_match_reference_functions(name):
    """corrcoef / corrcoef_loss / GaussianHistogram / histogram_imitation.loss of the REFERENCE (stored as
    ref.*) against the oracle's formulas on the same mu."""
    g = _golden(name)
    mu = torch.from_numpy(g["enc.mu"])
    target = torch.from_numpy(g["hist_target"])
    assert np.allclose(so.corrcoef(mu.t()).numpy(), g["ref.corr"], atol=1e-6)
    assert abs(float(so.corr_loss(mu)) - float(g["ref.corr_loss"])) < 1e-6
    hist = torch.stack([so.soft_hist(mu[:, d]) for d in range(mu.shape[1])]).numpy()
    assert np.allclose(hist, g["ref.hist"], rtol=1e-5, atol=1e-7)
    assert abs(float(so.hist_loss(mu, target)) - float(g["ref.hist_loss"])) < 1e-4 * abs(float(g["ref.hist_loss"]))


def test_corrcoef_matches_numpy():
    """The reference pins corrcoef to NumPy in its docstring example (ref util.py:488-494)."""
    rng = np.random.RandomState(0)
    x = rng.randn(5, 120)
    assert np.allclose(np.corrcoef(x), so.corrcoef(torch.from_numpy(x)).numpy())


def test_oracle_full_width_losses():
    """Full-width (nch 64) SRGAN, batch 2: losses and encoder output against the reference golden."""
    g = _golden("srgan_full")
    c, sds, tr, x, label, errs = _run_oracle("srgan_full")
    for got, ref in zip(errs, g["errs"]):
        assert abs(float(got) - ref) <= 2e-5 * max(1.0, abs(ref)), (float(got), ref)
    assert _rel(tr.stats["mu"].numpy(), g["enc.mu"]) < 1e-5
    for pname in ("resBlocks.0.c1.weight", "down_convs.0.weight", "up_convs.2.weight"):
        ref = g["G0.grad." + pname]
        d = so.digest(tr.record["G0.grad"][pname]).numpy()
        assert abs(d[0] - ref[0]) <= 2e-3 * abs(ref[0]), pname


def test_oracle_cbbnorm_matches_reference_golden():
    """oracle.cbbn (restated _CBBNorm.forward, ref pyfiles/model.py:121-148) against outputs, gradients and running
    statistics recorded from the unmodified reference (oracle/make_golden_norms.py): two training calls, one eval."""
    g = np.load(os.path.join(cases.GOLDEN, "cbbnorm.npz"))
    w, b = torch.tensor(g["weight"], requires_grad=True), torch.tensor(g["bias"], requires_grad=True)
    lw, lb = torch.tensor(g["lin_w"], requires_grad=True), torch.tensor(g["lin_b"], requires_grad=True)
    rm, rv = torch.zeros(w.shape[0]), torch.ones(w.shape[0])
    for call in range(3):
        pre = "call%d." % call
        x = torch.tensor(g[pre + "x"], requires_grad=True)
        con = torch.tensor(g[pre + "con"], requires_grad=True)
        y, rm, rv = so.cbbn(x, con, w, b, lw, lb, rm, rv, training=bool(g[pre + "training"]))
        grads = torch.autograd.grad((y * torch.tensor(g[pre + "probe"])).sum(), [x, con, w, b, lw, lb])
        np.testing.assert_allclose(y.detach().numpy(), g[pre + "y"], rtol=2e-5, atol=2e-5)
        for got, key in zip(grads, ("dx", "dcon", "dweight", "dbias", "dlin_w", "dlin_b")):
            ref = g[pre + key]
            assert np.abs(got.numpy() - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max()), key
        np.testing.assert_allclose(rm.numpy(), g[pre + "running_mean"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rv.numpy(), g[pre + "running_var"], rtol=1e-5, atol=1e-6)
