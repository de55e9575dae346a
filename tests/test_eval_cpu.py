"""CPU: the f4 oracle (oracle/eval_oracle.py) against its pins.
  * notebook-04 classifier step against tests/golden/classifier.npz, recorded from the UNMODIFIED reference
    (Encoder_classifier + the training cell of notebook 04) by oracle/make_golden_eval.py;
  * PRDC: the integer-count formulation the CUDA kernels implement == the literal restatement of prdc==0.2 on
    sklearn distances == tests/golden/prdc.npz;
  * the evaluation module's import surface (names of the reference's pyfiles/evaluation.py)."""
import os

import numpy as np
import pytest
import torch

import cases
import eval_oracle as eo


def _golden(name):
    return dict(np.load(os.path.join(cases.GOLDEN, name + ".npz")))


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_classifier_oracle_matches_reference_golden():
    g = _golden("classifier")
    sd = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("init/")}
    torch.set_num_threads(8)
    loss, y, acc, grads, new = eo.classifier_step(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["label"]), lr=1e-4)
    assert abs(loss - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    assert _rel(y.numpy(), g["y"]) < 1e-6
    worst = 0.0
    for k, v in grads.items():
        worst = max(worst, _rel(v.numpy(), g["grad/" + k]))
    assert worst < 2e-4, worst                        # fp32 summation-order noise of the two autograd graphs
    for k, v in new.items():
        # the first Adam step moves every weight by ~lr * sign(g): compare the update, not the weight
        du, dr = v.numpy() - sd[k].numpy(), g["after/" + k] - g["init/" + k]
        assert np.abs(du - dr).max() <= 2.5e-5, k     # lr = 1e-4; entries with g ~ 0 may flip sign
        assert np.mean(np.abs(du - dr) > 1e-6) < 0.02, k


def test_prdc_counts_equal_literal_prdc_and_golden():
    g = _golden("prdc")
    for tag in ("a", "b"):
        real, fake, k = g[tag + "/real"], g[tag + "/fake"], int(g[tag + "/k"])
        cnt = eo.prdc_counts(real, fake, k)
        for q in ("col_hits_real", "row_hits_fake", "row_min_in"):
            assert np.array_equal(cnt[q], g[tag + "/" + q]), (tag, q)
        met = eo.metrics_from_counts(cnt, k)
        lit = eo.compute_prdc_literal(real, fake, k)
        for i, q in enumerate(("precision", "recall", "density", "coverage")):
            assert abs(met[q] - lit[q]) < 1e-12 and abs(met[q] - g[tag + "/metrics"][i]) < 1e-12, (tag, q)


@pytest.mark.parametrize("shape", [(50, 37, 16, 3, 11), (33, 64, 5, 1, 12), (20, 20, 64, 7, 13)])
def test_prdc_counts_equal_literal_prdc_random(shape):
    n, m, d, k, seed = shape
    real, fake = eo.synthetic_features(n, m, d, seed)
    lit = eo.compute_prdc_literal(real, fake, k)
    met = eo.compute_prdc(real, fake, k)
    for q in lit:
        assert abs(met[q] - lit[q]) < 1e-12, (shape, q, met, lit)


def test_prdc_edge_cases():
    # identical sets: every sample is inside every ball that contains its own position
    real, _ = eo.synthetic_features(24, 24, 6, 5, duplicates=0)
    m = eo.compute_prdc(real, real.copy(), 3)
    assert m["precision"] == 1.0 and m["recall"] == 1.0 and m["coverage"] == 1.0 and m["density"] >= 1.0
    # disjoint sets: nothing is covered
    far = real + 1000.0
    m = eo.compute_prdc(real, far, 3)
    assert m == dict(precision=0.0, recall=0.0, density=0.0, coverage=0.0)


def test_evaluation_import_surface():
    model, util, nb = cases.use_product_modules()
    import evaluation as ev
    for name in ("vgg_model", "GAN_evaluation", "evaluation_init", "compute_prdc"):
        assert hasattr(ev, name), name
    store = ev.evaluation_init(["vgg-initialization"], (0, 1), {"precision": None, "recall": None})
    assert store == {"vgg-initialization": {s: {t: {"precision": [], "recall": []} for t in (0, 1)} for s in (0, 1)}}
    store["vgg-initialization"][0][1]["precision"].append(1.0)           # independent lists
    assert store["vgg-initialization"][1][0]["precision"] == []
    assert hasattr(nb, "Classifier_training") and hasattr(nb, "do_test")
