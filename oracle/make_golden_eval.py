"""Generate tests/golden/classifier.npz and tests/golden/prdc.npz -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):   python oracle/make_golden_eval.py

classifier.npz: the UNMODIFIED reference's Encoder_classifier (pyfiles/model.py:484-508) driven exactly like the
training cell of notebook 04 (cells 18 and 22): net.apply(weights_init), criterion = nn.CrossEntropyLoss(),
optimizer = optim.Adam(net.parameters(), lr), one iteration (zero_grad, forward, loss, backward, step) on a seeded
synthetic batch; stored: initial weights, batch, labels, loss, output probabilities, every gradient, every weight after
the step.
prdc.npz: seeded synthetic features, the metrics of the literal restatement of prdc==0.2 (sklearn distances) and the
integer counts of the formulation the kernels implement (oracle/eval_oracle.py).
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import eval_oracle  # noqa: E402
import ref_harness  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
CLS = dict(nch=8, num_cls=4, ndim=8, classes=4, batch=6, lr=1e-4, seed=7)


def classifier_golden():
    ref_model, ref_util, _ = ref_harness.load_reference()
    c = CLS
    torch.set_num_threads(8)
    torch.manual_seed(c["seed"])
    np.random.seed(c["seed"])
    net = ref_model.Encoder_classifier(3, c["ndim"], c["nch"], c["num_cls"], "instance", c["classes"])
    net.apply(ref_util.weights_init)                  # notebook 04 cell 18 (matches no class name: default init kept)
    init = {k: v.detach().clone() for k, v in net.state_dict().items()}
    criterion = torch.nn.CrossEntropyLoss()
    optimizer = torch.optim.Adam(net.parameters(), lr=c["lr"])
    g = torch.Generator().manual_seed(c["seed"] + 1)
    x = torch.rand(c["batch"], 3, 128, 128, generator=g) * 2 - 1
    label = torch.randint(0, c["classes"], (c["batch"],), generator=g)
    net.train()
    optimizer.zero_grad()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")               # F.softmax without dim (the reference's call)
        y = net(x)
    loss = criterion(y, label)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    optimizer.step()
    after = {k: v.detach().clone() for k, v in net.state_dict().items()}
    out = {"x": x.numpy(), "label": label.numpy(), "loss": np.float32(loss.item()), "y": y.detach().numpy()}
    for k, v in init.items():
        out["init/" + k] = v.numpy()
    for k, v in grads.items():
        out["grad/" + k] = v.numpy()
    for k, v in after.items():
        out["after/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "classifier.npz"), **out)
    print("classifier.npz: loss %.6f, %d tensors" % (loss.item(), len(init)))


def prdc_golden():
    out = {}
    for tag, (n, m, d, k, seed) in {"a": (96, 80, 24, 5, 3), "b": (40, 64, 8, 3, 4)}.items():
        real, fake = eval_oracle.synthetic_features(n, m, d, seed)
        lit = eval_oracle.compute_prdc_literal(real, fake, k)
        cnt = eval_oracle.prdc_counts(real, fake, k)
        met = eval_oracle.metrics_from_counts(cnt, k)
        assert all(abs(lit[q] - met[q]) < 1e-12 for q in lit), (tag, lit, met)
        out[tag + "/real"], out[tag + "/fake"], out[tag + "/k"] = real, fake, np.int64(k)
        for q in ("col_hits_real", "row_hits_fake", "row_min_in"):
            out[tag + "/" + q] = cnt[q]
        out[tag + "/metrics"] = np.array([lit[q] for q in ("precision", "recall", "density", "coverage")])
        print("prdc.npz case %s:" % tag, lit)
    np.savez_compressed(os.path.join(GOLDEN, "prdc.npz"), **out)


if __name__ == "__main__":
    classifier_golden()
    prdc_golden()
