"""CPU oracle of the notebook-04 classifier step and of the PRDC evaluation -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product
(style-restricted_gan_b200/) never does.

Pinned:
  * classifier step: tests/golden/classifier.npz, recorded by oracle/make_golden_eval.py from the UNMODIFIED
    reference (Encoder_classifier of pyfiles/model.py:484-508 + the training cell of notebook 04);
  * PRDC: the dependency `prdc==0.2` (Docker/requirements.txt:13) is not vendored in /root/reference and not installed
    here, so its published algorithm (Naeem et al., ICML 2020; prdc/prdc.py of that release) is restated twice:
    `compute_prdc_literal` follows the package line by line on sklearn.metrics.pairwise_distances (float Euclidean
    distances, argpartition), `prdc_counts` is the integer formulation the CUDA kernels implement (fp64 squared
    distances); tests/test_eval_cpu.py checks that both give identical metrics, and tests/golden/prdc.npz holds
    features + counts recorded by oracle/make_golden_eval.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

import srgan_oracle as so


# ------------------------------------------------------------------------------------------------ notebook 04
def classifier_forward(sd, x, num_cls=4):
    """Encoder_classifier.forward: softmax(fcclass(pool(lrelu(trunk(x))))).  ref pyfiles/model.py:502-508 (F.softmax
    with its implicit dim, = 1 for a 2-D input)."""
    f = so._encoder_trunk(sd, x, num_cls)
    return F.softmax(F.linear(f, sd["fcclass.weight"], sd["fcclass.bias"]), dim=1)


def classifier_step(sd, x, label, lr=1e-4, num_cls=4):
    """One iteration of the training cell of notebook 04 (cell 22): zero_grad, y = net(x), loss =
    CrossEntropyLoss(y, label) (on the softmax outputs, as the reference does), backward, Adam(lr) step (default betas,
    notebook cell 18).  Returns (loss, probabilities, accuracy, gradients, updated weights)."""
    w = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    y = classifier_forward(w, x, num_cls)
    loss = F.cross_entropy(y, label)
    names = list(w.keys())
    grads = torch.autograd.grad(loss, [w[k] for k in names])
    new = {}
    b1, b2, eps = 0.9, 0.999, 1e-8
    for k, g in zip(names, grads):                  # first Adam step from zero moments (torch.optim.Adam defaults)
        m = (1 - b1) * g
        v = (1 - b2) * g * g
        mh = m / (1 - b1)
        vh = v / (1 - b2)
        new[k] = (w[k].detach() - lr * mh / (vh.sqrt() + eps)).clone()
    acc = float((y.detach().argmax(dim=1) == label).float().mean())
    return float(loss.detach()), y.detach(), acc, dict(zip(names, [g.detach() for g in grads])), new


# ------------------------------------------------------------------------------------------------ PRDC
def compute_prdc_literal(real_features, fake_features, nearest_k):
    """prdc.compute_prdc of prdc==0.2, line by line (sklearn pairwise Euclidean distances, k-th value by argpartition).
    ref call site: GAN_evaluation.get_prdc pyfiles/evaluation.py:98-110."""
    from sklearn.metrics import pairwise_distances

    def pair(x, y=None):
        return pairwise_distances(x, x if y is None else y, metric="euclidean")

    def kth_value(unsorted, k, axis=-1):
        idx = np.argpartition(unsorted, k, axis=axis)[..., :k]
        return np.take_along_axis(unsorted, idx, axis=axis).max(axis=axis)

    def nn_radii(f, k):
        return kth_value(pair(f), k=k + 1, axis=-1)

    r_real = nn_radii(real_features, nearest_k)
    r_fake = nn_radii(fake_features, nearest_k)
    d = pair(real_features, fake_features)
    precision = (d < np.expand_dims(r_real, axis=1)).any(axis=0).mean()
    recall = (d < np.expand_dims(r_fake, axis=0)).any(axis=1).mean()
    density = (1.0 / float(nearest_k)) * (d < np.expand_dims(r_real, axis=1)).sum(axis=0).mean()
    coverage = (d.min(axis=1) < r_real).mean()
    return dict(precision=float(precision), recall=float(recall), density=float(density), coverage=float(coverage))


def pairdist2(a, b):
    """Squared Euclidean distances in float64 from float32 features (differences formed in float64)."""
    a = np.asarray(a, dtype=np.float32).astype(np.float64)
    b = np.asarray(b, dtype=np.float32).astype(np.float64)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
    for i in range(a.shape[0]):                       # row blocks keep the temporary small
        df = a[i][None, :] - b
        out[i] = np.einsum("jk,jk->j", df, df)
    return out


def prdc_counts(real_features, fake_features, nearest_k):
    """The integer counts the CUDA kernels return (include/srgan_b200.h srgan_prdc_counts) and the squared radii."""
    k = int(nearest_k)
    r2 = []
    for f in (real_features, fake_features):
        d2 = pairdist2(f, f)
        r2.append(np.sort(d2, axis=1)[:, k])          # (k+1)-th smallest of the row, the zero self distance included
    d2 = pairdist2(real_features, fake_features)
    col = (d2 < r2[0][:, None]).sum(axis=0).astype(np.int64)
    row = (d2 < r2[1][None, :]).sum(axis=1).astype(np.int64)
    rmin = (d2.min(axis=1) < r2[0]).astype(np.int64)
    return dict(col_hits_real=col, row_hits_fake=row, row_min_in=rmin, r2_real=r2[0], r2_fake=r2[1])


def metrics_from_counts(c, nearest_k):
    col, row, rmin = c["col_hits_real"], c["row_hits_fake"], c["row_min_in"]
    return dict(precision=float((col > 0).mean()), recall=float((row > 0).mean()),
                density=float((1.0 / float(nearest_k)) * col.mean()), coverage=float(rmin.mean()))


def compute_prdc(real_features, fake_features, nearest_k):
    return metrics_from_counts(prdc_counts(real_features, fake_features, nearest_k), nearest_k)


def synthetic_features(n_real, n_fake, dim, seed=0, shift=0.6, duplicates=3):
    """Two overlapping Gaussian clouds with a few exact duplicates (ties in the k-NN selection) - float32."""
    g = np.random.RandomState(seed)
    real = g.randn(n_real, dim).astype(np.float32)
    fake = (g.randn(n_fake, dim) * 1.2 + shift / np.sqrt(dim)).astype(np.float32)
    for i in range(min(duplicates, n_real - 1)):
        real[-1 - i] = real[i]
    for i in range(min(duplicates, n_fake - 1, n_real)):
        fake[-1 - i] = real[i + 1]                    # fakes that coincide with a real sample: distance exactly 0
    return real, fake
