"""Host-side helpers and loss functions of SingleGAN / Style-Restricted GAN, B200-native.

Drop-in for the reference's `pyfiles/util.py`: same function / class names and signatures.
The loss functions (section "Loss" at the bottom; ref pyfiles/util.py:457-553) run on the fused
latent-loss / reduction kernels of libsrgan_b200.so; everything above them is boundary glue that
runs on the host exactly like the reference's (label bookkeeping, image conversion, plotting).
matplotlib and prdc are imported lazily: neither is needed by the training step.
"""
import glob  # noqa: F401  (re-exported: the notebooks rely on `from util import *`)
import itertools  # noqa: F401
import os  # noqa: F401
import pickle
import shutil  # noqa: F401

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (re-exported for notebooks that rely on `from util import *`)

import srgan_ops as ops


def compute_prdc(*args, **kwargs):
    """Lazy forwarder to `prdc.compute_prdc` (evaluation only; not part of the training step)."""
    from prdc import compute_prdc as _impl
    return _impl(*args, **kwargs)


# ------------------------------------------------------------------------------- tensors / files
def cuda2numpy(x):
    """Device (or CPU) tensor -> detached numpy array."""
    return x.detach().to("cpu").numpy()


def cuda2cpu(x):
    """Device tensor -> detached CPU tensor."""
    return x.detach().to("cpu")


def pickle_save(data, path):
    with open(path, "wb") as f:
        pickle.dump(data, f)


def pickle_load(path):
    with open(path, mode="rb") as f:
        return pickle.load(f)


def min_max(x, axis=None, mean0=False, get_param=False):
    lo = x.min(axis=axis, keepdims=True)
    hi = x.max(axis=axis, keepdims=True)
    out = (x - lo) / (hi - lo + 1e-8)
    if mean0:
        out = out * 2 - 1
    return (out, lo, hi) if get_param else out


class MinMax(object):
    """Transform: rescale an image tensor to [0,1] (or [-1,1] when mean0)."""

    def __init__(self, mean0=True):
        self.mean0 = mean0

    def __call__(self, img):
        return torch.Tensor(min_max(cuda2numpy(img), mean0=self.mean0))

    def __repr__(self):
        return self.__class__.__name__


def image_from_output(output):
    """[N,C,H,W] (or [C,H,W]) tensor -> list of PIL images, each min-max stretched to 8 bit."""
    from PIL import Image
    if len(output.shape) == 3:
        output = output.unsqueeze(0)
    arr = cuda2numpy(output)
    images = []
    for a in arr:
        a = np.transpose(a, (1, 2, 0))
        a = np.tile(a, (1, 1, int(3 / a.shape[2])))
        a = min_max(a) * 2 ** 8
        a[a > 255] = 255
        images.append(Image.fromarray(np.uint8(a)))
    return images


class ToPIL(object):
    def __call__(self, img):
        return image_from_output(torch.reshape(img, (1,) + tuple(img.shape)))[0]

    def __repr__(self):
        return self.__class__.__name__


def weights_init(m):
    """Kept for the notebooks' `net.apply(weights_init)`.  Exactly like the reference (pyfiles/util.py:193-203)
    it matches LOWER-CASE substrings against class names such as `Conv2d`, so it never fires: the effective
    initialisation is PyTorch's default."""
    classname = m.__class__.__name__
    for key in ("conv", "linear", "batchnorm"):
        if classname.find(key) != -1:
            m.weight.data.normal_(0.0, 0.02)
            m.bias.data.fill_(0)
            break


_ref_tables = {}


def class_encode(label, device, ref_class):
    """Rows of `ref_class` (usually one-hot) selected by `label` -> float32 [len(label), dim] on `device`.
    ref pyfiles/util.py:205-234.  The lookup table is cached per device and indexed there, so a CUDA
    `label` works (the reference indexes a CPU table, which modern PyTorch rejects for CUDA indices)."""
    ref_class = np.asarray(ref_class)
    dev = torch.device(device)
    key = (ref_class.shape, ref_class.tobytes(), str(dev))
    table = _ref_tables.get(key)
    if table is None:
        table = torch.tensor(ref_class, dtype=torch.float32).to(dev)
        _ref_tables[key] = table
    idx = ops.to_device_async(label, dev, torch.long)
    return table[idx].view(-1, ref_class.shape[1])


def load_classifier(net, classifier_path, device):
    """Copy a pretrained classifier checkpoint into the encoder (missing fcmean/fcvar keys are expected)."""
    state = torch.load(classifier_path, map_location=device)
    print(net.load_state_dict(state, strict=False))
    return net


def get_target(label, classes, to_tensor=False, to_cuda=False, whole=False, shuffle=True):
    """For every source label, the other class ids (or all of them when `whole`), each row shuffled with
    the NumPy global RNG -- same draws, in the same order, as the reference (pyfiles/util.py:268-319)."""
    if hasattr(label, "detach"):
        label = label.to("cpu").detach().numpy()
    label = np.asarray(label)
    n_cls = len(classes)
    grid = np.tile(np.arange(n_cls), (label.shape[0], 1))
    if whole:
        target = grid
    else:
        others = np.array(1 - np.eye(n_cls)[label], dtype=bool)
        target = np.reshape(grid[others], (-1, n_cls - 1))
    if shuffle:
        for i in range(target.shape[0]):
            np.random.shuffle(target[i, :])
    if to_tensor:
        target = torch.Tensor(target)
        if to_cuda:
            target = target.to("cuda")
    return target


# ------------------------------------------------------------------------------- out of scope (evaluation)
def _out_of_scope(name):
    def stub(*args, **kwargs):
        raise NotImplementedError(
            "%s is an evaluation / plotting helper outside the training step this package replaces (SURVEY.md 2: out "
            "of scope); import it from the reference's pyfiles/util.py" % name)
    stub.__name__ = name
    return stub


# the names stay importable because the notebooks do `from util import *`
get_random_dataset = _out_of_scope("get_random_dataset")
plot_correlation_matrix = _out_of_scope("plot_correlation_matrix")
save_gif = _out_of_scope("save_gif")
plot_confusion_matrix = _out_of_scope("plot_confusion_matrix")


# =============================================================================== Loss
def _is_plain_mse(criterion):
    return isinstance(criterion, nn.MSELoss) and criterion.reduction == "mean"


def get_loss_D(outputs, target, criterion, device="cuda"):
    """Mean over the discriminator scales of criterion(output, constant target) (LSGAN when criterion is
    nn.MSELoss).  ref pyfiles/util.py:457-462."""
    loss = 0.0
    for output in outputs:
        if _is_plain_mse(criterion):
            loss = loss + ops.mse_const(output, target)           # fused (x - target)^2 mean, no target tensor
        else:
            loss = loss + criterion(output, torch.full(output.shape, target, device=output.device))
    return loss / len(outputs)


def get_domainloss_D(outputs_class, true_label, criterion_class):
    """Mean over scales of criterion_class(class probabilities, reference label).  ref pyfiles/util.py:464-468."""
    loss = 0.0
    for output_class in outputs_class:
        if _is_plain_mse(criterion_class):
            loss = loss + ops.mse(output_class, true_label)
        else:
            loss = loss + criterion_class(output_class, true_label)
    return loss / len(outputs_class)


def corrcoef(x):
    """Row-wise correlation matrix, like `np.corrcoef(x)`: x [D, n] -> [D, D], clamped to [-1, 1].
    ref pyfiles/util.py:470-511 (and its docstring example against NumPy)."""
    return ops.corrcoef(x)


def corrcoef_loss(m, device):
    """sum |corrcoef(m) - I| / (D (D-1)); ref pyfiles/util.py:513-517.  One fused kernel."""
    losses, _ = ops.latent_losses(m.t(), flags=ops.LAT_CORR)
    return losses[1]


class GaussianHistogram(nn.Module):
    """Differentiable (Gaussian-kernel) histogram; ref pyfiles/util.py:521-537."""

    def __init__(self, bins, min, max, sigma):
        super().__init__()
        self.bins, self.min, self.max, self.sigma = bins, min, max, sigma
        self.delta = float(max - min) / float(bins)
        self.centers = float(min) + self.delta * (torch.arange(bins).float() + 0.5)

    def forward(self, x):
        self.centers = self.centers.to(x.device)
        return ops.soft_histogram(x, self.bins, self.min, self.max, self.sigma)


class histogram_imitation():
    """KL(target || soft-histogram of each latent dimension), summed over dimensions.
    ref pyfiles/util.py:539-553.  The target is the soft histogram of `target_num` standard-normal samples
    drawn from the CPU default generator at construction time (same RNG consumption as the reference)."""

    def __init__(self, device, bins=50, range_max=10, sigma=0.2, target_num=100000):
        self.device = device
        self.gausshist = GaussianHistogram(bins=bins, min=-range_max, max=range_max, sigma=sigma)
        samples = torch.randn(target_num, 1)
        hist = self.gausshist(samples[:, 0].to(device))
        self.target = (hist / hist.sum() + 1e-8).detach().contiguous()

    def loss(self, x):
        g = self.gausshist
        losses, _ = ops.latent_losses(x, target=self.target, bins=g.bins, hmin=g.min, hmax=g.max, sigma=g.sigma,
                                      flags=ops.LAT_HIST)
        return losses[2]
