"""Host-side helpers and loss functions of SingleGAN / Style-Restricted GAN, B200-native.

Drop-in for the reference's `pyfiles/util.py`: same function / class names and signatures.
The loss functions (section "Loss" at the bottom; ref pyfiles/util.py:457-553) run on the fused
latent-loss / reduction kernels of libsrgan_b200.so; everything above them is boundary glue that
runs on the host exactly like the reference's (label bookkeeping, image conversion, plotting).
matplotlib and prdc are imported lazily: neither is needed by the training step.
"""
import glob
import itertools
import os
import pickle
import shutil

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (re-exported for notebooks that rely on `from util import *`)

import srgan_ops as ops


def _plt():
    import matplotlib.pyplot as plt
    return plt


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


def get_random_dataset(dataset, num, random=True, random_seed=0):
    if not random:
        np.random.seed(random_seed)
    index = np.random.choice(np.arange(len(dataset)), num, False)
    # (the reference draws `index` and then takes the first `num` items; kept)
    return torch.cat([dataset[i][0].unsqueeze(0) for i in range(len(index))], dim=0)


# ------------------------------------------------------------------------------- plotting (host only)
def plot_correlation_matrix(cm, save=False, save_dir="", save_name=""):
    plt = _plt()
    plt.figure(figsize=(10, 8))
    plt.imshow(cm, interpolation="nearest", cmap=plt.get_cmap("Blues"))
    plt.colorbar()
    half = cm.max() / 2
    for i, j in itertools.product(range(cm.shape[0]), range(cm.shape[1])):
        plt.text(j, i, str(round(cm[i, j], 4)), horizontalalignment="center", fontsize=12,
                 color="white" if cm[i, j] > half else "black")
    plt.tight_layout()
    if save:
        plt.savefig(fname=save_dir + save_name, format="png", bbox_inches="tight")
    plt.show()


def save_gif(data_list, gif_path, title, save_dir="contempolary_images/", fig_size=(8, 8), font_title=24,
             duration=100):
    from PIL import Image
    plt = _plt()
    shutil.rmtree(save_dir, ignore_errors=True)
    os.makedirs(save_dir, exist_ok=True)
    for i, frame in enumerate(data_list):
        fig = plt.figure(figsize=fig_size)
        ax = fig.add_subplot(1, 1, 1)
        ax.imshow(frame)
        plt.title(title, fontsize=font_title)
        plt.tick_params(labelbottom=False, labelleft=False, labelright=False, labeltop=False)
        plt.savefig(save_dir + f"{str(i).zfill(3)}.png", dpi=64, facecolor="lightgray", bbox_inches="tight",
                    format="png")
        plt.close()
    frames = [Image.open(f) for f in sorted(glob.glob(save_dir + "*.png"))]
    frames[0].save(gif_path, save_all=True, append_images=frames[1:], duration=duration, loop=0)
    shutil.rmtree(save_dir, ignore_errors=True)


def plot_confusion_matrix(cm, target_names, title="Confusion matrix", cmap=None, normalize=True):
    plt = _plt()
    accuracy = np.trace(cm) / float(np.sum(cm))
    plt.figure(figsize=(10, 8))
    plt.imshow(cm, interpolation="nearest", cmap=cmap if cmap is not None else plt.get_cmap("Blues"))
    plt.title(title)
    plt.colorbar()
    if target_names is not None:
        ticks = np.arange(len(target_names))
        plt.xticks(ticks, target_names, rotation=45)
        plt.yticks(ticks, target_names)
    if normalize:
        cm = cm.astype("float") / cm.sum(axis=1)[:, np.newaxis]
    thresh = cm.max() / 1.5 if normalize else cm.max() / 2
    fmt = "{:0.4f}" if normalize else "{:,}"
    for i, j in itertools.product(range(cm.shape[0]), range(cm.shape[1])):
        plt.text(j, i, fmt.format(cm[i, j]), horizontalalignment="center",
                 color="white" if cm[i, j] > thresh else "black")
    plt.tight_layout()
    plt.ylabel("True label")
    plt.xlabel("Predicted label\naccuracy={:0.4f}; misclass={:0.4f}".format(accuracy, 1 - accuracy))
    plt.show()


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
