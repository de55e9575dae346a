"""CelebA file listing for the training notebooks (`from dataset import get_class_label, FaceDataset`, notebook cell 1).
Host-side only: it decides WHICH images form the train / val / test splits of each class and opens them with PIL; the
pixels reach the GPU through the notebook's own DataLoader.  Mirrors ref pyfiles/dataset.py:11-141 (same constructor,
same ordering rules) without its removed-NumPy-alias dependency (`np.int`, F11)."""
import glob
import itertools

import numpy as np
import torch.utils.data

from util import *  # noqa: F401,F403  (the notebooks rely on the transitive re-export; pickle_load comes from here)


def get_class_label(n_class_type):
    """All sign patterns of `n_class_type` binary attributes, (1, 1, ..) first: the class index is the position in
    this list (ref pyfiles/dataset.py:11-18)."""
    return sorted(itertools.product((-1, 1), repeat=n_class_type), reverse=True)


class FaceDataset(torch.utils.data.Dataset):
    """`root`: image directory prefix; `label_root`: prefix of the pickled attribute tables (2-D string arrays:
    column 0 file name, other columns "1" / "-1"); `dataset_label`: {"class": columns whose sign pattern defines the
    class, "existed": columns that must be "1", "delete": columns that must be "-1"}; `classes`: class indices to
    include, in output order; `data_type`: "train" | "val" | "test".

    Per class the matching `<root><name>.png` paths of all tables are sorted; train = the first
    min(train_num, n - val_num - test_num), val = the next val_num, test = the last test_num (like the reference,
    test_num = 0 selects everything).  Items are (transform(RGB image), class index)."""

    def __init__(self, root, label_root, transform, dataset_label, classes, data_type="train", train_num=2000,
                 val_num=500, test_num=500):
        self.transform = transform
        self.images, self.labels = [], []
        patterns = get_class_label(len(dataset_label["class"]))
        tables = [np.asarray(pickle_load(path)) for path in glob.glob(label_root + "*")]  # noqa: F405

        def all_equal(info, columns, value):
            if len(columns) == 0:
                return np.ones(info.shape[0], dtype=bool)
            return np.all(info[:, np.asarray(columns)] == value, axis=1)

        per_class, per_class_labels = {}, []
        for i in range(len(classes)):
            paths = []
            for info in tables:
                keep = all_equal(info, dataset_label["delete"], "-1") & all_equal(info, dataset_label["existed"], "1")
                kept = info[keep]
                hit = np.ones(kept.shape[0], dtype=bool)
                for j, column in enumerate(dataset_label["class"]):
                    hit &= kept[:, column] == str(patterns[i][j])
                paths += [root + str(name).split(".")[0] + ".png" for name in kept[hit, 0]]
            paths.sort()
            n_train = min(train_num, len(paths) - val_num - test_num)
            if data_type == "train":
                paths = paths[:n_train]
            elif data_type == "val":
                paths = paths[n_train:n_train + val_num]
            elif data_type == "test":
                paths = paths[-test_num:]
            per_class[i] = paths
            per_class_labels.append(np.array([i] * len(paths)))
        for c in classes:
            self.images += per_class[c]
            self.labels += list(per_class_labels[c])

    def __getitem__(self, index):
        from PIL import Image
        with open(self.images[index], "rb") as f:
            image = Image.open(f).convert("RGB")
        if self.transform is not None:
            image = self.transform(image)
        return image, self.labels[index]

    def __len__(self):
        return len(self.images)


# ------------------------------------------------------------------------------------------- GPU-side transform
def resample_coeffs(in_size, out_size):
    """Pillow's fixed-point tables for an 8-bit BILINEAR (triangle filter, support scaled by the down-sampling factor)
    resize of one axis from `in_size` to `out_size` samples: (ksize, bounds[out,2] int32 = (first tap, tap count),
    coeffs[out,ksize] int32 with 22 fractional bits).  Same arithmetic, in the same order, as Pillow's
    precompute_coeffs / normalize_coeffs_8bpc (src/libImaging/Resample.c), which is what
    `transforms.Resize((128, 128))` runs on the PIL image (ref notebook 01 cell 9)."""
    import math
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
    inv = 1.0 / filterscale
    one = float(1 << 22)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        first = max(int(center - support + 0.5), 0)
        count = min(int(center + support + 0.5), in_size) - first
        w = [max(0.0, 1.0 - abs((x + first - center + 0.5) * inv)) for x in range(count)]
        total = 0.0
        for v in w:
            total += v
        if total != 0.0:
            w = [v / total for v in w]
        bounds[xx] = (first, count)
        coeffs[xx, :count] = [int(-0.5 + v * one) if v < 0 else int(0.5 + v * one) for v in w]
    return ksize, bounds, coeffs


class GpuFaceTransform(object):
    """The notebooks' `transform["train"]` / `transform["test"]` (ref notebook 01 cell 9: CenterCrop((178, 178)) ->
    Resize((128, 128)) -> RandomHorizontalFlip(p=0.5) -> ToTensor() -> MinMax(True)) for a whole batch of decoded
    images in ONE kernel launch on the GPU, bit-identical to the torchvision / Pillow CPU path.

        tf = GpuFaceTransform(train=True)
        x = tf(batch_u8)        # uint8 [B, H, W, 3] (host or device) -> float32 [B, 3, 128, 128], channels-last, cuda

    Flip decisions are drawn on the host exactly like torchvision does for consecutive samples (one `torch.rand(1)`
    per image from the default CPU generator), so a seeded run sees the same augmentation as the reference pipeline.
    There is no CPU fallback."""

    def __init__(self, crop=178, size=128, p_flip=0.5, train=True, device="cuda"):
        self.crop, self.size, self.p_flip, self.train, self.device = int(crop), int(size), float(p_flip), train, device
        self._tables = None

    def _device_tables(self, dev):
        if self._tables is None or self._tables[0] != dev:
            k, b, c = resample_coeffs(self.crop, self.size)
            to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            self._tables = (dev, k, to(b), to(c))
        return self._tables[1:]

    def draw_flips(self, n):
        """RandomHorizontalFlip.forward: `if torch.rand(1) < self.p`, once per image, in batch order."""
        if not self.train or self.p_flip <= 0:
            return None
        return torch.tensor([1 if float(torch.rand(1)) < self.p_flip else 0 for _ in range(n)], dtype=torch.uint8)

    def __call__(self, batch_u8, flips=None):
        import srgan_ops as ops
        x = torch.as_tensor(batch_u8)
        if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3:
            raise ValueError("expected a uint8 batch of decoded RGB images [B, H, W, 3]")
        B, H, W, _ = x.shape
        if flips is None:
            flips = self.draw_flips(B)
        dev = torch.device(self.device if x.device.type != "cuda" else x.device)
        if dev.type != "cuda":
            raise ops.SrganKernelError("GpuFaceTransform needs a CUDA device; there is no CPU fallback")
        if x.device.type != "cuda":
            x = (x if x.is_pinned() else x.contiguous().pin_memory()).to(dev, non_blocking=True)
        x = x.contiguous()
        k, bounds, coeffs = self._device_tables(x.device)
        f = None if flips is None else torch.as_tensor(flips, dtype=torch.uint8).to(x.device, non_blocking=True)
        y = torch.empty((B, 3, self.size, self.size), dtype=torch.float32, device=x.device,
                        memory_format=torch.channels_last)
        if B == 0:
            return y
        with torch.cuda.device(x.device):
            ops._call("srgan_face_transform", ops._p(x), B, H, W, self.crop, self.size, ops._p(coeffs), ops._p(bounds),
                      k, ops._p(coeffs), ops._p(bounds), k, ops._p(f), ops._p(y), ops._stream())
        return y


def raw_uint8_collate(items):
    """collate_fn for a DataLoader over FaceDataset(transform=None): stacks the decoded PIL images of one batch into
    one pinned uint8 tensor [B, H, W, 3] (what GpuFaceTransform consumes) and the labels into a LongTensor."""
    imgs = torch.from_numpy(np.stack([np.asarray(im.convert("RGB")) for im, _ in items]))
    labels = torch.as_tensor([int(lab) for _, lab in items], dtype=torch.long)
    return (imgs.pin_memory() if torch.cuda.is_available() else imgs), labels
