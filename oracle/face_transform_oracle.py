"""CPU ORACLE of the notebooks' image pre-processing -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in numpy integer / float32 arithmetic, what the reference's `transform["train"]` does to one decoded CelebA
image (ref notebook/01-train_Conventional_SingleGAN.ipynb cell 9; the same cell exists in notebooks 02/03/05):

    transforms.CenterCrop((178, 178)) -> transforms.Resize((128, 128)) -> transforms.RandomHorizontalFlip(p=0.5)
    -> transforms.ToTensor() -> MinMax(True)                                  (MinMax: ref pyfiles/util.py:108-116,148-153)

`Resize` on a PIL image is Pillow's two-pass (horizontal, then vertical) 8-bit resampling with a triangle filter whose
support grows with the down-scaling factor, fixed-point coefficients (22 fractional bits) and a uint8 intermediate image
(Pillow src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
Vertical_8bpc; Pillow is a dependency of the reference, not vendored in it: the published algorithm is restated here).

Parity status: PINNED.  tests/test_face_transform_cpu.py compares this restatement bit for bit with torchvision + Pillow
(both present in this image) on synthetic PNG files, and with tests/golden/face_transform.npz recorded from them by
oracle/make_golden_face.py.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size, out_size):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter (support 1.0) over the whole axis.
    Returns (ksize, bounds[out_size, 2] = (first input index, tap count), coeffs[out_size, ksize] int32)."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size          # box coordinates are C floats
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.array([max(0.0, 1.0 - abs((x + xmin - center + 0.5) * ss)) for x in range(xmax)], dtype=np.float64)
        ww = float(sum(w.tolist()))                                          # left-to-right double sum, as in C
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)), (0.5 + kk * (1 << PRECISION_BITS)))
    return ksize, bounds, np.trunc(fixed).astype(np.int32)


def _resample_axis0(img, out_size):
    """8-bit resampling along axis 0 of a [n, m, c] uint8 array."""
    ksize, bounds, coef = resample_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for j in range(n):
            acc += src[x0 + j] * int(coef[xx, j])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def center_crop_origin(height, width, crop):
    """torchvision.transforms.functional.center_crop: (top, left) of the crop window."""
    if crop > height or crop > width:
        raise ValueError("crop larger than the image (torchvision would pad): not part of the CelebA pipeline")
    return int(round((height - crop) / 2.0)), int(round((width - crop) / 2.0))


def face_transform(img_u8, crop=178, size=128, flip=False):
    """One decoded RGB image [H, W, 3] uint8 -> float32 [3, size, size] in [-1, 1] (the item a DataLoader batches)."""
    h, w, _ = img_u8.shape
    top, left = center_crop_origin(h, w, crop)
    c = img_u8[top:top + crop, left:left + crop]
    # horizontal pass first (uint8 intermediate of shape [crop, size]), then vertical
    hpass = np.transpose(_resample_axis0(np.transpose(c, (1, 0, 2)), size), (1, 0, 2))
    r = _resample_axis0(hpass, size)                                          # [size, size, 3] uint8
    if flip:
        r = r[:, ::-1]
    t = np.transpose(r, (2, 0, 1)).astype(np.float32) / np.float32(255.0)     # ToTensor
    lo, hi = t.min(), t.max()
    den = np.float32(np.float32(hi - lo) + np.float32(1e-8))
    return ((t - lo) / den) * np.float32(2.0) - np.float32(1.0)               # MinMax(True)
