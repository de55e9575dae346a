"""Record tests/golden/face_transform.npz from torchvision + Pillow (the libraries the reference's notebooks call:
notebook/01-train_Conventional_SingleGAN.ipynb cell 9) -- TEST INFRASTRUCTURE.  Run in the build container:
    python oracle/make_golden_face.py
Stores 3 synthetic 218x178 RGB images (uint8), their flip flags and the float32 [3,128,128] results of
CenterCrop(178) -> Resize(128) -> hflip -> ToTensor -> MinMax(True)."""
import os
import sys

import numpy as np
import torch
from PIL import Image
import torchvision.transforms as T
import torchvision.transforms.functional as TF

HERE = os.path.dirname(os.path.abspath(__file__))


def synthetic_images():
    rng = np.random.RandomState(7)
    yy, xx = np.mgrid[0:218, 0:178]
    a = (rng.rand(218, 178, 3) * 255).astype(np.uint8)                                        # white noise
    b = np.stack([yy * 255 / 217, xx * 255 / 177, (yy + xx) % 256], -1).astype(np.uint8)      # ramps / sawtooth
    c = (127 + 100 * np.sin(yy / 7.0)[..., None] * np.cos(xx / 5.0)[..., None] + rng.randn(218, 178, 3) * 9)
    c = np.clip(c, 30, 220).astype(np.uint8)                                                  # smooth, limited range
    return np.stack([a, b, c])


def reference_pipeline(img_u8, flip):
    pil = Image.fromarray(img_u8)
    pil = T.Resize((128, 128))(T.CenterCrop((178, 178))(pil))
    if flip:
        pil = TF.hflip(pil)
    t = T.ToTensor()(pil).numpy()
    lo, hi = t.min(keepdims=True), t.max(keepdims=True)
    return ((t - lo) / (hi - lo + 1e-8)) * 2 - 1          # ref pyfiles/util.py:108-116 with mean0=True


if __name__ == "__main__":
    imgs = synthetic_images()
    flips = np.array([0, 1, 1], dtype=np.uint8)
    out = np.stack([reference_pipeline(im, f) for im, f in zip(imgs, flips)]).astype(np.float32)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "face_transform.npz")
    np.savez_compressed(path, images=imgs, flips=flips, out=out)
    print("wrote", path, out.shape, os.path.getsize(path))
