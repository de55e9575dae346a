"""GPU: the input-pipeline kernel (srgan_face_transform through dataset.GpuFaceTransform) against the torchvision /
Pillow CPU pipeline of the notebooks (ref notebook 01 cell 9, pyfiles/dataset.py:127-141) on synthetic PNG files and
against the committed golden vectors.  Bar: BIT-EXACT (crop, resize bytes, flip, ToTensor, MinMax)."""
import os

import numpy as np
import pytest
import torch

import cases
import face_transform_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tf(train):
    cases.use_product_modules()
    import dataset
    return dataset.GpuFaceTransform(train=train, device=DEV)


def test_kernel_matches_golden_vectors_bit_for_bit():
    g = np.load(os.path.join(cases.GOLDEN, "face_transform.npz"))
    y = _tf(True)(torch.from_numpy(g["images"]), flips=torch.from_numpy(g["flips"]))
    assert y.shape == (3, 3, 128, 128) and y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y.cpu(), torch.from_numpy(g["out"]))


def test_kernel_matches_torchvision_on_png_files(tmp_path):
    T = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    model, util, _ = cases.use_product_modules()
    rng = np.random.RandomState(5)
    B = 9
    arrs = []
    for k in range(B):
        a = (rng.rand(218, 178, 3) * 255).astype(np.uint8)
        if k % 3 == 1:
            a = np.clip(a // 3 + 60, 0, 255).astype(np.uint8)      # narrow range: MinMax really rescales
        if k % 3 == 2:
            a[:] = 77
            a[100, 50] = (78, 77, 77)                              # almost constant image: tiny denominator
        Image.fromarray(a).save(str(tmp_path / ("%d.png" % k)))
    pils = []
    for k in range(B):
        with open(str(tmp_path / ("%d.png" % k)), "rb") as f:
            pils.append(Image.open(f).convert("RGB"))
    train = T.Compose([T.CenterCrop((178, 178)), T.Resize((128, 128)), T.RandomHorizontalFlip(p=0.5), T.ToTensor(),
                       util.MinMax(True)])
    torch.manual_seed(21)
    ref = torch.stack([train(p) for p in pils])                    # consumes one torch.rand(1) per image
    batch = torch.from_numpy(np.stack([np.asarray(p) for p in pils]))
    torch.manual_seed(21)
    got = _tf(True)(batch)
    assert torch.equal(got.cpu(), ref)
    test = T.Compose([T.CenterCrop((178, 178)), T.Resize((128, 128)), T.ToTensor(), util.MinMax(True)])
    ref_t = torch.stack([test(p) for p in pils])
    assert torch.equal(_tf(False)(batch.to(DEV)).cpu(), ref_t)     # device-resident input, no flip
    # ragged input size (crop origin rounding) against the oracle
    odd = (rng.rand(2, 223, 187, 3) * 255).astype(np.uint8)
    y = _tf(False)(torch.from_numpy(odd))
    for i in range(2):
        assert np.array_equal(y[i].cpu().numpy(), fo.face_transform(odd[i], 178, 128, False))
    assert _tf(True)(torch.zeros(0, 218, 178, 3, dtype=torch.uint8)).shape == (0, 3, 128, 128)


def test_transform_feeds_the_generator_and_throughput():
    """The output is the channels-last batch the generator's stem reads (no layout pass); prints images/s of the
    transform for a resident uint8 batch of 256 CelebA-sized images (SURVEY 8 f2 target: >= 1e4 images/s)."""
    model, util, _ = cases.use_product_modules()
    import srgan_ops as ops
    tf = _tf(True)
    batch = torch.randint(0, 256, (256, 218, 178, 3), dtype=torch.uint8, device=DEV)
    flips = torch.randint(0, 2, (256,), dtype=torch.uint8)
    y = tf(batch, flips=flips)
    assert ops._dense_nhwc(y) and float(y.min()) == -1.0
    G = model.SingleGenerator(3, 8, 2, 2, 1, "instance", num_con=12).to(DEV)
    assert G(y[:2], torch.randn(2, 12, device=DEV)).shape == (2, 3, 128, 128)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tf(batch, flips=flips)
    e1.record()
    torch.cuda.synchronize()
    ips = 256 * 10 / (e0.elapsed_time(e1) * 1e-3)
    mb = 256 * (178 * 178 * 3 + 128 * 128 * 3 * 4) / 1e6
    print("face transform: %.0f images/s (%.1f us per batch of 256, %.0f GB/s algorithmic)" %
          (ips, e0.elapsed_time(e1) * 100, mb * 10 / e0.elapsed_time(e1)))
    assert ips > 1e4
