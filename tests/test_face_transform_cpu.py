"""CPU: the input-pipeline oracle (oracle/face_transform_oracle.py) against the golden vectors recorded from
torchvision + Pillow (oracle/make_golden_face.py), against torchvision live on real PNG files, and the product's
host-side coefficient tables (dataset.resample_coeffs) against the oracle's.  Everything is byte / integer work or a
fixed sequence of correctly rounded fp32 operations: the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

import cases
import face_transform_oracle as fo

GOLD = os.path.join(cases.GOLDEN, "face_transform.npz")


def test_oracle_matches_golden_vectors_bit_for_bit():
    g = np.load(GOLD)
    for img, flip, ref in zip(g["images"], g["flips"], g["out"]):
        got = fo.face_transform(img, 178, 128, bool(flip))
        assert got.dtype == np.float32 and got.shape == (3, 128, 128)
        assert np.array_equal(got, ref)
        assert got.min() == -1.0 and abs(float(got.max()) - 1.0) < 1e-6


def test_oracle_matches_torchvision_on_png_files(tmp_path):
    """Live: synthetic PNG files -> PIL decode -> the notebook's transform (cell 9) vs the oracle on the decoded bytes;
    also an odd input size (crop origin rounding) and the test-time transform (no flip)."""
    T = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    model, util, _ = cases.use_product_modules()
    rng = np.random.RandomState(3)
    for k, (h, w) in enumerate(((218, 178), (221, 181), (178, 178))):
        arr = (rng.rand(h, w, 3) * 255).astype(np.uint8)
        arr[:, :, 1] = np.clip(arr[:, :, 1] // 2 + 40, 0, 255)
        path = str(tmp_path / ("img%d.png" % k))
        Image.fromarray(arr).save(path)
        with open(path, "rb") as f:
            pil = Image.open(f).convert("RGB")
        tf = T.Compose([T.CenterCrop((178, 178)), T.Resize((128, 128)), T.ToTensor(), util.MinMax(True)])
        ref = tf(pil).numpy()
        got = fo.face_transform(np.asarray(pil), 178, 128, False)
        assert np.array_equal(got, ref), (h, w)
        # crop and resize commute with the mirror only when the crop window is centred on whole pixels
        if (w - 178) % 2 == 0:
            assert np.array_equal(fo.face_transform(np.asarray(pil), 178, 128, True), ref[:, :, ::-1])


def test_product_coefficient_tables_equal_the_oracles():
    cases.use_product_modules()
    import dataset
    for n_in, n_out in ((178, 128), (218, 128), (128, 128), (100, 160), (64, 7)):
        k0, b0, c0 = fo.resample_coeffs(n_in, n_out)
        k1, b1, c1 = dataset.resample_coeffs(n_in, n_out)
        assert k0 == k1 and np.array_equal(b0, b1) and np.array_equal(c0, c1), (n_in, n_out)
        assert np.all(np.abs(c1.sum(1) - (1 << 22)) <= c1.shape[1])       # rows sum to one (up to rounding)


def test_flip_draws_consume_the_rng_like_torchvision():
    T = pytest.importorskip("torchvision.transforms")
    cases.use_product_modules()
    import dataset
    tf = dataset.GpuFaceTransform(train=True, device="cpu")
    torch.manual_seed(11)
    mine = tf.draw_flips(16).tolist()
    torch.manual_seed(11)
    flipper = T.RandomHorizontalFlip(p=0.5)
    probe = torch.arange(6.0).view(1, 2, 3)
    theirs = [int(not torch.equal(flipper(probe), probe)) for _ in range(16)]
    assert mine == theirs
    assert dataset.GpuFaceTransform(train=False).draw_flips(4) is None
    with pytest.raises(Exception):
        tf(torch.zeros(2, 218, 178, 3, dtype=torch.uint8))          # no CPU fallback
