"""CPU: the notebooks' dataset surface (`get_class_label`, `FaceDataset`, ref pyfiles/dataset.py) on a synthetic
CelebA-shaped directory: attribute tables pickled as string arrays, PNG files.  Where the reference checkout exists
(build container) the splits are compared with the reference's own class, item by item."""
import importlib.util
import os
import pickle
import sys

import numpy as np
import pytest

import cases
import ref_harness


def _make_celeba(tmp, n=60, seed=0):
    from PIL import Image
    rng = np.random.RandomState(seed)
    root, label_root = str(tmp / "img") + os.sep, str(tmp / "lab") + os.sep
    os.makedirs(root)
    os.makedirs(label_root)
    names = ["%06d.jpg" % (i + 1) for i in range(n)]
    for nm in names:
        Image.fromarray(rng.randint(0, 255, (8, 8, 3), dtype=np.uint8)).save(root + nm.split(".")[0] + ".png")
    attrs = rng.choice(["1", "-1"], size=(n, 5))
    table = np.concatenate([np.array(names)[:, None], attrs], axis=1)        # col 0 name, cols 1..5 attributes
    for part, rows in enumerate((table[: n // 2], table[n // 2:])):
        with open(label_root + "part%d.pkl" % part, "wb") as f:
            pickle.dump(rows, f)
    return root, label_root


def _product_dataset():
    cases.use_product_modules()
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "style-restricted_gan_b200",
                        "pyfiles")
    spec = importlib.util.spec_from_file_location("srgan_product_dataset", os.path.join(here, "dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_get_class_label_order():
    ds = _product_dataset()
    assert ds.get_class_label(1) == [(1,), (-1,)]
    assert ds.get_class_label(2) == [(1, 1), (1, -1), (-1, 1), (-1, -1)]
    assert len(ds.get_class_label(3)) == 8 and ds.get_class_label(3)[0] == (1, 1, 1)


def test_face_dataset_splits(tmp_path):
    ds = _product_dataset()
    root, label_root = _make_celeba(tmp_path)
    spec = {"class": [1, 2], "existed": [3], "delete": [4]}
    classes = (0, 1, 2, 3)
    kw = dict(train_num=4, val_num=1, test_num=1)
    parts = {t: ds.FaceDataset(root, label_root, None, spec, classes, t, **kw) for t in ("train", "val", "test")}
    table = np.concatenate([pickle.load(open(label_root + "part%d.pkl" % p, "rb")) for p in (0, 1)])
    pats = ds.get_class_label(2)
    for t, d in parts.items():
        assert len(d) == len(d.images) == len(d.labels)
        for path, lab in zip(d.images, d.labels):
            row = table[table[:, 0] == os.path.basename(path).replace(".png", ".jpg")][0]
            assert row[3] == "1" and row[4] == "-1"                          # "existed" and "delete" filters
            assert (int(row[1]), int(row[2])) == pats[int(lab)]              # the class is the sign pattern
    for c in classes:                                                        # splits of a class are disjoint, ordered
        tr = [p for p, l in zip(parts["train"].images, parts["train"].labels) if l == c]
        va = [p for p, l in zip(parts["val"].images, parts["val"].labels) if l == c]
        te = [p for p, l in zip(parts["test"].images, parts["test"].labels) if l == c]
        assert tr == sorted(tr) and len(tr) <= 4 and len(va) <= 1 and len(te) <= 1
        assert not set(tr) & set(va) and not set(tr) & set(te)
    img, lab = parts["train"][0]
    assert img.mode == "RGB" and img.size == (8, 8) and int(lab) == int(parts["train"].labels[0])
    loose = {"class": [1, 2], "existed": [], "delete": []}
    flipped = ds.FaceDataset(root, label_root, lambda im: np.asarray(im)[:, ::-1], loose, (1, 0), "train", **kw)
    assert len(flipped) == 8
    assert [int(l) for l in flipped.labels] == sorted([int(l) for l in flipped.labels], reverse=True)   # `classes` order
    assert flipped[0][0].shape == (8, 8, 3)


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present")
def test_face_dataset_matches_reference(tmp_path, monkeypatch):
    ds = _product_dataset()
    root, label_root = _make_celeba(tmp_path, n=80, seed=3)
    if not hasattr(np, "int"):
        monkeypatch.setattr(np, "int", int, raising=False)       # the reference uses the alias NumPy 1.24 removed
    ref_harness.load_reference()
    saved = sys.modules.get("util")
    sys.modules["util"] = ref_harness.load_reference()[1]        # the reference's dataset.py does `from util import *`
    try:
        spec = importlib.util.spec_from_file_location(
            "srgan_reference_dataset", os.path.join(ref_harness.REF_ROOT, "pyfiles", "dataset.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        if saved is None:
            sys.modules.pop("util", None)
        else:
            sys.modules["util"] = saved
    assert ds.get_class_label(3) == ref.get_class_label(3)
    for spec_ in ({"class": [1, 2], "existed": [], "delete": []}, {"class": [2], "existed": [3, 5], "delete": [4]},
                  {"class": [1, 2, 3], "existed": [], "delete": [5]}):
        ncls = 2 ** len(spec_["class"])
        for t in ("train", "val", "test"):
            for kw in (dict(train_num=5, val_num=2, test_num=2), dict(train_num=1000, val_num=1, test_num=3)):
                a = ds.FaceDataset(root, label_root, None, spec_, tuple(range(ncls)), t, **kw)
                b = ref.FaceDataset(root, label_root, None, spec_, tuple(range(ncls)), t, **kw)
                assert a.images == b.images, (spec_, t, kw)
                assert [int(x) for x in a.labels] == [int(x) for x in b.labels]
