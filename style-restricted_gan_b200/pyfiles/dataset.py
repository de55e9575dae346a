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
