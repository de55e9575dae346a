"""PRDC evaluation of translated images - drop-in for the reference's `pyfiles/evaluation.py`.

Same names and call pattern (`GAN_evaluation(feature_extractor, device, classes, reference)`, `.preprocess`,
`.get_feature`, `.get_prdc(true, pred, nearest_k=5, preprocess=True)`, `evaluation_init`, `vgg_model`).  What runs
where:
  * `compute_prdc` - the reference imports it from the dependency `prdc==0.2` (ref pyfiles/evaluation.py:7,107);
    here it is `srgan_ops.compute_prdc`: fp64 squared-distance, k-th-neighbour and counting kernels of this package
    (csrc/eval.cu); the four metrics are ratios of the integer counts the kernels return;
  * the feature extractor is torchvision's VGG19-bn exactly as in the reference (ref :12-35, 41-63): library code,
    imported lazily, not part of the hot path.  `feature_extractor="identity"` (new) skips it and measures PRDC on the
    flattened tensors, which is what the GPU tests use (no torchvision weights on the GPU box).
"""
import numpy as np
import torch
import torch.nn as nn

import srgan_ops as ops
from util import *  # noqa: F401,F403

compute_prdc = ops.compute_prdc


class vgg_model():
    """ref pyfiles/evaluation.py:12-35: features -> avgpool -> first six classifier layers ("feature") or the whole
    network ("score")."""

    def __init__(self, model):
        self.feature_extractor = nn.Sequential(*(list(model.features.children()) + list(model.avgpool.children())))
        self.fcs = nn.Sequential(*list(model.classifier.children())[:6])
        self.model = model

    def get(self, x, output_type="score"):
        with torch.no_grad():
            if output_type == "feature":
                return self.fcs(torch.flatten(self.feature_extractor(x), 1))
            if output_type == "score":
                return self.model(x)
        return None


class GAN_evaluation():
    def __init__(self, feature_extractor="vgg-initialization", device="cpu", classes=tuple(range(4)),
                 reference=tuple(range(4))):
        self.fe = feature_extractor
        self.device = device
        self.model = None
        self.transform = None
        if "vgg" in self.fe:
            import torchvision.models as models
            import torchvision.transforms as transforms
            if "ImageNet" in self.fe:
                model = models.vgg19_bn(pretrained=True).to(device)
            elif "initialization" in self.fe:
                model = models.vgg19_bn(pretrained=False).to(device)
                model.apply(weights_init)  # noqa: F405
            elif "CelebA" in self.fe:
                model = models.vgg19_bn(pretrained=False).to(device)
                model.classifier[6] = nn.Linear(in_features=4096, out_features=len(classes))
                path = "../data/parameters/B/facial_recognizer_vgg_lr5e-05_epoch126.pth"
                model.load_state_dict(torch.load(path, map_location=device))
                model = model.to(device)
            else:
                raise ValueError("unknown feature extractor %r" % (self.fe,))
            model.eval()
            self.model = vgg_model(model)
            self.transform = transforms.Compose([
                transforms.Resize((128, 128)),
                transforms.Resize((224, 224)),
                transforms.ToTensor(),
                transforms.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)),
            ])
        elif self.fe != "identity":
            raise ValueError("unknown feature extractor %r" % (self.fe,))

    def preprocess(self, tensor):
        if self.transform is None:
            return tensor
        images = []
        for i in range(tensor.shape[0]):
            image = image_from_output(tensor[i:i + 1])[0]  # noqa: F405
            images.append(self.transform(image).numpy())
        return torch.Tensor(np.array(images))

    def get_feature(self, tensor, batch=32, get_attention=False, thres=0.5):
        """[num, F] features; stays on the device (the reference returns numpy, `compute_prdc` here takes either)."""
        feats = []
        for lo in range(0, tensor.shape[0], batch):
            data = tensor[lo:lo + batch].to(self.device)
            f = data if self.model is None else self.model.get(data, "feature")
            feats.append(f.reshape(data.shape[0], -1).float())
        return torch.cat(feats, dim=0) if feats else torch.zeros((0, 0))

    def get_prdc(self, true, pred, nearest_k=5, preprocess=True, thres=0.5, batch=32):
        self.run_preprocess = preprocess
        if preprocess:
            true = self.preprocess(true)
            pred = self.preprocess(pred)
        f1 = self.get_feature(true, batch)
        f2 = self.get_feature(pred, batch)
        if f1.shape[1] == 0:
            return {"precision": None, "recall": None, "density": None, "coverage": None}
        return compute_prdc(real_features=f1, fake_features=f2, nearest_k=nearest_k)


def evaluation_init(fe_list, classes, metrics):
    """ref pyfiles/evaluation.py:112-124: nested result store [extractor][source][target][metric] -> []."""
    return {fe: {s: {t: {m: [] for m in metrics.keys()} for t in classes} for s in classes} for fe in fe_list}
