"""Drive the UNMODIFIED reference (shinshoji01/Style-Restricted_GAN) on CPU -- TEST INFRASTRUCTURE.

Only usable where the reference checkout exists (the build container: /root/reference).  Nothing that runs
on the GPU box imports this module; it exists to generate tests/golden/* (oracle/make_golden.py) and to let
CPU tests compare the product's host logic / the oracle restatement with the real reference in-process.

Recipe = SURVEY.md Appendix D: stub matplotlib / prdc (imported at module top by the reference's util.py but
unused by the training step), import pyfiles/ under private module names (the product ships modules with
the same names), and wrap optG/optE.step in torch.autograd._unsafe_preserve_version_counter to reproduce the
torch-1.4 `.data` update semantics the reference was written against (F7).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("SRGAN_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "pyfiles"))


_mods = None


def load_reference():
    """Returns (model, util, util_notebook) modules of the reference, imported under private names."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REF_ROOT)
    for n in ("matplotlib", "matplotlib.pyplot", "prdc"):
        if n not in sys.modules:
            try:
                importlib.import_module(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(sys.modules["prdc"], "compute_prdc"):
        sys.modules["prdc"].compute_prdc = lambda **k: None

    saved = {k: sys.modules.get(k) for k in ("util", "model", "util_notebook")}
    out = {}
    try:
        for name in ("util", "model", "util_notebook"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, "pyfiles", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod          # the reference does `from util import *`
            spec.loader.exec_module(mod)
            out[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    for name, mod in out.items():
        sys.modules["srgan_reference_" + name] = mod
    _mods = (out["model"], out["util"], out["util_notebook"])
    return _mods


def torch14_step(opt):
    """Make `opt.step()` behave like torch 1.4 (`p.data` updates do not bump the version counter)."""
    orig = opt.step

    def step(*a, **k):
        ps = tuple(p for g in opt.param_groups for p in g["params"])
        with torch.autograd._unsafe_preserve_version_counter(ps):
            return orig(*a, **k)
    opt.step = step


def capture_steps(opt, net, key, record):
    """Record gradients (and, via `after`, weights) of `net` at every `opt.step()`; numbering per key."""
    orig = opt.step
    count = {"n": 0}

    def step(*a, **k):
        name = "%s%d" % (key, count["n"])
        record[name + ".grad"] = {n: (None if p.grad is None else p.grad.detach().clone())
                                  for n, p in net.named_parameters()}
        r = orig(*a, **k)
        record[name + ".weight"] = {n: p.detach().clone() for n, p in net.named_parameters()}
        count["n"] += 1
        return r
    opt.step = step


PROPOSED = dict(cycle=5, idt=5, reg=0.5, idt_reg=0.5, KL=0, batch_KL=10, corr_enc=100, hist=100)
CONVENTIONAL = dict(cycle=5, idt=5, reg=0.5, idt_reg=0, KL=0.1, batch_KL=0, corr_enc=0, hist=0)


def synthetic_batch(batch, seed=123, n_class=4, get_target=None):
    """CelebA-shaped synthetic batch: x ~ U(-1,1) [B,3,128,128], labels in {0..n_class-1}, target
    labels through the reference's `get_target` (NumPy RNG seeded with `seed`)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, 128, 128, generator=g) * 2 - 1
    src = torch.randint(0, n_class, (batch,), generator=g)
    np.random.seed(seed)
    tgt = torch.tensor(get_target(src, tuple(range(n_class)), whole=False)[:, 0], dtype=torch.long)
    return x, {"source": src, "target": tgt}
