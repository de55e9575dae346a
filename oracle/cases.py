"""Parity cases shared by oracle/make_golden.py and tests/ -- TEST INFRASTRUCTURE.

A case fixes: the trainer kind, network widths, batch, unrolled k, the lambda set, and the seeds.  The same
recipe is replayed three ways: by the unmodified reference (make_golden.py, build container only), by the
CPU oracle restatement (srgan_oracle.OracleTrainer) and by the product on a B200.
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYFILES = os.path.join(REPO, "style-restricted_gan_b200", "pyfiles")
GOLDEN = os.path.join(REPO, "tests", "golden")

PROPOSED = dict(cycle=5, idt=5, reg=0.5, idt_reg=0.5, KL=0, batch_KL=10, corr_enc=100, hist=100)
CONVENTIONAL = dict(cycle=5, idt=5, reg=0.5, idt_reg=0, KL=0.1, batch_KL=0, corr_enc=0, hist=0)

CASES = {
    # notebook 03/05 recipe on narrow networks (full-tensor digests stay small)
    "srgan_small": dict(kind="srgan", nch=8, dis_nch=8, enc_nch=8, res_num=2, batch=4, k=2,
                        lbd=dict(PROPOSED, **{"class": 1}), feature="mu", seed=0),
    # notebook 02: class-conditioned encoder, single discriminator with class head
    "single_solo_small": dict(kind="single_solo", nch=8, dis_nch=8, enc_nch=8, res_num=2, batch=4, k=2,
                              lbd=dict(PROPOSED, **{"class": 1}), feature="mu", seed=1),
    # notebook 01 (BASELINE config 1 recipe): conventional KL, one discriminator per class, k=1
    "single_multi_small": dict(kind="single_multi", nch=8, dis_nch=8, enc_nch=8, res_num=2, batch=6, k=1,
                               lbd=dict(CONVENTIONAL), feature="latent", seed=2),
    # notebook 05 recipe (BASELINE config 4): encoder trunk "pretrained" and frozen while the optimizer is built, so
    # optE = Adam(lr 1e-3) over fcmean / fcvar only (05-train... cell 22); the trunk is melted again afterwards
    "srgan_frozen_small": dict(kind="srgan", nch=8, dis_nch=8, enc_nch=8, res_num=2, batch=4, k=2,
                               lbd=dict(PROPOSED, **{"class": 1}), feature="mu", seed=4, frozen=True),
    # notebook 03 at FULL width (nch 64, 6 residual blocks), batch 2, k=1: exercises the tensor-core shapes
    "srgan_full": dict(kind="srgan", nch=64, dis_nch=64, enc_nch=64, res_num=6, batch=2, k=1,
                       lbd=dict(PROPOSED, **{"class": 1}), feature="mu", seed=3),
}
N_CLASS, NDIM = 4, 8


def use_product_modules():
    """Put the product's drop-in pyfiles/ first on sys.path and return (model, util, util_notebook)."""
    if PYFILES not in sys.path:
        sys.path.insert(0, PYFILES)
    import model
    import util
    import util_notebook
    if not os.path.samefile(os.path.dirname(model.__file__), PYFILES):
        raise RuntimeError("a foreign `model` module shadows the product's pyfiles/")
    return model, util, util_notebook


def build_nets(model_mod, case, device="cpu"):
    """Construct G, D (or list of D), E in the notebooks' order (nb01/02/03 cell 20) with default init."""
    c = CASES[case] if isinstance(case, str) else case
    m = model_mod
    ref_dim = N_CLASS
    G = m.SingleGenerator(3, c["nch"], 2, 2, c["res_num"], "instance", num_con=ref_dim + NDIM)
    if c["kind"] == "single_multi":
        D = [m.SingleDiscriminator_original_multi(3, c["dis_nch"], 2, 4, "instance") for _ in range(N_CLASS)]
    else:
        D = m.SingleDiscriminator_solo_multi(3, c["dis_nch"], 2, 4, "instance", ref_dim)
    if c["kind"] == "srgan":
        E = m.Encoder(3, NDIM, c["enc_nch"], 4, "instance", ref_dim, device)
    else:
        E = m.Encoder_original(3, NDIM, c["enc_nch"], 4, "instance", ref_dim, device)
    return G, D, E


def build_trainer(nb_mod, case, nets, device, adam=torch.optim.Adam):
    """`adam`: optimizer class for the notebook-05 encoder optimizer (the notebook uses torch.optim.Adam; the product
    arm of bench.py passes srgan_ops.FusedAdam, same signature, so the step can be replayed as a CUDA graph)."""
    c = CASES[case] if isinstance(case, str) else case
    crit = torch.nn.MSELoss()
    ref_label = np.eye(N_CLASS)
    G, D, E = nets
    if c["kind"] == "srgan":
        optE = None
        if c.get("frozen"):
            keys = frozen_keys(E)
            E.freeze_melt(keys, "freeze")
            optE = adam(filter(lambda p: p.requires_grad, E.parameters()), lr=0.001, betas=(0.5, 0.999))
            E.freeze_melt(keys, "melt")
        sg = nb_mod.SRGAN_training([G, D, E], [None, None, optE], [crit, torch.nn.MSELoss()], c["lbd"], c["k"],
                                   device, ref_label, c["batch"], c["feature"], NDIM)
    else:
        single = c["kind"] == "single_solo"
        sg = nb_mod.SingleGAN_training([G, D, E], [None, None, None],
                                       [crit, torch.nn.MSELoss() if single else None], c["lbd"], c["k"], device,
                                       ref_label, NDIM, tuple(range(N_CLASS)), c["batch"], c["feature"], single)
    sg.opt_sche_initialization()
    return sg


def synthetic_batch(batch, get_target, seed=123):
    """CelebA-shaped synthetic batch."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, 128, 128, generator=g) * 2 - 1
    src = torch.randint(0, N_CLASS, (batch,), generator=g)
    np.random.seed(seed)
    tgt = torch.tensor(get_target(src, tuple(range(N_CLASS)), whole=False)[:, 0], dtype=torch.long)
    return x, {"source": src, "target": tgt}


def state_dicts(nets):
    G, D, E = nets
    sd = lambda n: {k: v.detach().clone() for k, v in n.state_dict().items()}
    return sd(G), ([sd(d) for d in D] if isinstance(D, (list, tuple)) else sd(D)), sd(E)


def frozen_keys(E):
    """state_dict keys of the classifier part of an Encoder (= Encoder_classifier's keys: everything but the
    fcmean / fcvar heads), what notebook 05 hands to freeze_melt."""
    return [k for k in E.state_dict().keys() if not k.startswith(("fcmean", "fcvar"))]


def build_oracle(case, sds, oracle_mod):
    c = CASES[case] if isinstance(case, str) else case
    g_sd, d_sd, e_sd = sds
    kw = {}
    if c.get("frozen"):
        kw = dict(e_trainable={k for k in e_sd if k.startswith(("fcmean", "fcvar"))}, lr_e=1e-3)
    return oracle_mod.OracleTrainer(c["kind"], g_sd, d_sd, e_sd, c["lbd"], c["k"], np.eye(N_CLASS), c["batch"],
                                    c["feature"], NDIM, g_cfg=(2, c["res_num"]), **kw)
