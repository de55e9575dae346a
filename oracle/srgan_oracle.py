"""CPU ORACLE for the SRGAN / SingleGAN training step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, fp32) restatement of the reference algorithm
(shinshoji01/Style-Restricted_GAN, pyfiles/model.py + util.py + util_notebook.py), written
functionally over `state_dict`-style parameter dictionaries.  Only `tests/`, `__graft_entry__.smoke()`
and the CPU-baseline legs of `bench.py` may import this module; the product (style-restricted_gan_b200/)
never does.

Parity status: PINNED.  `oracle/make_golden.py` runs the UNMODIFIED reference (imported from
/root/reference in the build container, with the import stubs and the torch-1.4 version-counter shim of
SURVEY.md Appendix D) on seeded inputs and stores losses / latent statistics / gradient digests under
tests/golden/; `tests/test_oracle_golden.py` checks this restatement against them.  The reference
itself ships no tests or golden vectors ("parity unpinned" by the reference's own suite).

Each function cites the reference lines it restates.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------ nets


def _cbin(sd, pre, x, con, eps=1e-5):
    """CBINorm2d, affine=True: (IN(x) + tanh(Linear(con))) * weight + bias.  ref model.py:54-67."""
    t = torch.tanh(F.linear(con, sd[pre + "ConBias.0.weight"], sd[pre + "ConBias.0.bias"]))
    h = F.instance_norm(x, eps=eps) + t[:, :, None, None]
    return h * sd[pre + "weight"][None, :, None, None] + sd[pre + "bias"][None, :, None, None]



def cbbn(x, con, weight, bias, lin_w, lin_b, running_mean, running_var, training=True, momentum=0.1, eps=1e-5):
    """CBBNorm2d, affine=True (ref model.py:121-148), written out instead of calling F.batch_norm:
    out = (x - mean_c) / sqrt(var_c + eps) with batch statistics (biased variance) when training, running statistics
    otherwise; result = (out - mean_hw(out) + tanh(Linear(con))) * weight + bias.
    Returns (y, new_running_mean, new_running_var) -- running_var is updated with the UNBIASED variance."""
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
        n = x.numel() // x.shape[1]
        new_rm = (1 - momentum) * running_mean + momentum * mean.detach()
        new_rv = (1 - momentum) * running_var + momentum * var.detach() * n / max(n - 1, 1)
    else:
        mean, var, new_rm, new_rv = running_mean, running_var, running_mean, running_var
    out = (x - mean[None, :, None, None]) * torch.rsqrt(var + eps)[None, :, None, None]
    t = torch.tanh(F.linear(con, lin_w, lin_b))
    y = (out - out.mean(dim=(2, 3), keepdim=True) + t[:, :, None, None]) * weight[None, :, None, None] \
        + bias[None, :, None, None]
    return y, new_rm, new_rv


def generator_forward(sd, x, c, num_cls=2, res_num=6):
    """SingleGenerator.forward.  ref model.py:236-249 (layers :203-234, residual block :188-201)."""
    for i in range(num_cls + 1):
        w = sd["down_convs.%d.weight" % i]
        x = F.conv2d(x, w, stride=1 if i == 0 else 2, padding=3 if i == 0 else 1)
        x = F.relu(_cbin(sd, "down_cnorms.%d." % i, x, c))
    for b in range(res_num):
        p = "resBlocks.%d." % b
        h = F.relu(_cbin(sd, p + "cn1.", F.conv2d(x, sd[p + "c1.weight"], padding=1), c))
        h = _cbin(sd, p + "cn2.", F.conv2d(h, sd[p + "c2.weight"], padding=1), c)
        x = h + x
    for i in range(num_cls):
        x = F.conv_transpose2d(x, sd["up_convs.%d.weight" % i], stride=2, padding=1)
        x = F.relu(F.instance_norm(x, eps=1e-5))
    return torch.tanh(F.conv2d(x, sd["up_convs.%d.weight" % num_cls], padding=3))


def _tower(sd, pre, x, num_cls, with_head):
    """conv4x4/s2 + LeakyReLU(0.01) x num_cls (+ patch head).  ref model.py:255-279,294-316."""
    for i in range(num_cls):
        x = F.leaky_relu(F.conv2d(x, sd[pre + "down_convs.%d.weight" % (2 * i)], stride=2, padding=1), 0.01)
    if with_head:
        j = 2 * num_cls
        x = F.conv2d(x, sd[pre + "down_convs.%d.weight" % j], sd[pre + "down_convs.%d.bias" % j], padding=1)
    return x


def _down3(x):
    return F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False)


def discriminator_solo_forward(sd, x, num_cls=4, n_class=4):
    """SingleDiscriminator_solo_multi.forward -> ([patch1, patch2], [probs1, probs2]).  ref model.py:339-346."""
    f1 = _tower(sd, "discriminator1.", x, num_cls, False)
    f2 = _tower(sd, "discriminator2.", _down3(x), num_cls, False)
    o1 = F.conv2d(f1, sd["last_layer1.weight"], sd["last_layer1.bias"], padding=1)
    o2 = F.conv2d(f2, sd["last_layer2.weight"], sd["last_layer2.bias"], padding=1)
    c1 = F.softmax(F.conv2d(f1, sd["classification_layer1.0.weight"], sd["classification_layer1.0.bias"]), dim=1)
    c2 = F.softmax(F.conv2d(f2, sd["classification_layer2.0.weight"], sd["classification_layer2.0.bias"]), dim=1)
    return [o1, o2], [c1.reshape(-1, n_class), c2.reshape(-1, n_class)]


def discriminator_original_forward(sd, x, num_cls=4):
    """SingleDiscriminator_original_multi.forward -> [patch1, patch2].  ref model.py:289-292."""
    return [_tower(sd, "discriminator1.", x, num_cls, True), _tower(sd, "discriminator2.", _down3(x), num_cls, True)]


def _reflect_conv3(x, w):
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w)


def _encoder_trunk(sd, x, num_cls, cond=None):
    """first_layer + BasicBlock(_classification) x num_cls.  ref model.py:352-376,413-437."""
    x = F.conv2d(x, sd["first_layer.weight"], sd["first_layer.bias"], stride=2, padding=1)
    for i in range(num_cls):
        p = "layers.%d." % i
        if cond is None:
            h = F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.2)
        else:
            h = F.leaky_relu(_cbin(sd, p + "cnorm1.", x, cond), 0.2)
        h = _reflect_conv3(h, sd[p + "conv1.weight"])
        if cond is None:
            h = F.leaky_relu(F.instance_norm(h, eps=1e-5), 0.2)
        else:
            h = F.leaky_relu(_cbin(sd, p + "cnorm2.", h, cond), 0.2)
        h = F.avg_pool2d(_reflect_conv3(h, sd[p + "cmp.0.weight"]), 2, 2)
        s = F.conv2d(F.avg_pool2d(x, 2, 2), sd[p + "shortcut.1.weight"], sd[p + "shortcut.1.bias"])
        x = h + s
    return F.adaptive_avg_pool2d(F.leaky_relu(x, 0.2), 1).flatten(1)


def _reparam(mu, logvar):
    """z = eps * exp(logvar/2) + mu with eps from the CPU default generator.  ref model.py:398-402,459-463."""
    eps = torch.empty(mu.shape, dtype=torch.float32).normal_()
    return eps * torch.exp(0.5 * logvar) + mu


def encoder_forward(sd, x, num_cls=4):
    """Encoder.forward -> (z, mu, logvar, class_logits, None).  ref model.py:474-482."""
    f = _encoder_trunk(sd, x, num_cls)
    mu = F.linear(f, sd["fcmean.weight"], sd["fcmean.bias"])
    logvar = F.linear(f, sd["fcvar.weight"], sd["fcvar.bias"])
    z = _reparam(mu, logvar)
    cls = F.linear(f, sd["fcclass.weight"], sd["fcclass.bias"])
    return z, mu, logvar, cls, None


def encoder_original_forward(sd, x, c, num_cls=4):
    """Encoder_original.forward -> (z, mu, logvar).  ref model.py:404-411."""
    f = _encoder_trunk(sd, x, num_cls, cond=c)
    mu = F.linear(f, sd["fcmean.weight"], sd["fcmean.bias"])
    logvar = F.linear(f, sd["fcvar.weight"], sd["fcvar.bias"])
    return _reparam(mu, logvar), mu, logvar


# ------------------------------------------------------------------------------------------------ losses


def lsgan_loss(outputs, target):
    """get_loss_D with nn.MSELoss: mean over scales of mean((o - target)^2).  ref util.py:457-462."""
    return sum(((o - target) ** 2).mean() for o in outputs) / len(outputs)


def domain_loss(probs, onehot):
    """get_domainloss_D with nn.MSELoss.  ref util.py:464-468."""
    return sum(((p - onehot) ** 2).mean() for p in probs) / len(probs)


def conventional_kl(mu, logvar):
    """-0.5 * sum(1 + logvar - mu^2 - exp(logvar)).  ref util_notebook.py:302,632."""
    return -0.5 * torch.sum(1 + logvar - mu ** 2 - logvar.exp())


def batch_kl(mu, n_batch):
    """ref util_notebook.py:316-318,646-648 (unbiased variance scaled once more by n_batch/(n_batch-1))."""
    var = torch.var(mu, dim=0) * n_batch / (n_batch - 1)
    mean = torch.mean(mu, dim=0)
    return -0.5 * torch.sum(1 + torch.log(var) - mean ** 2 - var)


def corrcoef(x):
    """Row-wise correlation matrix (np.corrcoef convention), clamped.  ref util.py:470-511."""
    xm = x - x.mean(1, keepdim=True)
    c = xm @ xm.t() / (x.shape[1] - 1)
    sd = torch.sqrt(torch.diag(c))
    c = c / sd[None, :] / sd[:, None]
    return torch.clamp(c, -1.0, 1.0)


def corr_loss(mu):
    """corrcoef_loss(mu.T).  ref util.py:513-517."""
    d = mu.shape[1]
    return torch.sum(torch.abs(corrcoef(mu.t()) - torch.eye(d))) / (d * (d - 1))


def soft_hist(x, bins=50, lo=-10.0, hi=10.0, sigma=0.2):
    """GaussianHistogram.forward of a vector.  ref util.py:521-537."""
    delta = float(hi - lo) / float(bins)
    centers = float(lo) + delta * (torch.arange(bins).float() + 0.5)
    u = x[None, :] - centers[:, None]
    k = torch.exp(-0.5 * (u / sigma) ** 2) / (sigma * np.sqrt(np.pi * 2)) * delta
    return k.sum(dim=1)


def hist_target(target_num=100000, **kw):
    """histogram_imitation.__init__: consumes `target_num` normals of the CPU generator.  ref util.py:543-545."""
    h = soft_hist(torch.randn(target_num, 1)[:, 0], **kw)
    return h / h.sum() + 1e-8


def hist_loss(mu, target, **kw):
    """histogram_imitation.loss: sum_d KL(target || p_d).  ref util.py:547-553."""
    total = 0
    for d in range(mu.shape[1]):
        h = soft_hist(mu[:, d], **kw)
        p = h / h.sum() + 1e-8
        total = total + F.kl_div(p.log(), target, reduction="sum")
    return total


def l1(a, b):
    return torch.mean(torch.abs(a - b))


# ------------------------------------------------------------------------------------------------ step


class _Net(object):
    """A parameter dictionary + forward function + Adam(0.5, 0.999), stepping with torch-1.4 `.data`
    semantics (the version counter of the parameters is preserved, so a later backward through an older
    graph sees the NEW weights with the OLD activations -- SURVEY F7)."""

    def __init__(self, state_dict, lr=1e-4, trainable=None):
        self.sd = {k: v.detach().clone().float().requires_grad_(True) for k, v in state_dict.items()}
        names = list(self.sd) if trainable is None else [k for k in self.sd if k in trainable]
        self.opt = torch.optim.Adam([self.sd[k] for k in names], lr=lr, betas=(0.5, 0.999))

    def zero_grad(self):
        for v in self.sd.values():
            v.grad = None

    def step(self):
        ps = tuple(self.sd.values())
        with torch.autograd._unsafe_preserve_version_counter(ps):
            self.opt.step()

    def grads(self):
        return {k: (None if v.grad is None else v.grad.detach().clone()) for k, v in self.sd.items()}

    def weights(self):
        return {k: v.detach().clone() for k, v in self.sd.items()}

    def force_weights(self, new):
        """Overwrite the parameters in place WITHOUT bumping version counters (teacher forcing in tests)."""
        ps = tuple(self.sd.values())
        with torch.no_grad(), torch.autograd._unsafe_preserve_version_counter(ps):
            for k, v in new.items():
                self.sd[k].copy_(v)


class OracleTrainer(object):
    """Restatement of SRGAN_training / SingleGAN_training.train().  ref util_notebook.py:28-734.

    kind: "srgan" (Encoder, solo-multi D), "single_solo" (Encoder_original, solo-multi D; notebook 02)
          or "single_multi" (Encoder_original, one original-multi D per class; notebook 01).
    `record` (optional dict) receives gradients / weights captured at every optimizer step.
    """

    def __init__(self, kind, g_sd, d_sd, e_sd, lbd, k, ref_label, n_batch, encoded_feature="latent", ndim=8,
                 classes=(0, 1, 2, 3), g_cfg=(2, 6), d_num_cls=4, e_num_cls=4, lr=(1e-4, 1e-4, 1e-4),
                 e_trainable=None, lr_e=None):
        self.kind, self.lbd, self.k = kind, lbd, k
        self.ref = torch.tensor(np.asarray(ref_label), dtype=torch.float32)
        self.n_batch, self.feature, self.ndim, self.classes = n_batch, encoded_feature, ndim, tuple(classes)
        self.g_cfg, self.d_num_cls, self.e_num_cls = g_cfg, d_num_cls, e_num_cls
        self.G = _Net(g_sd, lr[0])
        self.D = [_Net(s, lr[1]) for s in d_sd] if kind == "single_multi" else _Net(d_sd, lr[1])
        self.E = _Net(e_sd, lr[2] if lr_e is None else lr_e, e_trainable)
        # like the reference, the histogram target is built at construction and draws from the CPU RNG
        self.target = hist_target() if lbd["hist"] > 0 else None
        self.record = None
        self.after_phase1 = None      # test hook: callable(trainer) run right after optG/optE.step() of phase 1

    # -- forward helpers
    def _onehot(self, label):
        return self.ref[torch.as_tensor(label).long()]

    def _g(self, x, label, style):
        return generator_forward(self.G.sd, x, torch.cat([self._onehot(label), style], 1), *self.g_cfg)

    def _e(self, x, label):
        if self.kind == "srgan":
            return encoder_forward(self.E.sd, x, self.e_num_cls)
        return encoder_original_forward(self.E.sd, x, self._onehot(label), self.e_num_cls)

    def _style(self, info):
        return info[0] if self.feature == "latent" else info[1]

    def _d_solo(self, x):
        return discriminator_solo_forward(self.D.sd, x, self.d_num_cls, self.ref.shape[1])

    def _rec(self, key, net):
        if self.record is not None:
            self.record[key + ".grad"] = net.grads()

    def _rec_w(self, key, net):
        if self.record is not None:
            self.record[key + ".weight"] = net.weights()

    # -- discriminator update.  ref :188-251 / :563-594
    def _update_d(self, src, lab, it):
        z = torch.randn(src.shape[0], self.ndim)
        self.target_image, self.c_rand = self._g(src, lab["target"], z), z
        fake = self.target_image.detach()
        if self.kind != "single_multi":
            self.D.zero_grad()
            out, cls = self._d_solo(src)
            err = lsgan_loss(out, 1.0) + domain_loss(cls, self._onehot(lab["source"])) * self.lbd["class"]
            out, _ = self._d_solo(fake)
            err = err + lsgan_loss(out, 0.0)
            err.backward()
            self._rec("D%d" % it, self.D)
            self.D.step()
            self._rec_w("D%d" % it, self.D)
            return err
        for i in self.classes:
            err = 0
            self.D[i].zero_grad()
            real = src[torch.as_tensor(lab["source"]) == i]
            if real.shape[0]:
                err = err + lsgan_loss(discriminator_original_forward(self.D[i].sd, real, self.d_num_cls), 1.0)
            fk = fake[torch.as_tensor(lab["target"]) == i]
            if fk.shape[0]:
                err = err + lsgan_loss(discriminator_original_forward(self.D[i].sd, fk, self.d_num_cls), 0.0)
            if torch.is_tensor(err):
                err.backward()
            self._rec("Dc%d_%d" % (i, it), self.D[i])
            self.D[i].step()
            self._rec_w("Dc%d_%d" % (i, it), self.D[i])
        return err

    # -- generator / encoder update.  ref :253-367 / :596-694
    def _update_ge(self, src, lab):
        lbd = self.lbd
        self.G.zero_grad()
        self.E.zero_grad()
        info = self._e(src, lab["source"])
        recon = self._g(self.target_image, lab["source"], self._style(info))
        if self.kind == "single_multi":
            err_g = 0
            for i in self.classes:
                fk = self.target_image[torch.as_tensor(lab["target"]) == i]
                if fk.shape[0]:
                    err_g = err_g + lsgan_loss(discriminator_original_forward(self.D[i].sd, fk, self.d_num_cls),
                                               1.0) / len(self.classes)
        else:
            out, cls = self._d_solo(self.target_image)
            err_g = lsgan_loss(out, 1.0) + domain_loss(cls, self._onehot(lab["target"])) * lbd["class"]
        cyc = l1(src, recon)
        err_g = err_g + cyc * lbd["cycle"]
        err_e, rep = 0, cyc * lbd["cycle"]
        mu, logvar = info[1], info[2]
        if lbd["KL"] > 0:
            t = conventional_kl(mu, logvar) * lbd["KL"]
            err_e, rep = err_e + t, rep + t
        if lbd["idt"] > 0:
            info2 = self._e(src, lab["source"])
            idt = l1(src, self._g(src, lab["source"], self._style(info2)))
            err_g, rep = err_g + idt * lbd["idt"], rep + idt * lbd["idt"]
        self.stats = {"mu": mu.detach().clone(), "logvar": logvar.detach().clone()}
        if lbd["batch_KL"] > 0:
            t = batch_kl(mu, self.n_batch) * lbd["batch_KL"]
            err_e, rep = err_e + t, rep + t
            if lbd["corr_enc"] > 0:
                t = corr_loss(mu) * lbd["corr_enc"]
                err_e, rep = err_e + t, rep + t
            if lbd["hist"] > 0:
                t = hist_loss(mu, self.target) * lbd["hist"]
                err_e, rep = err_e + t, rep + t
        err_g.backward(retain_graph=True)
        if torch.is_tensor(err_e):
            err_e.backward(retain_graph=True)
        self._rec("G0", self.G)
        self._rec("E0", self.E)
        self.G.step()
        self.E.step()
        self._rec_w("G0", self.G)
        self._rec_w("E0", self.E)
        if self.after_phase1 is not None:
            self.after_phase1(self)
        # phase 2
        self.G.zero_grad()
        self.E.zero_grad()
        tmu = self._e(self.target_image, lab["target"])[1]
        err_x = l1(self.c_rand, tmu) * lbd["reg"]
        if lbd["idt_reg"] * lbd["idt"] > 0:
            if self.kind == "srgan":
                info3 = self._e(src, lab["source"])
                img = self._g(src, lab["source"], self._style(info3))
                reg = l1(info3[1], self._e(img, lab["source"])[1])
            else:
                z = torch.randn(src.shape[0], self.ndim)
                img = self._g(src, lab["source"], z)
                reg = l1(z, self._e(img, lab["source"])[1])
            err_x = err_x + reg * lbd["idt_reg"] * (lbd["idt"] / lbd["cycle"])
        err_x.backward()
        self._rec("G1", self.G)
        self._rec("E_final", self.E)
        self.G.step()
        self._rec_w("G1", self.G)
        return err_g + err_x, rep

    def train(self, src, lab):
        err_d0 = None
        for it in range(self.k):
            e = self._update_d(src, lab, it)
            if it == 0:
                err_d0 = e
        err_g, err_e = self._update_ge(src, lab)
        return [err_g, err_d0, err_e]


def latent_statistics(mu, n_batch, target=None):
    """Mean, (scaled) variance, correlation matrix and soft histograms of a latent batch -- the quantities the
    fused GPU kernel exposes -- computed with the oracle formulas above."""
    out = {
        "mean": mu.mean(0),
        "var": torch.var(mu, dim=0) * n_batch / (n_batch - 1),
        "corr": corrcoef(mu.t()),
        "hist": torch.stack([soft_hist(mu[:, d]) for d in range(mu.shape[1])]),
        "batch_kl": batch_kl(mu, n_batch),
        "corr_loss": corr_loss(mu),
    }
    if target is not None:
        out["hist_loss"] = hist_loss(mu, target)
    return out


def digest(t, samples=32):
    """Compact fingerprint of a tensor: [L2 norm, dot with a fixed probe, `samples` strided elements]."""
    v = t.detach().double().reshape(-1)
    n = v.numel()
    probe = torch.sin(torch.arange(n, dtype=torch.float64) * 0.37 + 0.11)
    idx = torch.linspace(0, n - 1, samples).long()
    return torch.cat([v.norm()[None], (v * probe).sum()[None], v[idx]]).float()


def math_isfinite(x):
    return math.isfinite(float(x))
