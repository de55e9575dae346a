"""Training-step drivers of SingleGAN (notebooks 01/02) and Style-Restricted GAN (notebooks 03/05).

Drop-in for the reference's `pyfiles/util_notebook.py`: `SingleGAN_training` / `SRGAN_training` keep their
constructor signatures, attributes (`G, D, E, optG/D/E, scheG/D/E, lbd, k, n_batch, hi, target_image,
c_rand, ...`) and methods (`opt_sche_initialization, G_transformation, update_D, update_GandE,
UnrolledUpdate, train`), and `train()` returns `[errG, errD, errE]` as 0-dim tensors.

The step is the reference's algorithm (ref pyfiles/util_notebook.py:28-734), including the behaviours that
are easy to "fix" by accident:
  * the discriminator really takes `k` Adam steps per `train()` -- the reference's Unrolled-GAN roll-back
    loads a state_dict that aliases the live parameters, i.e. it is a no-op (ref :721,727);
  * `corr_enc` and `hist` are only evaluated when `batch_KL > 0` (ref :644-662);
  * batch-KL scales the unbiased variance by n_batch/(n_batch-1) once more, with n_batch the configured
    batch size (ref :646);
  * phase 2 (`errG_ex`) back-propagates through the generator graph built in `update_D` AFTER
    `optG.step()` changed the weights in place: gradients use the new weights with the old activations
    (torch-1.4 semantics, see srgan_ops docstring);
  * all noise (z, eps) is drawn from the CPU default generator (ref :179,554; model.py:400,461).

New here: one-process-per-GPU data parallelism.  When `torch.distributed` is initialised every rank holds a
replica, works on its slice of the global batch, gradients are all-reduced (mean) after each backward, and
the latent batch statistics are computed on the ALL-GATHERED mu so batch-KL / correlation / histogram
losses equal their single-GPU global-batch values bit for bit.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

import srgan_ops as ops
from util import *  # noqa: F401,F403


def get_adjustable_parameters(notebook_no=1):
    """The hyper-parameter table notebook 01 iterates over (the other notebooks have none)."""
    if notebook_no != 1:
        return None
    import pandas as pd
    rows = [["conventionalKL", 1, 0], ["preposedKL", 1, 0], ["preposedKL", 5, 0.5]]
    return pd.DataFrame(np.array(rows), columns=["restriction_type", "unrolled_k", "idt_reg"])


# ------------------------------------------------------------------------------------------- data parallel
_BATCH_FAKES_MAX = int(os.environ.get("SRGAN_BATCH_FAKES_MAX", "128"))     # largest (k-1) x batch that is generated in one pass
_NO_BATCH_FAKES = os.environ.get("SRGAN_DBG_NO_BATCH_FAKES", "0") != "0"    # bring-up: one generator pass per D update
_SPLIT_D = os.environ.get("SRGAN_DBG_SPLIT_D", "0") != "0"           # bring-up: D(real) and D(fake) as two passes, like the reference
_REENCODE = os.environ.get("SRGAN_DBG_REENCODE", "0") != "0"     # bring-up: second encoder pass of phase 1, like the reference
_SPLIT_BACKWARD = os.environ.get("SRGAN_DBG_SPLIT_BACKWARD", "0") != "0"     # bring-up: the reference's two calls
_GEN_PIPELINE = os.environ.get("SRGAN_GEN_PIPELINE", "0") != "0"    # generator passes of the D updates on a second stream


def _world():
    return ops.dp_rank_world()


def _allreduce_mean_(t):
    """In-place mean over ranks of a dense tensor."""
    _, world = _world()
    if world == 1:
        return t
    if dist.get_backend() == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.div_(world)
    return t


def _sync_grads(module, optimizer=None):
    """All-reduce (mean) the gradients of `module`: one collective over a flat buffer."""
    _, world = _world()
    if world == 1:
        return
    flat_bufs = None
    if isinstance(optimizer, ops.FusedAdam):
        grads = [p.grad for p in module.parameters() if p.grad is not None]
        bufs = optimizer.flat_grads()
        # only usable when the gradients already live in the optimizer's flat buffers
        if grads and all(any(g.untyped_storage().data_ptr() == b.untyped_storage().data_ptr() for b in bufs)
                         for g in grads):
            flat_bufs = bufs
    if flat_bufs is not None:
        for b in flat_bufs:
            _allreduce_mean_(b)
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) if g.is_contiguous() else g.contiguous(memory_format=torch.channels_last)
                      .permute(0, 2, 3, 1).reshape(-1) for g in grads])
    _allreduce_mean_(flat)
    o = 0
    for g in grads:
        n = g.numel()
        if g.is_contiguous():
            g.copy_(flat[o:o + n].view_as(g))
        else:
            K, C, R, S = g.shape
            g.copy_(flat[o:o + n].view(K, R, S, C).permute(0, 3, 1, 2))
        o += n


def _zero_grads(module, optimizer=None):
    """`module.zero_grad()` of the reference (pyfiles/util_notebook.py:577,613-614,670-671).  With a FusedAdam the
    gradients of its parameters are zeroed in the optimizer's flat buffer and stay views of it, so autograd
    accumulates straight into the buffer that the all-reduce and the Adam kernel read (no per-tensor copies);
    parameters the optimizer does not own (notebook 05: the melted encoder trunk) are reset as usual."""
    if isinstance(optimizer, ops.FusedAdam):
        optimizer.zero_grad()
        owned = optimizer.owned_ids()
        for p in module.parameters():
            if id(p) not in owned:
                p.grad = None
    else:
        module.zero_grad()


def _unwrap(net):
    """The notebooks hand over nn.DataParallel wrappers; compute runs on the wrapped module."""
    return net.module if isinstance(net, torch.nn.DataParallel) else net


class _EncodedStyle(list):
    """`[latent, mu, logvar, ...]` exactly as the reference returns it (a plain list subclass)."""


# ------------------------------------------------------------------------------------------- shared engine
class _UnrolledTrainer(object):
    """Machinery shared by both trainers.  Subclasses define how the encoder and discriminator are called."""

    # ---- construction -------------------------------------------------------------------------------
    def _common_init(self, net, opt, criterion, lbd, unrolled_k, device, ref_label, batch_size,
                     encoded_feature, ndim):
        # nn.DataParallel wrappers (the notebooks pass them) stay visible as sg.G / sg.D / sg.E -- the notebooks
        # save `sg.G.module.state_dict()` -- but compute runs on the wrapped modules: parallelism here is one
        # process per GPU over torch.distributed, not single-process scatter/gather.
        self._nG, self._nE = _unwrap(self.G), _unwrap(self.E)
        self._nD = [_unwrap(d) for d in self.D] if isinstance(self.D, (list, tuple)) else _unwrap(self.D)
        self.optG, self.optD, self.optE = opt[0], opt[1], opt[2]
        self.scheG, self.scheD, self.scheE = None, None, None
        self.criterion, self.criterion_class = criterion
        self.lbd = lbd
        self.k = unrolled_k
        self.device = device
        self.ref_label = ref_label
        self.n_batch = batch_size
        self.encoded_feature = encoded_feature
        self.ndim = ndim
        self.source_image = None
        self.target_image = None
        self.recon_image = None
        self.label = None
        self.c_rand = None
        self.enc_info = None
        self.target_cenc = None
        self.latent_stats = None        # statistics blob of the last restriction-loss evaluation (ops.latent_stats_views)
        if lbd["hist"] > 0:
            self.hi = histogram_imitation(device)
        # Three step shortcuts (one no-grad generator pass for the k-1 early fakes, D(real) and D(fake) as one pass,
        # reuse of the source batch's encoder output in phase 1) are exact only while no layer couples the samples
        # of a batch: with norm_type="batch" (CBBNorm2d / BatchNorm2d) the batch statistics and the running-statistics
        # updates would differ from the reference's call pattern, so the reference's pattern is executed instead.
        import model as _model
        coupled = (_model._CBBNorm, torch.nn.modules.batchnorm._BatchNorm)
        nets_d = self._nD if isinstance(self._nD, (list, tuple)) else [self._nD]
        self._coupled = {name: any(isinstance(m, coupled) for n in nets for m in n.modules())
                         for name, nets in (("G", [self._nG]), ("E", [self._nE]), ("D", nets_d))}
        if isinstance(self._nD, (list, tuple)) and _world()[1] > 1:
            raise NotImplementedError(
                "one discriminator per class (singleD=False) is not data-parallel: the per-class sub-batches differ "
                "between ranks, so the D[i] replicas would drift apart; run notebook 01 on one GPU or use singleD=True")

    @staticmethod
    def _adam(module, lr):
        return ops.FusedAdam(module.parameters(), lr=lr, betas=(0.5, 0.999))

    @staticmethod
    def _sched(optimizer):
        return torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=0.95)

    # ---- collectives overlapped with compute (data parallel only) --------------------------------------
    # The gradient all-reduce of a network is followed by that network's Adam step and then by the next forward that
    # needs the new weights; whatever the step can compute in between without those weights hides the collective:
    #   * D's all-reduce + Adam of update i      ||  the generator pass that makes the fake batch of update i + 1
    #                                                (after the last update: encoder + generator pass of phase 1);
    #   * the all-gather of mu (batch statistics) ||  the reconstruction pass of G and the D pass of phase 1;
    #   * the all-reduce of the reported losses   ||  the backward pass of phase 2.
    # Collective + dependent optimizer step run on a side stream; `_comm_join` makes the compute stream wait for it
    # right before the first kernel that touches the network again (its next forward or the zeroing of its gradient
    # buffer).  Under CUDA-graph capture the fork / join become graph edges.  Results are unchanged: the same kernels
    # run on the same data, only their order relative to independent work differs.
    _OVERLAP = os.environ.get("SRGAN_DBG_NO_COMM_OVERLAP", "0") == "0"

    def _comm_fork(self):
        """Side stream that has waited for everything issued so far on the compute stream, or None (single process /
        CPU tensors / overlap disabled): the caller then runs the collective inline."""
        if not self._OVERLAP or _world()[1] == 1 or not torch.cuda.is_available():
            return None
        dev = torch.device(self.device)
        if dev.type != "cuda":
            return None
        if getattr(self, "_comm_stream", None) is None:
            self._comm_stream = torch.cuda.Stream(dev)
        self._comm_stream.wait_stream(torch.cuda.current_stream(dev))
        self._comm_pending = True
        return self._comm_stream

    def _comm_join(self):
        if getattr(self, "_comm_pending", False):
            torch.cuda.current_stream(torch.device(self.device)).wait_stream(self._comm_stream)
            self._comm_pending = False

    def _reduce_and_step(self, pairs):
        """All-reduce the gradients of every (module, optimizer) in `pairs` and step the optimizers - on the side
        stream when one is available (see above), else inline."""
        side = self._comm_fork()
        if side is None:
            for module, opt in pairs:
                _sync_grads(module, opt)
            for _, opt in pairs:
                opt.step()
            return
        with torch.cuda.stream(side):
            for module, opt in pairs:
                _sync_grads(module, opt)
            for _, opt in pairs:
                opt.step()

    # ---- small helpers --------------------------------------------------------------------------------
    def _onehot(self, label):
        return class_encode(label, self.device, self.ref_label)

    def _style_of(self, info):
        if self.encoded_feature == "latent":
            return info[0]
        if self.encoded_feature == "mu":
            return info[1]
        raise ValueError("encoded_feature must be 'latent' or 'mu'")

    def _generate(self, label, image, style):
        cond = torch.cat([self._onehot(label), style], 1)
        return self._nG(image, cond)

    def G_transformation(self, target_label, source_image, encoder=False, ref_image=None):
        """Translate `source_image` to `target_label`.  Style = encoder output of `ref_image` when `encoder`,
        else fresh N(0,1) noise.  Returns (image, info): info is [latent, mu, logvar(, ...)] or the noise."""
        if encoder:
            info = self._encode(ref_image, target_label)
            style = self._style_of(info)
        else:
            style = ops.host_normal(source_image.shape[0], self.ndim, self.device)
            info = style
        return self._generate(target_label, source_image, style), info

    def _latent_restriction(self, info):
        """KL / batch-KL / correlation / histogram terms (already weighted by their lambdas) on the encoder
        output of the source batch, all from ONE fused kernel.  ref :299-332 / :629-662."""
        lbd = self.lbd
        flags = 0
        if lbd["KL"] > 0:
            flags |= ops.LAT_KL
        if lbd["batch_KL"] > 0:
            flags |= ops.LAT_BKL
            if lbd["corr_enc"] > 0:
                flags |= ops.LAT_CORR
            if lbd["hist"] > 0:
                flags |= ops.LAT_HIST
        if not flags:
            return {}
        mu, logvar = info[1], info[2]
        rank, world = _world()
        mu_all = lv_all = None
        if world > 1:
            pre = getattr(self, "_gathered", None)
            if pre is not None and pre[0] is info:
                self._comm_join()                       # issued right after the encoder pass (_prefetch_latent_stats)
                mu_all, lv_all = pre[1], pre[2]
                self._gathered = None
            else:
                mu_all, lv_all = self._gather_latent(mu, logvar, flags)
        hi = getattr(self, "hi", None)
        kw = {}
        if flags & ops.LAT_HIST:
            g = hi.gausshist
            kw = dict(target=hi.target, bins=g.bins, hmin=g.min, hmax=g.max, sigma=g.sigma)
        losses, self.latent_stats = ops.latent_losses(mu, logvar if flags & ops.LAT_KL else None, n_cfg=self.n_batch,
                                                      flags=flags, mu_all=mu_all, logvar_all=lv_all,
                                                      row0=rank * mu.shape[0], **kw)
        terms = {}
        if flags & ops.LAT_KL:
            terms["KL"] = losses[3] * lbd["KL"]
        if flags & ops.LAT_BKL:
            terms["batch_KL"] = losses[0] * lbd["batch_KL"]
        if flags & ops.LAT_CORR:
            terms["corr_enc"] = losses[1] * lbd["corr_enc"]
        if flags & ops.LAT_HIST:
            terms["hist"] = losses[2] * lbd["hist"]
        return terms

    def _restriction_flags(self):
        lbd, flags = self.lbd, 0
        if lbd["KL"] > 0:
            flags |= ops.LAT_KL
        if lbd["batch_KL"] > 0:
            flags |= ops.LAT_BKL
        return flags

    @staticmethod
    def _gather_latent(mu, logvar, flags, side=None):
        """The statistics are those of the GLOBAL batch: all-gather mu (B_global x ndim floats; logvar too for the
        conventional KL term).  Buffers are allocated on the calling (compute) stream, which also consumes them; only
        the collectives run on `side`."""
        world = _world()[1]
        kl = bool(flags & ops.LAT_KL)
        mu_loc = mu.detach().contiguous()
        lv_loc = logvar.detach().contiguous() if kl else None
        mu_all = torch.empty((mu.shape[0] * world, mu.shape[1]), dtype=mu.dtype, device=mu.device)
        lv_all = torch.empty_like(mu_all) if kl else None

        def run():
            dist.all_gather_into_tensor(mu_all, mu_loc)
            if kl:
                dist.all_gather_into_tensor(lv_all, lv_loc)
        if side is None:
            run()
        else:
            with torch.cuda.stream(side):
                run()
        return mu_all, lv_all

    def _prefetch_latent_stats(self, info):
        """Start the all-gather of the encoder output as soon as it exists; `_latent_restriction` picks it up."""
        flags = self._restriction_flags()
        if not flags or _world()[1] == 1:
            return
        mu_loc = info[1].detach().contiguous()          # materialised on the compute stream before the fork
        side = self._comm_fork()
        if side is None:
            return
        mu_all, lv_all = self._gather_latent(mu_loc, info[2], flags, side)
        self._gathered = (info, mu_all, lv_all)

    # ---- the step --------------------------------------------------------------------------------------
    def _early_fakes(self):
        """Translated batches for the first k-1 discriminator updates.  The generator does not change during the k
        updates of `UnrolledUpdate` and these batches are only ever used detached (the reference builds and drops
        their graphs, pyfiles/util_notebook.py:716-722), so they come from ONE no-grad generator pass over the k-1
        noise draws - drawn in the reference's order: nothing else consumes the CPU generator between two
        `update_D` calls.  Returns a list of (image, noise), or None when there is nothing to batch."""
        if _NO_BATCH_FAKES or self.k < 2 or isinstance(self._nD, (list, tuple)) or self._coupled["G"]:
            return None
        src, lab, n = self.source_image, self.label["target"], self.k - 1
        B = src.shape[0]
        if B * n > _BATCH_FAKES_MAX:
            return None          # measured: +7 % images/s at batch 8, -0.6 % at batch 64 (the kernels are full anyway)
        with torch.no_grad():
            styles = [ops.host_normal(B, self.ndim, self.device) for _ in range(n)]
            onehot = self._onehot(lab)
            cond = torch.cat([torch.cat([onehot, z], 1) for z in styles], 0)
            images = self._nG(torch.cat([src] * n, 0), cond)
        return [(images[i * B:(i + 1) * B], styles[i]) for i in range(n)]

    def _pipelined_fakes(self):
        """The k generator passes of the k discriminator updates, issued up front on a second stream: the generator
        does not change during `UnrolledUpdate`'s loop and the discriminator does not consume the CPU generator, so
        pass i + 1 (full-GPU convolutions) can run next to discriminator update i, whose many small kernels leave SMs
        idle.  Same kernels, same data, same noise order - results are bit-identical to the sequential schedule; under
        CUDA-graph capture the two streams become parallel branches of the graph.  The LAST pass keeps its autograd
        graph (phase 2 of `update_GandE` back-propagates through it - on this stream, autograd follows the forward).
        Returns [(image, noise, event)] or None (CPU tensors, one discriminator per class, batch-coupled norms,
        SRGAN_GEN_PIPELINE=0)."""
        if not _GEN_PIPELINE or self.k < 2 or isinstance(self._nD, (list, tuple)) or self._coupled["G"]:
            return None
        src = self.source_image
        if not (torch.is_tensor(src) and src.is_cuda):
            return None
        dev = src.device
        main = torch.cuda.current_stream(dev)
        gen = getattr(self, "_gen_stream", None)
        if gen is None:
            gen = self._gen_stream = torch.cuda.Stream(dev)
        gen.wait_stream(main)
        src.record_stream(gen)
        out = []
        with torch.cuda.stream(gen):
            for i in range(self.k):
                with torch.set_grad_enabled(i == self.k - 1):
                    image, noise = self.G_transformation(self.label["target"], src, False)
                ev = torch.cuda.Event()
                ev.record(gen)
                out.append((image, noise, ev))
        return out

    def _take_fake(self, item):
        """Make the compute stream wait for a pipelined generator pass; returns (image, noise)."""
        image, noise, ev = item
        main = torch.cuda.current_stream(image.device)
        main.wait_event(ev)
        image.record_stream(main)
        noise.record_stream(main)
        return image, noise

    def _unrolled_D(self):
        """The k discriminator updates of `UnrolledUpdate`; returns the first update's loss."""
        early = self._early_fakes()
        piped = self._pipelined_fakes() if not early else None
        errorD = None
        for i in range(self.k):
            if piped is not None:
                fake = self._take_fake(piped[i])
            else:
                fake = early[i] if early and i < self.k - 1 else None
            errD = self.update_D(keep_graph=(i == self.k - 1), fake=fake)
            if i == 0:
                errorD = errD
                # (the reference snapshots D.state_dict() here and reloads it after the loop; the snapshot aliases the
                #  live parameters, so nothing is rolled back -- reproduced by doing nothing)
        return errorD

    def _real_and_fake(self, src, fake):
        """[real ; fake] as one batch in a buffer that lives across updates: two contiguous copies instead of a
        torch.cat of two channels-last 3-channel tensors (5 x 83 us per step in the launch list).  The discriminator
        pass that read the previous contents has been back-propagated before the buffer is written again."""
        fake = ops.to_nhwc(fake)
        if not (src.is_cuda and src.dtype == fake.dtype and src.shape == fake.shape):
            return torch.cat([src, fake], 0)
        B = src.shape[0]
        buf = getattr(self, "_d_both", None)
        if buf is None or buf.shape[0] != 2 * B or buf.shape[1:] != src.shape[1:] or buf.device != src.device:
            buf = self._d_both = torch.empty((2 * B,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device,
                                             memory_format=torch.channels_last)
        buf[:B].copy_(src)
        buf[B:].copy_(fake)
        return buf

    def _solo_D_loss(self, fake):
        """LSGAN + class loss of the single (solo-multi) discriminator on the real batch and on `fake`
        (ref pyfiles/util_notebook.py:582-589).  The discriminator has no batch-coupled layer (convolutions, LeakyReLU,
        average pooling), so D(real) and D(fake) are evaluated as ONE pass over the concatenated batch: every
        output is bit-identical to the two separate passes, the weight gradients differ only in summation order, and
        the many small layers of D run at twice the batch (fewer, better filled launches)."""
        src = self.source_image
        if _SPLIT_D or self._coupled["D"]:
            output, output_class = self._nD(src)
            out_fake, _ = self._nD(fake)
        else:
            B = src.shape[0]
            out_all, cls_all = self._nD(self._real_and_fake(src, fake))
            output, output_class = [o[:B] for o in out_all], [c[:B] for c in cls_all]
            out_fake = [o[B:].clone() for o in out_all]      # fresh storage: the loss kernels need 16-byte alignment
        errD = get_loss_D(output, 1., self.criterion, self.device) + \
            get_domainloss_D(output_class, self._onehot(self.label["source"]), self.criterion_class) \
            * self.lbd["class"]
        return errD + get_loss_D(out_fake, 0., self.criterion, self.device)

    def update_GandE(self):
        """Phase 1: one step of G and E on the SingleGAN losses; phase 2: one more step of G alone on the
        latent-regression losses.  Returns [errG, errE_output]."""
        lbd = self.lbd
        src, lab = self.source_image, self.label
        _zero_grads(self._nG, self.optG)
        _zero_grads(self._nE, self.optE)

        # == G_transformation(lab["source"], self.target_image, True, src), with the all-gather of the encoder output
        # started between the encoder and the generator pass.  Nothing here reads D: the all-reduce + Adam step of the
        # last discriminator update may still be running on the side stream ...
        enc_info = self._encode(src, lab["source"])
        self._prefetch_latent_stats(enc_info)
        recon_image = self._generate(lab["source"], self.target_image, self._style_of(enc_info))
        self._comm_join()                               # ... and is waited for here, before D's next forward
        errG = self._fool_D(self.target_image, lab["target"])
        cyc = ops.l1_mean(src, recon_image)
        errG = errG + cyc * lbd["cycle"]
        errE_output = cyc * lbd["cycle"]

        # The restriction terms do not touch the RNG, so evaluating them in one fused launch here (after the
        # identity pass) is equivalent to the reference's interleaving; they are ADDED in the reference's order.
        if lbd["idt"] > 0:
            if _REENCODE or self._coupled["E"]:
                identity_image, _ = self.G_transformation(lab["source"], src, True, src)
            else:
                # The reference encodes the source batch a second time here (pyfiles/util_notebook.py:637-641).  The
                # encoder is deterministic and its weights have not changed since `enc_info` was computed a few lines
                # up, so mu / logvar would come out bit-identical: reuse them and only draw the fresh reparametrisation
                # noise (same CPU-RNG consumption).  Both uses then back-propagate through ONE encoder graph - one
                # encoder forward and one backward traversal less per step, same gradients up to fp32 summation order.
                z2 = self._nE.reparametrize(enc_info[1], enc_info[2])
                info2 = _EncodedStyle([z2] + list(enc_info[1:]))
                identity_image = self._generate(lab["source"], src, self._style_of(info2))
            idt = ops.l1_mean(src, identity_image)
            errG = errG + idt * lbd["idt"]
        terms = self._latent_restriction(enc_info)
        errE = 0
        if "KL" in terms:
            errE = errE + terms["KL"]
            errE_output = errE_output + terms["KL"]
        if lbd["idt"] > 0:
            errE_output = errE_output + idt * lbd["idt"]
        for key in ("batch_KL", "corr_enc", "hist"):
            if key in terms:
                errE = errE + terms[key]
                errE_output = errE_output + terms[key]
        _, world = _world()
        # Every rank evaluates the GLOBAL restriction loss but differentiates only its own rows: those
        # gradients must be summed over ranks while _sync_grads averages -> pre-scale by the world size.
        restrict_bp = errE * world if (world > 1 and torch.is_tensor(errE)) else errE

        # The reference calls errG.backward(retain_graph=True) and then errE.backward(retain_graph=True)
        # (pyfiles/util_notebook.py:664-665): the second call walks the encoder graph of the source batch again.
        # Gradients are linear in the loss, so ONE backward of the sum leaves the same G / E / D gradients (up to
        # fp32 summation order) and saves an encoder backward traversal per step.
        if _SPLIT_BACKWARD:
            errG.backward(retain_graph=True)
            if torch.is_tensor(restrict_bp):
                restrict_bp.backward(retain_graph=True)
        else:
            total = errG + restrict_bp if torch.is_tensor(restrict_bp) else errG
            with ops.direct_param_grads():
                total.backward(retain_graph=True)
        self._reduce_and_step([(self._nG, self.optG), (self._nE, self.optE)])
        self._comm_join()
        hook = getattr(self, "_after_phase1", None)     # test hook (teacher forcing of the post-step weights)
        if hook is not None:
            hook(self)

        # ---- phase 2: G only (E receives gradients but is not stepped) ----
        _zero_grads(self._nG, self.optG)
        _zero_grads(self._nE, self.optE)
        target_mu = self._encode(self.target_image, lab["target"])[1]
        errG_ex = ops.l1_mean(self.c_rand, target_mu) * lbd["reg"]
        if lbd["idt_reg"] * lbd["idt"] > 0:
            errG_ex = errG_ex + self._identity_regression() * lbd["idt_reg"] * (lbd["idt"] / lbd["cycle"])
        out = [errG + errG_ex, errE_output]
        self._report_async(out)                         # the reported scalars are final: reduce them behind the backward
        with ops.direct_param_grads():
            errG_ex.backward()
        self._reduce_and_step([(self._nG, self.optG)])
        self._comm_join()
        return out

    def train(self, source_image, label):
        if getattr(self, "_graph", None) is not None and torch.is_tensor(source_image) and source_image.is_cuda:
            return self._train_graphed(source_image, label)
        self.source_image = ops.to_nhwc(source_image)
        self.label = label
        return self.UnrolledUpdate()

    # ---- CUDA-graph replay of the whole step ----------------------------------------------
    def enable_cuda_graph(self, warmup=2):
        """After `warmup` eager calls, `train()` captures the complete step (k discriminator updates, both
        generator/encoder phases, optimizer steps, gradient all-reduces) into ONE CUDA graph and replays it:
        a step then costs one graph launch plus the upload of its host-drawn noise instead of ~4000 kernel
        launches issued from Python.  The eager and the replayed step run the same kernels in the same order on
        the same noise (the CPU generator is consumed identically), so their results are bit-identical.
        Re-captured when the batch shape changes.  Not available with per-class discriminators (data-dependent
        sub-batch sizes)."""
        if isinstance(self._nD, (list, tuple)):
            raise NotImplementedError("CUDA-graph replay needs static shapes: not with one discriminator per class")
        for o in (self.optG, self.optD, self.optE):
            if not isinstance(o, ops.FusedAdam):
                raise NotImplementedError(
                    "CUDA-graph replay needs optimizers whose step is a capturable kernel: build them with "
                    "srgan_ops.FusedAdam (same constructor as torch.optim.Adam) or let opt_sche_initialization() do it")
        self._graph = dict(warmup=int(warmup), calls=0, state=None)
        return self

    def disable_cuda_graph(self):
        self._graph = None

    def _train_graphed(self, source_image, label):
        cfg = self._graph
        dev = source_image.device
        if cfg.get("stream") is None:
            cfg["stream"] = torch.cuda.Stream(dev)
        side = cfg["stream"]
        if cfg["calls"] < cfg["warmup"]:
            # warm-up on the capture stream: autograd's gradient-accumulation nodes remember the stream they were
            # created on, and a node that runs on the default stream would invalidate the capture
            cfg["calls"] += 1
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self.source_image = ops.to_nhwc(source_image)
                self.label = label
                out = self.UnrolledUpdate()
            torch.cuda.current_stream(dev).wait_stream(side)
            return out
        st = cfg["state"]
        key = (tuple(source_image.shape), str(dev))
        if st is None or st["key"] != key:
            st = dict(key=key, tape=ops.FeedTape(dev), graph=torch.cuda.CUDAGraph(),
                      x=ops.to_nhwc(source_image).clone(),
                      src=ops.to_device_async(label["source"], dev, torch.long).clone(),
                      tgt=ops.to_device_async(label["target"], dev, torch.long).clone())
            # nothing may keep the previous step's autograd graph (and its accumulation nodes) alive
            self.target_image = self.c_rand = self.source_image = self.label = None
            self._errD_first = self._reported = self._gathered = None
            cfg["state"] = None
            import gc
            gc.collect()
            torch.cuda.synchronize(dev)
            # the captured kernels keep raw addresses of their scratch: the graph gets its own set, allocated inside
            # the capture (graph memory pool) and held by `st` for as long as the graph lives
            st["scratch"] = ops.private_scratch()
            with ops.recording(st["tape"]), st["scratch"]:
                # data parallel: NCCL's watchdog thread polls CUDA events while this thread captures; only the
                # capturing thread may be restricted
                mode = "thread_local" if _world()[1] > 1 else "global"
                with torch.cuda.graph(st["graph"], stream=side, capture_error_mode=mode):
                    self.source_image = st["x"]
                    self.label = {"source": st["src"], "target": st["tgt"]}
                    st["out"] = self.UnrolledUpdate()
            cfg["state"] = st
        else:
            st["x"].copy_(ops.to_nhwc(source_image), non_blocking=True)
            st["src"].copy_(ops.to_device_async(label["source"], dev, torch.long), non_blocking=True)
            st["tgt"].copy_(ops.to_device_async(label["target"], dev, torch.long), non_blocking=True)
        st["tape"].upload()
        st["graph"].replay()
        return [e.clone() if torch.is_tensor(e) else e for e in st["out"]]

    def _report_async(self, g_and_e):
        """Called by update_GandE once errG / errE are final (before the phase-2 backward): average [errG, errD, errE]
        over ranks on the side stream.  `_report` returns the result."""
        self._reported = None
        errD = getattr(self, "_errD_first", None)
        if errD is None or _world()[1] == 1:
            return
        errs = [g_and_e[0], errD, g_and_e[1]]
        if not (self._OVERLAP and torch.cuda.is_available() and torch.device(self.device).type == "cuda"):
            return
        self._reported = (self._report_now(errs, overlap=True), errs)

    def _report(self, errs):
        """Average the reported scalars over ranks so they equal the global-batch values."""
        pre, self._reported = getattr(self, "_reported", None), None
        if pre is not None and pre[1][0] is errs[0] and pre[1][2] is errs[2]:
            self._comm_join()
            return pre[0]
        return self._report_now(errs)

    def _report_now(self, errs, overlap=False):
        _, world = _world()
        if world == 1:
            return errs
        idx = [i for i, e in enumerate(errs) if torch.is_tensor(e)]
        if not idx:
            return errs
        # ONE collective for all reported scalars (packed on the compute stream; the collective may run on the side)
        packed = torch.stack([errs[i].detach().float().reshape(()) for i in idx])
        side = self._comm_fork() if overlap else None
        if side is None:
            _allreduce_mean_(packed)
        else:
            with torch.cuda.stream(side):
                _allreduce_mean_(packed)
        out = list(errs)
        for j, i in enumerate(idx):
            out[i] = packed[j]
        return out


# ------------------------------------------------------------------------------------------- SingleGAN (nb 01/02)
class SingleGAN_training(_UnrolledTrainer):
    """SingleGAN with a class-conditioned encoder `E(image, onehot)`; `singleD=False` uses one patch
    discriminator per class, fed the samples of that class only.  ref pyfiles/util_notebook.py:28-417."""

    def __init__(self, net, opt, criterion, lbd, unrolled_k, device, ref_label, ndim,
                 classes, batch_size=64, encoded_feature="latent", singleD=False):
        self.G, self.D, self.E = net[0], net[1], net[2]
        self._common_init(net, opt, criterion, lbd, unrolled_k, device, ref_label, batch_size, encoded_feature,
                          ndim)
        self.classes = classes
        self.singleD = singleD

    def opt_sche_initialization(self, lr=[0.0001, 0.0001, 0.0001]):
        lr_G, lr_D, lr_E = lr
        if self.optG is None:
            self.optG = self._adam(self._nG, lr_G)
        self.scheG = self._sched(self.optG)
        if self.singleD:
            if self.optD is None:
                self.optD = self._adam(self._nD, lr_D)
            self.scheD = self._sched(self.optD)
        else:
            self.optD = [self._adam(self._nD[i], lr_D) for i in self.classes]
            self.scheD = [self._sched(o) for o in self.optD]
        if self.optE is None:
            self.optE = self._adam(self._nE, lr_E)
        self.scheE = self._sched(self.optE)

    def _encode(self, image, label):
        return _EncodedStyle(self._nE(image, self._onehot(label)))

    def _class_mask(self, which, i):
        return ops.to_device_async(self.label[which], self.device) == i

    def _fool_D(self, fake, fake_label):
        if self.singleD:
            output, output_class = self._nD(fake)
            return get_loss_D(output, 1., self.criterion, self.device) + \
                get_domainloss_D(output_class, self._onehot(fake_label), self.criterion_class) * self.lbd["class"]
        err = 0
        for i in self.classes:
            sub = fake[self._class_mask("target", i)]
            if sub.shape[0] != 0:
                err = err + get_loss_D(self._nD[i](sub), 1., self.criterion, self.device) / len(self.classes)
        return err

    def update_D(self, keep_graph=True, fake=None):
        if fake is not None:
            self.target_image, self.c_rand = fake          # pre-generated without a graph (_early_fakes)
        else:
            with torch.set_grad_enabled(keep_graph):
                self.target_image, self.c_rand = self.G_transformation(self.label["target"], self.source_image, False)
        fake = self.target_image.detach()
        if self.singleD:
            self._comm_join()                  # the previous update's all-reduce + Adam step (overlapped the G pass above)
            _zero_grads(self._nD, self.optD)
            errD = self._solo_D_loss(fake)
            with ops.direct_param_grads():
                errD.backward()
            self._reduce_and_step([(self._nD, self.optD)])
            return errD
        # one discriminator per class; like the reference, the LAST class's loss is what gets returned
        for i in self.classes:
            errD = 0
            _zero_grads(self._nD[i], self.optD[i])
            real = self.source_image[self._class_mask("source", i)]
            if real.shape[0] != 0:
                errD = errD + get_loss_D(self._nD[i](real), 1., self.criterion, self.device)
            sub = fake[self._class_mask("target", i)]
            if sub.shape[0] != 0:
                errD = errD + get_loss_D(self._nD[i](sub), 0., self.criterion, self.device)
            if torch.is_tensor(errD):
                with ops.direct_param_grads():
                    errD.backward()
            self.optD[i].step()
        return errD

    def _identity_regression(self):
        image, z = self.G_transformation(self.label["source"], self.source_image, False)
        mu = self._encode(image, self.label["source"])[1]
        return ops.l1_mean(z, mu)

    def UnrolledUpdate(self):
        errorD = self._unrolled_D()
        self._errD_first = errorD.detach() if torch.is_tensor(errorD) else errorD    # value only: no graph is kept
        errorG, errorE = self.update_GandE()
        self._comm_join()
        return self._report([errorG, errorD, errorE])


# ------------------------------------------------------------------------------------------- SRGAN (nb 03/05)
class SRGAN_training(_UnrolledTrainer):
    """Style-Restricted GAN: unconditional encoder `E(image) -> (z, mu, logvar, class_logits, None)`, one
    discriminator with patch + class heads.  ref pyfiles/util_notebook.py:419-734."""

    def __init__(self, net, opt, criterion, lbd, unrolled_k, device, ref_label,
                 batch_size=64, encoded_feature="latent", ndim=8):
        self.G, self.D, self.E = net[0].to(device), net[1].to(device), net[2].to(device)
        self._common_init(net, opt, criterion, lbd, unrolled_k, device, ref_label, batch_size, encoded_feature,
                          ndim)

    def opt_sche_initialization(self, lr=[0.0001, 0.0001, 0.0001]):
        lr_G, lr_D, lr_E = lr
        if self.optG is None:
            self.optG = self._adam(self._nG, lr_G)
        self.scheG = self._sched(self.optG)
        if self.optD is None:
            self.optD = self._adam(self._nD, lr_D)
        self.scheD = self._sched(self.optD)
        if self.optE is None:
            self.optE = self._adam(self._nE, lr_E)
        self.scheE = self._sched(self.optE)

    def _encode(self, image, label=None):
        return _EncodedStyle(self._nE(image))

    def _fool_D(self, fake, fake_label):
        output, output_class = self._nD(fake)
        return get_loss_D(output, 1., self.criterion, self.device) + \
            get_domainloss_D(output_class, self._onehot(fake_label), self.criterion_class) * self.lbd["class"]

    def update_D(self, keep_graph=True, fake=None):
        if fake is not None:
            self.target_image, self.c_rand = fake          # pre-generated without a graph (_early_fakes)
        else:
            with torch.set_grad_enabled(keep_graph):
                self.target_image, self.c_rand = self.G_transformation(self.label["target"], self.source_image, False)
        self._comm_join()                      # the previous update's all-reduce + Adam step (overlapped the G pass above)
        _zero_grads(self._nD, self.optD)
        errD = self._solo_D_loss(self.target_image.detach())
        with ops.direct_param_grads():
            errD.backward()
        self._reduce_and_step([(self._nD, self.optD)])
        return errD

    def _identity_regression(self):
        image, info = self.G_transformation(self.label["source"], self.source_image, True, self.source_image)
        mu = self._encode(image)[1]
        return ops.l1_mean(info[1], mu)

    def UnrolledUpdate(self):
        errorD = self._unrolled_D()
        self._errD_first = errorD.detach() if torch.is_tensor(errorD) else errorD    # value only: no graph is kept
        errorG, errorE = self.update_GandE()
        self._comm_join()
        return self._report([errorG, errorD, errorE])


# ------------------------------------------------------------------------------------------- checkpoints
def save_checkpoint(sg, path, **extra):
    """Everything a resumed run needs (the reference only ever saves `sg.G.module.state_dict()` etc. and
    cannot resume): network weights under the REFERENCE's state_dict keys (each entry loads into the reference's
    modules and vice versa), Adam moments / step counters / learning rates, scheduler state and the CPU RNG states
    the step draws its noise from."""
    def sd(net):
        if isinstance(net, (list, tuple)):
            return [sd(n) for n in net]
        return {k: v.detach().cpu() for k, v in _unwrap(net).state_dict().items()}

    def opt_state(o):
        if isinstance(o, (list, tuple)):
            return [opt_state(x) for x in o]
        if o is None:
            return None
        return {"fused": o.flat_state()} if isinstance(o, ops.FusedAdam) else {"torch": o.state_dict()}

    def sch_state(s_):
        if isinstance(s_, (list, tuple)):
            return [sch_state(x) for x in s_]
        return None if s_ is None else s_.state_dict()
    state = {"format": "srgan_b200.checkpoint.v1",
             "G": sd(sg.G), "D": sd(sg.D), "E": sd(sg.E),
             "optG": opt_state(sg.optG), "optD": opt_state(sg.optD), "optE": opt_state(sg.optE),
             "scheG": sch_state(sg.scheG), "scheD": sch_state(sg.scheD), "scheE": sch_state(sg.scheE),
             "rng": {"torch_cpu": torch.get_rng_state(), "numpy": np.random.get_state()},
             "config": {"lbd": dict(sg.lbd), "k": sg.k, "n_batch": sg.n_batch, "encoded_feature": sg.encoded_feature,
                        "ndim": sg.ndim},
             "extra": extra}
    if getattr(sg, "hi", None) is not None:
        state["hist_target"] = sg.hi.target.detach().cpu()      # drawn from the RNG at construction (util.py:543)
    torch.save(state, path)
    return path


def load_checkpoint(sg, path, restore_rng=True):
    """Inverse of save_checkpoint on a trainer built the same way (same nets, `opt_sche_initialization()` called).
    Returns the `extra` dictionary."""
    state = torch.load(path, map_location="cpu", weights_only=False)
    if state.get("format") != "srgan_b200.checkpoint.v1":
        raise ValueError("not a srgan_b200 checkpoint: %r" % (state.get("format"),))

    def load_net(net, sd_):
        if isinstance(net, (list, tuple)):
            for n, s_ in zip(net, sd_):
                load_net(n, s_)
        else:
            _unwrap(net).load_state_dict(sd_)

    def load_opt(o, st):
        if isinstance(o, (list, tuple)):
            for x, s_ in zip(o, st):
                load_opt(x, s_)
        elif o is not None and st is not None:
            if isinstance(o, ops.FusedAdam):
                o.load_flat_state(st["fused"])
            else:
                o.load_state_dict(st["torch"])

    def load_sch(s_, st):
        if isinstance(s_, (list, tuple)):
            for x, y in zip(s_, st):
                load_sch(x, y)
        elif s_ is not None and st is not None:
            s_.load_state_dict(st)
    load_net(sg.G, state["G"]); load_net(sg.D, state["D"]); load_net(sg.E, state["E"])
    load_opt(sg.optG, state["optG"]); load_opt(sg.optD, state["optD"]); load_opt(sg.optE, state["optE"])
    load_sch(sg.scheG, state["scheG"]); load_sch(sg.scheD, state["scheD"]); load_sch(sg.scheE, state["scheE"])
    if "hist_target" in state and getattr(sg, "hi", None) is not None:
        sg.hi.target = state["hist_target"].to(sg.hi.target.device)
    if restore_rng:
        torch.set_rng_state(state["rng"]["torch_cpu"])
        np.random.set_state(state["rng"]["numpy"])
    if getattr(sg, "_graph", None) is not None:
        sg._graph["state"] = None          # a captured step keeps pointing at the same buffers, but re-capture is cheap
    return state.get("extra", {})


# ------------------------------------------------------------------------------------------- visual check
def get_output_and_plot(sg, dataset, index, class_info, random_sample_num=5, *legacy, device="cuda"):
    """Grid of translations of one sample (source / target / reconstruction / identity, by encoder style and
    by random styles).  Notebooks 01/02 pass one extra positional argument that the current signature of
    the reference no longer has; it is accepted and, when it names a device, used as such."""
    for extra in legacy:
        if isinstance(extra, (str, torch.device)):
            device = extra
    if isinstance(random_sample_num, (str, torch.device)):
        device, random_sample_num = random_sample_num, 5
    import matplotlib.pyplot as plt
    classes, label_discription = class_info
    image0, label0 = dataset[index][0], dataset[index][1]
    n = random_sample_num
    with torch.no_grad():
        src = image0.view(1, 3, 128, 128).to(device)
        src_label = torch.tensor(label0).view(1,)
        others = torch.tensor(get_target(src_label, classes, whole=False, shuffle=False))
        tgt_label = others[:, 0:1]
        show = lambda t: image_from_output(cuda2cpu(t))

        by_enc, _ = sg.G_transformation(tgt_label, src, True, src)
        by_rand, _ = sg.G_transformation(tgt_label.repeat(1, n), src.repeat(n, 1, 1, 1), False)
        first = by_rand[0:1]
        recon_enc, _ = sg.G_transformation(src_label, first, True, src)
        idt_enc, _ = sg.G_transformation(src_label, src, True, src)
        per_class, _ = sg.G_transformation(others, src.repeat(len(classes) - 1, 1, 1, 1), False)
        recon_rand, _ = sg.G_transformation(src_label.repeat(n), first.repeat(n, 1, 1, 1), False)
        idt_rand, _ = sg.G_transformation(src_label.repeat(n), src.repeat(n, 1, 1, 1), False)

    rows, cols = n + 1, 4
    fig = plt.figure(figsize=(5 * cols, 5 * rows))

    def cell(pos, img, title):
        ax = fig.add_subplot(rows, cols, pos)
        ax.imshow(img)
        ax.set_title(title)

    cell(1, show(src)[0], "source")
    cell(2, show(by_enc)[0], "target by source condition")
    cell(3, show(recon_enc)[0], "recon by source condition")
    cell(4, show(idt_enc)[0], "identity image by source condition")
    for i, img in enumerate(show(per_class)):
        cell(4 * (i + 1) + 1, img, label_discription[others[0][i]])
    for col, batch, title in ((2, by_rand, "target by random latent"), (3, recon_rand, "recon by random latent"),
                              (4, idt_rand, "idt by random latent")):
        for i, img in enumerate(show(batch)):
            cell(4 * (i + 1) + col, img, title)
    return fig


def dic_init(get_edge=False):
    """Empty (data, label) containers of `get_samples` (ref pyfiles/util_notebook.py:848-856)."""
    return {"source": [], "target": [], "recon": []}, {"source": [], "target": []}


def get_samples(netG, netE, dataset, index, latent=None, classes=tuple(range(4)), ref_label=None,
                ndim=8, scale=1, image_type="pil", batch=32, device="cuda", conventional_E=False):
    """Translate sample `index` of `dataset` to every class in `classes` under the given style codes and re-encode
    the results (ref pyfiles/util_notebook.py:858-950).  `latent`: one [num, ndim] array shared by all classes or a
    list with one array per class value.  Returns (data, label): data["source"] the input, data["target"][c] the
    `num` translations to class c (PIL-ready arrays for image_type="pil", one tensor for "tensor"),
    label["latent"][c] the encoder means of those translations, one array per chunk of `batch` codes."""
    if image_type not in ("pil", "tensor"):
        raise ValueError("image_type must be 'pil' or 'tensor'")
    src = dataset[index][0].view(1, 3, 128, 128).to(device)
    data, label = dic_init(False)
    label["source"] = cuda2numpy(torch.tensor([dataset[index][1]]))
    data["source"] = image_from_output(src)[0] if image_type == "pil" else cuda2cpu(src)[0]
    netG.eval()
    netE.eval()
    as_dev = lambda v: torch.as_tensor(np.asarray(v), dtype=torch.float32).to(device)
    codes = [as_dev(v) for v in latent] if isinstance(latent, list) else [as_dev(latent)] * len(classes)
    num = codes[0].shape[0]
    label["latent"], data["target"] = {}, {}
    with torch.no_grad():
        for c in classes:
            onehot = class_encode(torch.tensor([c]), device, ref_label)
            mus, images = [], []
            for lo in range(0, num, batch):
                z = codes[c][lo:lo + batch]
                n = z.shape[0]
                out = netG(src.repeat(n, 1, 1, 1), torch.cat([onehot.repeat(n, 1), z], 1))
                mu = netE(out, onehot.repeat(n, 1))[1] if conventional_E else netE(out)[1]
                mus.append(cuda2numpy(mu))
                images.append(image_from_output(out) if image_type == "pil" else cuda2numpy(out))
            label["latent"][c] = mus
            if image_type == "pil":
                data["target"][c] = [im for chunk in images for im in chunk]
            else:
                data["target"][c] = torch.Tensor(np.concatenate(images, axis=0))
    if image_type == "tensor":
        data["source"] = torch.Tensor(np.asarray(data["source"])).unsqueeze(0)
    return data, label


# ------------------------------------------------------------------------------------------- notebook 04 (f4)
def do_test(net, testloader, device="cuda", mode="eval"):
    """ref notebook 04 cell 13: labels and outputs of `net` over a loader of (images, labels) batches."""
    if mode == "train":
        net.train()
    elif mode == "eval":
        net.eval()
    else:
        return None
    labels, outputs = [], []
    with torch.no_grad():
        for data in testloader:
            outputs.append(net(data[0].to(device)).detach().cpu().numpy())
            labels.append(np.asarray(data[1].detach().cpu().numpy() if torch.is_tensor(data[1]) else data[1]))
    return np.concatenate(labels).astype(np.float64), np.concatenate(outputs, axis=0)


class Classifier_training(object):
    """The training job of notebook 04 (`04_Facial_Recognition-Encoder.ipynb` cells 18 and 22) on the kernels of this
    package: `Encoder_classifier` forward / backward through the tcgen05 convolutions and fused norms, the
    cross-entropy kernel on the softmax outputs (the reference applies nn.CrossEntropyLoss to probabilities - kept),
    one fused Adam launch (lr, default betas) and `ExponentialLR(gamma)` stepped once per epoch.  Its result, the
    `state_dict` of the net, is what notebook 05 loads into `Encoder` (`freeze_melt`, keys of Appendix E).

        job = Classifier_training(net, lr=1e-4)            # net = Encoder_classifier(...).to(device)
        loss, acc = job.train_step(x, label)               # one iteration of the inner loop of cell 22
        job.fit(loader, epochs, valloader, test_interval)  # the whole cell: per-epoch means, validation accuracy
    """

    def __init__(self, net, lr=0.0001, gamma=0.99, device=None):
        self.net = net
        self.device = device if device is not None else next(net.parameters()).device
        self.criterion = ops.CrossEntropyLoss()
        self.optimizer = ops.FusedAdam(net.parameters(), lr=lr)
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, gamma=gamma)
        self.losses_epoch, self.acc_epoch, self.acc_test_list = [], [], []
        self.best_acc, self.best_epoch = 0, 0

    def train_step(self, x, label):
        """-> (loss, accuracy) as 0-dim CUDA tensors (no host synchronisation)."""
        self.net.train()
        x = ops.to_nhwc(x.to(self.device))
        label = label.to(self.device).long()
        _zero_grads(self.net, self.optimizer)
        y = self.net(x)
        loss = self.criterion(y, label)
        with ops.direct_param_grads():
            loss.backward()
        _sync_grads(self.net, self.optimizer)
        self.optimizer.step()
        acc = (y.detach().argmax(dim=1) == label).float().mean()
        return loss.detach(), acc

    def fit(self, dataloader, epoch_num, valloader=None, test_interval=3, on_epoch=None):
        for epoch in range(epoch_num):
            stats = [self.train_step(data[0], data[1]) for data in dataloader]
            self.scheduler.step()
            self.losses_epoch.append(float(torch.stack([s[0] for s in stats]).mean()))
            self.acc_epoch.append(float(torch.stack([s[1] for s in stats]).mean()))
            if valloader is not None and epoch % test_interval == 0:
                labels, outputs = do_test(self.net, valloader, self.device, "eval")
                acc_test = float((np.argmax(outputs, axis=1) == labels).mean())
                self.acc_test_list.append(acc_test)
                if self.best_acc < acc_test:
                    self.best_acc, self.best_epoch = acc_test, epoch
            if on_epoch is not None:
                on_epoch(self, epoch)
        return self.losses_epoch, self.acc_epoch, self.acc_test_list
