"""Generator / discriminator / style-encoder modules of SingleGAN and Style-Restricted GAN, B200-native.

Drop-in for the reference's `pyfiles/model.py`: same class names, constructor signatures, `forward`
signatures and return values, `state_dict` keys and shapes, and the same parameter-creation order
(so a given `torch.manual_seed` yields the same initial weights).  What differs is underneath:
every forward/backward runs hand-written sm_100a kernels through `srgan_ops` (C ABI in
include/srgan_b200.h).  Activations are kept channels-last (NHWC) between layers, conv filters
are stored KRSC, and norm + conditional bias + affine + activation (+ residual) are one fused kernel.

There is no CPU path: calling a module with CPU tensors raises `SrganKernelError`.
"""
import functools
import os

import torch
import torch.nn as nn

import srgan_ops as ops
from util import *  # noqa: F401,F403  (the notebooks import MinMax through this module)

CL = torch.channels_last


# --------------------------------------------------------------------------------------------
# parameter containers with kernel-backed forwards
# --------------------------------------------------------------------------------------------
class _KernelConv2d(nn.Conv2d):
    """nn.Conv2d parameters (default PyTorch init), filter stored KRSC, forward on our kernels.
    `act`/`slope` fuse the activation that follows the convolution into its epilogue."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.groups != 1 or self.dilation != (1, 1):
            raise NotImplementedError("grouped / dilated convolutions are not part of this model family")
        if self.stride[0] != self.stride[1] or self.padding[0] != self.padding[1]:
            raise NotImplementedError("asymmetric stride / padding")
        self.weight.data = self.weight.data.contiguous(memory_format=CL)

    def forward(self, x, act=ops.ACT_NONE, slope=0.0, out_dtype=None):
        return ops.conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0],
                          self.padding_mode, act, slope, out_dtype)

    def thin16_ok(self, N, H, W, need_dgrad=True):
        """Can this RGB layer run with its fat side in bf16 on [N, in_channels, H, W] inputs (engine bf16)?"""
        key = (N, H, W, need_dgrad)
        cache = self.__dict__.setdefault("_thin16_cache", {})
        if key not in cache:
            cache[key] = self.padding_mode == "zeros" and self.kernel_size[0] == self.kernel_size[1] and \
                ops.thin16_supported(N, H, W, self.in_channels, self.out_channels, self.kernel_size[0], self.stride[0],
                                     self.padding[0], need_dgrad)
        return cache[key]


class _KernelConvTranspose2d(nn.ConvTranspose2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.bias is not None or self.output_padding != (0, 0) or self.groups != 1:
            raise NotImplementedError("only bias-free, ungrouped transposed convolutions are used here")
        self.weight.data = self.weight.data.contiguous(memory_format=CL)

    def forward(self, x):
        return ops.conv_transpose2d(x, self.weight, self.stride[0], self.padding[0])


class _KernelLinear(nn.Linear):
    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class _KernelInstanceNorm2d(nn.Module):
    """nn.InstanceNorm2d(affine=False, track_running_stats=False): no parameters, no buffers."""

    def __init__(self, num_features, eps=1e-5, affine=False):
        super().__init__()
        if affine:
            raise NotImplementedError("plain InstanceNorm2d is only used with affine=False here")
        self.num_features, self.eps = num_features, eps

    def forward(self, x, act=ops.ACT_NONE, slope=0.0, out_dtype=None):
        return ops.instance_norm_act(x, eps=self.eps, act=act, slope=slope, out_dtype=out_dtype)


def _activation_of(module):
    """(act id, slope) of an nn activation module -- used to fuse it into the producing kernel."""
    if isinstance(module, nn.LeakyReLU):
        return ops.ACT_LRELU, module.negative_slope
    if isinstance(module, nn.ReLU):
        return ops.ACT_RELU, 0.0
    if isinstance(module, nn.Tanh):
        return ops.ACT_TANH, 0.0
    raise NotImplementedError(type(module).__name__)


# --------------------------------------------------------------------------------------------
# conditional norms   (ref: pyfiles/model.py:12-182)
# --------------------------------------------------------------------------------------------
class _CBINorm(nn.Module):
    """Conditional instance norm: IN(x) + tanh(Linear(con)), then the affine (weight, bias).
    Statistics always come from the input (track_running_stats=False is the only mode the model
    family uses; ref pyfiles/model.py:60,179)."""

    def __init__(self, num_features, num_con=8, eps=1e-5, momentum=0.1, affine=False, track_running_stats=False):
        super().__init__()
        if track_running_stats:
            raise NotImplementedError("CBINorm with running statistics is not used by SingleGAN/SRGAN")
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.track_running_stats = affine, track_running_stats
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        self.ConBias = nn.Sequential(nn.Linear(num_con, num_features), nn.Tanh())

    _version = 2          # what nn.InstanceNorm2d-derived modules of the reference record in state_dict metadata

    def _check_input_dim(self, input):
        raise NotImplementedError

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        """Checkpoints without version metadata (torch < 0.4) may carry running statistics although nothing tracks
        them: report and drop them, as the reference does (pyfiles/model.py:24-52)."""
        if local_metadata.get("version", None) is None and not self.track_running_stats:
            stale = [prefix + n for n in ("running_mean", "running_var") if prefix + n in state_dict]
            if stale:
                error_msgs.append("Unexpected running stats buffer(s) {} for {} with track_running_stats=False. "
                                  "Remove these keys from the state_dict (checkpoint saved before torch 0.4.0?)."
                                  .format(" and ".join('"%s"' % k for k in stale), type(self).__name__))
                for k in stale:
                    state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def forward(self, input, ConInfor, act=ops.ACT_NONE, slope=0.0, residual=None, out_dtype=None):
        self._check_input_dim(input)
        lin = self.ConBias[0]
        t = ops.cond_bias(ConInfor, lin.weight, lin.bias)
        return ops.instance_norm_act(input, self.weight, self.bias, t, residual, self.eps, act, slope, out_dtype)

    def extra_repr(self):
        return "{num_features}, eps={eps}, affine={affine}".format(**self.__dict__)


class CBINorm2d(_CBINorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError("expected 4D input (got {}D input)".format(input.dim()))


class _CBBNorm(nn.Module):
    """Conditional *batch* norm variant (ref pyfiles/model.py:75-171): batch norm without affine, minus its own
    spatial mean, plus tanh(Linear(con)), then weight / bias.  The notebooks all pass norm_type="instance"; this is
    what norm_type="batch" selects.  Data parallel: statistics are those of the global batch."""

    def __init__(self, num_features, num_con, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.track_running_stats = affine, track_running_stats
        if affine:
            self.weight = nn.Parameter(torch.empty(num_features))
            self.bias = nn.Parameter(torch.empty(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        if track_running_stats:
            self.register_buffer("running_mean", torch.zeros(num_features))
            self.register_buffer("running_var", torch.ones(num_features))
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        else:
            self.register_parameter("running_mean", None)
            self.register_parameter("running_var", None)
            self.register_parameter("num_batches_tracked", None)
        self.reset_parameters()
        self.ConBias = nn.Sequential(nn.Linear(num_con, num_features), nn.Tanh())
        self.avgpool = nn.AdaptiveAvgPool2d(1)

    def reset_running_stats(self):
        if self.track_running_stats:
            self.running_mean.zero_()
            self.running_var.fill_(1)
            self.num_batches_tracked.zero_()

    def reset_parameters(self):
        self.reset_running_stats()
        if self.affine:
            self.weight.data.uniform_()
            self.bias.data.zero_()

    _version = 2

    def _check_input_dim(self, input):
        raise NotImplementedError

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        """Checkpoints older than metadata version 2 have no `num_batches_tracked`: default it to 0
        (ref pyfiles/model.py:152-165)."""
        version = local_metadata.get("version", None)
        if (version is None or version < 2) and self.track_running_stats:
            key = prefix + "num_batches_tracked"
            if key not in state_dict:
                state_dict[key] = torch.tensor(0, dtype=torch.long)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def _factor(self):
        """exponential_average_factor of the reference (pyfiles/model.py:124-131); bumps num_batches_tracked."""
        if self.training and self.track_running_stats:
            self.num_batches_tracked += 1
            if self.momentum is None:
                return 1.0 / float(self.num_batches_tracked.item())
            return self.momentum
        return 0.0

    def forward(self, input, ConInfor, act=ops.ACT_NONE, slope=0.0, residual=None, out_dtype=None):
        if out_dtype not in (None, torch.float32):
            raise NotImplementedError("batch-statistics norms keep fp32 storage")
        self._check_input_dim(input)
        factor = self._factor()
        lin = self.ConBias[0]
        t = ops.cond_bias(ConInfor, lin.weight, lin.bias)
        return ops.batch_norm_act(input, self.weight, self.bias, t, residual, self.running_mean, self.running_var,
                                  self.training or not self.track_running_stats, factor, self.eps, True, act, slope)

    def extra_repr(self):
        return "{num_features}, eps={eps}, momentum={momentum}, affine={affine}, " \
               "track_running_stats={track_running_stats}".format(**self.__dict__)


class CBBNorm2d(_CBBNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError("expected 4D input (got {}D input)".format(input.dim()))


class _KernelBatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d(affine=True) for norm_type="batch" (ref get_norm_layer pyfiles/model.py:175): same parameters,
    buffers and state_dict keys; forward through the fused kernels (optional activation)."""

    def forward(self, x, act=ops.ACT_NONE, slope=0.0, out_dtype=None):
        if out_dtype not in (None, torch.float32):
            raise NotImplementedError("batch-statistics norms keep fp32 storage")
        self._check_input_dim(x)
        factor = 0.0
        if self.training and self.track_running_stats:
            self.num_batches_tracked += 1
            factor = 1.0 / float(self.num_batches_tracked.item()) if self.momentum is None else self.momentum
        training = self.training or not self.track_running_stats
        return ops.batch_norm_act(x, self.weight, self.bias, None, None, self.running_mean, self.running_var, training,
                                  factor, self.eps, False, act, slope)


def get_norm_layer(layer_type="instance", num_con=2):
    if layer_type == "batch":
        norm_layer = functools.partial(_KernelBatchNorm2d, affine=True)
        c_norm_layer = functools.partial(CBBNorm2d, affine=True, num_con=num_con)
    elif layer_type == "instance":
        norm_layer = functools.partial(_KernelInstanceNorm2d, affine=False)
        c_norm_layer = functools.partial(CBINorm2d, affine=True, num_con=num_con)
    else:
        raise NotImplementedError("normalization layer [%s] is not found" % layer_type)
    return norm_layer, c_norm_layer


# --------------------------------------------------------------------------------------------
# generator   (ref: pyfiles/model.py:188-249)
# --------------------------------------------------------------------------------------------
_NO_SKIP_FUSE = os.environ.get("SRGAN_DBG_NO_SKIP_FUSE", "0") != "0"     # bring-up: autograd adds the skip gradient


class SingleResidualBlock(nn.Module):
    def __init__(self, nch, c_norm_layer):
        super().__init__()
        self.c1 = _KernelConv2d(nch, nch, kernel_size=3, stride=1, padding=1, bias=False)
        self.cn1 = c_norm_layer(nch)
        self.c2 = _KernelConv2d(nch, nch, kernel_size=3, stride=1, padding=1, bias=False)
        self.cn2 = c_norm_layer(nch)

    def forward(self, x):
        data, con = x[0], x[1]
        if self.c1.bias is None and self.c1.padding_mode == "zeros" and not _NO_SKIP_FUSE:
            # c1 and the skip connection leave `data` through one autograd node: their two gradients are added in
            # the dgrad epilogue of c1 (no separate pass over the tensor)
            h1, skip = ops.conv2d_skip(data, self.c1.weight, self.c1.stride[0], self.c1.padding[0])
        else:
            h1, skip = self.c1(data), data
        h = self.cn1(h1, con, act=ops.ACT_RELU)
        # second norm has no activation; the skip connection is added inside the same kernel
        return self.cn2(self.c2(h), con, residual=skip), con


class SingleGenerator(nn.Module):
    def __init__(self, nch_in, nch, reduce=2, num_cls=3, res_num=6, norm_type="instance", num_con=2,
                 nch_out=None):
        super().__init__()
        nch_out = nch_in if nch_out is None else nch_out
        norm_layer, c_norm_layer = get_norm_layer(layer_type=norm_type, num_con=num_con)
        self.num_cls = num_cls
        k, s, p = 2 * reduce, reduce, int(reduce / 2)

        convs = [_KernelConv2d(nch_in, nch, kernel_size=7, stride=1, padding=3, bias=False)]
        cnorms = [c_norm_layer(nch)]
        for i in range(num_cls):
            convs.append(_KernelConv2d(nch * 2 ** i, nch * 2 ** (i + 1), kernel_size=k, stride=s, padding=p,
                                       bias=False))
            cnorms.append(c_norm_layer(nch * 2 ** (i + 1)))
        self.down_convs = nn.ModuleList(convs)
        self.down_cnorms = nn.ModuleList(cnorms)

        self.resBlocks = nn.Sequential(*[SingleResidualBlock(nch * 2 ** num_cls, c_norm_layer)
                                         for _ in range(res_num)])

        ups = [_KernelConvTranspose2d(nch * 2 ** num_cls, nch * 2 ** (num_cls - 1), kernel_size=k, stride=s,
                                      padding=p, bias=False)]
        norms = [norm_layer(nch * 2 ** (num_cls - 1))]
        for i in reversed(range(1, num_cls)):
            ups.append(_KernelConvTranspose2d(nch * 2 ** i, nch * 2 ** (i - 1), kernel_size=k, stride=s,
                                              padding=p, bias=False))
            norms.append(norm_layer(nch * 2 ** (i - 1)))
        ups.append(_KernelConv2d(nch, nch_out, kernel_size=7, stride=1, padding=3, bias=False))
        self.up_convs = nn.ModuleList(ups)
        self.up_norms = nn.ModuleList(norms)

    def _bf16_trunk(self):
        """Engine 'bf16': everything between the RGB stem and the RGB head keeps bf16 activations - the first
        conditional norm converts on its way out (fp32 in, bf16 out), the last up-path norm on its way back (bf16 in,
        fp32 out), so both thin RGB layers keep their fp32 kernels and no conversion pass exists.  Needs instance
        norms (no batch-coupled statistics in bf16) and trunk widths that are multiples of 64."""
        if not ops.bf16_trunk_enabled() or self.num_cls < 1:
            return False
        if self.down_convs[0].out_channels % 64:
            return False
        norms = list(self.down_cnorms) + list(self.up_norms)
        return all(isinstance(m, (_CBINorm, _KernelInstanceNorm2d)) for m in norms) and \
            all(isinstance(b.cn1, _CBINorm) and isinstance(b.cn2, _CBINorm) for b in self.resBlocks)

    def forward(self, x, c):
        lo = torch.bfloat16 if self._bf16_trunk() else None
        # "thin16": the RGB stem writes bf16 and the RGB head reads bf16 (TF32 / kind::f16 kernels with mixed storage),
        # so the two largest tensors of the generator (64 channels at full resolution) never exist in fp32; where the
        # shapes do not qualify the norms next to them convert (fp32 in / bf16 out and back), as in round 2's first cut
        head = self.up_convs[-1]
        N, _, H, W = x.shape
        thin = lo is not None and N > 0 and self.down_convs[0].thin16_ok(N, H, W) and head.thin16_ok(N, H, W)
        for i, (conv, cnorm) in enumerate(zip(self.down_convs, self.down_cnorms)):
            h = conv(x, out_dtype=lo) if (thin and i == 0) else conv(x)
            x = cnorm(h, c, act=ops.ACT_RELU, out_dtype=lo)
        x = self.resBlocks([x, c])[0]
        for i in range(self.num_cls):
            last = i == self.num_cls - 1
            x = self.up_norms[i](self.up_convs[i](x), act=ops.ACT_RELU,
                                 out_dtype=torch.float32 if (last and lo is not None and not thin) else None)
        return head(x, act=ops.ACT_TANH)


# --------------------------------------------------------------------------------------------
# discriminators   (ref: pyfiles/model.py:255-346)
# --------------------------------------------------------------------------------------------
def _tower_bf16_ok(stack, x):
    """Engine 'bf16': can this discriminator tower keep bf16 activations?  Its RGB stem must qualify for the thin16
    kernels (fp32 image -> bf16, wgrad on the bf16 gradient) and every following activated convolution must be
    bias-free with channel counts that are multiples of 64 (tcgen05 kind::f16 fprop / dgrad / wgrad); a trailing head
    convolution (1 or 4 logits) reads an fp32 copy.  The narrow second tower of the multi-scale discriminators
    (nch / 2 = 32) does not qualify and stays on the TF32 engine."""
    if not ops.bf16_trunk_enabled() or x.dtype != torch.float32 or x.shape[0] == 0:
        return False
    convs = [m for m in stack if isinstance(m, nn.Conv2d)]
    if len(convs) < 2 or convs[0].in_channels > 4:
        return False
    N, _, H, W = x.shape
    if not convs[0].thin16_ok(N, H, W, need_dgrad=False) or convs[0].out_channels % 64:
        return False
    mods = list(stack)
    for i, m in enumerate(mods):
        if not isinstance(m, nn.Conv2d) or m is convs[0]:
            continue
        activated = i + 1 < len(mods) and not isinstance(mods[i + 1], nn.Conv2d)
        if activated and (m.bias is not None or m.in_channels % 64 or m.out_channels % 64 or
                          m.padding_mode != "zeros"):
            return False
    return True


def _run_conv_stack(stack, x):
    """nn.Sequential of [conv, LeakyReLU, conv, LeakyReLU, ... (, conv)]: each activation is fused
    into the epilogue of the convolution before it.  Returns fp32 (a bf16 tower converts at its end)."""
    mods = list(stack)
    lo = torch.bfloat16 if _tower_bf16_ok(stack, x) else None
    i = 0
    while i < len(mods):
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        if nxt is not None and not isinstance(nxt, nn.Conv2d):
            act, slope = _activation_of(nxt)
            if lo is not None and i == 0:
                x = mods[i](x, act=act, slope=slope, out_dtype=lo)      # thin16 stem: fp32 image -> bf16
            else:
                x = mods[i](x, act=act, slope=slope)
            i += 2
        else:
            x = mods[i](ops.cast_f32(x))                                # un-activated head convolution: fp32
            i += 1
    return ops.cast_f32(x)


def _patch_tower(nch_in, nch, reduce, num_cls, with_head):
    layers = [_KernelConv2d(nch_in, nch, kernel_size=4, stride=2, padding=1, bias=False), nn.LeakyReLU()]
    dim_in = nch
    for _ in range(1, num_cls):
        dim_out = min(dim_in * 2, nch * 8)
        layers += [_KernelConv2d(dim_in, dim_out, kernel_size=2 * reduce, stride=reduce,
                                 padding=int(reduce / 2), bias=False), nn.LeakyReLU()]
        dim_in = dim_out
    if with_head:
        layers.append(_KernelConv2d(dim_in, 1, kernel_size=4, stride=1, padding=1, bias=True))
    return nn.Sequential(*layers)


class SingleDiscriminator_original(nn.Module):
    def __init__(self, nch_in, nch, reduce=2, num_cls=3, norm_type="instance", num_con=2):
        super().__init__()
        self.num_cls = num_cls
        self.down_convs = _patch_tower(nch_in, nch, reduce, num_cls, with_head=True)

    def forward(self, x):
        return _run_conv_stack(self.down_convs, x)


class _Pool3s2(nn.AvgPool2d):
    def forward(self, x):
        return ops.avg_pool3s2(x)


class SingleDiscriminator_original_multi(nn.Module):
    def __init__(self, nch_in, nch, reduce=2, num_cls=3, norm_type="instance", num_con=2):
        super().__init__()
        self.discriminator1 = SingleDiscriminator_original(nch_in, nch, reduce, num_cls, norm_type, num_con)
        self.down = _Pool3s2(3, stride=2, padding=[1, 1], count_include_pad=False)
        self.discriminator2 = SingleDiscriminator_original(nch_in, nch // 2, reduce, num_cls, norm_type, num_con)

    def forward(self, x):
        x = ops.to_nhwc(x)
        with ops.tower_streams(x) as ts:
            with ts.side():
                out2 = self.discriminator2(self.down(x))
            out1 = self.discriminator1(x)
            ts.join(out2)
        return [out1, out2]


class SingleDiscriminator_solo(nn.Module):
    def __init__(self, nch_in, nch, reduce=2, num_cls=3, norm_type="instance", num_con=2):
        super().__init__()
        self.num_cls = num_cls
        self.down_convs = _patch_tower(nch_in, nch, reduce, num_cls, with_head=False)

    def forward(self, x):
        return _run_conv_stack(self.down_convs, x)


class SingleDiscriminator_solo_multi(nn.Module):
    def __init__(self, nch_in, nch, reduce=2, num_cls=3, norm_type="instance", n_class=4):
        super().__init__()
        self.n_class = n_class
        self.discriminator1 = SingleDiscriminator_solo(nch_in, nch, reduce, num_cls, norm_type, None)
        self.down = _Pool3s2(3, stride=2, padding=[1, 1], count_include_pad=False)
        self.discriminator2 = SingleDiscriminator_solo(nch_in, nch // 2, reduce, num_cls, norm_type, None)

        dim_in = min(nch * 2 ** num_cls, nch * 8)
        self.last_layer1 = _KernelConv2d(dim_in, 1, kernel_size=4, stride=1, padding=1, bias=True)
        self.last_layer2 = _KernelConv2d(dim_in // 2, 1, kernel_size=4, stride=1, padding=1, bias=True)
        head1 = _KernelConv2d(dim_in, n_class, kernel_size=8, stride=1, padding=0, bias=True)
        head2 = _KernelConv2d(dim_in // 2, n_class, kernel_size=4, stride=1, padding=0, bias=True)
        self.classification_layer1 = nn.Sequential(head1, nn.Softmax(dim=1))
        self.classification_layer2 = nn.Sequential(head2, nn.Softmax(dim=1))

    def _classify(self, head, feat):
        logits = head[0](feat)                       # [B, n_class, h, w]  (h = w = 1 for 128x128 inputs)
        B, J, h, w = logits.shape
        if h * w == 1:
            return ops.softmax_rows(logits.reshape(B, J))
        # general spatial size: softmax over channels at every position, then the reference's view(-1, n_class)
        rows = ops._raw_to_nhwc(logits).permute(0, 2, 3, 1).reshape(-1, J)
        probs = ops.softmax_rows(rows).view(B, h, w, J).permute(0, 3, 1, 2)
        return probs.reshape(-1, J)

    def forward(self, x):
        x = ops.to_nhwc(x)
        with ops.tower_streams(x) as ts:
            with ts.side():                      # the narrow tower and its heads: small kernels, second stream
                disout2 = self.discriminator2(self.down(x))
                output2 = self.last_layer2(disout2)
                out_class2 = self._classify(self.classification_layer2, disout2).view(-1, self.n_class)
            disout1 = self.discriminator1(x)
            output1 = self.last_layer1(disout1)
            out_class1 = self._classify(self.classification_layer1, disout1).view(-1, self.n_class)
            ts.join(output2, out_class2)
        return [output1, output2], [out_class1, out_class2]


# --------------------------------------------------------------------------------------------
# encoders   (ref: pyfiles/model.py:352-508)
# --------------------------------------------------------------------------------------------
class _Pool2(nn.AvgPool2d):
    def forward(self, x):
        return ops.avg_pool2(x)


def _block_tail(nch_in, nch_out):
    cmp = nn.Sequential(
        _KernelConv2d(nch_in, nch_out, kernel_size=3, stride=1, padding=1, bias=False, padding_mode="reflect"),
        _Pool2(2, 2))
    shortcut = nn.Sequential(
        _Pool2(2, 2),
        _KernelConv2d(nch_in, nch_out, kernel_size=1, stride=1, padding=0, bias=True))
    return cmp, shortcut


def _encoder_block_dtype(block, norm1, norm2):
    """Engine 'bf16': the block's big tensors (first norm's output, both reflect-padded tensors, conv1's output, the
    second norm's output and the pre-pool output of the compression conv) are bf16 - the first norm converts on its
    way out; the block's input / output (pooled, 4x smaller) and the 1x1 shortcut stay fp32.  Needs instance norms and
    a width that is a multiple of 64."""
    if not ops.bf16_trunk_enabled() or block.conv1.in_channels % 64:
        return None
    if not all(isinstance(m, (_CBINorm, _KernelInstanceNorm2d)) for m in (norm1, norm2)):
        return None
    return torch.bfloat16


def _block_forward(x, h, cmp, shortcut):
    """out = avgpool2(cmp_conv(h)) + shortcut_conv(avgpool2(x)); pool + add is one kernel."""
    return ops.avg_pool2_add(cmp[0](h), shortcut[1](shortcut[0](x)))


class BasicBlock(nn.Module):
    def __init__(self, nch_in, nch_out, c_norm_layer=None):
        super().__init__()
        self.cnorm1 = c_norm_layer(nch_in)
        self.nl1 = nn.LeakyReLU(0.2)
        self.conv1 = _KernelConv2d(nch_in, nch_in, kernel_size=3, stride=1, padding=1, bias=False,
                                   padding_mode="reflect")
        self.cnorm2 = c_norm_layer(nch_in)
        self.nl2 = nn.LeakyReLU(0.2)
        self.cmp, self.shortcut = _block_tail(nch_in, nch_out)

    def forward(self, input):
        x, d = input
        a1, s1 = _activation_of(self.nl1)
        a2, s2 = _activation_of(self.nl2)
        lo = _encoder_block_dtype(self, self.cnorm1, self.cnorm2)
        h = self.conv1(self.cnorm1(x, d, act=a1, slope=s1, out_dtype=lo))
        h = self.cnorm2(h, d, act=a2, slope=s2)
        return [_block_forward(x, h, self.cmp, self.shortcut), d]


class _EncoderBase(nn.Module):
    def reparametrize(self, mu, logvar):
        # the noise comes from the CPU default generator, like the reference (pyfiles/model.py:400,461):
        # same seed => same eps as the reference, on any device
        eps = ops.host_normal(mu.shape[0], mu.shape[1], mu.device)
        return ops.reparametrize(mu, logvar, eps)


class Encoder_original(_EncoderBase):
    def __init__(self, nch_in, nch_out, nch=64, num_cls=3, norm_type="instance", num_con=2, device="cpu"):
        super().__init__()
        _, c_norm_layer = get_norm_layer(layer_type=norm_type, num_con=num_con)
        self.num_cls = num_cls
        self.device = device
        self.first_layer = _KernelConv2d(nch_in, nch, kernel_size=7, stride=2, padding=1, bias=True)
        blocks, in_nch = [], nch
        for _ in range(num_cls):
            out_nch = in_nch * 2
            blocks.append(BasicBlock(in_nch, out_nch, c_norm_layer))
            in_nch = out_nch
        self.layers = nn.Sequential(*blocks)
        self.last_layer = nn.Sequential(nn.LeakyReLU(0.2), nn.AdaptiveAvgPool2d(1))
        self.fcmean = _KernelLinear(out_nch, nch_out)
        self.fcvar = _KernelLinear(out_nch, nch_out)

    def forward(self, x, c):
        feat = self.layers([self.first_layer(x), c])[0]
        pooled = ops.lrelu_gap(feat, self.last_layer[0].negative_slope)
        mu = self.fcmean(pooled)
        logvar = self.fcvar(pooled)
        return self.reparametrize(mu, logvar), mu, logvar


class BasicBlock_classification(nn.Module):
    def __init__(self, nch_in, nch_out, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(nch_in)
        self.nl1 = nn.LeakyReLU(0.2)
        self.conv1 = _KernelConv2d(nch_in, nch_in, kernel_size=3, stride=1, padding=1, bias=False,
                                   padding_mode="reflect")
        self.norm2 = norm_layer(nch_in)
        self.nl2 = nn.LeakyReLU(0.2)
        self.cmp, self.shortcut = _block_tail(nch_in, nch_out)

    def forward(self, input):
        x = input
        a1, s1 = _activation_of(self.nl1)
        a2, s2 = _activation_of(self.nl2)
        lo = _encoder_block_dtype(self, self.norm1, self.norm2)
        h = self.conv1(self.norm1(x, act=a1, slope=s1, out_dtype=lo))
        h = self.norm2(h, act=a2, slope=s2)
        return _block_forward(x, h, self.cmp, self.shortcut)


def _classification_trunk(nch_in, nch, num_cls, norm_layer):
    first = _KernelConv2d(nch_in, nch, kernel_size=7, stride=2, padding=1, bias=True)
    blocks, in_nch = [], nch
    for _ in range(num_cls):
        out_nch = in_nch * 2
        blocks.append(BasicBlock_classification(in_nch, out_nch, norm_layer))
        in_nch = out_nch
    return first, nn.Sequential(*blocks), out_nch


class Encoder(_EncoderBase):
    def __init__(self, nch_in, nch_out, nch=64, num_cls=3, norm_type="instance", num_con=2, device="cpu"):
        super().__init__()
        norm_layer, _ = get_norm_layer(layer_type=norm_type, num_con=num_con)
        self.num_cls = num_cls
        self.device = device
        self.first_layer, self.layers, out_nch = _classification_trunk(nch_in, nch, num_cls, norm_layer)
        self.last_layer = nn.Sequential(nn.LeakyReLU(0.2), nn.AdaptiveAvgPool2d(1))
        self.fcmean = _KernelLinear(out_nch, nch_out)
        self.fcvar = _KernelLinear(out_nch, nch_out)
        self.fcclass = _KernelLinear(out_nch, num_con)

    def freeze_melt(self, classifier_layers, mode="freeze"):
        names = list(self.state_dict().keys())
        for name, param in zip(names, self.parameters()):
            if name in classifier_layers:
                if mode == "freeze":
                    param.requires_grad = False
                elif mode == "melt":
                    param.requires_grad = True

    def forward(self, x):
        feat = self.layers(self.first_layer(x))
        # the reference pools the same activation three times (pyfiles/model.py:477-480); once is enough
        pooled = ops.lrelu_gap(feat, self.last_layer[0].negative_slope)
        mu = self.fcmean(pooled)
        logvar = self.fcvar(pooled)
        c_code = self.reparametrize(mu, logvar)
        class_output = self.fcclass(pooled)
        return c_code, mu, logvar, class_output, None


class Encoder_classifier(nn.Module):
    def __init__(self, nch_in, nch_out, nch=64, num_cls=3, norm_type="instance", num_con=2):
        super().__init__()
        norm_layer, _ = get_norm_layer(layer_type=norm_type, num_con=num_con)
        self.num_cls = num_cls
        self.first_layer, self.layers, out_nch = _classification_trunk(nch_in, nch, num_cls, norm_layer)
        self.last_layer = nn.Sequential(nn.LeakyReLU(0.2), nn.AdaptiveAvgPool2d(1))
        self.fcclass = _KernelLinear(out_nch, num_con)

    def forward(self, x):
        feat = self.layers(self.first_layer(x))
        pooled = ops.lrelu_gap(feat, self.last_layer[0].negative_slope)
        return ops.softmax_rows(self.fcclass(pooled))
