"""Differentiable operators of the SRGAN training step, backed by libsrgan_b200.so.

Every function here launches hand-written sm_100a kernels through the C ABI
(include/srgan_b200.h); PyTorch supplies device memory, streams and the autograd tape only.
CUDA fp32 tensors are required -- there is no CPU path.

Activations are handled as logical NCHW tensors stored channels-last (NHWC); convolution
filters as logical [K,C,R,S] parameters stored channels-last (KRSC).

Backward passes read parameters LIVE (they are not saved on the tape).  This reproduces the
torch-1.4 semantics the reference was trained with: `optG.step()` between the two backward
passes of `update_GandE` (ref: pyfiles/util_notebook.py:666 then :689) mutates the weights in
place, and the second backward uses the updated weights with the activations saved earlier.
"""
import math
import os

import numpy as np
import torch

import _srgan_lib as L
from _srgan_lib import ConvDesc, SrganKernelError, check

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_FP32, ENGINE_TF32 = 0, 1, 2
_ENGINE_NAMES = {"auto": ENGINE_AUTO, "fp32": ENGINE_FP32, "tf32": ENGINE_TF32}
DT_F32, DT_BF16 = 0, 1
BF16 = torch.bfloat16

CL = torch.channels_last
abi_calls = 0          # number of C-ABI kernel entry points invoked (bench: gpu_launches)


def set_conv_engine(name):
    """'auto' (tcgen05 TF32 on fp32 storage where the shape qualifies), 'fp32' (exact FFMA), 'tf32' (force tcgen05)
    or 'bf16': like 'auto', and the generator keeps the activations of its trunk (everything between the RGB stem and
    the RGB head) in bfloat16 and runs those convolutions on tcgen05 kind::f16 with bf16 shadows of the fp32 master
    weights (fp32 accumulation, fp32 statistics, fp32 gradients and optimizer state)."""
    global _engine, _bf16
    _bf16 = name == "bf16"
    _engine = ENGINE_AUTO if _bf16 else _ENGINE_NAMES[name]


def get_conv_engine():
    return "bf16" if _bf16 else {v: k for k, v in _ENGINE_NAMES.items()}[_engine]


def bf16_trunk_enabled():
    return _bf16


_engine, _bf16 = ENGINE_AUTO, False
set_conv_engine(os.environ.get("SRGAN_CONV_ENGINE", "auto").lower())


def _lib():
    return L.load()


def _stream():
    """Raw handle of the current CUDA stream.  torch.cuda.current_stream() builds a Stream object through several
    Python layers (15 us per call, 28 ms per training step); the two C calls below cost under a microsecond."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _req(*ts):
    """Operand check of every public operator: CUDA fp32 tensors that live on the CURRENT device.  Kernels are
    launched on the current device's current stream with raw pointers, so a tensor of another device would be an
    illegal address (or a silent peer access): callers select the device with torch.cuda.set_device() /
    `with torch.cuda.device(...)` (one process per GPU does this once)."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda or (t.dtype != torch.float32 and t.dtype != BF16):
            raise SrganKernelError(
                "srgan_b200 operators need CUDA float32 tensors (bfloat16 inside the bf16 trunk; got %s on %s); there "
                "is no CPU fallback" % (t.dtype, t.device))
        if cur is None:
            cur = torch._C._cuda_getDevice()
        if t.device.index != cur:
            raise SrganKernelError(
                "srgan_b200 operators launch on the current CUDA device (cuda:%d) but got a tensor on %s: call "
                "torch.cuda.set_device(%d) or wrap the step in `with torch.cuda.device(%d)`"
                % (cur, t.device, t.device.index, t.device.index))


def _p(t):
    return None if t is None else t.data_ptr()


def _call(name, *args):
    global abi_calls
    abi_calls += 1
    check(getattr(_lib(), name)(*args), name)


class _ScratchSet(object):
    """Kernel scratch of one launch context: the conv / norm workspace (split-K partials, transposed filters, norm
    partials) and the reduction scratch with its ticket counter.  Kernels get RAW ADDRESSES of these buffers, so a
    buffer must outlive everything that was launched - or captured - with it: a set never frees a block it has
    handed out (`keep`), and a captured CUDA graph owns a private set allocated inside the capture (from the graph's
    memory pool), see `private_scratch`."""

    def __init__(self):
        self.ws, self.red, self.ctr, self.keep = {}, {}, {}, []

    # Every buffer is keyed by (device, stream): two streams that run kernels concurrently (the two towers of the
    # multi-scale discriminators, see tower_streams) must not share scratch.

    def counters(self, dev, nbytes):
        """Zero-initialised int32 ticket counters of the norm kernels (they leave them zero, see
        srgan_inorm_fwd_mixed)."""
        key = (dev, _stream())
        c = self.ctr.get(key)
        if c is None or c.numel() * 4 < nbytes:
            c = torch.zeros(max(nbytes // 4 + 1, 1 << 16), dtype=torch.int32, device=dev)
            self.ctr[key] = c
            if self is not _global_scratch:
                self.keep.append(c)
        return c

    def workspace(self, dev, nbytes):
        key = (dev, _stream())
        buf = self.ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
            self.ws[key] = buf
            if self is not _global_scratch:
                self.keep.append(buf)          # addresses are baked into the graph: nothing is ever released
        return buf

    def red_scratch(self, dev):
        key = (dev, _stream())
        s = self.red.get(key)
        if s is None:
            # zero-initialised ticket counter; inside a capture this is a captured memset, replayed with the graph
            s = torch.zeros(_lib().srgan_reduce_scratch_bytes(0) // 4, dtype=torch.float32, device=dev)
            self.red[key] = s
        return s


_global_scratch = _ScratchSet()
_scratch_set = _global_scratch


class private_scratch(object):
    """with private_scratch() as s: every operator launched inside uses `s` (a fresh _ScratchSet) instead of the
    process-wide one.  The trainers wrap the CUDA-graph capture of a step in it and keep `s` with the graph, so an
    eager call that later grows the global workspace cannot free memory the graph still writes to."""

    def __init__(self, existing=None):
        self.set = existing or _ScratchSet()

    def __enter__(self):
        global _scratch_set
        self.prev, _scratch_set = _scratch_set, self.set
        return self.set

    def __exit__(self, *exc):
        global _scratch_set
        _scratch_set = self.prev


def _workspace(dev, nbytes):
    return _scratch_set.workspace(dev, nbytes)


# ---- two independent sub-networks on two streams (the towers of the multi-scale discriminators): the narrow tower's
# kernels are small (8 - 64 CTAs, launch bound) and fit next to the wide tower's on other SMs.  Autograd runs every
# backward node on the stream of its forward and synchronises across streams, so the split carries over to backward.
# Measured (same box, ABAB, batch 64, CUDA graph): 51.9 / 51.3 ms with, 52.5 / 52.7 ms without; losses bit-identical.
TOWER_STREAMS = os.environ.get("SRGAN_TOWER_STREAMS", "1") != "0"
_tower_side = {}


class tower_streams(object):
    """with tower_streams(x) as ts: ... ; with ts.side(): y2 = f2(x) ; y1 = f1(x) ; ts.join(y2)"""

    def __init__(self, *inputs):
        self.inputs = [t for t in inputs if torch.is_tensor(t) and t.is_cuda]
        self.enabled = TOWER_STREAMS and bool(self.inputs)

    def __enter__(self):
        if self.enabled:
            dev = self.inputs[0].device
            self.main = torch.cuda.current_stream(dev)
            st = _tower_side.get(dev)
            if st is None:
                st = _tower_side[dev] = torch.cuda.Stream(dev)
            self.side_stream = st
            st.wait_stream(self.main)
            for t in self.inputs:
                t.record_stream(st)
        return self

    def __exit__(self, *exc):
        return False

    def side(self):
        import contextlib
        return torch.cuda.stream(self.side_stream) if self.enabled else contextlib.nullcontext()

    def join(self, *outputs):
        if self.enabled:
            self.main.wait_stream(self.side_stream)
            for t in outputs:
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(self.main)


def _red_scratch(dev):
    return _scratch_set.red_scratch(dev)


def _norm_counters(dev, N, C):
    return _scratch_set.counters(dev, _lib().srgan_inorm_mixed_counters(N, C))


# ----------------------------------------------------------------------------- host RNG
_dp_forced = None


def dp_rank_world():
    """(rank, world) of the data-parallel job; (0, 1) without torch.distributed or inside `single_process()`."""
    if _dp_forced is not None:
        return _dp_forced
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class single_process(object):
    """with single_process(): the trainers and operators behave as a single-GPU job (no collectives, the whole batch
    is local) although torch.distributed is initialised.  bench.py uses it on rank 0 to re-run a data-parallel step
    as the single-GPU global-batch step it must reproduce."""

    def __enter__(self):
        global _dp_forced
        self.prev, _dp_forced = _dp_forced, (0, 1)

    def __exit__(self, *exc):
        global _dp_forced
        _dp_forced = self.prev


class FeedTape(object):
    """Host-produced inputs of a training step that is captured into a CUDA graph.  While `recording`, every host
    feed (noise draw, Adam scalars) registers a pinned staging buffer, a static device buffer and a `fill` callable
    instead of copying; `upload()` re-runs the fills IN PROGRAM ORDER (so the CPU generator is consumed exactly as
    in the eager step) and enqueues the host-to-device copies ahead of the graph launch."""

    def __init__(self, device="cuda", staging_floats=1 << 21):
        self.entries = []
        self.recording = False
        self._done = None
        # Both arenas are allocated BEFORE the capture starts.  Pinned staging: cudaHostAlloc is not a capturable
        # call.  Device side: the feeds are uploaded ahead of the graph launch, so their memory must not come from
        # the graph's private pool, where an earlier temporary of the captured step may share the address and
        # overwrite the feed before its consumer runs.
        self._arena = torch.empty(staging_floats, dtype=torch.float32).pin_memory()
        self._dev_arena = torch.empty(staging_floats, dtype=torch.float32, device=device)
        self._used = self._dev_used = 0

    def device_buffer(self, *shape):
        n = 1
        for d in shape:
            n *= int(d)
        n_pad = -(-n // 64) * 64              # 256-byte aligned slices
        if self._dev_used + n_pad > self._dev_arena.numel():
            raise SrganKernelError("FeedTape: device arena exhausted")
        buf = self._dev_arena[self._dev_used:self._dev_used + n].view(*shape)
        self._dev_used += n_pad
        return buf

    def staging(self, *shape):
        n = 1
        for d in shape:
            n *= int(d)
        n_pad = -(-n // 16) * 16
        if self._used + n_pad > self._arena.numel():
            raise SrganKernelError("FeedTape: staging arena exhausted")
        buf = self._arena[self._used:self._used + n].view(*shape)
        self._used += n_pad
        return buf

    def add(self, pinned, dev, fill, src=None):
        self.entries.append((pinned, dev, fill, pinned if src is None else src))
        return dev

    def upload(self):
        if self._done is not None:
            self._done.synchronize()          # the previous upload has left the staging buffers
        for pinned, dev, fill, src in self.entries:
            fill(pinned)
            dev.copy_(src, non_blocking=True)
        self._done = torch.cuda.Event()
        self._done.record()


_tape = None


class recording(object):
    """with recording(tape): host feeds are registered on `tape` (used while capturing a CUDA graph)."""

    def __init__(self, tape):
        self.tape = tape

    def __enter__(self):
        global _tape
        self.tape.recording = True
        _tape = self.tape
        return self.tape

    def __exit__(self, *exc):
        global _tape
        self.tape.recording = False
        _tape = None


def host_normal(rows, dim, device):
    """Standard-normal [rows, dim] from the CPU default generator, moved to `device` (the reference draws all
    of its noise this way: pyfiles/util_notebook.py:179,554, pyfiles/model.py:400,461).
    Data parallel: every rank draws the noise of the GLOBAL batch (same seed => same stream) and keeps its
    own rows, so an N-GPU run consumes the RNG exactly like the single-GPU global-batch run."""
    rank, world = dp_rank_world()
    pin = torch.device(device).type == "cuda"
    if _tape is not None and _tape.recording:
        pinned = _tape.staging(rows * world, dim)
        dev = _tape.device_buffer(rows, dim)
        return _tape.add(pinned, dev, lambda buf: torch.randn(buf.shape, out=buf),
                         pinned[rank * rows:(rank + 1) * rows])
    # pinned staging + non-blocking copy: a pageable .to(device) synchronises the stream, i.e. drains the device
    # pipeline at every noise draw (24 times per training step)
    z = torch.randn(rows * world, dim, pin_memory=pin)
    if world > 1:
        z = z[rank * rows:(rank + 1) * rows]
    return z.to(device, non_blocking=pin)


def to_device_async(t, device, dtype=None):
    """Host tensor / array -> device without synchronising the stream (pinned staging, non-blocking copy).  Tensors
    already on the device are only cast."""
    t = torch.as_tensor(t)
    dev = torch.device(device)
    if t.device.type == "cpu" and dev.type == "cuda":
        if dtype is not None:
            t = t.to(dtype)
        return t.pin_memory().to(dev, non_blocking=True)
    return t.to(device=dev, dtype=dtype) if dtype is not None else t.to(dev)


# ----------------------------------------------------------------------------- layout
def _dense_nhwc(x):
    return x.dim() == 4 and x.is_contiguous(memory_format=CL)


def _raw_to_nhwc(x):
    """No-autograd conversion of a 4-D tensor to channels-last storage with our kernel."""
    if _dense_nhwc(x):
        return x
    if x.dtype != torch.float32:
        raise SrganKernelError("bf16 activations only exist in channels-last storage (produced by these kernels); "
                               "got strides %s" % (tuple(x.stride()),))
    if not x.is_contiguous():
        x = x.contiguous()
    N, C, H, W = x.shape
    y = torch.empty((N, C, H, W), dtype=x.dtype, device=x.device, memory_format=CL)
    if x.numel():
        _call("srgan_nchw_to_nhwc", _p(x), _p(y), N, C, H, W, _stream())
    return y


def _raw_to_nchw(x):
    if x.is_contiguous():
        return x
    if not _dense_nhwc(x):
        return x.contiguous()
    N, C, H, W = x.shape
    y = torch.empty((N, C, H, W), dtype=x.dtype, device=x.device)
    if x.numel():
        _call("srgan_nhwc_to_nchw", _p(x), _p(y), N, C, H, W, _stream())
    return y


class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _raw_to_nhwc(x)

    @staticmethod
    def backward(ctx, g):
        return _raw_to_nchw(g)


def to_nhwc(x):
    """Logical NCHW tensor -> same values stored channels-last (differentiable, no-op if already)."""
    _req(x)
    if x.dim() != 4:
        raise ValueError("expected 4D input (got {}D input)".format(x.dim()))
    if _dense_nhwc(x):
        return x
    return _ToNHWC.apply(x)


def _empty_nhwc(N, C, H, W, like, dtype=None):
    return torch.empty((N, C, H, W), dtype=like.dtype if dtype is None else dtype, device=like.device,
                       memory_format=CL)


def _dt(t):
    return DT_BF16 if t.dtype == BF16 else DT_F32


# ----------------------------------------------------------------------------- convolution
def _desc(N, H, W, C, K, R, S, stride, pad):
    d = ConvDesc()
    d.N, d.H, d.W, d.C, d.K, d.R, d.S = N, H, W, C, K, R, S
    d.stride, d.pad = stride, pad
    d.P = (H + 2 * pad - R) // stride + 1
    d.Q = (W + 2 * pad - S) // stride + 1
    d.xs_n = d.xs_h = d.xs_w = d.xs_c = 0
    return d


def _krsc(w):
    """Filter parameter [K,C,R,S] -> tensor whose storage is KRSC (no copy when already so)."""
    if w.is_contiguous(memory_format=CL):
        return w
    return _raw_to_nhwc(w.detach())


def cast_bf16(src, dst):
    """dst (bf16, same storage order and size as src) = round-to-nearest-even(src fp32)."""
    _call("srgan_cast_f32_bf16", _p(src), _p(dst), src.numel(), _stream())


def _shadow16(w):
    """bf16 shadow of an fp32 filter parameter, in the parameter's own storage order (KRSC for 4-D filters).
    Parameters owned by a FusedAdam share one flat bf16 buffer that the optimizer refreshes after every step (the
    backward passes read weights LIVE, see the module docstring, and so do they read the live shadow); any other
    in-place change of the parameter (load_state_dict, copy_) bumps its version counter and refreshes the shadow on
    next use."""
    sh = getattr(w, "_srgan_bf16", None)
    ver, ptr = w._version, w.data_ptr()
    if sh is not None and sh[1] == ver and sh[2] == ptr:
        return sh[0]
    owner = getattr(w, "_srgan_owner", None)
    with torch.no_grad():
        if owner is not None and owner[0]["p"].data_ptr() <= ptr < owner[0]["p"].data_ptr() + owner[0]["p"].numel() * 4:
            st, view, o, k = owner
            if st.get("p16") is None:
                st["p16"] = torch.empty(st["p"].numel(), dtype=BF16, device=st["p"].device)
                cast_bf16(st["p"], st["p16"])
            else:
                cast_bf16(st["p"][o:o + k], st["p16"][o:o + k])
            t = view(st["p16"])
        else:
            src = _krsc(w) if w.dim() == 4 else w.detach().contiguous()
            t = sh[0] if sh is not None and sh[0].shape == w.shape else torch.empty_like(src, dtype=BF16)
            cast_bf16(src, t)
    w._srgan_bf16 = (t, ver, ptr)
    return t


def _conv_weight(w, like):
    """The filter tensor a convolution on `like`-typed activations reads: the KRSC fp32 parameter, or its bf16 shadow."""
    return _shadow16(w) if like.dtype == BF16 else _krsc(w)


def _need16(d, p):
    if not _lib().srgan_conv2d_bf16_supported(d, p):
        raise SrganKernelError(
            "bf16 convolution (%s) not available for N=%d C=%d K=%d %dx%d stride %d: the bf16 trunk needs channel "
            "counts that are multiples of 64" % (("fprop", "dgrad", "wgrad")[p], d.N, d.C, d.K, d.R, d.S, d.stride))


# Per-tile statistics written by the epilogue of the last bf16 FORWARD convolution, waiting for the instance norm that
# consumes its output: (data_ptr of the output, [N, rows, K, 2] fp32, rows).  See srgan_conv2d_fprop_bf16.
# OPT-IN (SRGAN_TILE_STATS=1).  Measured on B200 (profiles/r2k_*): the residual-block convolution goes from 63.5 to
# 78.6 us with the statistics in its epilogue (the kernel is bound by shared-memory bandwidth - TMA fill + MMA operand
# reads - and the column sums need the same port: a shuffle butterfly made it 115.7 us), which is what the saved
# statistics kernel costs (13.9 us + launch): the training step does not move (60.41 vs 60.45 ms, same box).  The norm
# itself becomes one pass (fold + apply: 25.7 vs 33.0 us).  Kept for planes where the conv has smem headroom.
_tile_stats = None
_TILE_STATS = os.environ.get("SRGAN_TILE_STATS", "0") != "0"
# SRGAN_TILE_STATS=2: only for convolutions with <= 128 output channels (64- / 128-wide tiles: tensor pipe 27 - 42 %
# busy, i.e. with epilogue headroom - unlike the 256-wide residual trunk)
_TILE_STATS_MAXK = 128 if os.environ.get("SRGAN_TILE_STATS", "0") == "2" else 1 << 30


def set_tile_stats(on):
    """Enable / disable the conv-epilogue statistics path (see above); returns the previous setting."""
    global _TILE_STATS
    prev, _TILE_STATS = _TILE_STATS, bool(on)
    return prev


def _tile_stats_buffer(d, p, out, plain):
    """fp32 [N, rows, K_out, 2] when the pass can deliver tile statistics for `out` (and the epilogue is plain)."""
    if not _TILE_STATS or not plain or out.shape[1] > _TILE_STATS_MAXK:
        return None, 0
    rows = _lib().srgan_conv2d_bf16_stat_rows(d, p)
    if rows <= 0:
        return None, 0
    return torch.empty((out.shape[0], rows, out.shape[1], 2), dtype=torch.float32, device=out.device), rows


def _take_tile_stats(x):
    """The pending tile statistics if they describe `x` (consumed), else None."""
    global _tile_stats
    ts, _tile_stats = _tile_stats, None
    if ts is not None and ts[0] == x.data_ptr() and ts[1].shape[0] == x.shape[0] and ts[1].shape[2] == x.shape[1]:
        return ts[1], ts[2]
    return None


def _fprop(d, x, w, bias, act, slope, out_dtype=None, want_stats=False):
    """w: the filter in the storage `x` asks for (see _conv_weight).  want_stats (bf16, forward passes): also produce
    the tile statistics for the instance norm that follows."""
    global _tile_stats
    if x.dtype == BF16:
        y = _empty_nhwc(d.N, d.K, d.P, d.Q, x)
        if y.numel() == 0:
            return y
        _need16(d, 0)
        ts, rows = _tile_stats_buffer(d, 0, y, bias is None and act == ACT_NONE) if want_stats else (None, 0)
        _call("srgan_conv2d_fprop_bf16", d, _p(x), _p(w), _p(bias), _p(y), act, slope, _p(ts), _stream())
        _tile_stats = (y.data_ptr(), ts, rows) if ts is not None else None
        return y
    y = _empty_nhwc(d.N, d.K, d.P, d.Q, x)
    if y.numel() == 0:
        return y
    lib = _lib()
    nb = lib.srgan_conv2d_workspace(d, 0, _engine)
    ws = _workspace(x.device, nb) if nb else None
    _call("srgan_conv2d_fprop", d, _p(x), _p(w), _p(bias), _p(y), act, slope, _engine, _p(ws), nb, _stream())
    return y


def _dgrad(d, dy, w, like, addend=None, want_stats=False):
    """dx = dgrad(dy) (+ addend, fused into the epilogue where the engine supports it).  want_stats: see _fprop (the
    forward pass of a transposed convolution is a dgrad)."""
    global _tile_stats
    dx = _empty_nhwc(d.N, d.C, d.H, d.W, like, dtype=dy.dtype)
    if dx.numel() == 0:
        return dx
    lib = _lib()
    if dy.dtype == BF16:
        _need16(d, 1)
        nb = lib.srgan_conv2d_bf16_workspace(d, 1)
        ws = _workspace(dy.device, nb) if nb else None
        fused = addend is not None and d.stride == 1
        if addend is not None:
            addend = _raw_to_nhwc(addend)
            if addend.dtype != BF16:
                raise SrganKernelError("dgrad: the skip gradient must have the storage type of dy")
        ts, rows = _tile_stats_buffer(d, 1, dx, addend is None) if want_stats else (None, 0)
        _call("srgan_conv2d_dgrad_bf16", d, _p(dy), _p(w), _p(addend) if fused else None, _p(dx), _p(ts), _p(ws), nb,
              _stream())
        _tile_stats = (dx.data_ptr(), ts, rows) if ts is not None else None
        if addend is not None and not fused:
            dx.add_(addend)
        return dx
    nb = lib.srgan_conv2d_workspace(d, 1, _engine)
    ws = _workspace(dy.device, nb) if nb else None
    if addend is not None:
        addend = _raw_to_nhwc(addend)
        if lib.srgan_conv2d_dgrad_add_supported(d, _engine):
            _call("srgan_conv2d_dgrad_add", d, _p(dy), _p(w), _p(addend), _p(dx), _engine, _p(ws), nb, _stream())
            return dx
    _call("srgan_conv2d_dgrad", d, _p(dy), _p(w), _p(dx), _engine, _p(ws), nb, _stream())
    if addend is not None:
        dx.add_(addend)
    return dx


_direct_grads = False
_NO_DIRECT_GRADS = os.environ.get("SRGAN_DBG_NO_DIRECT_GRADS", "0") != "0"      # bring-up: autograd accumulates every gradient


# Later contributions (a generator / encoder that runs several times inside one loss) are parked in up to
# _MAX_PENDING zeroed copies of the optimizer's flat gradient buffer - contribution k + 1 of a parameter goes into copy
# k at the parameter's offset - and folded into the gradient buffer by ONE launch per optimizer when the backward pass
# ends (srgan_grad_fold, arrival order: bit-identical to autograd's tensor-by-tensor adds; ~400 small launches less per
# step).  0 restores the tensor-by-tensor accumulation.
_MAX_PENDING = max(0, min(2, int(os.environ.get("SRGAN_PENDING_GRADS", "2"))))
_pending_states = []          # flat states (FusedAdam) with parked contributions, in first-use order


class direct_param_grads(object):
    """with direct_param_grads(): backward kernels write the gradient contributions of a parameter straight into the
    optimizer's flat buffers (FusedAdam.zero_grad() arms the parameters) and hand autograd None for them - no
    per-parameter accumulation kernels: the FIRST contribution goes into the gradient buffer itself, the next
    _MAX_PENDING ones into parking copies that are folded in when the block exits; anything beyond accumulates as
    usual.  Only for backward() calls that follow a FusedAdam.zero_grad() and accumulate into .grad (the trainers);
    never around torch.autograd.grad()."""

    def __enter__(self):
        global _direct_grads
        self.prev, _direct_grads = _direct_grads, not _NO_DIRECT_GRADS

    def __exit__(self, *exc):
        global _direct_grads
        _direct_grads = self.prev
        if not self.prev:
            fold_pending_grads()


def fold_pending_grads():
    """Fold the parked gradient contributions into the flat gradient buffers (one launch per optimizer group)."""
    while _pending_states:
        st = _pending_states.pop()
        used, st["pend_used"] = st.get("pend_used", 0), 0
        if used:
            pend = st["pend"]
            _call("srgan_grad_fold", _p(st["g"]), _p(pend[0]), _p(pend[1]) if used > 1 else None, st["g"].numel(),
                  _stream())


def _grad_sink(p, wanted, channels_last=False):
    """The flat-buffer view to write this parameter's gradient into (overwrite semantics), or None."""
    if not (wanted and _direct_grads) or p is None:
        return None
    k = getattr(p, "_srgan_arrivals", None)
    if k is None:
        return None
    if k == 0:
        g = p.grad
        if g is None or g.shape != p.shape:
            return None
        if not (g.is_contiguous(memory_format=CL) if (channels_last and g.dim() == 4) else g.is_contiguous()):
            p._srgan_arrivals = None
            return None
        p._srgan_arrivals = 1
        return g
    owner = getattr(p, "_srgan_owner", None)
    if k > _MAX_PENDING or owner is None:
        return None
    st, view = owner[0], owner[1]
    pend = st.setdefault("pend", [])
    while len(pend) < k:
        pend.append(torch.zeros_like(st["g"]))
    views = p.__dict__.setdefault("_srgan_pend_views", {})
    t = views.get(k)
    if t is None or t.data_ptr() < pend[k - 1].data_ptr() or \
            t.data_ptr() >= pend[k - 1].data_ptr() + pend[k - 1].numel() * 4:
        t = views[k] = view(pend[k - 1])
    if st.get("pend_used", 0) == 0:
        _pending_states.append(st)
    st["pend_used"] = max(st.get("pend_used", 0), k)
    p._srgan_arrivals = k + 1
    return t


def _wgrad(d, x, dy, want_w, want_b, dw_out=None, db_out=None):
    """dw_out / db_out: write the gradients there (dense KRSC / [K]) instead of into fresh tensors."""
    dw = dw_out if dw_out is not None else (
        torch.empty((d.K, d.C, d.R, d.S), dtype=torch.float32, device=x.device, memory_format=CL) if want_w else None)
    db = db_out if db_out is not None else (
        torch.empty((d.K,), dtype=torch.float32, device=x.device) if want_b else None)
    if d.N == 0:
        if dw is not None:
            dw.zero_()
        if db is not None:
            db.zero_()
        return dw, db
    lib = _lib()
    if x.dtype == BF16 or dy.dtype == BF16:
        if x.dtype != dy.dtype or want_b or db is not None:
            raise SrganKernelError("bf16 wgrad: x and dy must both be bf16 and the layer bias-free")
        _need16(d, 2)
        nb = lib.srgan_conv2d_bf16_workspace(d, 2)
        ws = _workspace(x.device, nb) if nb else None
        if dw is not None:
            _call("srgan_conv2d_wgrad_bf16", d, _p(x), _p(dy), _p(dw), _p(ws), nb, _stream())
        return dw, None
    nb = lib.srgan_conv2d_workspace(d, 2, _engine)
    ws = _workspace(x.device, nb) if nb else None
    _call("srgan_conv2d_wgrad", d, _p(x), _p(dy), _p(dw), _p(db), _engine, _p(ws), nb, _stream())
    return dw, db


def wgrad_plan(d):
    """(pixel splits, CTAs) of the tcgen05 wgrad launch for this layer (introspection for tests / tools)."""
    import ctypes
    s_, c_ = ctypes.c_int(0), ctypes.c_int(0)
    check(_lib().srgan_conv2d_wgrad_plan(ctypes.byref(d), ctypes.addressof(s_), ctypes.addressof(c_)),
          "srgan_conv2d_wgrad_plan")
    return s_.value, c_.value


def _act_bwd(dy, y, act, slope):
    if act == ACT_NONE:
        return dy
    if dy.dtype != y.dtype:
        raise SrganKernelError("act_bwd: dy is %s, the activation output was %s" % (dy.dtype, y.dtype))
    dz = torch.empty_like(y)
    if y.dtype == BF16:           # the bf16 discriminator tower (LeakyReLU fused into the conv epilogues)
        _call("srgan_act_bwd_bf16", _p(dy), _p(y), _p(dz), y.numel(), act, slope, _stream())
    else:
        _call("srgan_act_bwd", _p(dy), _p(y), _p(dz), y.numel(), act, slope, _stream())
    return dz


class _CastF32Fn(torch.autograd.Function):
    """bf16 -> fp32 (values unchanged); the gradient comes back rounded to bf16."""

    @staticmethod
    def forward(ctx, x):
        y = torch.empty_like(x, dtype=torch.float32)
        if x.numel():
            _call("srgan_cast_bf16_f32", _p(x), _p(y), x.numel(), _stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _raw_to_nhwc(dy) if dy.dim() == 4 else dy.contiguous()
        dx = torch.empty_like(dy, dtype=BF16)
        if dy.numel():
            cast_bf16(dy, dx)
        return dx


def cast_f32(x):
    """A bf16 activation as fp32 (no-op for fp32 inputs): the boundary between a bf16 tower and fp32 heads."""
    if x.dtype == torch.float32:
        return x
    _req(x)
    return _CastF32Fn.apply(_raw_to_nhwc(x) if x.dim() == 4 else x.contiguous())


class _Conv2dFn(torch.autograd.Function):
    """y = act(conv2d(x, w) + b); ref: nn.Conv2d (+LeakyReLU / Tanh that follow it), pyfiles/model.py."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, act, slope):
        x = _raw_to_nhwc(x)
        N, C, H, W = x.shape
        K, C2, R, S = weight.shape
        if C2 != C:
            raise ValueError("conv2d: input has %d channels, filter expects %d" % (C, C2))
        d = _desc(N, H, W, C, K, R, S, stride, pad)
        y = _fprop(d, x, _conv_weight(weight, x), bias, act, slope, want_stats=True)
        ctx.d, ctx.act, ctx.slope = d, act, slope
        ctx.weight, ctx.bias, ctx.has_bias = weight, bias, bias is not None
        ctx.save_for_backward(x, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dz = _act_bwd(_raw_to_nhwc(dy), y, ctx.act, ctx.slope)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _dgrad(ctx.d, dz, _conv_weight(ctx.weight, dz), x)       # live weight (see module docstring)
        want_w = ctx.needs_input_grad[1]
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if want_w or want_b:
            sink_w, sink_b = _grad_sink(ctx.weight, want_w, True), _grad_sink(ctx.bias, want_b)
            dw, db = _wgrad(ctx.d, x, dz, want_w, want_b, sink_w, sink_b)
            dw = None if sink_w is not None else dw
            db = None if sink_b is not None else db
        return dx, dw, db, None, None, None, None


_THIN16 = os.environ.get("SRGAN_THIN16", "1") != "0"      # A/B switch: 0 keeps the RGB layers of the bf16 engine in fp32


def thin16_supported(N, H, W, C, K, R, stride, pad, need_dgrad=True):
    """Can this RGB layer (C <= 4 or K <= 4) run with its fat side in bf16 (srgan_conv2d_*_thin16)?"""
    if not _THIN16 or not (C <= 4 or K <= 4):
        return False
    d = _desc(N, H, W, C, K, R, R, stride, pad)
    lib = _lib()
    passes = (0, 1, 2) if need_dgrad else (0, 2)
    return all(lib.srgan_conv2d_thin16_supported(d, p) == 1 for p in passes)


class _Conv2dThin16Fn(torch.autograd.Function):
    """The RGB stem / head of the bf16 trunk: y = act(conv2d(x, w) + b) with the 3-channel side in fp32 and the fat
    side in bf16 (stem: fp32 image -> bf16 activation; head: bf16 activation -> fp32 image).  ref: first and last
    convolution of SingleGenerator, pyfiles/model.py:280-318."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, act, slope):
        x = _raw_to_nhwc(x)
        N, C, H, W = x.shape
        K, C2, R, S = weight.shape
        if C2 != C:
            raise ValueError("conv2d: input has %d channels, filter expects %d" % (C, C2))
        thin_in = C <= 4
        if (x.dtype == BF16) == thin_in:
            raise SrganKernelError("thin16 conv: the 3-channel side must be fp32 and the fat side bf16")
        d = _desc(N, H, W, C, K, R, S, stride, pad)
        lib = _lib()
        if lib.srgan_conv2d_thin16_supported(d, 0) != 1:
            raise SrganKernelError("thin16 conv: shape not supported (N=%d C=%d K=%d %dx%d stride %d)"
                                   % (N, C, K, R, S, stride))
        y = _empty_nhwc(N, K, d.P, d.Q, x, dtype=BF16 if thin_in else torch.float32)
        if y.numel():
            nb = lib.srgan_conv2d_thin16_workspace(d, 0)
            ws = _workspace(x.device, nb) if nb else None
            _call("srgan_conv2d_fprop_thin16", d, _p(x), _p(_krsc(weight)), _p(bias), _p(y), act, slope, _p(ws), nb,
                  _stream())
        ctx.d, ctx.act, ctx.slope = d, act, slope
        ctx.weight, ctx.bias, ctx.has_bias = weight, bias, bias is not None
        ctx.save_for_backward(x, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        d = ctx.d
        lib = _lib()
        dz = _act_bwd(_raw_to_nhwc(dy), y, ctx.act, ctx.slope)
        dx = dw = db = None
        if d.N == 0:
            return (torch.zeros_like(x) if ctx.needs_input_grad[0] else None), None, None, None, None, None, None
        if ctx.needs_input_grad[0]:
            if lib.srgan_conv2d_thin16_supported(d, 1) == 1:
                dx = torch.empty_like(x)
                nb = lib.srgan_conv2d_thin16_workspace(d, 1)
                ws = _workspace(x.device, nb) if nb else None
                _call("srgan_conv2d_dgrad_thin16", d, _p(dz), _p(_krsc(ctx.weight)), _p(dx), _p(ws), nb, _stream())
            elif d.C <= 4 and dz.dtype == BF16:
                # strided stems (discriminator, encoder): the image gradient goes through an fp32 copy of dz and the
                # fp32 engine's dgrad (rare: only passes that back-propagate into the image)
                dz32 = torch.empty_like(dz, dtype=torch.float32)
                _call("srgan_cast_bf16_f32", _p(dz), _p(dz32), dz.numel(), _stream())
                dx = _dgrad(d, dz32, _krsc(ctx.weight), x)
            else:
                raise SrganKernelError("thin16 conv: no input-gradient kernel for this shape")
        want_w = ctx.needs_input_grad[1]
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if want_w or want_b:
            sink_w, sink_b = _grad_sink(ctx.weight, want_w, True), _grad_sink(ctx.bias, want_b)
            dw = sink_w if sink_w is not None else (
                torch.empty((d.K, d.C, d.R, d.S), dtype=torch.float32, device=x.device, memory_format=CL)
                if want_w else None)
            db = sink_b if sink_b is not None else (
                torch.empty((d.K,), dtype=torch.float32, device=x.device) if want_b else None)
            nb = lib.srgan_conv2d_thin16_workspace(d, 2)
            ws = _workspace(x.device, nb) if nb else None
            _call("srgan_conv2d_wgrad_thin16", d, _p(x), _p(dz), _p(dw), _p(db), _p(ws), nb, _stream())
            dw = None if sink_w is not None else dw
            db = None if sink_b is not None else db
        return dx, dw, db, None, None, None, None


class _Conv2dSkipFn(torch.autograd.Function):
    """(y, x) = (conv2d(x, w), x): the first convolution of a residual block together with the skip connection that
    leaves the same tensor (ref SingleResidualBlock.forward pyfiles/model.py:196-201).  Both gradients of x arrive in
    ONE backward call, so the skip gradient is added in the dgrad epilogue instead of by autograd's separate add."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad):
        x = _raw_to_nhwc(x)
        N, C, H, W = x.shape
        K, C2, R, S = weight.shape
        if C2 != C:
            raise ValueError("conv2d: input has %d channels, filter expects %d" % (C, C2))
        d = _desc(N, H, W, C, K, R, S, stride, pad)
        y = _fprop(d, x, _conv_weight(weight, x), None, ACT_NONE, 0.0, want_stats=True)
        ctx.d, ctx.weight = d, weight
        ctx.save_for_backward(x)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dskip):
        x, = ctx.saved_tensors
        dz = _raw_to_nhwc(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = _dgrad(ctx.d, dz, _conv_weight(ctx.weight, dz), x, addend=dskip)      # live weight (see module docstring)
        if ctx.needs_input_grad[1]:
            sink = _grad_sink(ctx.weight, True, True)
            dw, _ = _wgrad(ctx.d, x, dz, True, False, sink)
            dw = None if sink is not None else dw
        return dx, dw, None, None


class _ConvTranspose2dFn(torch.autograd.Function):
    """ref: nn.ConvTranspose2d pyfiles/model.py:227,230.  weight [Cin,Cout,R,S] stored channels-last is
    exactly the KRSC filter of the mirrored convolution (K=Cin, C=Cout): forward = its dgrad."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad):
        x = _raw_to_nhwc(x)
        N, Cin, H, W = x.shape
        Cin2, Cout, R, S = weight.shape
        if Cin2 != Cin:
            raise ValueError("conv_transpose2d: channel mismatch")
        Ho = (H - 1) * stride - 2 * pad + R
        Wo = (W - 1) * stride - 2 * pad + S
        d = _desc(N, Ho, Wo, Cout, Cin, R, S, stride, pad)     # mirrored conv: (Ho,Wo,Cout) -> (H,W,Cin)
        assert d.P == H and d.Q == W
        y = _dgrad(d, x, _conv_weight(weight, x), x, want_stats=True)
        ctx.d, ctx.weight = d, weight
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _raw_to_nhwc(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = _fprop(ctx.d, dy, _conv_weight(ctx.weight, dy), None, ACT_NONE, 0.0)
        if ctx.needs_input_grad[1]:
            sink = _grad_sink(ctx.weight, True, True)
            dw, _ = _wgrad(ctx.d, dy, x, True, False, sink)
            dw = None if sink is not None else dw
        return dx, dw, None, None


class _ReflectPadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pad):
        x = _raw_to_nhwc(x)
        N, C, H, W = x.shape
        y = _empty_nhwc(N, C, H + 2 * pad, W + 2 * pad, x)
        _call("srgan_reflect_pad_fwd_bf16" if x.dtype == BF16 else "srgan_reflect_pad_fwd", _p(x), _p(y), N, H, W, C,
              pad, _stream())
        ctx.shape, ctx.pad = (N, C, H, W), pad
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W = ctx.shape
        dy = _raw_to_nhwc(dy)
        dx = _empty_nhwc(N, C, H, W, dy)
        _call("srgan_reflect_pad_bwd_bf16" if dy.dtype == BF16 else "srgan_reflect_pad_bwd", _p(dy), _p(dx), N, H, W, C,
              ctx.pad, _stream())
        return dx, None


def conv2d(x, weight, bias=None, stride=1, padding=0, padding_mode="zeros", act=ACT_NONE, slope=0.0, out_dtype=None):
    """out_dtype=torch.bfloat16 on an fp32 3-channel input, or a bf16 input of a <= 4-filter layer: the RGB layers of
    the bf16 trunk (_Conv2dThin16Fn); otherwise the output has the storage type of x."""
    _req(x, weight, bias)
    if x.dim() != 4:
        raise ValueError("expected 4D input (got {}D input)".format(x.dim()))
    if padding_mode == "reflect" and padding > 0:
        x = _ReflectPadFn.apply(x, int(padding))
        padding = 0
    elif padding_mode not in ("zeros", "reflect"):
        raise NotImplementedError("padding_mode %r" % (padding_mode,))
    stem16 = out_dtype == BF16 and x.dtype == torch.float32
    head16 = x.dtype == BF16 and weight.shape[0] <= 4
    if stem16 or head16:
        return _Conv2dThin16Fn.apply(x, weight, bias, int(stride), int(padding), int(act), float(slope))
    if out_dtype is not None and out_dtype != x.dtype:
        raise SrganKernelError("conv2d: out_dtype %s on a %s input is only available for the RGB stem" % (out_dtype, x.dtype))
    return _Conv2dFn.apply(x, weight, bias, int(stride), int(padding), int(act), float(slope))


def conv2d_skip(x, weight, stride=1, padding=0):
    """Returns (conv2d(x, weight), x'): x' is x, routed through the same autograd node (see _Conv2dSkipFn)."""
    _req(x, weight)
    if x.dim() != 4:
        raise ValueError("expected 4D input (got {}D input)".format(x.dim()))
    return _Conv2dSkipFn.apply(x, weight, int(stride), int(padding))


def conv_transpose2d(x, weight, stride=1, padding=0):
    _req(x, weight)
    return _ConvTranspose2dFn.apply(x, weight, int(stride), int(padding))


class _LinearFn(torch.autograd.Function):
    """y = x @ w.T + b as a 1x1 convolution over a [N,1,1,F] activation (ref: nn.Linear fcmean/fcvar/fcclass
    pyfiles/model.py:395-396,455-457)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        N, F = x.shape
        J = weight.shape[0]
        d = _desc(N, 1, 1, F, J, 1, 1, 1, 0)
        y = _fprop(d, x, weight.contiguous(), bias, ACT_NONE, 0.0).view(N, J)
        ctx.d, ctx.weight, ctx.bias, ctx.has_bias = d, weight, bias, bias is not None
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        d = ctx.d
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _dgrad(d, dy, ctx.weight.contiguous(), x).view(d.N, d.C)
        want_w = ctx.needs_input_grad[1]
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if want_w or want_b:
            sink_w, sink_b = _grad_sink(ctx.weight, want_w), _grad_sink(ctx.bias, want_b)
            dw, db = _wgrad(d, x, dy, want_w, want_b, sink_w, sink_b)
            dw = None if (sink_w is not None or dw is None) else dw.view(d.K, d.C)
            db = None if sink_b is not None else db
        return dx, dw, db


def linear(x, weight, bias=None):
    _req(x, weight, bias)
    return _LinearFn.apply(x, weight, bias)


# ----------------------------------------------------------------------------- normalisation
class _CondBiasFn(torch.autograd.Function):
    """t = tanh(con @ W.T + b); ref: ConBias pyfiles/model.py:16-19,57."""

    @staticmethod
    def forward(ctx, con, weight, bias):
        con = con.contiguous()
        N, J = con.shape
        C = weight.shape[0]
        t = torch.empty((N, C), dtype=torch.float32, device=con.device)
        if N:
            _call("srgan_condbias_fwd", _p(con), _p(weight), _p(bias), _p(t), N, J, C, _stream())
        ctx.weight, ctx.bias = weight, bias
        ctx.save_for_backward(con, t)
        return t

    @staticmethod
    def backward(ctx, dt):
        con, t = ctx.saved_tensors
        w = ctx.weight
        N, J = con.shape
        C = w.shape[0]
        dt = dt.contiguous()
        sink_w, sink_b = _grad_sink(w, ctx.needs_input_grad[1]), _grad_sink(ctx.bias, ctx.needs_input_grad[2])
        dw = sink_w if sink_w is not None else (torch.empty_like(w) if ctx.needs_input_grad[1] else None)
        db = sink_b if sink_b is not None else (
            torch.empty((C,), dtype=torch.float32, device=w.device) if ctx.needs_input_grad[2] else None)
        dcon = torch.empty_like(con) if ctx.needs_input_grad[0] else None
        _call("srgan_condbias_bwd", _p(dt), _p(t), _p(con), _p(w), _p(dw), _p(db), _p(dcon), N, J, C, _stream())
        return dcon, (None if sink_w is not None else dw), (None if sink_b is not None else db)


def cond_bias(con, weight, bias):
    _req(con, weight, bias)
    if not weight.is_contiguous():
        raise SrganKernelError("cond_bias: weight must be contiguous")
    return _CondBiasFn.apply(con, weight, bias)


class _InstanceNormFn(torch.autograd.Function):
    """y = act(((x-mean)*rstd + cbias) * gamma + beta) (+ residual).
    ref: CBINorm2d.forward pyfiles/model.py:54-67, nn.InstanceNorm2d(affine=False) :178, ReLU/LeakyReLU/add.
    `out_dtype`: storage type of y (and of the residual): fp32 or bf16 independently of x's (srgan_inorm_*_mixed);
    dx comes back in x's storage type, dy arrives in y's."""

    @staticmethod
    def forward(ctx, x, gamma, beta, cbias, residual, eps, act, slope, out_dtype):
        x = _raw_to_nhwc(x)
        N, C, H, W = x.shape
        out_dtype = x.dtype if out_dtype is None else out_dtype
        if residual is not None:
            residual = _raw_to_nhwc(residual)
            if residual.dtype != out_dtype:
                raise SrganKernelError("instance_norm_act: the residual must have the output's storage type")
        if cbias is not None:
            cbias = cbias.contiguous()
        y = torch.empty_like(x, dtype=out_dtype)
        mean = torch.empty((N, C), dtype=torch.float32, device=x.device)
        rstd = torch.empty((N, C), dtype=torch.float32, device=x.device)
        mixed = x.dtype != torch.float32 or out_dtype != torch.float32
        tiles = _take_tile_stats(x)
        if x.numel():
            if mixed:
                nb = _lib().srgan_inorm_mixed_workspace(N, H * W, C)
                ws = _workspace(x.device, nb)
                if tiles is not None:
                    # the convolution that produced x left per-tile sums: no statistics pass over x
                    _call("srgan_inorm_stats_from_tiles", _p(tiles[0]), tiles[1], N, H * W, C, eps, _p(mean), _p(rstd),
                          _stream())
                _call("srgan_inorm_fwd_mixed", _p(x), _dt(x), _p(y), _dt(y), _p(mean), _p(rstd), _p(gamma), _p(beta),
                      _p(cbias), _p(residual), N, H * W, C, eps, act, slope, int(tiles is not None), _p(ws), nb,
                      _p(_norm_counters(x.device, N, C)), _stream())
            else:
                nb = _lib().srgan_inorm_workspace(N, H * W, C)
                ws = _workspace(x.device, nb)
                _call("srgan_inorm_fwd", _p(x), _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(cbias),
                      _p(residual), N, H * W, C, eps, act, slope, _p(ws), nb, _stream())
        ctx.gamma, ctx.beta = gamma, beta
        ctx.act, ctx.slope = act, slope
        ctx.has_res = residual is not None
        ctx.out_dtype, ctx.mixed = out_dtype, mixed
        ctx.save_for_backward(x, mean, rstd, cbias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, cbias = ctx.saved_tensors
        gamma, beta = ctx.gamma, ctx.beta
        N, C, H, W = x.shape
        dy = _raw_to_nhwc(dy)
        if dy.dtype != ctx.out_dtype:
            raise SrganKernelError("instance_norm_act backward: dy is %s, the forward output was %s"
                                   % (dy.dtype, ctx.out_dtype))
        dx = torch.empty_like(x)
        s1 = torch.empty((N, C), dtype=torch.float32, device=x.device)
        s2 = torch.empty((N, C), dtype=torch.float32, device=x.device)
        if x.numel():
            if ctx.mixed:
                nb = _lib().srgan_inorm_mixed_workspace(N, H * W, C)
                ws = _workspace(x.device, nb)
                _call("srgan_inorm_bwd_mixed", _p(dy), _dt(dy), _p(x), _dt(x), _p(mean), _p(rstd), _p(gamma), _p(beta),
                      _p(cbias), _p(dx), _p(s1), _p(s2), N, H * W, C, ctx.act, ctx.slope, _p(ws), nb,
                      _p(_norm_counters(x.device, N, C)), _stream())
            else:
                nb = _lib().srgan_inorm_workspace(N, H * W, C)
                ws = _workspace(x.device, nb)
                _call("srgan_inorm_bwd", _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(cbias), _p(dx),
                      _p(s1), _p(s2), N, H * W, C, ctx.act, ctx.slope, _p(ws), nb, _stream())
        dgamma = dbeta = dcb = None
        need_g = gamma is not None and ctx.needs_input_grad[1]
        need_b = beta is not None and ctx.needs_input_grad[2]
        need_c = cbias is not None and ctx.needs_input_grad[3]
        if need_g or need_b or need_c:
            sink_g, sink_b = _grad_sink(gamma, need_g), _grad_sink(beta, need_b)
            dgamma = sink_g if sink_g is not None else (torch.empty_like(gamma) if need_g else None)
            dbeta = sink_b if sink_b is not None else (torch.empty_like(beta) if need_b else None)
            dcb = torch.empty_like(cbias) if need_c else None
            _call("srgan_inorm_param_grads", _p(s1), _p(s2), _p(gamma), _p(cbias), _p(dgamma), _p(dbeta), _p(dcb),
                  N, C, _stream())
            dgamma = None if sink_g is not None else dgamma
            dbeta = None if sink_b is not None else dbeta
        dres = dy if ctx.has_res and ctx.needs_input_grad[4] else None
        return (dx if ctx.needs_input_grad[0] else None), dgamma, dbeta, dcb, dres, None, None, None, None


def instance_norm_act(x, gamma=None, beta=None, cbias=None, residual=None, eps=1e-5, act=ACT_NONE, slope=0.0,
                      out_dtype=None):
    _req(x, gamma, beta, cbias, residual)
    if x.dim() != 4:
        raise ValueError("expected 4D input (got {}D input)".format(x.dim()))
    if x.shape[1] % 8:
        raise SrganKernelError("instance_norm_act: channel count must be a multiple of 8 (got %d)" % x.shape[1])
    if residual is not None and act != ACT_NONE:
        raise ValueError("residual add is only fused with act=none")
    if out_dtype not in (None, torch.float32, BF16):
        raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")
    return _InstanceNormFn.apply(x, gamma, beta, cbias, residual, float(eps), int(act), float(slope), out_dtype)


def _gather_rows(t):
    """[N_loc, C] table -> (all-gathered [N_all, C] table, first local row).  Ranks hold equal batch slices."""
    rank, world = dp_rank_world()
    if world == 1:
        return t, 0
    import torch.distributed as dist
    out = torch.empty((t.shape[0] * world, t.shape[1]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous())
    return out, rank * t.shape[0]


class _BatchNormFn(torch.autograd.Function):
    """Batch-statistics norms.  cond=True: CBBNorm2d, y = act(((x - mean_hw(x)) * r_c + cbias) * gamma + beta)
    (ref _CBBNorm.forward pyfiles/model.py:121-148: BN without affine, minus its own spatial mean, plus the
    conditional bias, then weight/bias); cond=False: nn.BatchNorm2d(affine=True) (ref get_norm_layer :175).
    Data parallel: the [N, C] tables are all-gathered, so statistics and gradients equal the global-batch ones."""

    @staticmethod
    def forward(ctx, x, gamma, beta, cbias, residual, running_mean, running_var, training, momentum, eps, cond, act,
                slope, sync):
        x = _raw_to_nhwc(x)
        N, C, H, W = x.shape
        if residual is not None:
            residual = _raw_to_nhwc(residual)
        if cbias is not None:
            cbias = cbias.contiguous()
        dev = x.device
        y = torch.empty_like(x)
        mean = torch.empty((N, C), dtype=torch.float32, device=dev)
        rstd = torch.empty((N, C), dtype=torch.float32, device=dev)
        bmean = torch.empty((C,), dtype=torch.float32, device=dev)
        if x.numel():
            mnc = torch.empty((N, C), dtype=torch.float32, device=dev)
            m2 = torch.empty((N, C), dtype=torch.float32, device=dev)
            nb = _lib().srgan_inorm_workspace(N, H * W, C)
            ws = _workspace(dev, nb)
            _call("srgan_bnorm_image_stats", _p(x), _p(mnc), _p(m2), N, H * W, C, _p(ws), nb, _stream())
            if sync and training:
                both, n0 = _gather_rows(torch.cat([mnc, m2], 1))
                mnc_all, m2_all = both[:, :C].contiguous(), both[:, C:].contiguous()
            else:
                mnc_all, m2_all, n0 = mnc, m2, 0
            _call("srgan_bnorm_batch_stats", _p(mnc_all), _p(m2_all), mnc_all.shape[0], n0, N, H * W, C, eps,
                  int(cond), int(training), _p(running_mean), _p(running_var), momentum, _p(mean), _p(rstd),
                  _p(bmean), _stream())
            _call("srgan_bnorm_apply", _p(x), _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(cbias), _p(residual),
                  N, H * W, C, act, slope, _stream())
        ctx.gamma, ctx.beta = gamma, beta
        ctx.act, ctx.slope, ctx.cond, ctx.training, ctx.sync = act, slope, cond, training, sync
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, mean, rstd, bmean, cbias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, bmean, cbias = ctx.saved_tensors
        gamma, beta = ctx.gamma, ctx.beta
        N, C, H, W = x.shape
        dev = x.device
        dy = _raw_to_nhwc(dy)
        dx = torch.empty_like(x)
        s1 = torch.empty((N, C), dtype=torch.float32, device=dev)
        s2 = torch.empty((N, C), dtype=torch.float32, device=dev)
        if x.numel():
            nb = _lib().srgan_inorm_workspace(N, H * W, C)
            ws = _workspace(dev, nb)
            _call("srgan_bnorm_bwd_sums", _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(cbias), _p(s1),
                  _p(s2), N, H * W, C, ctx.act, ctx.slope, _p(ws), nb, _stream())
            if ctx.sync and ctx.training:
                both, n0 = _gather_rows(torch.cat([s1, s2], 1))
                s1_all, s2_all = both[:, :C].contiguous(), both[:, C:].contiguous()
            else:
                s1_all, s2_all, n0 = s1, s2, 0
            m1 = torch.empty((N, C), dtype=torch.float32, device=dev)
            m2 = torch.empty((N, C), dtype=torch.float32, device=dev)
            _call("srgan_bnorm_bwd_coeffs", _p(s1_all), _p(s2_all), _p(mean), _p(rstd), _p(bmean), s1_all.shape[0], n0,
                  N, H * W, C, int(ctx.cond), int(ctx.training), _p(m1), _p(m2), _stream())
            _call("srgan_bnorm_bwd_apply", _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(cbias), _p(m1),
                  _p(m2), _p(dx), N, H * W, C, ctx.act, ctx.slope, _stream())
        dgamma = dbeta = dcb = None
        need_g = gamma is not None and ctx.needs_input_grad[1]
        need_b = beta is not None and ctx.needs_input_grad[2]
        need_c = cbias is not None and ctx.needs_input_grad[3]
        if need_g or need_b or need_c:
            sink_g, sink_b = _grad_sink(gamma, need_g), _grad_sink(beta, need_b)
            dgamma = sink_g if sink_g is not None else (torch.empty_like(gamma) if need_g else None)
            dbeta = sink_b if sink_b is not None else (torch.empty_like(beta) if need_b else None)
            dcb = torch.empty_like(cbias) if need_c else None
            _call("srgan_inorm_param_grads", _p(s1), _p(s2), _p(gamma), _p(cbias), _p(dgamma), _p(dbeta), _p(dcb),
                  N, C, _stream())
            dgamma = None if sink_g is not None else dgamma
            dbeta = None if sink_b is not None else dbeta
        dres = dy if ctx.has_res and ctx.needs_input_grad[4] else None
        return ((dx if ctx.needs_input_grad[0] else None), dgamma, dbeta, dcb, dres) + (None,) * 9


def batch_norm_act(x, gamma=None, beta=None, cbias=None, residual=None, running_mean=None, running_var=None,
                   training=True, momentum=0.1, eps=1e-5, cond=False, act=ACT_NONE, slope=0.0, sync=True):
    """BatchNorm2d (cond=False) / CBBNorm2d (cond=True, cbias = tanh(Linear(con))) with fused activation.
    Updates running_mean / running_var in place when training."""
    _req(x, gamma, beta, cbias, residual, running_mean, running_var)
    if x.dim() != 4:
        raise ValueError("expected 4D input (got {}D input)".format(x.dim()))
    if x.shape[1] % 8:
        raise SrganKernelError("batch_norm_act: channel count must be a multiple of 8 (got %d)" % x.shape[1])
    if residual is not None and act != ACT_NONE:
        raise ValueError("residual add is only fused with act=none")
    if not training and (running_mean is None or running_var is None):
        raise ValueError("evaluation mode needs running statistics")
    return _BatchNormFn.apply(x, gamma, beta, cbias, residual, running_mean, running_var, bool(training),
                              float(momentum), float(eps), bool(cond), int(act), float(slope), bool(sync))


# ----------------------------------------------------------------------------- pooling
def _pool_fn(fwd_name, bwd_name, out_hw):
    class _Pool(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            x = _raw_to_nhwc(x)
            N, C, H, W = x.shape
            P, Q = out_hw(H, W)
            y = _empty_nhwc(N, C, P, Q, x)
            if y.numel():
                _call(fwd_name, _p(x), _p(y), N, H, W, C, _stream())
            ctx.shape = (N, C, H, W)
            return y

        @staticmethod
        def backward(ctx, dy):
            N, C, H, W = ctx.shape
            dy = _raw_to_nhwc(dy)
            dx = _empty_nhwc(N, C, H, W, dy)
            if dx.numel():
                _call(bwd_name, _p(dy), _p(dx), N, H, W, C, _stream())
            return dx
    return _Pool


_AvgPool2 = _pool_fn("srgan_avgpool2_fwd", "srgan_avgpool2_bwd", lambda H, W: (H // 2, W // 2))
_AvgPool3s2 = _pool_fn("srgan_avgpool3s2_fwd", "srgan_avgpool3s2_bwd",
                       lambda H, W: ((H - 1) // 2 + 1, (W - 1) // 2 + 1))


def avg_pool2(x):
    """nn.AvgPool2d(2, 2); ref pyfiles/model.py:365,368,426,429."""
    _req(x)
    return _AvgPool2.apply(x)


def avg_pool3s2(x):
    """nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=False); ref pyfiles/model.py:286,324."""
    _req(x)
    return _AvgPool3s2.apply(x)


class _AvgPool2AddFn(torch.autograd.Function):
    """a may be bf16 (the big pre-pool tensor of an encoder block under the bf16 engine); b and the result are fp32."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _raw_to_nhwc(a), _raw_to_nhwc(b)
        N, C, H, W = a.shape
        if b.dtype != torch.float32:
            raise SrganKernelError("avg_pool2_add: the shortcut operand is fp32")
        y = torch.empty_like(b)
        if tuple(b.shape) != (N, C, H // 2, W // 2):
            raise ValueError("avg_pool2_add: shape mismatch")
        if y.numel():
            _call("srgan_avgpool2_add_fwd_mixed" if a.dtype == BF16 else "srgan_avgpool2_add_fwd", _p(a), _p(b), _p(y),
                  N, H, W, C, _stream())
        ctx.shape, ctx.a_dtype = (N, C, H, W), a.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W = ctx.shape
        dy = _raw_to_nhwc(dy)
        da = None
        if ctx.needs_input_grad[0]:
            da = _empty_nhwc(N, C, H, W, dy, dtype=ctx.a_dtype)
            if da.numel():
                _call("srgan_avgpool2_bwd_mixed" if ctx.a_dtype == BF16 else "srgan_avgpool2_bwd", _p(dy), _p(da), N, H,
                      W, C, _stream())
        return da, (dy if ctx.needs_input_grad[1] else None)


def avg_pool2_add(a, b):
    """avgpool2(a) + b -- tail of the encoder block (ref pyfiles/model.py:374-375,435-436)."""
    _req(a, b)
    return _AvgPool2AddFn.apply(a, b)


class _LReluGapFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, slope):
        x = _raw_to_nhwc(x)
        N, C, H, W = x.shape
        f = torch.empty((N, C), dtype=torch.float32, device=x.device)
        if f.numel():
            _call("srgan_lrelu_gap_fwd", _p(x), _p(f), N, H * W, C, slope, _stream())
        ctx.slope = slope
        ctx.save_for_backward(x)
        return f

    @staticmethod
    def backward(ctx, df):
        (x,) = ctx.saved_tensors
        N, C, H, W = x.shape
        dx = torch.empty_like(x)
        if dx.numel():
            _call("srgan_lrelu_gap_bwd", _p(df.contiguous()), _p(x), _p(dx), N, H * W, C, ctx.slope, _stream())
        return dx, None


def lrelu_gap(x, slope=0.2):
    """LeakyReLU(slope) then AdaptiveAvgPool2d(1), flattened to [N,C]; ref pyfiles/model.py:394,454."""
    _req(x)
    return _LReluGapFn.apply(x, float(slope))


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        if a.dim() == 4:
            a, b = _raw_to_nhwc(a), _raw_to_nhwc(b)
        else:
            a, b = a.contiguous(), b.contiguous()
        y = torch.empty_like(a)
        if y.numel():
            _call("srgan_add", _p(a), _p(b), _p(y), y.numel(), _stream())
        return y

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    _req(a, b)
    if a.shape != b.shape:
        raise ValueError("add: shape mismatch")
    return _AddFn.apply(a, b)


# ----------------------------------------------------------------------------- heads
class _SoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, J = x.shape
        y = torch.empty_like(x)
        if N:
            _call("srgan_softmax_fwd", _p(x), _p(y), N, J, _stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        N, J = y.shape
        dx = torch.empty_like(y)
        if N:
            _call("srgan_softmax_bwd", _p(dy.contiguous()), _p(y), _p(dx), N, J, _stream())
        return dx


def softmax_rows(x):
    """softmax over dim 1 of a [N,J] tensor; ref: nn.Softmax() (implicit dim=1) pyfiles/model.py:333-334."""
    _req(x)
    return _SoftmaxFn.apply(x)


class _CrossEntropyFn(torch.autograd.Function):
    """mean_n (logsumexp(x_n) - x_n[label_n]); ref: nn.CrossEntropyLoss() in notebook 04 cells 18 / 22."""

    @staticmethod
    def forward(ctx, x, label):
        x = x.contiguous()
        N, J = x.shape
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        rows = torch.empty((N,), dtype=torch.float32, device=x.device)
        _call("srgan_cross_entropy_fwd", _p(x), _p(label), _p(loss), _p(rows), N, J, _stream())
        ctx.save_for_backward(x, label)
        return loss

    @staticmethod
    def backward(ctx, gout):
        x, label = ctx.saved_tensors
        N, J = x.shape
        dx = torch.empty_like(x)
        _call("srgan_cross_entropy_bwd", _p(x), _p(label), _p(gout.contiguous()), _p(dx), N, J, _stream())
        return dx, None


def cross_entropy(x, label):
    """nn.CrossEntropyLoss()(x, label) (mean reduction) for [N, J] fp32 scores and int64 labels."""
    _req(x)
    if x.dim() != 2 or x.shape[0] == 0:
        raise ValueError("cross_entropy: expected a non-empty [N, J] tensor")
    if not (label.is_cuda and label.dtype == torch.int64 and label.shape == (x.shape[0],)):
        raise SrganKernelError("cross_entropy: label must be a CUDA int64 tensor of shape [N]")
    if label.device != x.device:
        raise SrganKernelError("cross_entropy: label lives on another device")
    return _CrossEntropyFn.apply(x, label.contiguous())


class CrossEntropyLoss(torch.nn.Module):
    """Drop-in for the `criterion = nn.CrossEntropyLoss()` of notebook 04 on the kernels of this package."""

    def forward(self, x, label):
        return cross_entropy(x, label)


def prdc_counts(real_features, fake_features, nearest_k):
    """The integer counts behind precision / recall / density / coverage (prdc.compute_prdc, ref
    pyfiles/evaluation.py:98-110): {"col_hits_real" [M], "row_hits_fake" [N], "row_min_in" [N]} as int32 CUDA tensors,
    plus the squared k-NN radii.  Features: CUDA fp32 [N, D] and [M, D]."""
    a, b = real_features.contiguous(), fake_features.contiguous()
    _req(a, b)
    if a.dtype != torch.float32 or b.dtype != torch.float32:
        raise SrganKernelError("prdc: features must be float32")
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1] or a.shape[1] == 0:
        raise ValueError("prdc: expected [N, D] and [M, D] features")
    N, D = a.shape
    M = b.shape[0]
    k = int(nearest_k)
    if not (0 < k < N and k < M):
        raise ValueError("prdc: need 0 < nearest_k < number of samples")
    dev, st = a.device, _stream()
    f64 = dict(dtype=torch.float64, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    radii = []
    for t in (a, b):
        n = t.shape[0]
        d2 = torch.empty((n, n), **f64)
        r = torch.empty((n,), **f64)
        _call("srgan_prdc_pairdist2", _p(t), _p(t), _p(d2), n, n, D, st)
        _call("srgan_prdc_kth_radius", _p(d2), _p(r), n, k, st)
        radii.append(r)
    d2 = torch.empty((N, M), **f64)
    _call("srgan_prdc_pairdist2", _p(a), _p(b), _p(d2), N, M, D, st)
    col = torch.empty((M,), **i32)
    row = torch.empty((N,), **i32)
    rmin = torch.empty((N,), **i32)
    _call("srgan_prdc_counts", _p(d2), _p(radii[0]), _p(radii[1]), _p(col), _p(row), _p(rmin), N, M, st)
    return {"col_hits_real": col, "row_hits_fake": row, "row_min_in": rmin, "r2_real": radii[0], "r2_fake": radii[1]}


def compute_prdc(real_features, fake_features, nearest_k):
    """prdc.compute_prdc on the GPU: dict(precision, recall, density, coverage) of Python floats.  numpy / CPU inputs
    are moved to the current CUDA device (the reference hands over numpy feature arrays)."""
    def dev(t):
        t = torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t, dtype=torch.float32)
        return t if t.is_cuda else t.cuda()
    a, b = dev(real_features), dev(fake_features)
    c = prdc_counts(a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1), nearest_k)
    col, row, rmin = (c[k_].cpu().numpy().astype(np.int64) for k_ in ("col_hits_real", "row_hits_fake", "row_min_in"))
    return dict(precision=float((col > 0).mean()), recall=float((row > 0).mean()),
                density=float((1.0 / float(nearest_k)) * col.mean()), coverage=float(rmin.mean()))


class _ReparamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, eps):
        mu, logvar, eps = mu.contiguous(), logvar.contiguous(), eps.contiguous()
        z = torch.empty_like(mu)
        if z.numel():
            _call("srgan_reparam_fwd", _p(mu), _p(logvar), _p(eps), _p(z), z.numel(), _stream())
        ctx.save_for_backward(logvar, eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        logvar, eps = ctx.saved_tensors
        dz = dz.contiguous()
        dmu = torch.empty_like(dz) if ctx.needs_input_grad[0] else None
        dlv = torch.empty_like(dz) if ctx.needs_input_grad[1] else None
        if dz.numel():
            _call("srgan_reparam_bwd", _p(dz), _p(logvar), _p(eps), _p(dmu), _p(dlv), dz.numel(), _stream())
        return dmu, dlv, None


def reparametrize(mu, logvar, eps):
    """z = eps * exp(0.5*logvar) + mu; ref pyfiles/model.py:398-402,459-463."""
    _req(mu, logvar, eps)
    return _ReparamFn.apply(mu, logvar, eps)


# ----------------------------------------------------------------------------- losses
def _same_layout(a, b):
    """Bring two same-shape tensors to one dense storage order (channels-last for 4-D)."""
    if a.dim() == 4:
        return _raw_to_nhwc(a), _raw_to_nhwc(b)
    return a.contiguous(), b.contiguous()


class _L1MeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _same_layout(a, b)
        out = torch.empty((), dtype=torch.float32, device=a.device)
        _call("srgan_l1_mean_fwd", _p(a), _p(b), a.numel(), _p(out), _p(_red_scratch(a.device)), _stream())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        _call("srgan_l1_mean_bwd", _p(a), _p(b), _p(g.contiguous()), _p(da), _p(db), a.numel(), _stream())
        return da, db


def l1_mean(a, b):
    """torch.mean(torch.abs(a - b)); ref pyfiles/util_notebook.py:295,309,348,359,625,639,676,686."""
    _req(a, b)
    if a.shape != b.shape:
        raise ValueError("l1_mean: shape mismatch %s vs %s" % (tuple(a.shape), tuple(b.shape)))
    return _L1MeanFn.apply(a, b)


class _MseConstFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target):
        x = _raw_to_nhwc(x) if x.dim() == 4 else x.contiguous()
        out = torch.empty((), dtype=torch.float32, device=x.device)
        _call("srgan_mse_const_fwd", _p(x), target, x.numel(), _p(out), _p(_red_scratch(x.device)), _stream())
        ctx.target = target
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        _call("srgan_mse_const_bwd", _p(x), ctx.target, _p(g.contiguous()), _p(dx), x.numel(), _stream())
        return dx, None


def mse_const(x, target):
    """nn.MSELoss()(x, full_like(x, target)); ref get_loss_D pyfiles/util.py:457-462."""
    _req(x)
    return _MseConstFn.apply(x, float(target))


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _same_layout(a, b)
        out = torch.empty((), dtype=torch.float32, device=a.device)
        _call("srgan_mse_fwd", _p(a), _p(b), a.numel(), _p(out), _p(_red_scratch(a.device)), _stream())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        _call("srgan_mse_bwd", _p(a), _p(b), _p(g.contiguous()), _p(da), _p(db), a.numel(), _stream())
        return da, db


def mse(a, b):
    """nn.MSELoss()(a, b); ref get_domainloss_D pyfiles/util.py:464-468."""
    _req(a, b)
    if a.shape != b.shape:
        raise ValueError("mse: shape mismatch")
    return _MseFn.apply(a, b)


# ----------------------------------------------------------------------------- latent batch losses
LAT_BKL, LAT_CORR, LAT_HIST, LAT_KL = 1, 2, 4, 8


def latent_out_floats(D, bins):
    return 4 + 4 * D + D * D + D * bins


def latent_stats_views(out, D, bins):
    """Named views into the statistics blob written by the latent-loss kernel."""
    o = 4
    v = {}
    for name, n in (("mean", D), ("var", D), ("cdiag", D), ("corr", D * D), ("hist", D * bins), ("hsum", D)):
        v[name] = out[o:o + n]
        o += n
    v["corr"] = v["corr"].view(D, D)
    v["hist"] = v["hist"].view(D, bins)
    return v


class _LatentLossFn(torch.autograd.Function):
    """[batch-KL, corr, hist, KL] of a latent batch; see srgan_latent_losses_fwd in the header.
    `mu_all` (optional, no grad) is the all-gathered global batch whose rows [row0,row0+n_local) are `mu`."""

    @staticmethod
    def forward(ctx, mu, logvar, mu_all, logvar_all, row0, n_cfg, target, bins, hmin, hmax, sigma, flags):
        mu = mu.contiguous()
        full = mu if mu_all is None else mu_all.contiguous()
        lv_full = None
        if flags & LAT_KL:
            lv_full = (logvar if logvar_all is None else logvar_all).contiguous()
        n, D = full.shape
        out = torch.empty((latent_out_floats(D, bins),), dtype=torch.float32, device=mu.device)
        _call("srgan_latent_losses_fwd", _p(full), _p(lv_full), n, D, n_cfg, _p(target), bins, hmin, hmax, sigma,
              flags, _p(out), _stream())
        ctx.cfg = (n, D, n_cfg, bins, hmin, hmax, sigma, flags, row0, mu.shape[0])
        ctx.save_for_backward(full, lv_full, target, out)
        ctx.mark_non_differentiable(out)
        return out[:4].clone(), out

    @staticmethod
    def backward(ctx, g4, _gout):
        full, lv_full, target, out = ctx.saved_tensors
        n, D, n_cfg, bins, hmin, hmax, sigma, flags, row0, rows = ctx.cfg
        dmu = torch.empty((rows, D), dtype=torch.float32, device=full.device)
        dlv = torch.empty((rows, D), dtype=torch.float32, device=full.device) \
            if (flags & LAT_KL) and ctx.needs_input_grad[1] else None
        _call("srgan_latent_losses_bwd", _p(full), _p(lv_full), n, D, n_cfg, _p(target), bins, hmin, hmax, sigma,
              flags, _p(out), _p(g4.contiguous()), _p(dmu), _p(dlv), row0, rows, _stream())
        return (dmu,) + (dlv,) + (None,) * 10


def latent_losses(mu, logvar=None, n_cfg=2.0, target=None, bins=50, hmin=-10.0, hmax=10.0, sigma=0.2,
                  flags=LAT_BKL | LAT_CORR | LAT_HIST, mu_all=None, logvar_all=None, row0=0):
    """Returns (losses[4] = [batch-KL, corr, hist, KL], stats blob)."""
    _req(mu, logvar, target, mu_all, logvar_all)
    if (flags & LAT_HIST) and target is None:
        raise ValueError("latent_losses: the histogram term needs a target")
    if not (flags & LAT_HIST):
        bins = 1
    return _LatentLossFn.apply(mu, logvar, mu_all, logvar_all, int(row0), float(n_cfg), target, int(bins),
                               float(hmin), float(hmax), float(sigma), int(flags))


class _CorrcoefFn(torch.autograd.Function):
    """corrcoef of the ROWS of x [D, n] (np.corrcoef convention); ref pyfiles/util.py:470-511."""

    @staticmethod
    def forward(ctx, x):
        mu = x.t().contiguous()            # kernel layout: [n samples][D]
        n, D = mu.shape
        out = torch.empty((latent_out_floats(D, 1),), dtype=torch.float32, device=x.device)
        _call("srgan_latent_losses_fwd", _p(mu), None, n, D, 2.0, None, 1, 0.0, 1.0, 1.0, LAT_CORR, _p(out),
              _stream())
        ctx.save_for_backward(mu, out)
        return latent_stats_views(out, D, 1)["corr"].clone()

    @staticmethod
    def backward(ctx, dc):
        mu, out = ctx.saved_tensors
        n, D = mu.shape
        dmu = torch.empty_like(mu)
        _call("srgan_corrcoef_bwd", _p(mu), n, D, _p(out), _p(dc.contiguous()), _p(dmu), _stream())
        return dmu.t()


def corrcoef(x):
    _req(x)
    if x.dim() != 2:
        raise ValueError("corrcoef expects a 2D tensor")
    return _CorrcoefFn.apply(x)


class _SoftHistFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bins, hmin, hmax, sigma):
        x = x.contiguous()
        h = torch.empty((bins,), dtype=torch.float32, device=x.device)
        _call("srgan_softhist_fwd", _p(x), x.numel(), bins, hmin, hmax, sigma, _p(h), _stream())
        ctx.cfg = (bins, hmin, hmax, sigma)
        ctx.save_for_backward(x)
        return h

    @staticmethod
    def backward(ctx, dh):
        (x,) = ctx.saved_tensors
        bins, hmin, hmax, sigma = ctx.cfg
        dx = torch.empty_like(x)
        _call("srgan_softhist_bwd", _p(x), _p(dh.contiguous()), x.numel(), bins, hmin, hmax, sigma, _p(dx),
              _stream())
        return dx, None, None, None, None


def soft_histogram(x, bins, hmin, hmax, sigma):
    """Gaussian-kernel soft histogram of a vector; ref GaussianHistogram.forward pyfiles/util.py:532-537."""
    _req(x)
    if x.dim() != 1:
        raise ValueError("soft_histogram expects a 1D tensor")
    return _SoftHistFn.apply(x, int(bins), float(hmin), float(hmax), float(sigma))


# ----------------------------------------------------------------------------- optimizer
class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (ref: optim.Adam(..., betas=(0.5, 0.999)) pyfiles/util_notebook.py:117-131,
    500-507) with one kernel launch per parameter group: parameters, gradients and both moments of a
    group live in flat buffers; `p.data` / `p.grad` are re-pointed at views of them."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._flat = {}

    def _flatten(self, gi, group):
        ps = [p for p in group["params"] if p.requires_grad]
        if not ps:
            return None
        _req(*ps)
        # every parameter starts on a 256-byte boundary: TMA tensor maps need 16-byte aligned bases and the
        # glue kernels use float4 loads; the padding is zero and stays zero under Adam
        n = sum(-(-p.numel() // 64) * 64 for p in ps)
        dev = ps[0].device
        fp = torch.zeros(n, dtype=torch.float32, device=dev)
        fg = torch.zeros(n, dtype=torch.float32, device=dev)
        views = []
        o = 0
        for p in ps:
            k = p.numel()
            cl = p.dim() == 4 and p.is_contiguous(memory_format=CL) and not p.is_contiguous()
            shape = tuple(p.shape)

            def view(flat, o=o, k=k, cl=cl, shape=shape):
                if cl:
                    K, C, R, S = shape
                    return flat[o:o + k].view(K, R, S, C).permute(0, 3, 1, 2)
                return flat[o:o + k].view(shape)
            if not (cl or p.is_contiguous()):
                raise SrganKernelError("FusedAdam: parameters must be dense")
            with torch.no_grad():
                view(fp).copy_(p.data)
                if p.grad is not None:
                    view(fg).copy_(p.grad)
            p.data = view(fp)
            p.grad = view(fg)
            views.append((p, view))
            o += -(-k // 64) * 64
        st = dict(p=fp, g=fg, m=torch.zeros_like(fp), v=torch.zeros_like(fp), step=0, views=views, params=ps,
                  p16=None)
        o = 0
        for p, view in views:
            p._srgan_owner = (st, view, o, p.numel())          # see _shadow16
            o += -(-p.numel() // 64) * 64
        self._flat[gi] = st
        return st

    @staticmethod
    def _refresh_shadow(st):
        """Keep the bf16 shadow of the flat parameter buffer (created on first use by a bf16 convolution) in step
        with the weights the Adam kernel just wrote."""
        if st["p16"] is not None:
            cast_bf16(st["p"], st["p16"])

    def flat_grads(self):
        """Flat gradient buffers (one per group) -- what a data-parallel caller all-reduces."""
        out = []
        for gi, group in enumerate(self.param_groups):
            st = self._flat.get(gi) or self._flatten(gi, group)
            if st is not None:
                out.append(st["g"])
        return out

    def flat_state(self):
        """Adam moments and step counters per parameter group (CPU tensors), for checkpoints.  The weights are the
        modules' state_dict; the moments are stored in the optimizer's flat layout, which is a pure function of the
        parameter list."""
        out = {}
        for gi, group in enumerate(self.param_groups):
            st = self._flat.get(gi) or self._flatten(gi, group)
            if st is not None:
                out[gi] = {"m": st["m"].detach().cpu(), "v": st["v"].detach().cpu(), "step": int(st["step"]),
                           "lr": float(group["lr"])}
        return out

    def load_flat_state(self, state):
        for gi, group in enumerate(self.param_groups):
            st = self._flat.get(gi) or self._flatten(gi, group)
            if st is None:
                continue
            src = state[gi] if gi in state else state[str(gi)]
            if src["m"].numel() != st["m"].numel():
                raise ValueError("FusedAdam.load_flat_state: parameter layout differs from the checkpoint")
            st["m"].copy_(src["m"])
            st["v"].copy_(src["v"])
            st["step"] = int(src["step"])
            group["lr"] = float(src["lr"])

    def owned_ids(self):
        """ids of the parameters whose data / grad live in this optimizer's flat buffers."""
        out = set()
        for gi, group in enumerate(self.param_groups):
            st = self._flat.get(gi) or self._flatten(gi, group)
            if st is not None:
                out.update(id(p) for p in st["params"])
        return out

    def zero_grad(self, set_to_none=False):
        for gi, group in enumerate(self.param_groups):
            st = self._flat.get(gi) or self._flatten(gi, group)
            if st is None:
                continue
            st["g"].zero_()
            for p, view in st["views"]:
                p.grad = view(st["g"])
                p._srgan_arrivals = 0          # see direct_param_grads / _grad_sink

    @torch.no_grad()
    def step(self, closure=None):
        for gi, group in enumerate(self.param_groups):
            st = self._flat.get(gi) or self._flatten(gi, group)
            if st is None:
                continue
            # gradients written by autograd into fresh tensors (p.grad re-assigned) are folded back
            for p, view in st["views"]:
                gv = view(st["g"])
                if p.grad is None:
                    gv.zero_()
                elif p.grad.data_ptr() != gv.data_ptr():
                    gv.copy_(p.grad)
                    p.grad = gv
            if _tape is not None and _tape.recording:
                # captured into a CUDA graph: the step-dependent scalars come from device memory, refreshed (and the
                # step counter advanced) by FeedTape.upload() before every replay
                def fill(buf, st=st, group=group):
                    st["step"] += 1
                    b1, b2 = group["betas"]
                    # same roundings as srgan_adam_step: float betas widened to double for pow, results cast back
                    b1d, b2d = float(np.float32(b1)), float(np.float32(b2))
                    buf[0], buf[1], buf[2], buf[3] = group["lr"], b1, b2, group["eps"]
                    buf[4] = float(np.float32(1.0) - np.float32(b1d ** st["step"]))
                    buf[5] = float(np.float32(math.sqrt(1.0 - b2d ** st["step"])))
                hyper = _tape.add(_tape.staging(8), _tape.device_buffer(8), fill)
                _call("srgan_adam_step_dev", _p(st["p"]), _p(st["g"]), _p(st["m"]), _p(st["v"]), st["p"].numel(),
                      _p(hyper), _stream())
                self._refresh_shadow(st)
                continue
            st["step"] += 1
            b1, b2 = group["betas"]
            _call("srgan_adam_step", _p(st["p"]), _p(st["g"]), _p(st["m"]), _p(st["v"]), st["p"].numel(),
                  group["lr"], b1, b2, group["eps"], st["step"], _stream())
            self._refresh_shadow(st)
