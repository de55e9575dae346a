"""The six thin16 launches of the generator's RGB stem / head at batch 64, once each after a warm-up (for ncu):
ncu --set full --import-source on -k regex:'conv_umma_kernel|wgrad_umma|thinout2' python tools/thin16_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "style-restricted_gan_b200", "pyfiles"))
import srgan_ops as ops  # noqa: E402

DEV, CL, BF = "cuda", torch.channels_last, torch.bfloat16
N, H = int(os.environ.get("PROBE_BATCH", "64")), 128
lib = ops._lib()
dev = torch.device(DEV, torch.cuda.current_device())
img = (torch.rand(N, 3, H, H, device=DEV) * 2 - 1).contiguous(memory_format=CL)
fat16 = torch.randn(N, 64, H, H, device=DEV).to(BF).contiguous(memory_format=CL)
ws_ = torch.randn(64, 3, 7, 7, device=DEV).contiguous(memory_format=CL)
wh_ = torch.randn(3, 64, 7, 7, device=DEV).contiguous(memory_format=CL)
ds, dh = ops._desc(N, H, H, 3, 64, 7, 7, 1, 3), ops._desc(N, H, H, 64, 3, 7, 7, 1, 3)
out16, thin = torch.empty_like(fat16), torch.empty_like(img)
dws, dwh = torch.empty_like(ws_), torch.empty_like(wh_)


def ws(d, p):
    nb = lib.srgan_conv2d_thin16_workspace(d, p)
    return ops._workspace(dev, nb), nb


calls = {
    "stem_fprop": lambda: ops._call("srgan_conv2d_fprop_thin16", ds, ops._p(img), ops._p(ws_), None, ops._p(out16), 0, 0.0, *map_ws(ds, 0)),
    "stem_dgrad": lambda: ops._call("srgan_conv2d_dgrad_thin16", ds, ops._p(fat16), ops._p(ws_), ops._p(thin), *map_ws(ds, 1)),
    "stem_wgrad": lambda: ops._call("srgan_conv2d_wgrad_thin16", ds, ops._p(img), ops._p(fat16), ops._p(dws), None, *map_ws(ds, 2)),
    "head_fprop": lambda: ops._call("srgan_conv2d_fprop_thin16", dh, ops._p(fat16), ops._p(wh_), None, ops._p(thin), 0, 0.0, *map_ws(dh, 0)),
    "head_dgrad": lambda: ops._call("srgan_conv2d_dgrad_thin16", dh, ops._p(img), ops._p(wh_), ops._p(out16), *map_ws(dh, 1)),
    "head_wgrad": lambda: ops._call("srgan_conv2d_wgrad_thin16", dh, ops._p(fat16), ops._p(img), ops._p(dwh), None, *map_ws(dh, 2)),
}


def map_ws(d, p):
    w, nb = ws(d, p)
    return ops._p(w), nb, ops._stream()


only = os.environ.get("PROBE_ONLY", "")
for name, fn in calls.items():
    if only and name not in only.split(","):
        continue
    fn()
    torch.cuda.synchronize()
    print("ran", name)
