"""Per-layer timing of every convolution geometry of the SRGAN step (fprop / dgrad / wgrad) and of the
fused norm kernels, through the C ABI, CUDA events, L2 flushed between launches.

Usage: python tools/conv_bench.py [--batch 64] [--reps 5] [--only G.res]
Prints one row per (layer, pass): GFLOP, microseconds, TFLOP/s, and the weight of that launch in one
SRGAN nb03 step (launch count x time), so the table ranks what to optimise."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402,F401  (sets sys.path)
import torch  # noqa: E402

# name, H(=W) in, C, K, R, stride, pad, transposed, (fprop, dgrad, wgrad) launches per SRGAN k=5 step
# G: 8 fwd, 5 bwd traversals; D: 11 fwd, 11 bwd; E: 5 fwd, 6 bwd  
LAYERS = [
    ("G.down0 3>64 k7",        128, 3,   64,  7, 1, 3, False, (8, 2, 5)),
    ("G.down1 64>128 k4s2",    128, 64,  128, 4, 2, 1, False, (8, 5, 5)),
    ("G.down2 128>256 k4s2",   64,  128, 256, 4, 2, 1, False, (8, 5, 5)),
    ("G.res 256>256 k3",       32,  256, 256, 3, 1, 1, False, (96, 60, 60)),
    ("G.up0 T256>128 k4s2",    32,  256, 128, 4, 2, 1, True,  (8, 5, 5)),
    ("G.up1 T128>64 k4s2",     64,  128, 64,  4, 2, 1, True,  (8, 5, 5)),
    ("G.up2 64>3 k7",          128, 64,  3,   7, 1, 3, False, (8, 5, 5)),
    ("D1.0 3>64 k4s2",         128, 3,   64,  4, 2, 1, False, (11, 1, 11)),
    ("D1.1 64>128 k4s2",       64,  64,  128, 4, 2, 1, False, (11, 11, 11)),
    ("D1.2 128>256 k4s2",      32,  128, 256, 4, 2, 1, False, (11, 11, 11)),
    ("D1.3 256>512 k4s2",      16,  256, 512, 4, 2, 1, False, (11, 11, 11)),
    ("D1.patch 512>1 k4p1",    8,   512, 1,   4, 1, 1, False, (11, 11, 11)),
    ("D1.class 512>4 k8",      8,   512, 4,   8, 1, 0, False, (11, 11, 11)),
    ("D2.0 3>32 k4s2",         64,  3,   32,  4, 2, 1, False, (11, 1, 11)),
    ("D2.1 32>64 k4s2",        32,  32,  64,  4, 2, 1, False, (11, 11, 11)),
    ("D2.2 64>128 k4s2",       16,  64,  128, 4, 2, 1, False, (11, 11, 11)),
    ("D2.3 128>256 k4s2",      8,   128, 256, 4, 2, 1, False, (11, 11, 11)),
    ("D2.patch 256>1 k4p1",    4,   256, 1,   4, 1, 1, False, (11, 11, 11)),
    ("D2.class 256>4 k4",      4,   256, 4,   4, 1, 0, False, (11, 11, 11)),
    ("E.first 3>64 k7s2p1",    128, 3,   64,  7, 2, 1, False, (5, 3, 6)),
    ("E.l0.conv1 64 k3 @64",   64,  64,  64,  3, 1, 0, False, (5, 6, 6)),      # reflect-padded input 62+2
    ("E.l0.cmp 64>128 k3",     64,  64,  128, 3, 1, 0, False, (5, 6, 6)),
    ("E.l0.short 64>128 k1",   31,  64,  128, 1, 1, 0, False, (5, 6, 6)),
    ("E.l1.conv1 128 k3 @33",  33,  128, 128, 3, 1, 0, False, (5, 6, 6)),
    ("E.l1.cmp 128>256 k3",    33,  128, 256, 3, 1, 0, False, (5, 6, 6)),
    ("E.l1.short 128>256 k1",  15,  128, 256, 1, 1, 0, False, (5, 6, 6)),
    ("E.l2.conv1 256 k3 @17",  17,  256, 256, 3, 1, 0, False, (5, 6, 6)),
    ("E.l2.cmp 256>512 k3",    17,  256, 512, 3, 1, 0, False, (5, 6, 6)),
    ("E.l2.short 256>512 k1",  7,   256, 512, 1, 1, 0, False, (5, 6, 6)),
    ("E.l3.conv1 512 k3 @9",   9,   512, 512, 3, 1, 0, False, (5, 6, 6)),
    ("E.l3.cmp 512>1024 k3",   9,   512, 1024, 3, 1, 0, False, (5, 6, 6)),
    ("E.l3.short 512>1024 k1", 3,   512, 1024, 1, 1, 0, False, (5, 6, 6)),
]

NORMS = [   # name, C, H, launches fwd, bwd per step
    ("IN 64@128",  64,  128, 16, 10),
    ("IN 128@64",  128, 64,  16, 10),
    ("IN 256@32",  256, 32,  104, 65),
    ("IN 64@62",   64,  62,  10, 12),
    ("IN 128@31",  128, 31,  10, 12),
    ("IN 256@15",  256, 15,  10, 12),
    ("IN 512@7",   512, 7,   10, 12),
]


def timed(fn, reps, flush):
    fn()
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--engine", default="auto")
    a = ap.parse_args()
    import srgan_ops as ops
    ops.set_conv_engine(a.engine)
    dev = "cuda:0"
    B = a.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    CL = torch.channels_last
    total = 0.0
    rows = []
    print("%-26s %-6s %8s %9s %8s %5s %9s  %s" % ("layer", "pass", "GFLOP", "us", "TF/s", "n", "ms/step", "engine"))
    for name, H, C, K, R, stride, pad, transposed, counts in LAYERS:
        if a.only and a.only not in name:
            continue
        if transposed:
            # ConvTranspose2d Cin=C -> Cout=K on HxH input: mirrored conv has C'=K, K'=C on (2H)x(2H)
            Ho = (H - 1) * stride - 2 * pad + R
            d = ops._desc(B, Ho, Ho, K, C, R, R, stride, pad)
            x = torch.randn(B, K, Ho, Ho, device=dev).contiguous(memory_format=CL)     # mirrored conv input
            dy = torch.randn(B, C, H, H, device=dev).contiguous(memory_format=CL)
            w = (torch.randn(C, K, R, R, device=dev) * 0.05).contiguous(memory_format=CL)
            passes = (("fwd=dgrad", lambda: ops._dgrad(d, dy, w, dy), counts[0]),
                      ("bwd=fprop", lambda: ops._fprop(d, x, w, None, 0, 0.0), counts[1]),
                      ("wgrad", lambda: ops._wgrad(d, x, dy, True, False), counts[2]))
            engines = [ops._lib().srgan_conv2d_engine(d, p) for p in (1, 0, 2)]
        else:
            d = ops._desc(B, H, H, C, K, R, R, stride, pad)
            x = torch.randn(B, C, H, H, device=dev).contiguous(memory_format=CL)
            dy = torch.randn(B, K, d.P, d.Q, device=dev).contiguous(memory_format=CL)
            w = (torch.randn(K, C, R, R, device=dev) * 0.05).contiguous(memory_format=CL)
            passes = (("fprop", lambda: ops._fprop(d, x, w, None, 0, 0.0), counts[0]),
                      ("dgrad", lambda: ops._dgrad(d, dy, w, x), counts[1]),
                      ("wgrad", lambda: ops._wgrad(d, x, dy, True, False), counts[2]))
            engines = [ops._lib().srgan_conv2d_engine(d, p) for p in (0, 1, 2)]
        gf = 2.0 * B * d.P * d.Q * K * C * R * R / 1e9 if not transposed else 2.0 * B * d.P * d.Q * d.K * d.C * R * R / 1e9
        for (pname, fn, n), eng in zip(passes, engines):
            us = timed(fn, a.reps, flush)
            ms_step = n * us / 1e3
            total += ms_step
            rows.append((ms_step, name, pname))
            print("%-26s %-9s %8.2f %9.1f %8.1f %5d %9.2f  %s" % (name, pname, gf, us, gf / us * 1e3, n, ms_step,
                                                              {1: "ffma", 2: "umma"}.get(eng, str(eng))))
        del x, dy, w
    print("conv total per step (isolated, L2-cold): %.1f ms" % total)
    ntotal = 0.0
    print("%-14s %-5s %9s %9s %6s %9s" % ("norm", "pass", "MB", "us", "GB/s", "ms/step"))
    for name, C, H, nf, nb in NORMS:
        if a.only and a.only not in name:
            continue
        x = torch.randn(B, C, H, H, device=dev).contiguous(memory_format=CL).requires_grad_(True)
        g = torch.ones(C, device=dev)
        b = torch.zeros(C, device=dev)
        cb = torch.randn(B, C, device=dev)
        y = ops.instance_norm_act(x, g, b, cb, None, 1e-5, ops.ACT_RELU, 0.0)
        dy = torch.randn_like(y)
        mb = x.numel() * 4 / 1e6
        us = timed(lambda: ops.instance_norm_act(x.detach(), g, b, cb, None, 1e-5, ops.ACT_RELU, 0.0), a.reps, flush)
        print("%-14s %-5s %9.1f %9.1f %6.0f %9.2f" % (name, "fwd", 2 * mb, us, 2 * mb / us * 1e3, nf * us / 1e3))
        ntotal += nf * us / 1e3
        us = timed(lambda: torch.autograd.grad(y, x, dy, retain_graph=True), a.reps, flush)
        print("%-14s %-5s %9.1f %9.1f %6.0f %9.2f" % (name, "bwd", 3 * mb, us, 3 * mb / us * 1e3, nb * us / 1e3))
        ntotal += nb * us / 1e3
    print("norm total per step (isolated): %.1f ms" % ntotal)
    rows.sort(reverse=True)
    print("top:", [(round(r[0], 2), r[1], r[2]) for r in rows[:12]])


if __name__ == "__main__":
    main()
