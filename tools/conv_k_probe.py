"""Kernel time of the bf16 trunk convolution as a function of the reduction length (filter size 1 / 3 / 5 at the same
256 -> 256 channels, 32x32, batch 64): separates the per-tile fixed cost (epilogue, pipeline fill) from the per-stage
cost.  Usage: python tools/conv_k_probe.py   (SRGAN_CONV_PAIRS=0|1 selects the kernel)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402,F401
import torch  # noqa: E402


def main():
    import srgan_ops as ops
    dev, CL = "cuda:0", torch.channels_last
    N, C, K, H = 64, 256, 256, 32
    x = torch.randn(N, C, H, H, device=dev).to(torch.bfloat16).contiguous(memory_format=CL)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = []
    for R in (1, 3, 5):
        w = (torch.randn(K, C, R, R, device=dev) * 0.02).to(torch.bfloat16).contiguous(memory_format=CL)
        d = ops._desc(N, H, H, C, K, R, R, 1, R // 2)
        y = torch.empty((N, K, H, H), dtype=torch.bfloat16, device=dev).contiguous(memory_format=CL)
        fn = lambda: ops._call("srgan_conv2d_fprop_bf16", d, ops._p(x), ops._p(w), None, ops._p(y), 0, 0.0, None,
                               ops._stream())
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        # back-to-back launches (no flush): steady state, launch overhead amortised
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / 20 * 1e3
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        stages = R * R * C // 64
        fl = 2.0 * N * H * H * K * C * R * R
        res.append((R, stages, warm, ts[len(ts) // 2]))
        print("R=%d stages/tile=%3d  back-to-back %.1f us (%.0f TF/s)   cold %.1f us" %
              (R, stages, warm, fl / warm / 1e6, ts[len(ts) // 2]))
    (r1, s1, w1, _), (r3, s3, w3, _), (r5, s5, w5, _) = res
    per_stage = (w5 - w3) / (s5 - s3)
    print("per stage: %.3f us per launch = %.0f clk per stage per CTA (3.46 tiles per CTA, 1.9 GHz); fixed: %.1f us" %
          (per_stage, per_stage / 3.46 * 1900, w3 - per_stage * s3))


if __name__ == "__main__":
    main()
