"""One-pass (cluster) instance norm against the two-kernel path of norm8.cu on the residual-trunk plane
(256 ch @32x32 bf16) over a range of batch sizes: time per launch (CUDA events, L2 flushed, and back to back in a
loop of 20 without flush), clusters the device holds at once.  Separates the fixed cost of a cluster launch from the
per-image cost and shows whether the kernel is limited by resident clusters (steps in N) or by bandwidth."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "style-restricted_gan_b200", "pyfiles"))
import srgan_ops as ops  # noqa: E402

DEV = "cuda"
CL = torch.channels_last
BF = torch.bfloat16
lib = ops._lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, reps=15, cold=True):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def loop(fn, n=20):
    fn(); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) * 1e3 / n


def graph_loop(fn, n=20):
    """n launches captured in one CUDA graph (no host cost between them); the plane stays in L2 when it fits"""
    fn(); torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn()
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                fn()
    g.replay(); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) * 1e3 / n


def main():
    planes = [(256, 32, 32)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for (C, H, W) in planes:
        lib.srgan_inorm_onepass_enable(1)
        print("plane %d ch @%dx%d bf16: max resident clusters %d" %
              (C, H, W, lib.srgan_inorm_onepass_max_clusters(H * W, C, ops.DT_BF16)))
        batches = [int(v) for v in os.environ.get("PROBE_N", "1,4,8,16,18,24,32,48,64,96,128").split(",")]
        for N in batches:
            x = torch.randn(N, C, H, W, device=DEV).to(BF).contiguous(memory_format=CL)
            dy = torch.randn(N, C, H, W, device=DEV).to(BF).contiguous(memory_format=CL)
            g = torch.randn(C, device=DEV)
            b = torch.randn(C, device=DEV)
            cb = torch.randn(N, C, device=DEV)
            y, dx = torch.empty_like(x), torch.empty_like(x)
            mean, rstd = torch.empty((N, C), device=DEV), torch.empty((N, C), device=DEV)
            s1, s2 = torch.empty((N, C), device=DEV), torch.empty((N, C), device=DEV)
            nb = lib.srgan_inorm_mixed_workspace(N, H * W, C)
            ws = ops._workspace(x.device, nb)
            ctr = ops._norm_counters(x.device, N, C)

            def fwd():
                ops._call("srgan_inorm_fwd_mixed", ops._p(x), ops._dt(x), ops._p(y), ops._dt(y), ops._p(mean),
                          ops._p(rstd), ops._p(g), ops._p(b), ops._p(cb), None, N, H * W, C, 1e-5, ops.ACT_RELU, 0.0, 0,
                          ops._p(ws), nb, ops._p(ctr), ops._stream())

            def bwd():
                ops._call("srgan_inorm_bwd_mixed", ops._p(dy), ops._dt(dy), ops._p(x), ops._dt(x), ops._p(mean),
                          ops._p(rstd), ops._p(g), ops._p(b), ops._p(cb), ops._p(dx), ops._p(s1), ops._p(s2), N, H * W, C,
                          ops.ACT_RELU, 0.0, ops._p(ws), nb, ops._p(ctr), ops._stream())
            row = []
            for on in (1, 0):
                lib.srgan_inorm_onepass_enable(on)
                row.append((timed(fwd), graph_loop(fwd), timed(bwd), graph_loop(bwd)))
            lib.srgan_inorm_onepass_enable(0)
            print("N %4d  one pass: fwd %6.1f us cold %6.1f graph | bwd %6.1f cold %6.1f graph   two kernels: fwd %6.1f "
                  "cold %6.1f graph | bwd %6.1f cold %6.1f graph" % ((N,) + row[0] + row[1]))


if __name__ == "__main__":
    main()
