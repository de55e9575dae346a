"""What a plain device copy reaches at the sizes of the norm planes, under the same timing protocol as
tools/conv_bench.py (L2 flushed with a 256 MB memset before every launch) and without the flush.
Gives the practical ceiling for the 1-read + 1-write norm kernels."""
import torch

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps, do_flush):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for mb in (12.8, 29.5, 67.1, 134.2, 268.4, 1073.7):
    n = int(mb * 1e6 / 4)
    x = torch.randn(n, device=dev)
    y = torch.empty_like(x)
    for fl in (True, False):
        us = timed(lambda: y.copy_(x), 7, fl)
        print("copy %8.1f MB  flush=%d  %8.1f us  %6.0f GB/s (read+write)" % (mb, fl, us, 2 * mb / us * 1e3))
    us = timed(lambda: x.sum(), 7, True)
    print("sum  %8.1f MB  flush=1  %8.1f us  %6.0f GB/s (read)" % (mb, us, mb / us * 1e3))
