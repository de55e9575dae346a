"""Bring-up probe for the tcgen05 conv engine: each case runs in its own process (a trap kills the context),
compares the tf32 engine with torch's fp64 conv and prints error statistics.  Usage: python tools/umma_probe.py"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "style-restricted_gan_b200", "pyfiles"))

# name, kind, N, C, H, W, K, R, stride, pad
CASES = [
    ("gemm_1x1_c32_k16", "fprop", 1, 32, 8, 16, 16, 1, 1, 0),
    ("gemm_1x1_c64_k16", "fprop", 1, 64, 8, 16, 16, 1, 1, 0),
    ("gemm_1x1_c32_k256", "fprop", 1, 32, 8, 16, 256, 1, 1, 0),
    ("gemm_1x1_c256_k64_m4096", "fprop", 4, 256, 32, 32, 64, 1, 1, 0),
    ("res3x3", "fprop", 2, 256, 32, 32, 256, 3, 1, 1),
    ("res3x3_dgrad", "dgrad", 2, 256, 32, 32, 256, 3, 1, 1),
    ("down4x4s2", "fprop", 2, 64, 128, 128, 128, 4, 2, 1),
    ("down4x4s2_dgrad", "dgrad", 2, 64, 128, 128, 128, 4, 2, 1),
    ("out7x7_k3", "fprop", 2, 64, 128, 128, 3, 7, 1, 3),
    ("stem7x7_dgrad_c3", "dgrad", 2, 3, 128, 128, 64, 7, 1, 3),
    ("enc_valid3x3_62", "fprop", 2, 64, 64, 64, 128, 3, 1, 0),
    ("enc_valid3x3_62_dgrad", "dgrad", 2, 64, 64, 64, 128, 3, 1, 0),
    ("enc_tail_7", "fprop", 3, 512, 9, 9, 1024, 3, 1, 0),
    ("d_c3_s2", "fprop", 2, 256, 16, 16, 512, 4, 2, 1),
    ("d_c3_s2_dgrad", "dgrad", 2, 256, 16, 16, 512, 4, 2, 1),
    ("shortcut1x1_31", "fprop", 2, 64, 31, 31, 128, 1, 1, 0),
    ("convT", "convT", 2, 256, 32, 32, 128, 4, 2, 1),
    ("wgrad_1x1_c32_k128", "wgrad", 1, 32, 4, 8, 128, 1, 1, 0),
    ("wgrad_1x1_c64_k256", "wgrad", 2, 64, 16, 16, 256, 1, 1, 0),
    ("wgrad_res3x3", "wgrad", 4, 256, 32, 32, 256, 3, 1, 1),
    ("wgrad_down4x4s2", "wgrad", 2, 64, 128, 128, 128, 4, 2, 1),
    ("wgrad_out7x7_k3", "wgrad", 2, 64, 128, 128, 3, 7, 1, 3),
    ("wgrad_enc_valid_62", "wgrad", 2, 64, 64, 64, 128, 3, 1, 0),
    ("wgrad_enc_tail7", "wgrad", 3, 512, 9, 9, 1024, 3, 1, 0),
    ("wgrad_d_c3", "wgrad", 2, 256, 16, 16, 512, 4, 2, 1),
    ("wgrad_dclass", "wgrad", 2, 512, 8, 8, 4, 8, 1, 0),
]


def run_case(i):
    import torch
    import torch.nn.functional as F
    import srgan_ops as ops
    name, kind, N, C, H, W, K, R, stride, pad = CASES[i]
    g = torch.Generator().manual_seed(i)
    CL = torch.channels_last
    ops.set_conv_engine("auto")
    if kind == "convT":
        x = torch.randn(N, C, H, W, generator=g).cuda().contiguous(memory_format=CL)
        w = (torch.randn(C, K, R, R, generator=g) * (C * 4) ** -0.5).cuda().contiguous(memory_format=CL)
        y = ops.conv_transpose2d(x, w, stride, pad)
        ref = F.conv_transpose2d(x.double(), w.double(), None, stride, pad)
    elif kind == "fprop":
        x = torch.randn(N, C, H, W, generator=g).cuda().contiguous(memory_format=CL)
        w = (torch.randn(K, C, R, R, generator=g) * (C * R * R) ** -0.5).cuda().contiguous(memory_format=CL)
        b = torch.randn(K, generator=g).cuda()
        y = ops.conv2d(x, w, b, stride, pad)
        ref = F.conv2d(x.double(), w.double(), b.double(), stride, pad)
    elif kind == "wgrad":
        x = torch.randn(N, C, H, W, generator=g).cuda().contiguous(memory_format=CL)
        wr = (torch.randn(K, C, R, R, generator=g) * (C * R * R) ** -0.5).cuda().double().requires_grad_(True)
        yr = F.conv2d(x.double(), wr, None, stride, pad)
        gy = torch.randn(yr.shape, generator=g).cuda().contiguous(memory_format=CL)
        d = ops._desc(N, H, W, C, K, R, R, stride, pad)
        y, _ = ops._wgrad(d, x, gy, True, False)
        yr.backward(gy.double())
        ref = wr.grad
    else:
        x = torch.randn(N, C, H, W, generator=g).cuda().contiguous(memory_format=CL).requires_grad_(True)
        w = (torch.randn(K, C, R, R, generator=g) * (C * R * R) ** -0.5).cuda().contiguous(memory_format=CL)
        xr = x.detach().double().requires_grad_(True)
        yr = F.conv2d(xr, w.double(), None, stride, pad)
        gy = torch.randn(yr.shape, generator=g).cuda().contiguous(memory_format=CL)
        d = ops._desc(N, H, W, C, K, R, R, stride, pad)
        y = ops._dgrad(d, gy, w, x)
        yr.backward(gy.double())
        ref = xr.grad
    torch.cuda.synchronize()
    err = (y.double() - ref).abs()
    rel = float((y.double() - ref).norm() / ref.norm())
    dd = ops._desc(N, H, W, C, K, R, R, stride, pad) if kind != "convT" else None
    eng = ops._lib().srgan_conv2d_engine(dd, {"dgrad": 1, "wgrad": 2}.get(kind, 0)) if dd is not None else -9
    print("%-26s eng %d rel-L2 %.3e  max-abs %.3e  ref-rms %.3e  nan %d" % (
        name, eng, rel, float(err.max()), float(ref.pow(2).mean().sqrt()), int(torch.isnan(y).sum())), flush=True)
    if rel > 5e-3:
        yf, rf = y.detach().permute(0, 2, 3, 1).reshape(-1, y.shape[1]), ref.permute(0, 2, 3, 1).reshape(-1, y.shape[1])
        print("   got[0,:8]", [round(float(v), 4) for v in yf[0, :8]])
        print("   ref[0,:8]", [round(float(v), 4) for v in rf[0, :8]])
        rowerr = (yf.double() - rf).norm(dim=1) / rf.norm(dim=1).clamp_min(1e-9)
        colerr = (yf.double() - rf).norm(dim=0) / rf.norm(dim=0).clamp_min(1e-9)
        print("   bad rows: %d / %d (first %s)  bad cols: %d / %d (first %s)" % (
            int((rowerr > 1e-2).sum()), rowerr.numel(), (rowerr > 1e-2).nonzero()[:6].flatten().tolist(),
            int((colerr > 1e-2).sum()), colerr.numel(), (colerr > 1e-2).nonzero()[:6].flatten().tolist()))
    return rel


def sweep_wgrad():
    """Descriptor / TMA-swizzle variants for the MN-major wgrad operands on two small cases."""
    variants = [dict(), dict(SBO="1024"), dict(LAYOUT="2", TMASWZ="3", SBO="1024"), dict(LAYOUT="2", TMASWZ="3"),
                dict(LAYOUT="1", TMASWZ="3"), dict(LBO="512", SBO="4096"), dict(LBO="1024", SBO="4096"),
                dict(TMASWZ="5"), dict(TMASWZ="6", LAYOUT="1")]
    idx = [i for i, c in enumerate(CASES) if c[0] in ("wgrad_1x1_c32_k128", "wgrad_1x1_c64_k256", "wgrad_res3x3")]
    for v in variants:
        env = dict(os.environ)
        for k, val in v.items():
            env["SRGAN_DBG_WGRAD_" + k] = val
        for i in idx:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True,
                                   text=True, timeout=120, env=env)
                line = (r.stdout.strip().splitlines() or ["(no output) rc=%d %s" % (r.returncode, r.stderr.strip()[-200:])])[0]
            except subprocess.TimeoutExpired:
                line = "TIMEOUT"
            print(str(v), "|", line, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        sweep_wgrad()
    elif len(sys.argv) > 1:
        run_case(int(sys.argv[1]))
    else:
        first = int(os.environ.get("PROBE_FROM", "0"))
        for i in range(first, len(CASES)):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True,
                                   text=True, timeout=180)
                out = r.stdout.strip() or ("(no output) rc=%d %s" % (r.returncode, r.stderr.strip()[-400:]))
                if r.returncode != 0 and r.stdout.strip():
                    out += "\n   rc=%d %s" % (r.returncode, r.stderr.strip()[-300:])
            except subprocess.TimeoutExpired:
                out = "%s TIMEOUT" % CASES[i][0]
            print(out, flush=True)
