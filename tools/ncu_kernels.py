"""One launch of every kernel the roofline discussion names, at the production shapes (batch 64), between
cudaProfilerStart/Stop - for ONE `ncu --set full --profile-from-start off` capture per round:

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/rXX_kernels \
      python tools/ncu_kernels.py [--engine auto|bf16] [--batch 64]

Each op is run twice untimed first (tensor maps, workspaces, smem attributes), then once inside the profiled range.
Read with `python tools/ncu_brief.py gpurun_out/rXX_kernels.ncu-rep`."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402,F401  (sets sys.path)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import srgan_ops as ops
    import util
    ops.set_conv_engine(a.engine)
    dev, B, CL = "cuda:0", a.batch, torch.channels_last
    act_dtype = torch.bfloat16 if a.engine == "bf16" else torch.float32
    jobs = []

    def conv_job(name, C, H, K, R, stride, pad, transposed=False):
        if transposed:
            Ho = (H - 1) * stride - 2 * pad + R
            d = ops._desc(B, Ho, Ho, K, C, R, R, stride, pad)
            x = torch.randn(B, K, Ho, Ho, device=dev).to(act_dtype).contiguous(memory_format=CL)
            dy = torch.randn(B, C, H, H, device=dev).to(act_dtype).contiguous(memory_format=CL)
            w = (torch.randn(C, K, R, R, device=dev) * 0.05).contiguous(memory_format=CL)
            jobs.append((name + " fwd(=dgrad)", lambda: ops._dgrad(d, dy, ops._conv_weight(w, dy), dy)))
            jobs.append((name + " bwd(=fprop)", lambda: ops._fprop(d, x, ops._conv_weight(w, x), None, 0, 0.0)))
            jobs.append((name + " wgrad", lambda: ops._wgrad(d, x, dy, True, False)))
        else:
            d = ops._desc(B, H, H, C, K, R, R, stride, pad)
            thin = C <= 4 or K <= 4            # the RGB stem / head keep fp32 tensors on both sides (norms convert)
            xd = torch.float32 if thin else act_dtype
            yd = torch.float32 if thin else act_dtype
            x = torch.randn(B, C, H, H, device=dev).to(xd).contiguous(memory_format=CL)
            dy = torch.randn(B, K, d.P, d.Q, device=dev).to(yd).contiguous(memory_format=CL)
            w = (torch.randn(K, C, R, R, device=dev) * 0.05).contiguous(memory_format=CL)
            jobs.append((name + " fprop", lambda: ops._fprop(d, x, ops._conv_weight(w, x), None, 0, 0.0)))
            jobs.append((name + " dgrad", lambda: ops._dgrad(d, dy, ops._conv_weight(w, dy), x)))
            jobs.append((name + " wgrad", lambda: ops._wgrad(d, x, dy, True, False)))

    conv_job("G.res 256>256 k3 @32", 256, 32, 256, 3, 1, 1)
    conv_job("G.down1 64>128 k4s2 @128", 64, 128, 128, 4, 2, 1)
    conv_job("G.up1 T128>64 k4s2 @64", 128, 64, 64, 4, 2, 1, transposed=True)
    conv_job("G.up2 64>3 k7 @128 (thin output)", 64, 128, 3, 7, 1, 3)
    conv_job("G.down0 3>64 k7 @128 (thin input)", 3, 128, 64, 7, 1, 3)
    conv_job("E.first 3>64 k7s2 @128", 3, 128, 64, 7, 2, 1)
    conv_job("E.l3.cmp 512>1024 k3 @9", 512, 9, 1024, 3, 1, 0)
    conv_job("D1.1 64>128 k4s2 @64 (fp32 storage, TF32)", 64, 64, 128, 4, 2, 1) if a.engine != "bf16" else None

    def norm_job(name, C, H):
        x = torch.randn(B, C, H, H, device=dev).to(act_dtype).contiguous(memory_format=CL).requires_grad_(True)
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        cb = torch.randn(B, C, device=dev)
        y = ops.instance_norm_act(x, g, b, cb, None, 1e-5, ops.ACT_RELU, 0.0)
        dy = torch.randn_like(y)
        jobs.append((name + " fwd", lambda: ops.instance_norm_act(x.detach(), g, b, cb, None, 1e-5, ops.ACT_RELU, 0.0)))
        jobs.append((name + " bwd", lambda: torch.autograd.grad(y, x, dy, retain_graph=True)))

    norm_job("IN 256@32", 256, 32)
    norm_job("IN 64@128", 64, 128)
    norm_job("IN 512@7", 512, 7)

    hi = util.histogram_imitation(dev)
    gh = hi.gausshist
    mu = torch.randn(B, 8, device=dev, requires_grad=True)
    w4 = torch.tensor([10.0, 100.0, 100.0, 0.0], device=dev)

    def latent():
        losses, _ = ops.latent_losses(mu, None, n_cfg=B, target=hi.target, bins=gh.bins, hmin=gh.min, hmax=gh.max,
                                      sigma=gh.sigma, flags=ops.LAT_BKL | ops.LAT_CORR | ops.LAT_HIST)
        torch.autograd.grad(losses, mu, w4)
    jobs.append(("latent-loss pair", latent))

    p = torch.nn.Parameter(torch.randn(14_000_000, device=dev))
    opt = ops.FusedAdam([p], lr=1e-4, betas=(0.5, 0.999))
    opt.zero_grad()
    p.grad.normal_()
    jobs.append(("fused Adam, 14 M parameters", opt.step))

    a1, b1 = torch.rand(B, 3, 128, 128, device=dev), torch.rand(B, 3, 128, 128, device=dev)
    jobs.append(("L1 mean 3x128x128", lambda: ops.l1_mean(a1, b1)))

    jobs = [j for j in jobs if a.only in j[0]]
    for name, fn in jobs:
        fn()
        fn()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for name, fn in jobs:
        fn()
        torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled:", [n for n, _ in jobs])


if __name__ == "__main__":
    main()
