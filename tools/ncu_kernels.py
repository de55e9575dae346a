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
            thin = C <= 4 or K <= 4            # fp32 on both sides: the TF32-engine kernels of the RGB layers / the E and D stems
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

    if a.engine == "bf16":
        # thin16: the RGB stem / head with a bf16 fat side
        BF = torch.bfloat16
        img = (torch.rand(B, 3, 128, 128, device=dev) * 2 - 1).contiguous(memory_format=CL)
        ws_ = torch.nn.Parameter((torch.randn(64, 3, 7, 7, device=dev) * 0.05).contiguous(memory_format=CL))
        wh_ = torch.nn.Parameter((torch.randn(3, 64, 7, 7, device=dev) * 0.05).contiguous(memory_format=CL))
        fat = torch.randn(B, 64, 128, 128, device=dev).to(BF).contiguous(memory_format=CL).requires_grad_(True)
        imgg = img.clone().requires_grad_(True)
        ys = ops.conv2d(imgg, ws_, None, 1, 3, out_dtype=BF)
        yh = ops.conv2d(fat, wh_, None, 1, 3, act=ops.ACT_TANH)
        gs, gh_ = torch.randn_like(ys), torch.randn_like(yh)
        jobs.append(("thin16 stem fprop (fp32 image -> bf16)", lambda: ops.conv2d(img, ws_, None, 1, 3, out_dtype=BF)))
        jobs.append(("thin16 stem dgrad + wgrad", lambda: torch.autograd.grad(ys, [imgg, ws_], gs, retain_graph=True)))
        jobs.append(("thin16 head fprop (bf16 -> fp32 image, tanh)", lambda: ops.conv2d(fat.detach(), wh_, None, 1, 3, act=ops.ACT_TANH)))
        jobs.append(("thin16 head dgrad + wgrad", lambda: torch.autograd.grad(yh, [fat, wh_], gh_, retain_graph=True)))

        # the one-pass (cluster) instance norm, opt-in
        xo = torch.randn(B, 256, 32, 32, device=dev).to(BF).contiguous(memory_format=CL).requires_grad_(True)
        g1, b1_ = torch.ones(256, device=dev), torch.zeros(256, device=dev)
        cb1 = torch.randn(B, 256, device=dev)
        lib = ops._lib()

        def onepass(fn):
            def run():
                prev = lib.srgan_inorm_onepass_enable(1)
                try:
                    return fn()
                finally:
                    lib.srgan_inorm_onepass_enable(prev)
            return run
        prev = lib.srgan_inorm_onepass_enable(1)
        yo = ops.instance_norm_act(xo, g1, b1_, cb1, None, 1e-5, ops.ACT_RELU, 0.0)
        lib.srgan_inorm_onepass_enable(prev)
        dyo = torch.randn_like(yo)
        jobs.append(("IN 256@32 one-pass cluster fwd", onepass(
            lambda: ops.instance_norm_act(xo.detach(), g1, b1_, cb1, None, 1e-5, ops.ACT_RELU, 0.0))))
        jobs.append(("IN 256@32 one-pass cluster bwd", onepass(lambda: torch.autograd.grad(yo, xo, dyo, retain_graph=True))))

    # f4: cross entropy of the notebook-04 job (batch 512) and PRDC on 2048 x 2048 features of 4096 dimensions
    xe = torch.randn(512, 4, device=dev, requires_grad=True)
    le = torch.randint(0, 4, (512,), device=dev)
    jobs.append(("cross entropy 512x4 fwd + bwd", lambda: torch.autograd.grad(ops.cross_entropy(xe, le), xe)))
    fr, ff = torch.randn(2048, 4096, device=dev), torch.randn(2048, 4096, device=dev) * 1.1
    jobs.append(("PRDC 2048 x 2048 x 4096, k = 5", lambda: ops.prdc_counts(fr, ff, 5)))

    # gradient fold of the generator's flat buffers (14 M parameters, two parked contributions)
    gf = [torch.zeros(14_000_000, device=dev) for _ in range(3)]
    jobs.append(("gradient fold, 14 M parameters, 2 parked contributions", lambda: ops._call(
        "srgan_grad_fold", ops._p(gf[0]), ops._p(gf[1]), ops._p(gf[2]), gf[0].numel(), ops._stream())))

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
