"""One small instance of every kernel family, for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`
where that tool is available; also a quick coverage smoke test of the library."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402,F401
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    import cases
    import srgan_ops as ops
    dev = "cuda:0"
    CL = torch.channels_last
    g = torch.Generator().manual_seed(0)

    def rnd(*s):
        return torch.randn(*s, generator=g).to(dev)
    # convolutions: generic, strided, transposed, thin input / output, 1x1, heads, odd sizes
    for (N, C, H, K, R, st, pad) in ((2, 256, 32, 256, 3, 1, 1), (2, 64, 32, 128, 4, 2, 1), (3, 3, 32, 64, 7, 1, 3),
                                     (2, 64, 32, 3, 7, 1, 3), (2, 3, 32, 64, 4, 2, 1), (2, 64, 15, 128, 1, 1, 0),
                                     (2, 512, 8, 1, 4, 1, 1), (2, 512, 8, 4, 8, 1, 0), (1, 24, 13, 40, 3, 2, 1),
                                     (3, 128, 9, 256, 3, 1, 0), (2, 3, 64, 64, 7, 2, 1)):
        x = rnd(N, C, H, H).contiguous(memory_format=CL).requires_grad_(True)
        w = (rnd(K, C, R, R) * 0.05).contiguous(memory_format=CL).requires_grad_(True)
        b = rnd(K).requires_grad_(True)
        y = ops.conv2d(x, w, b, st, pad, "zeros", ops.ACT_LRELU, 0.01)
        y.sum().backward()
    for (N, Ci, H, Co) in ((2, 256, 16, 128), (2, 128, 16, 64)):
        x = rnd(N, Ci, H, H).contiguous(memory_format=CL).requires_grad_(True)
        w = (rnd(Ci, Co, 4, 4) * 0.05).contiguous(memory_format=CL).requires_grad_(True)
        ops.conv_transpose2d(x, w, 2, 1).sum().backward()
    x = rnd(2, 64, 16, 16).contiguous(memory_format=CL).requires_grad_(True)
    w = (rnd(64, 64, 3, 3) * 0.05).contiguous(memory_format=CL).requires_grad_(True)
    y, skip = ops.conv2d_skip(x, w, 1, 1)
    (y.sum() + skip.sum()).backward()
    # norms
    for (N, C, H) in ((3, 256, 32), (2, 64, 62), (5, 512, 7), (2, 24, 10), (1, 8, 3)):
        x = rnd(N, C, H, H).contiguous(memory_format=CL).requires_grad_(True)
        ga, be, cb = rnd(C).requires_grad_(True), rnd(C).requires_grad_(True), rnd(N, C).requires_grad_(True)
        ops.instance_norm_act(x, ga, be, cb, None, 1e-5, ops.ACT_RELU, 0.0).sum().backward()
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        for cond in (True, False):
            x2 = x.detach().clone().requires_grad_(True)
            ops.batch_norm_act(x2, ga, be, cb if cond else None, None, rm, rv, True, 0.1, 1e-5, cond,
                               ops.ACT_LRELU, 0.2).sum().backward()
    # one full small training step (pools, pads, heads, losses, Adam)
    model, util, nb = cases.use_product_modules()
    c = dict(cases.CASES["srgan_small"], batch=2, k=1)
    torch.manual_seed(0)
    np.random.seed(0)
    nets = tuple(n.to(dev) for n in cases.build_nets(model, c, dev))
    sg = cases.build_trainer(nb, c, nets, dev)
    xb, lab = cases.synthetic_batch(2, util.get_target)
    errs = sg.train(xb.to(dev), {"source": lab["source"].to(dev), "target": lab["target"]})
    torch.cuda.synchronize()
    print("sanitize_small ok", [float(e) for e in errs])


if __name__ == "__main__":
    main()
