import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "style-restricted_gan_b200", "pyfiles"))
import torch, torch.nn.functional as F
import srgan_ops as ops
for C in (3, 16):
  for H, W in ((128, 128), (9, 7), (4, 4)):
    x = torch.randn(2, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = ops.avg_pool3s2(x)
    xr = x.detach().clone().requires_grad_(True)
    yr = F.avg_pool2d(xr, 3, 2, 1, count_include_pad=False)
    g = torch.randn_like(yr)
    y.backward(g.contiguous(memory_format=torch.channels_last)); yr.backward(g)
    e1 = (x.grad - xr.grad).abs().max().item()
    x.grad = None
    y = ops.avg_pool3s2(x); y.backward(g.contiguous())
    e2 = (x.grad - xr.grad).abs().max().item()
    print(C, H, W, "fwd", (y - yr).abs().max().item(), "bwd(cl dy)", e1, "bwd(nchw dy)", e2)
