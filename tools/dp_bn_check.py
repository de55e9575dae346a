"""Data-parallel check of the batch-statistics norms: every rank evaluates CBBNorm2d / BatchNorm2d on its slice of a
global batch (tables all-gathered between the kernel stages) and on the whole batch alone; outputs, input gradients and
(summed) parameter gradients must agree.

  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_bn_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402,F401  (path setup)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    import srgan_ops as ops
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    dist.init_process_group("nccl", device_id=torch.device(dev))
    CL = torch.channels_last
    ok = True
    for cond in (True, False):
        g = torch.Generator().manual_seed(42)
        N, C, H, W = 8 * world, 64, 16, 16
        x = (torch.randn(N, C, H, W, generator=g) * 1.3 + 0.2).to(dev).contiguous(memory_format=CL)
        probe = torch.randn(N, C, H, W, generator=g).to(dev).contiguous(memory_format=CL)
        gamma = (torch.rand(C, generator=g) + 0.5).to(dev)
        beta = (torch.randn(C, generator=g) * 0.1).to(dev)
        cb = torch.randn(N, C, generator=g).to(dev) if cond else None
        b = N // world
        sl = slice(rank * b, (rank + 1) * b)

        def run(xs, ps, cbs, sync):
            xs = xs.clone().requires_grad_(True)
            ga, be = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
            rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
            y = ops.batch_norm_act(xs, ga, be, cbs, None, rm, rv, True, 0.1, 1e-5, cond, ops.ACT_RELU, 0.0, sync)
            (y * ps).sum().backward()
            return y.detach(), xs.grad, ga.grad, be.grad, rm, rv
        yf, dxf, dgf, dbf, rmf, rvf = run(x, probe, cb, False)
        yh, dxh, dgh, dbh, rmh, rvh = run(x[sl].contiguous(memory_format=CL), probe[sl].contiguous(memory_format=CL),
                                          cb[sl].contiguous() if cond else None, True)
        dist.all_reduce(dgh)
        dist.all_reduce(dbh)
        errs = {"y": rel(yh, yf[sl]), "dx": rel(dxh, dxf[sl]), "dgamma": rel(dgh, dgf), "dbeta": rel(dbh, dbf),
                "running_mean": rel(rmh, rmf), "running_var": rel(rvh, rvf)}
        good = all(v < 2e-5 for v in errs.values())
        ok &= good
        if rank == 0:
            print("cond=%d world=%d" % (cond, world), {k: "%.2e" % v for k, v in errs.items()}, "OK" if good else "FAIL")
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_BN_CHECK", "PASS" if float(t) == 1.0 else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
