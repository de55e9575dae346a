"""CPU, world_size 2, gloo: the host-side data-parallel logic of the trainers (gradient mean all-reduce over a
flat buffer, global-batch noise slicing, reported-loss averaging).  No kernel is launched."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, _, nb = cases.use_product_modules()
        import srgan_ops as ops
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.Linear(5, 2))
        net[0].weight.data = net[0].weight.data.contiguous(memory_format=torch.channels_last)
        for i, p in enumerate(net.parameters()):
            g = torch.full_like(p, float(rank + 1)) + torch.arange(p.numel()).view_as(p).float() * (rank + 1)
            p.grad = g.contiguous(memory_format=torch.channels_last) if p.dim() == 4 else g
        nb._sync_grads(net, None)
        ok = True
        for p in net.parameters():
            exp = torch.full_like(p, 1.5) + torch.arange(p.numel()).view_as(p).float() * 1.5
            ok &= bool(torch.allclose(p.grad, exp))
        # noise: every rank draws the global batch and keeps its rows
        torch.manual_seed(7)
        mine = ops.host_normal(3, 8, "cpu")
        torch.manual_seed(7)
        full = torch.randn(3 * world, 8)
        ok &= bool(torch.equal(mine, full[rank * 3:(rank + 1) * 3]))
        # batch-norm tables: every rank ends up with the global [N_all, C] table and knows its first row
        tab = torch.arange(2 * 3, dtype=torch.float32).view(2, 3) + 100 * rank
        allt, row0 = ops._gather_rows(tab)
        exp_all = torch.cat([torch.arange(6, dtype=torch.float32).view(2, 3) + 100 * r for r in range(world)])
        ok &= bool(torch.equal(allt, exp_all)) and row0 == 2 * rank
        t = nb._UnrolledTrainer()
        rep = t._report([torch.tensor(float(rank)), 0, torch.tensor(2.0)])
        ok &= abs(float(rep[0]) - 0.5) < 1e-7 and rep[1] == 0 and float(rep[2]) == 2.0
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_grad_sync_noise_slicing_and_reporting_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(0) is True and out.get(1) is True


def test_single_process_is_identity():
    _, _, nb = cases.use_product_modules()
    import srgan_ops as ops
    assert ops.dp_rank_world() == (0, 1)
    torch.manual_seed(3)
    a = ops.host_normal(4, 8, "cpu")
    torch.manual_seed(3)
    assert torch.equal(a, torch.randn(4, 8))
    t = torch.ones(3)
    assert nb._allreduce_mean_(t) is t


def test_get_target_and_class_encode_match_reference_semantics():
    _, util, _ = cases.use_product_modules()
    lab = torch.tensor([0, 3, 1, 2, 2])
    np.random.seed(0)
    t = util.get_target(lab, (0, 1, 2, 3), whole=False)
    assert t.shape == (5, 3)
    for i, row in enumerate(t):
        assert sorted(row.tolist()) == sorted(set(range(4)) - {int(lab[i])})
    assert util.get_target(lab, (0, 1, 2, 3), whole=True, shuffle=False).tolist() == [[0, 1, 2, 3]] * 5
    one = util.class_encode(lab, "cpu", np.eye(4))
    assert one.dtype == torch.float32 and torch.equal(one, torch.eye(4)[lab])
