"""Data-parallel invariance check: an N-rank run (one process per GPU, NCCL) of one SRGAN step on a global
batch must reproduce the single-GPU run on the same global batch: losses, latent statistics, and the
gradients seen by every optimizer step.

  python tools/dp_check.py --save gpurun_out/dp_ref.pt                       (1 GPU, whole batch)
  torchrun --nproc-per-node N tools/dp_check.py --compare gpurun_out/dp_ref.pt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402,F401  (path setup)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def run(global_batch, width, k):
    import cases
    import srgan_ops as ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and 4 % world == 0 and os.environ.get("SRGAN_DBG_NORM_F32_PARTIALS", "0") != "0":
        # With fp32 slice partials (bring-up switch) the instance-norm statistics depend on the slice count, which the
        # grid planner derives from the number of images a rank holds; pin it to the 1-GPU run's so that what is
        # compared is the data-parallel logic.  The default fp64 partials over fixed atoms need no pin.
        os.environ.setdefault("SRGAN_DBG_NORM_CTAS_PER_SM", str(4 // world))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    case = dict(kind="srgan", nch=width, dis_nch=width, enc_nch=width, res_num=2 if width < 64 else 6,
                batch=global_batch, k=k, lbd=dict(cases.PROPOSED, **{"class": 1}), feature="mu", seed=0)
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(0)
    np.random.seed(0)
    G, D, E = cases.build_nets(model, case, dev)
    sg = cases.build_trainer(nb, case, (G.to(dev), D.to(dev), E.to(dev)), dev)
    x, lab = cases.synthetic_batch(global_batch, util.get_target)
    b = global_batch // world
    sl = slice(rank * b, (rank + 1) * b)
    rec = {}

    def capture(opt, net, key):
        orig = opt.step
        cnt = {"n": 0}

        def step(*a, **kw):
            rec["%s%d" % (key, cnt["n"])] = torch.cat([p.grad.detach().reshape(-1).float().cpu() if p.grad is not None
                                                       else torch.zeros(p.numel()) for p in net.parameters()])
            cnt["n"] += 1
            return orig(*a, **kw)
        opt.step = step
    capture(sg.optG, G, "G")
    capture(sg.optD, D, "D")
    capture(sg.optE, E, "E")
    torch.manual_seed(1)
    errs = sg.train(x[sl].to(dev), {"source": lab["source"][sl].to(dev), "target": lab["target"][sl]})
    torch.cuda.synchronize()
    rec["errs"] = torch.tensor([float(e) for e in errs], dtype=torch.float64)
    rec["abi_calls"] = ops.abi_calls
    return rec, rank, world


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--save")
    ap.add_argument("--compare")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--width", type=int, default=64)
    ap.add_argument("--k", type=int, default=2)
    a = ap.parse_args()
    rec, rank, world = run(a.batch, a.width, a.k)
    if a.save:
        torch.save(rec, a.save)
        print("saved", a.save, rec["errs"].tolist())
    if a.compare and rank == 0:
        ref = torch.load(a.compare)
        ok = True
        print("world %d  losses %s  vs 1-GPU %s" % (world, rec["errs"].tolist(), ref["errs"].tolist()))
        for got, r in zip(rec["errs"].tolist(), ref["errs"].tolist()):
            ok &= abs(got - r) <= 2e-4 * max(1.0, abs(r))
        for key in sorted(k for k in ref if k not in ("errs", "abi_calls")):
            rel = float((rec[key].double() - ref[key].double()).norm() / ref[key].double().norm().clamp_min(1e-30))
            # G1 = generator gradients of phase 2, taken AFTER the phase-1 Adam step: Adam turns 1e-4 relative
            # gradient differences into sign-level weight differences, so this quantity is ill-conditioned; two
            # fp32 CPU runs of the reference itself that differ only in thread count disagree by 8.5e-2
            # (SURVEY F12).  Everything up to that step is held to 2e-3 / 2e-2.
            # With bf16 storage the same sign-level weight differences are amplified by the 2^-9 rounding of every stored
            # activation (measured 2.6e-1 against 8.5e-2 for TF32): 4e-1.
            g1 = 4e-1 if os.environ.get("SRGAN_CONV_ENGINE", "").lower() == "bf16" else 2e-1
            tol = g1 if key in ("G1",) else (2e-2 if key.startswith("G") or key.startswith("E") else 2e-3)
            print("  grads at %-3s rel-L2 vs 1-GPU = %.3e (tol %.0e)" % (key, rel, tol))
            ok &= rel < tol
        print("DP_CHECK", "PASS" if ok else "FAIL")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
