"""Which PyTorch (aten) operators still run inside one SRGAN step, with call counts: the glue that is not ours.
Usage: python tools/aten_ops.py [--batch 64]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    a = ap.parse_args()
    import cases
    dev = "cuda:0"
    case = bench.build_case("srgan_nb03", a.batch)
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(0)
    np.random.seed(0)
    G, D, E = cases.build_nets(model, case, dev)
    sg = cases.build_trainer(nb, case, (G.to(dev), D.to(dev), E.to(dev)), dev)
    x, lab = cases.synthetic_batch(a.batch, util.get_target)
    x = x.to(dev)
    lab = {"source": lab["source"].to(dev), "target": lab["target"]}
    for _ in range(2):
        sg.train(x, lab)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=False) as prof:
        sg.train(x, lab)
        torch.cuda.synchronize()
    rows = [(e.count, e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total)
            for e in prof.key_averages() if e.key.startswith("aten::")]
    rows.sort(reverse=True)
    for cnt, key, dt in rows[:40]:
        print("%6d  %-40s device %.2f ms" % (cnt, key, dt / 1e3))


if __name__ == "__main__":
    main()
