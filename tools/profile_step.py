"""One SRGAN training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).
Usage: python tools/profile_step.py [--batch B] [--engine auto|fp32]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (sets sys.path for pyfiles/ and oracle/)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--warm", type=int, default=1)
    a = ap.parse_args()
    import cases
    import srgan_ops as ops
    ops.set_conv_engine(a.engine)
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
    for _ in range(a.warm):
        sg.train(x, lab)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    errs = sg.train(x, lab)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("step ok", [float(e) for e in errs], "abi calls", ops.abi_calls)


if __name__ == "__main__":
    main()
