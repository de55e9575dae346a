"""Host-side cost of one SRGAN step: cProfile over sg.train() (the device runs asynchronously; what is measured is
the Python / ctypes / autograd time that has to stay below the device time for the step to be GPU-bound).
Usage: python tools/host_profile.py [--batch 64]"""
import argparse
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
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
    for _ in range(3):
        sg.train(x, lab)
    torch.cuda.synchronize()
    # host time of a step when the device is not the bottleneck: batch 2 keeps kernels tiny
    t0 = time.perf_counter()
    sg.train(x, lab)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("host-side issue time %.1f ms, until device idle %.1f ms" % (t_host * 1e3, t_all * 1e3))
    pr = cProfile.Profile()
    pr.enable()
    sg.train(x, lab)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)
    st.sort_stats("cumulative").print_stats(30)


if __name__ == "__main__":
    main()
