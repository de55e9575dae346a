"""Generate tests/golden/cbbnorm.npz by running the UNMODIFIED reference CBBNorm2d on CPU -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):   python oracle/make_golden_norms.py

Two training-mode calls (running statistics evolve) followed by one evaluation-mode call of
model.CBBNorm2d(8, num_con=12) (ref pyfiles/model.py:75-171) on seeded inputs; stores inputs, parameters, outputs,
input / parameter gradients of sum(y * probe) and the running statistics after every call.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import ref_harness  # noqa: E402


def main():
    ref_model, _, _ = ref_harness.load_reference()
    torch.set_num_threads(4)
    torch.manual_seed(7)
    N, C, H, W, J = 5, 8, 6, 7, 12
    m = ref_model.CBBNorm2d(C, num_con=J)          # affine=True, track_running_stats=True, weight ~ U(0,1)
    with torch.no_grad():
        m.bias.uniform_(-0.3, 0.3)
    out = {"weight": m.weight.detach().numpy().copy(), "bias": m.bias.detach().numpy().copy(),
           "lin_w": m.ConBias[0].weight.detach().numpy().copy(), "lin_b": m.ConBias[0].bias.detach().numpy().copy()}
    for call, training in enumerate((True, True, False)):
        m.train(training)
        x = (torch.randn(N, C, H, W) * (1.0 + 0.5 * call) + 0.7 * call).requires_grad_(True)
        con = torch.randn(N, J).requires_grad_(True)
        probe = torch.randn(N, C, H, W)
        y = m(x, con)
        m.zero_grad()
        (y * probe).sum().backward()
        pre = "call%d." % call
        out[pre + "x"], out[pre + "con"], out[pre + "probe"] = x.detach().numpy(), con.detach().numpy(), probe.numpy()
        out[pre + "y"] = y.detach().numpy()
        out[pre + "dx"], out[pre + "dcon"] = x.grad.numpy(), con.grad.numpy()
        out[pre + "dweight"], out[pre + "dbias"] = m.weight.grad.numpy().copy(), m.bias.grad.numpy().copy()
        out[pre + "dlin_w"] = m.ConBias[0].weight.grad.numpy().copy()
        out[pre + "dlin_b"] = m.ConBias[0].bias.grad.numpy().copy()
        out[pre + "running_mean"] = m.running_mean.numpy().copy()
        out[pre + "running_var"] = m.running_var.numpy().copy()
        out[pre + "training"] = np.array(int(training))
    path = os.path.join(cases.GOLDEN, "cbbnorm.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
