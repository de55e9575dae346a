"""Generate tests/golden/<case>.npz by running the UNMODIFIED reference on CPU -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py [case ...]

For every case of oracle/cases.py:
  seed -> build reference nets + trainer (default init; the histogram target consumes the CPU RNG at
  construction, like in the notebooks) -> re-seed -> one `sg.train(x, label)` with the torch-1.4 optimizer
  shim -> store losses, the first encoder output (z, mu, logvar), latent statistics and digests
  (L2 norm, probe dot-product, 32 samples) of every parameter's initial value, of its gradient at each
  optimizer step and of its value after each step.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import ref_harness  # noqa: E402
import srgan_oracle  # noqa: E402


def run_reference(case):
    c = cases.CASES[case]
    ref_model, ref_util, ref_nb = ref_harness.load_reference()
    torch.set_num_threads(8)
    torch.manual_seed(c["seed"])
    np.random.seed(c["seed"])
    nets = cases.build_nets(ref_model, case)
    init = cases.state_dicts(nets)
    sg = cases.build_trainer(ref_nb, case, nets, "cpu")
    ref_harness.torch14_step(sg.optG)
    ref_harness.torch14_step(sg.optE)
    record = {}
    ref_harness.capture_steps(sg.optG, sg.G, "G", record)
    ref_harness.capture_steps(sg.optE, sg.E, "E", record)
    if isinstance(sg.optD, list):
        for i, o in enumerate(sg.optD):
            ref_harness.capture_steps(o, sg.D[i], "Dc%d_" % i, record)
    else:
        ref_harness.capture_steps(sg.optD, sg.D, "D", record)
    enc_out = []
    hook = sg.E.register_forward_hook(lambda m, i, o: enc_out.append([t.detach().clone() for t in o[:3]]))
    x, label = cases.synthetic_batch(c["batch"], ref_util.get_target)
    torch.manual_seed(c["seed"] + 1000)
    errs = sg.train(x, label)
    hook.remove()
    # E gradients that exist after train() (phase 2; E is not stepped there)
    record["E_final.grad"] = {n: (None if p.grad is None else p.grad.detach().clone())
                              for n, p in sg.E.named_parameters()}
    # the reference's own loss functions evaluated on the encoder output of the source batch
    mu = enc_out[0][1]
    ref_fn = {"corr": ref_util.corrcoef(mu.t()).detach(), "corr_loss": ref_util.corrcoef_loss(mu.t(), "cpu").detach()}
    if hasattr(sg, "hi"):
        ref_fn["hist_loss"] = sg.hi.loss(mu).detach()
        ref_fn["hist"] = torch.stack([sg.hi.gausshist(mu[:, d]) for d in range(mu.shape[1])]).detach()
    return dict(init=init, record=record, ref_fn=ref_fn, errs=[float(e) for e in errs], enc=enc_out[0], x=x, label=label,
                hist_target=getattr(getattr(sg, "hi", None), "target", None), n_batch=c["batch"])


def pack(res):
    out = {"errs": np.array(res["errs"], dtype=np.float64)}
    z, mu, logvar = res["enc"]
    out["enc.z"], out["enc.mu"], out["enc.logvar"] = z.numpy(), mu.numpy(), logvar.numpy()
    st = srgan_oracle.latent_statistics(mu, res["n_batch"], res["hist_target"])
    for k, v in st.items():
        out["stats." + k] = v.detach().numpy()
    for k, v in res["ref_fn"].items():
        out["ref." + k] = v.numpy()
    if res["hist_target"] is not None:
        out["hist_target"] = res["hist_target"].detach().numpy()
    g0, d0, e0 = res["init"]
    nets = {"G": g0, "E": e0}
    if isinstance(d0, list):
        for i, d in enumerate(d0):
            nets["Dc%d" % i] = d
    else:
        nets["D"] = d0
    for net, sd in nets.items():
        for k, v in sd.items():
            out["init.%s.%s" % (net, k)] = srgan_oracle.digest(v).numpy()
    for step, tensors in res["record"].items():
        for k, v in tensors.items():
            if v is not None:
                out["%s.%s" % (step, k)] = srgan_oracle.digest(v).numpy()
            else:
                out["%s.%s" % (step, k)] = np.zeros(0, dtype=np.float32)
    out["x.digest"] = srgan_oracle.digest(res["x"]).numpy()
    out["label.source"] = res["label"]["source"].numpy()
    out["label.target"] = res["label"]["target"].numpy()
    return out


def main(argv):
    names = argv or list(cases.CASES)
    os.makedirs(cases.GOLDEN, exist_ok=True)
    for name in names:
        res = run_reference(name)
        path = os.path.join(cases.GOLDEN, name + ".npz")
        np.savez_compressed(path, **pack(res))
        print("%-20s errs=%s -> %s (%.1f KB)" % (name, res["errs"], path, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main(sys.argv[1:])
