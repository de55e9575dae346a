"""Bring-up: run an eager step with torch.empty() filled with NaN (deterministic-algorithms mode) and report the first
module / op that turns finite inputs into NaN: finds kernels that read memory they never wrote."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import numpy as np, torch
import cases
import srgan_ops as ops
DEV = "cuda"
MODE = os.environ.get("DBG_MODE", "all")
if MODE in ("all", "det", "fill"):
    torch.use_deterministic_algorithms(True, warn_only=True)
    torch.utils.deterministic.fill_uninitialized_memory = MODE in ("all", "fill")
c = dict(cases.CASES["srgan_small"], batch=4, k=2)
model, util, nb = cases.use_product_modules()
torch.manual_seed(0); np.random.seed(0)
nets = tuple(n.to(DEV) for n in cases.build_nets(model, c, DEV))
torch.manual_seed(1)
sg = cases.build_trainer(nb, c, nets, DEV)
# poison the shared workspace too
if MODE in ("all", "ws"):
    ws = ops._workspace(torch.device("cuda:0"), 1 << 28)
    ws.view(torch.float32).fill_(float("nan"))
print("MODE", MODE)
seen = []
def hook(name):
    def f(mod, inp, out):
        def flat(o):
            if torch.is_tensor(o): return [o]
            if isinstance(o, (list, tuple)): return [t for e in o for t in flat(e)]
            return []
        bad_in = any(torch.isnan(t).any().item() for t in flat(inp) if t.is_floating_point())
        bad_out = any(torch.isnan(t).any().item() for t in flat(out) if t.is_floating_point())
        if bad_out and not bad_in and len(seen) < 12:
            seen.append(name); print("NaN introduced by", name, type(mod).__name__)
    return f
for nm, n in zip("GDE", nets):
    for mn, m in n.named_modules():
        if mn: m.register_forward_hook(hook(nm + "." + mn))
# wrap autograd functions' backward: check grads after backward instead
x, label = cases.synthetic_batch(c["batch"], util.get_target, seed=100)
errs = sg.train(x.to(DEV), {"source": label["source"].to(DEV), "target": label["target"]})
torch.cuda.synchronize()
print("losses", [float(e) for e in errs])
for nm, n in zip("GDE", nets):
    for pn, p in n.named_parameters():
        if p.grad is not None and torch.isnan(p.grad).any():
            print("NaN grad", nm, pn, tuple(p.shape)); break
    for pn, p in n.named_parameters():
        if torch.isnan(p).any():
            print("NaN param", nm, pn, tuple(p.shape)); break
