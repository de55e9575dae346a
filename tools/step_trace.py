"""In-step kernel times with warm caches: one SRGAN step under torch.profiler (CUPTI activity records, no replay,
no cache flush), aggregated per kernel.  Complements the ncu launch list, whose per-launch times are cold-cache.
Usage: python tools/step_trace.py [--batch 64] > gpurun_out/step_trace.log"""
import argparse
import collections
import os
import re
import sys

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
    for _ in range(2):
        sg.train(x, lab)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        sg.train(x, lab)
        torch.cuda.synchronize()
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    t0, t1 = None, None
    spans = []
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            if not ev.name.startswith("Optimizer.step"):
                spans.append((ev.time_range.start, ev.time_range.end, re.sub(r"[(<].*", "", ev.name)))
            name = re.sub(r"\(.*", "", ev.name)
            if "at::" in name:
                name = re.sub(r"<.*", "", name)
            tot[name] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            cnt[name] += 1
            s, e = ev.time_range.start, ev.time_range.end
            t0 = s if t0 is None else min(t0, s)
            t1 = e if t1 is None else max(t1, e)
    T = sum(tot.values())
    print("kernel time %.1f ms over %d launches; span first..last kernel %.1f ms" % (T / 1e3, sum(cnt.values()),
                                                                                   (t1 - t0) / 1e3))
    # idle gaps of the device between consecutive kernels (launch-bound stretches), attributed to the kernel that
    # the device was waiting for
    spans.sort()
    gap_by = collections.defaultdict(float)
    gaps = []
    end = spans[0][1]
    for s0, e0, nm in spans[1:]:
        if s0 > end:
            gap_by[nm] += s0 - end
            gaps.append((s0 - end, nm))
        end = max(end, e0)
    print("idle gaps: %.1f ms in total; %d gaps > 20 us" % (sum(g for g, _ in gaps) / 1e3, sum(1 for g, _ in gaps if g > 20)))
    for k, v in sorted(gap_by.items(), key=lambda kv: -kv[1])[:15]:
        print("  waiting for %-50s %.2f ms" % (k[:50], v / 1e3))
    print("  largest:", [(round(g), n[:30]) for g, n in sorted(gaps, reverse=True)[:12]])
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:45]:
        print("| `%s` | %d | %.2f | %.1f%% | %.1f |" % (k[:90], cnt[k], v / 1e3, 100 * v / T, v / cnt[k]))


if __name__ == "__main__":
    main()
