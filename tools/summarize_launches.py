"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv [title] > profiles/xxx.md"""
import collections
import csv
import re
import sys


def main(path, title):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        try:
            name, v, unit = row["Kernel Name"], float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        except Exception:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", name)
        if "at::" in name:
            name = re.sub(r"<.*", "", name)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    ours = sum(v for k, v in tot.items() if "srgan::" in k)
    print("# %s\n" % title)
    print("Source: `%s` (ncu --metrics gpu__time_duration.sum --clock-control none, one training step between "
          "cudaProfilerStart/Stop; per-launch times are cold-cache and serialised: compare SHARES).\n" % path)
    print("Total kernel time %.1f ms over %d launches; %.1f%% in this repo's kernels (`srgan::*`), the rest are "
          "PyTorch elementwise/copy kernels of the autograd tape.\n" % (T / 1e3, sum(cnt.values()), 100 * ours / T))
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:40]:
        print("| `%s` | %d | %.2f | %.1f%% | %.1f |" % (k[:90], cnt[k], v / 1e3, 100 * v / T, v / cnt[k]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "kernel launch summary")
