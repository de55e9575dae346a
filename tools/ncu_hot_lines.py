"""Top source lines of a kernel by warp-stall samples: python tools/ncu_hot_lines.py report.ncu-rep [N]
(reads `ncu -i report --page source --csv`; compile with -lineinfo and capture with --import-source on)."""
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    if not out.strip():
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    recs = []
    for r in rows:
        if "Source" in r and any("Sampl" in c for c in r):
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            recs.append(dict(zip(hdr, r)))
    if not hdr:
        print("no source table found; header candidates:", [r[:6] for r in rows[:5]])
        return
    key = next((c for c in hdr if "Sampling" in c and "All" in c), None) or next(c for c in hdr if "Sampl" in c)
    def val(d):
        try:
            return float(d[key].replace(",", ""))
        except Exception:
            return 0.0
    total = sum(val(d) for d in recs) or 1.0
    recs.sort(key=val, reverse=True)
    print("columns:", [c for c in hdr][:12])
    print("total samples (%s): %.0f over %d lines" % (key, total, len(recs)))
    for d in recs[:top]:
        src = d.get("Source", "")[:110]
        loc = d.get("Address", d.get("#", ""))
        print("%6.0f %5.1f%%  %s  %s" % (val(d), 100 * val(d) / total, loc, src))


if __name__ == "__main__":
    main()
