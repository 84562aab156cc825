#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches and time per
kernel, share of the total (cold-cache and serialised under the profiler: shares, not absolutes)."""
import csv
import sys
from collections import OrderedDict


def main(path, title=""):
    rows = [l for l in open(path) if l.startswith('"')]
    agg = OrderedDict()
    for r in csv.DictReader(rows):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        t = float(r["Metric Value"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6)
        n, tot = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, tot + t)
    total = sum(t for _, t in agg.values())
    print(title or path)
    print("launch shares (cold-cache, serialised under the profiler: shares, not absolutes)")
    for name, (n, t) in agg.items():
        print(f"{n:4d} x {name[:72]:72s} {t:12.3f} ms {100 * t / total:7.2f}%")
    print(f"total {total:.3f} ms over {sum(n for n, _ in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
