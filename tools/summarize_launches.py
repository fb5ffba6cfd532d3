#!/usr/bin/env python
"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: count, total and share.
usage: python tools/summarize_launches.py launches.csv [skip_first_n_launches]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
    rows = rows[skip:] if skip >= 0 else rows[skip:]
    agg = defaultdict(lambda: [0, 0.0])
    for name, us in rows:
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"[<(].*", "", name)
        agg[name][0] += 1
        agg[name][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {total / 1e3:.2f} ms (cold-cache, serialised under ncu)")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / 1e3:9.3f} ms {100 * us / total:5.1f}%  n={n:5d}  avg={us / n:8.1f} us  {name}")


if __name__ == "__main__":
    main()
