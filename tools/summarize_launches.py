#!/usr/bin/env python
"""Group an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`)
by kernel: count, total time, share, average, DRAM bytes. Only the LAST step is summarised: launches from the last
`first_kernel` launch (default conv0_ln_gelu = the first kernel of a step) to the end.
usage: python tools/summarize_launches.py launches.csv [first_kernel_substring]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    first = sys.argv[2] if len(sys.argv) > 2 else "conv0_ln_gelu"
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    per = {}
    order = []
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        if i not in per:
            per[i] = {"name": r["Kernel Name"], "us": 0.0, "rd": 0.0, "wr": 0.0}
            order.append(i)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        if r["Metric Name"] == "gpu__time_duration.sum":
            per[i]["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        else:
            b = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            per[i]["rd" if "read" in r["Metric Name"] else "wr"] = b
    rows = [per[i] for i in order]
    starts = [k for k, r in enumerate(rows) if first in r["name"]]
    rows = rows[starts[-1]:] if starts else rows
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows:
        n = r["name"].replace("void ", "").replace("b2s::<unnamed>::", "").replace("at::<unnamed>::", "at::")
        n = re.sub(r"\(.*", "", n)
        n = re.sub(r"std::array.*", "", n)[:70]
        agg[n][0] += 1
        agg[n][1] += r["us"]
        agg[n][2] += r["rd"] + r["wr"]
    total = sum(v[1] for v in agg.values())
    print(f"# {len(rows)} launches of the last step, {total / 1e3:.2f} ms serialised under ncu (cold-cache), "
          f"DRAM {sum(v[2] for v in agg.values()) / 1e9:.2f} GB")
    for n, (c, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / 1e3:9.3f} ms {100 * us / total:5.1f}%  n={c:5d}  avg={us / c:8.1f} us  dram={by / 1e9:7.3f} GB  {n}")


if __name__ == "__main__":
    main()
