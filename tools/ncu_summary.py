#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, avg, share)."""
import collections
import csv
import re
import sys


def main(path, title=""):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        name = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    if title:
        print("# " + title)
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:60s} n={c:4d} total={t:10.1f}us avg={t/c:9.1f}us share={t/tot:.3f}")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
