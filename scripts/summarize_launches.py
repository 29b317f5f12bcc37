#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total ms,
share of the captured window).  Usage: summarize_launches.py launches.csv [skip count] > profiles/xxx.md
(skip/count select a window of launches, e.g. one timed pass of bench.py)."""
import collections
import csv
import sys


def main(path, skip=0, count=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    for row in rows[skip:None if count is None else skip + count]:
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        k = row["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.3f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1], *[int(a) for a in sys.argv[2:4]])
