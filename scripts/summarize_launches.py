#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py launches.csv [out.md]"""
import collections
import csv
import sys


def main(path, out=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        agg[row["Kernel Name"]][0] += 1
        agg[row["Kernel Name"]][1] += v
    tot = sum(v[1] for v in agg.values())
    txt = ["| share | total ms | launches | avg us | kernel |", "|---:|---:|---:|---:|---|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        txt.append(f"| {100 * t / tot:.1f}% | {t / 1e3:.3f} | {c} | {t / c:.1f} | `{n[:110]}` |")
    txt.append(f"\ntotal {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")
    s = "\n".join(txt)
    print(s)
    if out:
        open(out, "a").write(s + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:3])
