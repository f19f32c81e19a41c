#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_xxx.txt"""
import collections
import csv
import sys


def main(path):
    hdr, agg, n = None, collections.defaultdict(lambda: [0, 0.0]), 0
    for r in csv.reader(open(path)):
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(d["Metric Unit"], 1.0)
        name = d["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print("# %s: %d launches, %.1f us of kernel time (cold-cache, serialised under ncu: compare SHARES)" % (path, n, tot / 1e3))
    print("%-48s %6s %12s %7s" % ("kernel", "count", "total_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s %6d %12.1f %6.1f%%" % (k[:48], v[0], v[1] / 1e3, 100 * v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
