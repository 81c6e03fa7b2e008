#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (times are cold-cache and
serialised under ncu: compare SHARES, not absolutes)."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        agg.setdefault(row["Kernel Name"][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("%-72s %5s %12s %12s %6s" % ("kernel", "n", "avg_us", "total_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-72s %5d %12.1f %12.1f %5.1f%%" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main(sys.argv[1])
