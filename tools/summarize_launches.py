#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of GPU time)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    print("%-34s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("%-34s %6d %12.1f %10.1f %6.1f%%" % (k[:34], cnt[k], v, v / cnt[k], 100 * v / T))
    print("%-34s %6d %12.1f" % ("TOTAL", sum(cnt.values()), T))


if __name__ == "__main__":
    main(sys.argv[1])
