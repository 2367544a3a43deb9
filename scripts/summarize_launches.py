"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel
summary (launches, total device time, share of the profiled run).

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    tot = defaultdict(lambda: [0, 0.0])
    with open(path, newline="") as f:
        rows = [r for r in f if r.startswith('"')]
    rd = csv.DictReader(rows)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") in ("us", "usecond"):
            ns *= 1e3
        elif r.get("Metric Unit") in ("ms", "msecond"):
            ns *= 1e6
        t = tot[name]
        t[0] += 1
        t[1] += ns
    total = sum(v[1] for v in tot.values())
    print("# source: %s" % path)
    print("# %d launches, %.3f ms total device time (ncu: serialised, cold cache; compare shares)" % (
        sum(v[0] for v in tot.values()), total / 1e6))
    print("%-48s %9s %14s %12s %7s" % ("kernel", "launches", "total_ms", "avg_us", "share"))
    for name, (cnt, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-48s %9d %14.3f %12.2f %6.1f%%" % (name[:48], cnt, ns / 1e6, ns / cnt / 1e3, 100.0 * ns / total))


if __name__ == "__main__":
    main(sys.argv[1])
