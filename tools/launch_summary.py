#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share.
    python tools/launch_summary.py gpurun_out/launches.csv [--md]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    md = "--md" in sys.argv
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        name = re.sub(r"<unnamed>::", "", re.sub(r"\(.*", "", row["Kernel Name"]))
        name = re.sub(r"^void ", "", name)[:80]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(("total GPU time %.2f ms over %d launches" % (T / 1e3, sum(cnt.values()))))
    if md:
        print("\n| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        if md:
            print("| `%s` | %d | %.2f | %.1f | %.1f%% |" % (k, cnt[k], v / 1e3, v / cnt[k], 100 * v / T))
        else:
            print("%-80s n=%5d total %9.2f ms avg %8.1f us share %5.1f%%" % (k, cnt[k], v / 1e3, v / cnt[k], 100 * v / T))


if __name__ == "__main__":
    main()
