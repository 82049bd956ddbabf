#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        tot[row["Kernel Name"]][0] += 1
        tot[row["Kernel Name"]][1] += v
    total = sum(t for _, t in tot.values())
    for n, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1]):
        print("%-64s n=%3d total=%9.1f us avg=%8.1f us  %4.1f%%" % (n[:64], c, t, t / c, 100 * t / total))


if __name__ == "__main__":
    main(sys.argv[1])
