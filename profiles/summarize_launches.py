#!/usr/bin/env python
"""Mean device time per kernel of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main(path):
    hdr = None
    agg = collections.OrderedDict()
    for r in csv.reader(open(path)):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
                ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
            continue
        if len(r) > vi and r[0].isdigit():
            agg.setdefault(r[ki][:78], []).append(float(r[vi].replace(",", "")))
    tot = 0.0
    for k, v in agg.items():
        print("%-80s n=%-3d mean=%8.1f us" % (k, len(v), sum(v) / len(v) / 1000))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
