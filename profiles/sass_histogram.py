#!/usr/bin/env python
"""Dynamic opcode histogram of one kernel from `ncu -i X.ncu-rep --page source --csv` output:
warp-level instructions executed per warp, by opcode, plus stall samples."""
import collections
import csv
import re
import sys


def main(path, warps=None):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot, ops, samp = 0, collections.Counter(), collections.Counter()
    first = None
    for r in rows[hi + 1:]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        n, s = int(r[ie]), int(r[isamp])
        if first is None:
            first = n
        tot += n
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ia])
        op = m.group(2) if m else "?"
        ops[op] += n
        samp[op] += s
    w = float(warps or first)
    print("warp-instructions %d, warps %d, per warp %.1f" % (tot, w, tot / w))
    ts = sum(samp.values())
    for op, n in ops.most_common(40):
        print("%-10s %8.1f per warp  %5.1f%% of samples" % (op, n / w, 100.0 * samp[op] / max(ts, 1)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
