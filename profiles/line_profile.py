#!/usr/bin/env python
"""Executed warp-instructions per SOURCE LINE of one kernel.

  cuobjdump -xelf all build/X.o ; nvdisasm -g -c X.sm_100a.cubin > X.asm      (line info: compile with -lineinfo)
  ncu -i R.ncu-rep --page source --csv > R_src.csv                              (one kernel launch)
  python profiles/line_profile.py R_src.csv X.asm <substring of the kernel's mangled name> [warps] [top]
"""
import collections
import csv
import re
import sys


def main(src_csv, asm, kernel_sub, warps=None, top=40):
    # address offset -> (file, line) from nvdisasm -g
    line_of, cur, inside = {}, None, False
    for ln in open(asm):
        if ln.startswith("\t.section\t.text."):
            inside = kernel_sub in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    ia, ie, iaddr, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Address"), hdr.index("# Samples")
    first, per_line, samples, tot = None, collections.Counter(), collections.Counter(), 0
    for r in rows[hi + 1:]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        addr = int(r[iaddr], 16)
        if first is None:
            first = addr
            w = float(warps or int(r[ie]))
        key = line_of.get(addr - first, ("?", 0))
        per_line[key] += int(r[ie])
        samples[key] += int(r[isamp])
        tot += int(r[ie])
    ts = float(sum(samples.values()) or 1)
    print("total %.1f warp-instructions per warp (%d warps)" % (tot / w, w))
    for key, n in per_line.most_common(int(top)):
        print("%8.1f  %5.1f%% smp  %s:%d" % (n / w, 100 * samples[key] / ts, key[0], key[1]))


if __name__ == "__main__":
    main(*sys.argv[1:])
