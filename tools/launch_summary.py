#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: time, launches, share."""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    ni = hdr.index("Metric Name") if "Metric Name" in hdr else -1
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= mi or (ni >= 0 and r[ni] != "gpu__time_duration.sum"):  # the list may carry other metrics per launch too
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("hxr::dev::", "")
        try:
            t = float(r[mi].replace(",", ""))
        except ValueError:
            continue
        t *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(r[ui], 1.0)
        agg.setdefault(name, [0.0, 0])
        agg[name][0] += t
        agg[name][1] += 1
    tot = sum(v[0] for v in agg.values())
    print("total %.3f ms over %d launches" % (tot, sum(v[1] for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("%-44s %9.3f ms %5d launches %5.1f%%" % (k[:44], v[0], v[1], 100 * v[0] / tot))


if __name__ == "__main__":
    main()
