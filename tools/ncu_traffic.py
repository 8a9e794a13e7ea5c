#!/usr/bin/env python
"""DRAM traffic of the KD-walk kernel per launch, from an ncu metrics CSV
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_walk --csv`).
Writes profiles/walk_traffic.json, which bench.py reports as roofline.traffic.
usage: ncu_traffic.py launches.csv "<command that was profiled>" [out.json]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ii, ki, ni, ui, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    per = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi or "k_walk" not in r[ki]:
            continue
        # the counting pass of bench.py (template argument COUNT = 1) renders 1 spp: its launches are not the timed ones
        if re.search(r"k_walk<(\(bool\))?[01], (\(bool\))?1,", r[ki]):
            continue
        per.setdefault(r[ii], {})[r[ni]] = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
    n = len(per)
    rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values())
    wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values())
    ms = sum(p.get("gpu__time_duration.sum", 0) for p in per.values())
    out = {"dram_bytes_per_launch": (rd + wr) / max(1, n), "dram_read_bytes_per_launch": rd / max(1, n), "dram_write_bytes_per_launch": wr / max(1, n),
           "launches": n, "avg_launch_ms_under_ncu": ms / max(1, n), "dram_gbs_under_ncu": (rd + wr) / max(1e-9, ms) / 1e6,
           "command": sys.argv[2] if len(sys.argv) > 2 else "", "source": os.path.basename(sys.argv[1]),
           "note": "the timed-size k_walk launches (closest-hit and shadow, every bounce level; the 1-spp counting pass excluded) of the profiled command; ncu serialises launches and runs them cold"}
    dst = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "walk_traffic_r2.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
