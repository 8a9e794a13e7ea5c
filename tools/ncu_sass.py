#!/usr/bin/env python
"""Per-region summary of the SASS hot spots of one kernel in an .ncu-rep (source page): where the issued instructions
go and how many lanes are active there. usage: ncu_sass.py report.ncu-rep [kernel-index] [--dump]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
    want = next((a[7:] for a in sys.argv if a.startswith("--name=")), None)  # --name=substring: the first kernel whose name contains it
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    kernels, cur = [], None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            cur = []
            kernels.append((r[1], cur))
        elif cur is not None:
            cur.append(r)
    if want:
        sel = [k for k in kernels if want in k[0]]
        if not sel:
            print("no kernel matching", want)
            return
        name, rs = sel[which if which < len(sel) else 0]
    else:
        name, rs = kernels[which]
    hdr, body = rs[0], rs[1:]
    iE, iT, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    rows = [(int(r[iE]), int(r[iT]), int(r[iS]), r[iSrc].strip()) for r in body]
    tot, totT, totS = sum(r[0] for r in rows), sum(r[1] for r in rows), sum(r[2] for r in rows)
    print(name[:80], "| sass", len(rows), "| warp inst", tot, "| thread inst", totT, "| avg lanes %.2f" % (totT / max(1, tot)), "| samples", totS)
    if "--dump" in sys.argv:
        for i, r in enumerate(rows):
            print("%4d %10d %5.1f %6d  %s" % (i, r[0], r[1] / max(1, r[0]), r[2], r[3][:100]))
        return
    start = 0
    for i in range(1, len(rows) + 1):
        if i == len(rows) or abs(rows[i][0] - rows[start][0]) > 0.02 * max(rows[start][0], 1):
            seg = rows[start:i]
            ex, th, sm = sum(r[0] for r in seg), sum(r[1] for r in seg), sum(r[2] for r in seg)
            if ex / tot > 0.008 or sm / max(1, totS) > 0.01:
                print("%4d-%4d n=%3d exec=%9d inst %5.1f%% lanes %4.1f samples %5.1f%%  %s" % (
                    start, i - 1, len(seg), seg[0][0], 100 * ex / tot, th / max(1, ex), 100 * sm / max(1, totS), seg[0][3][:60]))
            start = i


if __name__ == "__main__":
    main()
