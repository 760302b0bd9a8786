#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an `ncu --set full --import-source on` report:
   ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --print-source cuda,sass > k.csv ; python profiles/bylines.py k.csv [top]"""
import csv
import sys
from collections import defaultdict


def main(path, top=50):
    rows = list(csv.reader(open(path)))
    fileName = None
    hdr = None
    lastLine = -1
    per = defaultdict(lambda: [0, 0, 0, ""])  # (file, line) -> warp instr, samples, thread instr, text
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fileName = r[1].split("/")[-1]; hdr = None; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iT = hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != "":
            lastLine = int(r[0]); per[(fileName, lastLine)][3] = r[1].strip()
        key = (fileName, lastLine)
        try:
            per[key][0] += int(r[iI] or 0); per[key][1] += int(r[iS] or 0); per[key][2] += int(r[iT] or 0)
        except ValueError:
            continue
    tI = sum(v[0] for v in per.values()); tS = sum(v[1] for v in per.values())
    print(f"total warp instructions {tI}, samples {tS}")
    byfile = defaultdict(lambda: [0, 0])
    for (f, l), v in per.items():
        byfile[f][0] += v[0]; byfile[f][1] += v[1]
    for f, v in byfile.items():
        print(f"  {f:28s} instr {100 * v[0] / tI:5.1f}%  samples {100 * v[1] / tS:5.1f}%")
    keys = sorted(per, key=lambda k: -per[k][0])[:top]
    for k in sorted(keys):
        v = per[k]
        print(f"{k[0][:18]:18s} L{k[1]:4d} instr {100 * v[0] / tI:5.2f}% samp {100 * v[1] / tS:5.2f}% thr/instr {v[2] / max(1, v[0]):4.1f}  {v[3][:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50)
