#!/usr/bin/env python
"""traffic_json.py out.json key=file.csv:parcels ... -- DRAM bytes per launch of the stage kernels from
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` captures (one launch of each kernel),
in the form bench.py reads for `roofline.traffic` (key = "<workload>:<gas>")."""
import csv
import json
import sys

STAGE = {"moveKernel": "move", "gatherKernel": "sort", "collideLaneKernel": "collide", "sampleKernel": "sample"}


def one(path, parcels):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    out = {}
    per = {}
    for r in rows:
        per.setdefault((r[0], r[4]), {})[r[12]] = float(r[14])
    for (_, kernel), m in per.items():
        for k, stage in STAGE.items():
            if k in kernel and stage not in out and "dram__bytes_read.sum" in m:
                b = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
                out[stage] = {"parcels": parcels, "dram_bytes": b, "kernel": kernel.split("(")[0], "ns": m.get("gpu__time_duration.sum"),
                              "bytes_per_parcel": b / parcels}
    return out


def main():
    dst = sys.argv[1]
    res = {"capture": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, one launch of each "
                      "stage kernel in the 4th step of the default bench.py workloads (profiles/r02_traffic_*.csv), final build of round 2"}
    for spec in sys.argv[2:]:
        key, rest = spec.split("=")
        path, parcels = rest.rsplit(":", 1)
        res[key] = one(path, int(parcels))
    json.dump(res, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main()
