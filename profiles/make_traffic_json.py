#!/usr/bin/env python
"""profiles/*_ncu_full_summary.csv -> profiles/spectral_traffic.json: the DRAM bytes ONE launch of the dominant kernel
moves (dram__bytes_read.sum + dram__bytes_write.sum of an `ncu --set full` capture at batch 64), keyed by kernel name.
bench.py quotes it as `roofline.traffic` with the capture it came from (a number printed under ncu is never a bench
value; DRAM byte counts do not depend on the replay).

    python profiles/make_traffic_json.py spectral_warp_kernel=profiles/r02k_prof_spectral_f32x2_ncu_full_summary.csv ...
"""
import csv
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    out = {}
    for arg in sys.argv[1:]:
        name, path = arg.split("=")
        vals = {}
        for row in csv.reader(open(path)):
            if len(row) == 3 and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                vals[row[0]] = float(row[2]) * UNIT[row[1]]
        out[name] = {"dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"],
                     "batch": 64, "source": f"ncu --set full --clock-control none, {os.path.relpath(path, os.path.dirname(here))}"}
    json.dump(out, open(os.path.join(here, "spectral_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
