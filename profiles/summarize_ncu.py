#!/usr/bin/env python
"""ncu report -> compact (metric,unit,value) CSV per kernel:  python profiles/summarize_ncu.py rep.ncu-rep regex out.csv
Keeps the metrics the roofline / stall analysis in DESIGN.md and profiles/README.md quotes."""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration|dram__bytes|dram__throughput|sm__throughput|sm__cycles|smsp__cycles_active|"
                  r"smsp__inst_executed\.sum|smsp__issue_active|sm__warps_active|sm__inst_executed_pipe|"
                  r"smsp__average_warps_issue_stalled|l1tex__data_bank_conflicts|l1tex__data_pipe_lsu_wavefronts|"
                  r"l1tex__t_(sectors|requests)_pipe_lsu_mem_global|lts__t_sectors_srcunit_tex|launch__|"
                  r"smsp__cycles_elapsed\.avg\.per_second|sm__pipe_fma|smsp__thread_inst_executed)")


def main():
    rep, pat, out = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if re.search(pat, r[ki]):
            with open(out, "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["metric", "unit", "value"])
                w.writerow(["kernel", "", r[ki]])
                for h, u, v in zip(hdr, units, r):
                    if KEEP.match(h):
                        w.writerow([h, u, v])
            return
    sys.exit(f"no kernel matching {pat!r}")


if __name__ == "__main__":
    main()
