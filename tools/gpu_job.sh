#!/bin/bash
# tools/gpu_job.sh LABEL [tests|notests] [ncu-kernel-regex] -- one gpurun call: GPU tests, default bench, optional ncu capture.
# Everything lands in gpurun_out/ with the LABEL as suffix.  (Development helper, not product code.)
L=${1:-x}; T=${2:-tests}; K=${3:-}
mkdir -p gpurun_out
if [ "$T" = "tests" ]; then
  (time python -m pytest tests -m gpu -x -q) > gpurun_out/gputest_$L.log 2>&1
  tail -4 gpurun_out/gputest_$L.log
fi
python bench.py --no-cpu > gpurun_out/bench_$L.json 2> gpurun_out/bench_$L.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$L.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "clips/s", round(d["ms_per_step"],4), "ms/step; K1", round(d["roofline"]["ms_per_launch"],4), "rest", round(d["roofline"]["rest_of_step_ms"],4), "e2e", round(d["e2e"]["value"]), "sustained", d["sustained"]["ms_per_step"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$L.err").read()[-2000:])
PY
if [ -n "$K" ]; then
  ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 4 -c 1 -f -o gpurun_out/ncu_${L} python bench.py --no-cpu --steps 2 --warmup 3 --preroll 0 --sustain 0 > gpurun_out/ncu_$L.log 2>&1
  tail -2 gpurun_out/ncu_$L.log
fi
