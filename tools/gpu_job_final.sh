#!/bin/bash
# tools/gpu_job_final.sh LABEL -- what the driver runs at round end (GPU tests, smoke, default bench) + ncu of the LUFS kernel
L=${1:-x}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/gputest_$L.log 2>&1; tail -4 gpurun_out/gputest_$L.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$L.log 2>&1; tail -1 gpurun_out/smoke_$L.log
python bench.py > gpurun_out/bench_$L.json 2> gpurun_out/bench_$L.err; cut -c1-200 gpurun_out/bench_$L.json
ncu --set full --clock-control none --import-source on -k regex:kweight_energy_warp --launch-skip 4 -c 1 -f -o gpurun_out/ncu_$L python bench.py --no-cpu --steps 2 --warmup 3 --preroll 0 --sustain 0 > gpurun_out/ncu_$L.log 2>&1; tail -1 gpurun_out/ncu_$L.log
