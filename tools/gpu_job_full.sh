#!/bin/bash
# tools/gpu_job_full.sh LABEL -- GPU tests, smoke(), the default bench (with the cpu_baseline leg), the reference arm (short),
# bench_configs, the ncu launch list of the bench and one `ncu --set full` capture of the dominant kernel.
L=${1:-x}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/gputest_$L.log 2>&1; tail -4 gpurun_out/gputest_$L.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$L.log 2>&1; tail -1 gpurun_out/smoke_$L.log
python bench.py > gpurun_out/bench_$L.json 2> gpurun_out/bench_$L.err; cut -c1-200 gpurun_out/bench_$L.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$L.json 2> gpurun_out/bench_ref_$L.err; cut -c1-200 gpurun_out/bench_ref_$L.json
python bench_configs.py --no-cpu > gpurun_out/bench_configs_$L.jsonl 2> gpurun_out/bench_configs_$L.err; cut -c1-260 gpurun_out/bench_configs_$L.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$L.csv python bench.py --no-cpu --steps 2 --warmup 3 --preroll 0 --sustain 0 > gpurun_out/launches_$L.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectral_warp --launch-skip 4 -c 1 -f -o gpurun_out/ncu_$L python bench.py --no-cpu --steps 2 --warmup 3 --preroll 0 --sustain 0 > gpurun_out/ncu_$L.log 2>&1; tail -1 gpurun_out/ncu_$L.log
