#!/bin/bash
# tools/gpu_job_scale.sh LABEL -- on a 4-GPU box: the driver-style torchrun launch of bench.py at N = 2 and 4 (+ N = 1), and
# the reference arm under torchrun (rank 0 prints, the others exit 0).
L=${1:-x}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/scale1_$L.json 2> gpurun_out/scale1_$L.err
timeout 600 $TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/scale2_$L.json 2> gpurun_out/scale2_$L.err
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/scale4_$L.json 2> gpurun_out/scale4_$L.err
timeout 600 $TR --nproc-per-node 2 --master-port 29523 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/scale_ref2_$L.json 2> gpurun_out/scale_ref2_$L.err; echo "ref rc=$?"
for n in 1 2 4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale${n}_$L.json").read().strip().splitlines()[-1])
    print($n, round(d["value"]), d["ms_per_step"], d["per_rank_ms_per_step"], d.get("exchange"), d["cpu_baseline"] is None)
except Exception as e:
    print($n, "failed", e); print(open("gpurun_out/scale${n}_$L.err").read()[-1500:])
PY
done
cut -c1-200 gpurun_out/scale_ref2_$L.json
