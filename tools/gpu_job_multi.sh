#!/bin/bash
# tools/gpu_job_multi.sh LABEL -- one `gpurun --gpus 8` call: peer-exchange test on 2 GPUs, cfg4 on 4, cfg5 on 8, bench.py on 8.
L=${1:-x}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 600 python -m pytest tests/test_peer_exchange_gpu.py -x -q -m gpu) > gpurun_out/peer_test_$L.log 2>&1; tail -2 gpurun_out/peer_test_$L.log
timeout 900 $TR --nproc-per-node 4 --master-port 29511 bench_configs.py --gpus 4 --only cfg4 > gpurun_out/cfg4_4gpu_$L.jsonl 2> gpurun_out/cfg4_4gpu_$L.err; cat gpurun_out/cfg4_4gpu_$L.jsonl | cut -c1-400
timeout 900 $TR --nproc-per-node 8 --master-port 29512 bench_configs.py --gpus 8 --only cfg5 > gpurun_out/cfg5_8gpu_$L.jsonl 2> gpurun_out/cfg5_8gpu_$L.err; cat gpurun_out/cfg5_8gpu_$L.jsonl | cut -c1-400
timeout 900 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_8gpu_$L.json 2> gpurun_out/bench_8gpu_$L.err; cut -c1-300 gpurun_out/bench_8gpu_$L.json
