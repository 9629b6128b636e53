#!/bin/bash
# run with: gpurun --gpus 8 -- bash scripts/gpu_run_scale.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus8.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== bench 8 gpus (scale 27)"; timeout 900 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_8gpu.log | cut -c1-2500
echo "== bench 4 gpus (scale 26)"; timeout 600 $TR --nproc-per-node 4 --master-port 29552 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_4gpu.log | cut -c1-2500
