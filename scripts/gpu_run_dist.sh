#!/bin/bash
# run with: gpurun --gpus 2 -- bash scripts/gpu_run_dist.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== dist check scale 16"; timeout 600 $TR --master-port 29541 tests/dist_gpu_check.py 16 > gpurun_out/dist_check16.log 2>&1; echo rc=$?; grep -E "dtype=|oracle|DIST CHECK|Error|error" gpurun_out/dist_check16.log | head -20
echo "== dist check scale 16, tiny hub blocks + tail"; PGB_HSELL_BLOCK_COLS=1024 PGB_HSELL_BLOCKS=6 PGB_HSELL_MIN_ENTRIES=4 timeout 600 $TR --master-port 29545 tests/dist_gpu_check.py 16 > gpurun_out/dist_check16b.log 2>&1; echo rc=$?; grep -E "dtype=|oracle|DIST CHECK|Error|error" gpurun_out/dist_check16b.log | head -20
echo "== dist check scale 20"; timeout 600 $TR --master-port 29542 tests/dist_gpu_check.py 20 > gpurun_out/dist_check20.log 2>&1; echo rc=$?; grep -E "dtype=|DIST CHECK|Error|error" gpurun_out/dist_check20.log | head -20
echo "== bench 2 gpus (scale 25)"; timeout 900 $TR --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo rc=$?; tail -2 gpurun_out/bench_2gpu.log | cut -c1-2500
