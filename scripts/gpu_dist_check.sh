#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== dist check scale 16"; timeout 300 $TR --master-port 29541 tests/dist_gpu_check.py 16 > gpurun_out/dist_check16.log 2>&1; echo rc=$?; grep -E "dtype=|oracle|DIST CHECK|Error|error|peer" gpurun_out/dist_check16.log | head -20
echo "== dist check scale 20"; timeout 300 $TR --master-port 29543 tests/dist_gpu_check.py 20 > gpurun_out/dist_check20.log 2>&1; echo rc=$?; grep -E "dtype=|DIST CHECK|Error|error|peer" gpurun_out/dist_check20.log | head -20
echo "== bench 2 gpus (scale 25)"; timeout 600 $TR --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_2gpu.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value=%.0f e2e=%.0f ms_per_solve=%.2f iters=%.1f kernel_ms=%.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['conv_calls_per_solve'],d['roofline']['kernel_ms']))"
