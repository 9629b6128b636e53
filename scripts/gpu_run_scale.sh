#!/bin/bash
# run with: gpurun --gpus 8 -- bash scripts/gpu_run_scale.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus8.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('N=%d value=%.0f e2e=%.0f ms_per_solve=%.2f iters/solve=%.1f kernel_ms=%.3f build=%.1f'%(d['n_gpus'],d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['conv_calls_per_solve'],d['roofline']['kernel_ms'],d['config']['graph_build_s']))
except Exception as e: print('parse error', e)
"; }
echo "== bench 8 gpus (scale 27)"; timeout 600 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.log 2>&1; echo rc=$?; show gpurun_out/bench_8gpu.log
echo "== bench 4 gpus (scale 26)"; timeout 600 $TR --nproc-per-node 4 --master-port 29553 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_4gpu.log 2>&1; echo rc=$?; show gpurun_out/bench_4gpu.log
