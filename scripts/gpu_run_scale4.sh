#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { tail -1 $1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('N=%d value=%.0f e2e=%.0f ms_per_solve=%.2f iters/solve=%.1f kernel_ms=%.3f build=%.1f'%(d['n_gpus'],d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['conv_calls_per_solve'],d['roofline']['kernel_ms'],d['config']['graph_build_s']))
except Exception as e: print('parse error', e)
"; }
echo "== bench 8 gpus (scale 27) K auto (512)"; timeout 600 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu_K512.log 2>&1; echo rc=$?; show gpurun_out/bench_8gpu_K512.log
echo "== bench 8 gpus (scale 27) K cap 256"; PGB_HSELL_BLOCKS_CAP=256 timeout 600 $TR --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu_K256.log 2>&1; echo rc=$?; show gpurun_out/bench_8gpu_K256.log
