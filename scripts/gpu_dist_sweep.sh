#!/bin/bash
# kernel_ms / GTEPS of the N-GPU bench under hsell shape knobs: scripts/gpu_dist_sweep.sh <world>
W=${1:-4}
mkdir -p gpurun_out
port=29600
run() {
  name=$1; shift
  port=$((port+1))
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $port bench.py --gpus $W --steps 2 --warmup 2 --no-parity 2>/dev/null | tail -1 > gpurun_out/dsw_${W}_$name.json
  python -c "
import json
d=json.loads(open('gpurun_out/dsw_${W}_$name.json').read())
print('$name $*', 'value=%.0f kernel_ms=%.3f ms_per_iter=%.3f'%(d['value'], d['roofline']['kernel_ms'], d['ms_per_step']/d['config']['conv_calls_per_solve']))"
}
run base X=1
run cap512 PGB_HSELL_BLOCKS_CAP=512
run cap128 PGB_HSELL_BLOCKS_CAP=128
run wmin8 PGB_HSELL_TAIL_WINDOW_MIN=8
run wmin32 PGB_HSELL_TAIL_WINDOW_MIN=32
run tw8 PGB_HSELL_TAIL_WARPS=8
run min16 PGB_HSELL_MIN_ENTRIES=16
