#!/bin/bash
# run with: gpurun --gpus 2 -- bash scripts/gpu_run_dist2.sh
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== dist check scale 16 (peer, multicast if available)"; timeout 300 $TR --master-port 29541 tests/dist_gpu_check.py 16 > gpurun_out/dist_check16.log 2>&1; echo rc=$?; grep -E "dtype=|oracle|DIST CHECK|Error|error|peer" gpurun_out/dist_check16.log | head -20
echo "== dist check scale 16 (peer, unicast stores)"; PGB_PEER_MULTICAST=0 timeout 300 $TR --master-port 29542 tests/dist_gpu_check.py 16 > gpurun_out/dist_check16u.log 2>&1; echo rc=$?; grep -E "dtype=|oracle|DIST CHECK|Error|error|peer" gpurun_out/dist_check16u.log | head -20
echo "== dist check scale 20"; timeout 300 $TR --master-port 29543 tests/dist_gpu_check.py 20 > gpurun_out/dist_check20.log 2>&1; echo rc=$?; grep -E "dtype=|DIST CHECK|Error|error|peer" gpurun_out/dist_check20.log | head -20
for mode in "PGB_X=1" "PGB_PEER_MULTICAST=0" "PGB_PEER=0"; do
echo "== bench 2 gpus (scale 25) $mode"; env $mode timeout 600 $TR --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_$mode.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_2gpu_$mode.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('value=%.0f e2e=%.0f ms_per_solve=%.2f kernel_ms=%.3f exchange=%s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['config'].get('exchange')))
except Exception as e: print('parse error', e)
"
done
