#!/bin/bash
# Multi-GPU parity matrix: every exchange path of pygrank_b200/dist.py against the single-GPU engine (and the
# oracle at scale <= 18), NaN-poisoned buffers.  Usage: scripts/gpu_dist_parity.sh <world> [scales...]
W=${1:-2}; shift
SCALES=${@:-"16 22"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
port=29600
for scale in $SCALES; do
  for variant in ${VARIANTS:-"X=1" "PGB_PEER_MASK=1" "PGB_PEER_MASK=0" "PGB_PEER=0" "PGB_PEER_MULTICAST=1"}; do
    port=$((port+1))
    log=gpurun_out/distpar_w${W}_s${scale}_${variant//=/}.log
    env $variant timeout 600 $TR --master-port $port tests/dist_gpu_check.py $scale $( [ "$variant" != "X=1" ] && echo --no-shard ) > $log 2>&1
    echo "w=$W scale=$scale $variant rc=$? :: $(grep -E 'exchange:' $log | head -1) :: $(grep -E 'DIST CHECK|Error' $log | tail -1)"
    grep -E "FAIL" $log | head -5
  done
done
grep -h "DIST REPORT" gpurun_out/distpar_w${W}_*.log | sed 's/DIST REPORT //' > gpurun_out/distpar_w${W}.jsonl
