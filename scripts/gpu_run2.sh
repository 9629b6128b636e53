#!/bin/bash
mkdir -p gpurun_out
echo "== microbench"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/gather_bench scripts/gather_bench.cu && timeout 300 /tmp/gather_bench > gpurun_out/gather_bench.log 2>&1; cat gpurun_out/gather_bench.log
