#!/bin/bash
mkdir -p gpurun_out
echo "== hsell tests"; timeout 900 python -m pytest tests/test_hsell_gpu.py -x -q -m gpu > gpurun_out/tests_hsell.log 2>&1; echo rc=$?; tail -25 gpurun_out/tests_hsell.log
echo "== bench s22"; timeout 600 python bench.py --steps 3 --warmup 3 --scale 20 --no-cpu > gpurun_out/bench_hsell_s20.log 2>&1; echo rc=$?; tail -3 gpurun_out/bench_hsell_s20.log | cut -c1-2500
echo "== bench s24"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_hsell_s24.log 2>&1; echo rc=$?; tail -3 gpurun_out/bench_hsell_s24.log | cut -c1-2500
