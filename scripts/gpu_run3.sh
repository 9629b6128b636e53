#!/bin/bash
mkdir -p gpurun_out
echo "== tests (v2 kernel)"; timeout 1800 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/tests.log | head -40
echo "== bench scale 24 v2"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench24_v2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench24_v2.log
echo "== bench scale 24 v1"; PGB_KERNEL_VARIANT=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench24_v1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench24_v1.log | cut -c1-400
echo "== bench scale 24 v2 fp64"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --dtype f64 > gpurun_out/bench24_v2_f64.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench24_v2_f64.log | cut -c1-1600
echo "== ncu full v2 (scale 24)"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:warp_tile_kernel -s 12 -c 1 -o gpurun_out/prof_r1_v2 -f python bench.py --scale 24 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_v2.log 2>&1; echo "rc=$?"
