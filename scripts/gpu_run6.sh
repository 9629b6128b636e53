#!/bin/bash
mkdir -p gpurun_out
echo "== ncu full v3 mb3 (scale 24)"; PGB_LIB=$PWD/pygrank_b200/lib/libpgb200_mb3.so timeout 1200 ncu --set full --clock-control none --import-source on -k regex:item_stream_kernel -s 12 -c 1 -o gpurun_out/prof_r1_v3 -f python bench.py --scale 24 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_v3.log 2>&1; echo "rc=$?"
echo "== ncu probe"; timeout 1200 ncu --set full --clock-control none -k regex:gather_probe -s 2 -c 1 -o gpurun_out/prof_r1_probe -f python bench.py --scale 24 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_probe.log 2>&1; echo "rc=$?"
