#!/bin/bash
# ncu evidence for round 2: launch list of the default bench command + full captures of the two step kernels
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-plugin --no-panel > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hsell_gather_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_gather -f python bench.py --kernel-only > gpurun_out/ncu_full_${TAG}_gather.log 2>&1; echo "gather rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hsell_update_accum_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_update -f python bench.py --kernel-only > gpurun_out/ncu_full_${TAG}_update.log 2>&1; echo "update rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:hsell_gather_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_gather_f64 -f python bench.py --kernel-only --dtype f64 > gpurun_out/ncu_full_${TAG}_gather_f64.log 2>&1; echo "gather f64 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:hsell_update_accum_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_update_f64 -f python bench.py --kernel-only --dtype f64 > gpurun_out/ncu_full_${TAG}_update_f64.log 2>&1; echo "update f64 rc=$?"
