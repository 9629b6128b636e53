#!/bin/bash
# ncu evidence: launch list of the default bench command + full captures of the hsell gather and update kernels
TAG=${1:-r1b}
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hsell_gather_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_gather -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_${TAG}_gather.log 2>&1; echo "gather rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hsell_update_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_update -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_${TAG}_update.log 2>&1; echo "update rc=$?"
