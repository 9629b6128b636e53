#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r1f}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hsell_(gather|reduce|update)" -c 60 --csv --log-file gpurun_out/launches_hsell.csv python bench.py --kernel-only --steps 1 --warmup 1 > gpurun_out/ncu_list_hsell.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hsell_update_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_update -f python bench.py --kernel-only --steps 1 --warmup 1 > gpurun_out/ncu_full_${TAG}_update.log 2>&1; echo "update rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hsell_gather_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_gather -f python bench.py --kernel-only --steps 1 --warmup 1 > gpurun_out/ncu_full_${TAG}_gather.log 2>&1; echo "gather rc=$?"
