#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r1g}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hsell_update_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_update -f python bench.py --kernel-only --steps 1 --warmup 1 > gpurun_out/ncu_full_${TAG}_update.log 2>&1; echo "update rc=$?"
