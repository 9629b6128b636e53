#!/bin/bash
# round-1 evidence run: default bench line, BASELINE configs, ncu launch list / step list / full captures
TAG=${1:-r1}
mkdir -p gpurun_out
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_default.log | cut -c1-3000
echo "== configs"; timeout 1500 python scripts/run_configs.py --tag $TAG > gpurun_out/configs_$TAG.log 2>&1; echo rc=$?; tail -6 gpurun_out/configs_$TAG.log | cut -c1-1200
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hsell_(gather|reduce|update)" -c 90 --csv --log-file gpurun_out/launches_${TAG}_step.csv python bench.py --kernel-only --steps 1 --warmup 1 > gpurun_out/ncu_step_$TAG.log 2>&1; echo "step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hsell_gather_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_gather -f python bench.py --kernel-only --steps 1 --warmup 1 > gpurun_out/ncu_full_${TAG}_gather.log 2>&1; echo "gather rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hsell_update_kernel -s 6 -c 1 -o gpurun_out/prof_${TAG}_update -f python bench.py --kernel-only --steps 1 --warmup 1 > gpurun_out/ncu_full_${TAG}_update.log 2>&1; echo "update rc=$?"
