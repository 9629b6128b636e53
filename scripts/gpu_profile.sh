#!/bin/bash
# ncu evidence for profiles/: launch list of the default bench command + one full capture of the top kernel + the probe
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:item_stream_kernel -s 12 -c 1 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1; echo "full rc=$?"
timeout 1200 ncu --set full --clock-control none -k regex:gather_probe -s 2 -c 1 -o gpurun_out/prof_${TAG}_probe -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_probe_$TAG.log 2>&1; echo "probe rc=$?"
