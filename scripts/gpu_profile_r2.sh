#!/bin/bash
# ncu evidence for round 2: launch list of the default bench command + full captures of the step kernels (single-vector
# fp32 / fp64 and the panel path).  The .ncu-rep files are exported to CSV on the box and removed (gpurun_out/ must stay
# under 64 MiB); scripts/summarize_profile_r2.py reads the CSVs.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-plugin --no-panel > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
capture() {   # name, kernel regex, command...
  local name=$1 regex=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s 6 -c 1 -o gpurun_out/prof_${TAG}_$name -f "$@" > gpurun_out/ncu_full_${TAG}_$name.log 2>&1; echo "$name rc=$?"
  ncu -i gpurun_out/prof_${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$name.raw.csv 2>/dev/null
  python scripts/ncu_hot.py gpurun_out/prof_${TAG}_$name.ncu-rep 14 > gpurun_out/prof_${TAG}_$name.hot.txt 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_$name.ncu-rep
}
capture gather hsell_gather_kernel python bench.py --kernel-only
capture update hsell_update_accum_kernel python bench.py --kernel-only
capture gather_f64 hsell_gather_kernel python bench.py --kernel-only --dtype f64
capture update_f64 hsell_update_accum_kernel python bench.py --kernel-only --dtype f64
export PGB_PANEL=1 PROBE_REPS=0
capture panel_gather hsell_gather_kernel python scripts/panel_probe.py
capture panel_update hsell_update_panel_kernel python scripts/panel_probe.py
du -sh gpurun_out
