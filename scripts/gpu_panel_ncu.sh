#!/bin/bash
# ncu --set full captures of the two kernels of the hub-blocked panel step (RMAT-24, 4 x fp32 columns)
mkdir -p gpurun_out
export PGB_PANEL=1 PROBE_REPS=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hsell_gather_kernel -s 8 -c 1 -o gpurun_out/prof_r2_panel_gather -f python scripts/panel_probe.py > gpurun_out/ncu_panel_gather.log 2>&1; echo "gather rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hsell_update_panel_kernel -s 8 -c 1 -o gpurun_out/prof_r2_panel_update -f python scripts/panel_probe.py > gpurun_out/ncu_panel_update.log 2>&1; echo "update rc=$?"
ls -la gpurun_out/*.ncu-rep
