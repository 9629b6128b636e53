#!/bin/bash
# hub-blocked panel kernel: parity tests, probe, configs C1-C5 (round 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "batched or sweep or panel_max or custom_absorption" 2>&1 | tail -15 | tee gpurun_out/panel_tests.log
PGB_PANEL=1 timeout 300 python scripts/panel_probe.py 2>&1 | tail -1 | cut -c1-400
timeout 900 python scripts/run_configs.py --tag r2b > gpurun_out/configs_r2b.log 2>&1; grep -E "^C3|^C5" gpurun_out/configs_r2b.log
