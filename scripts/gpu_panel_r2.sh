#!/bin/bash
# hub-blocked panel kernel: parity tests, plugin-route tests, configs C1-C5 (round 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "batched or sweep or panel_max or custom_absorption" 2>&1 | tail -15 | tee gpurun_out/panel_tests.log
timeout 600 python -m pytest tests/test_plugin_dropin.py tests/test_hsell_gpu.py -x -q 2>&1 | tail -8 | tee gpurun_out/panel_tests2.log
timeout 900 python scripts/run_configs.py --tag r2b > gpurun_out/configs_r2b.log 2>&1; tail -5 gpurun_out/configs_r2b.log
