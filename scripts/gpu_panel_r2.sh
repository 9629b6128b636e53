#!/bin/bash
mkdir -p gpurun_out
for b in 1 0 1 0; do
  echo "bulk=$b f32: $(PGB_HSELL_BULK=$b timeout 300 python bench.py --kernel-only 2>&1 | tail -1 | cut -c1-60)"
done
for b in 1 0; do
  echo "bulk=$b f64: $(PGB_HSELL_BULK=$b timeout 300 python bench.py --kernel-only --dtype f64 2>&1 | tail -1 | cut -c1-60)"
  echo "bulk=$b panel: $(PGB_HSELL_BULK=$b PGB_PANEL=1 timeout 300 python scripts/panel_probe.py 2>&1 | tail -1 | python -c 'import sys,json; print(json.loads(sys.stdin.read())["ms_per_panel_step"])')"
done
timeout 900 python -m pytest tests/test_hsell_gpu.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
