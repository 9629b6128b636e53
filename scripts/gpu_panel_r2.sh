#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hsell_gpu.py tests/test_abi.py -x -q 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "weighted300" 2>&1 | tail -4
timeout 300 python scripts/weighted_probe.py 2>&1 | tail -1 | tee gpurun_out/weighted_probe.json
timeout 600 python bench.py --kernel-only 2>&1 | tail -1 | cut -c1-600
