#!/bin/bash
mkdir -p gpurun_out
echo "== batched tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "batched or propagate" > gpurun_out/tests_batched.log 2>&1; echo rc=$?; tail -15 gpurun_out/tests_batched.log
