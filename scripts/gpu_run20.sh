#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env $1 $2 $3 timeout 300 python bench.py --kernel-only --steps 1 --warmup 1 $EXTRA 2>&1 | tail -1 | cut -c1-400; }
echo "== hsell tests"; timeout 900 python -m pytest tests/test_hsell_gpu.py -x -q -m gpu 2>&1 | tail -3
run PGB_X=0
run PGB_HSELL_DEBUG_SKIP=3
run PGB_HSELL_MIN_ENTRIES=32
run PGB_HSELL_MIN_ENTRIES=48 PGB_HSELL_TAIL_WARPS=4
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
