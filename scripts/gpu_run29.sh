#!/bin/bash
mkdir -p gpurun_out
echo "== hsell tests"; timeout 900 python -m pytest tests/test_hsell_gpu.py -x -q -m gpu 2>&1 | tail -2
run() { echo "== $*"; env $1 $2 $3 timeout 300 python bench.py --kernel-only --steps 1 --warmup 1 $EXTRA 2>&1 | tail -1 | cut -c1-400; }
run PGB_X=0
run PGB_HSELL_TAIL_WARPS=4
run PGB_HSELL_TAIL_WARPS=6
run PGB_HSELL_TAIL_WARPS=12
