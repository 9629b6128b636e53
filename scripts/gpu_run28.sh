#!/bin/bash
mkdir -p gpurun_out
echo "== hsell tests"; timeout 900 python -m pytest tests/test_hsell_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { echo "== $*"; env $1 $2 $3 timeout 300 python bench.py --kernel-only --steps 1 --warmup 1 $EXTRA 2>&1 | tail -1 | cut -c1-400; }
run PGB_X=0
run PGB_HSELL_INLINE_TAIL=8
run PGB_HSELL_INLINE_TAIL=32
run PGB_HSELL_INLINE_TAIL=32 PGB_HSELL_BLOCKS=36
run PGB_HSELL_INLINE_TAIL=32 PGB_HSELL_BLOCKS=48
run PGB_HSELL_INLINE_TAIL=128 PGB_HSELL_BLOCKS=36
