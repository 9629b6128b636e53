#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env $1 $2 $3 timeout 300 python bench.py --kernel-only --steps 1 --warmup 1 $EXTRA 2>&1 | tail -1 | cut -c1-400; }
run PGB_X=0
run PGB_HSELL_DEBUG_SKIP=3
run PGB_HSELL_TAIL_WARPS=4
run PGB_HSELL_TAIL_WARPS=6
EXTRA="--dtype f64" run PGB_X=0
EXTRA="--scale 22" run PGB_X=0
EXTRA="--scale 20" run PGB_X=0
EXTRA="--scale 20" run PGB_KERNEL_VARIANT=3
EXTRA="--scale 18" run PGB_X=0
EXTRA="--scale 18" run PGB_KERNEL_VARIANT=3
