#!/bin/bash
# Round-2 sweep (c): tail gathers through the TEX pipe, RMAT-24 fp32 kernel-only
mkdir -p gpurun_out
run() {
  name=$1; rel=$2; shift 2
  env "$@" timeout 300 python bench.py --kernel-only --relabel $rel ${DT:+--dtype $DT} ${SCALE:+--scale $SCALE} > gpurun_out/sw_$name.log 2>&1
  echo "$name rel=$rel $* :: $(tail -1 gpurun_out/sw_$name.log | cut -c1-120)"
}
run tex0 hub PGB_HSELL_TEX=0
run tex1 hub PGB_HSELL_TEX=1
run tex1_tw4 hub PGB_HSELL_TAIL_WARPS=4
run tex1_tw8 hub PGB_HSELL_TAIL_WARPS=8
run tex1_tw12 hub PGB_HSELL_TAIL_WARPS=12
run tex1_tw16 hub PGB_HSELL_TAIL_WARPS=16
run tex1_deg degree X=1
run tex1_k32 hub PGB_HSELL_BLOCKS=32
run tex1_k48 hub PGB_HSELL_BLOCKS=48
run tex1_k96 hub PGB_HSELL_BLOCKS=96
run tex1_min64 hub PGB_HSELL_MIN_ENTRIES=64
run tex1_min16 hub PGB_HSELL_MIN_ENTRIES=16
run tex1_skiphub hub PGB_HSELL_DEBUG_SKIP=1
run tex1_skiptail hub PGB_HSELL_DEBUG_SKIP=2
DT=f64 run tex0_f64 hub PGB_HSELL_TEX=0
DT=f64 run tex1_f64 hub X=1
DT=f64 run tex1_f64_tw10 hub PGB_HSELL_TAIL_WARPS=10
