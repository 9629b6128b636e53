#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; rel=$2; shift 2
  env "$@" timeout 300 python bench.py --kernel-only --relabel $rel ${DT:+--dtype $DT} ${SCALE:+--scale $SCALE} > gpurun_out/sw_$name.log 2>&1
  echo "$name rel=$rel $* :: $(tail -1 gpurun_out/sw_$name.log | cut -c1-120)"
}
run f_acc0 degree PGB_HSELL_ACCUM=0 PGB_HSELL_TAIL_WARPS=5
run f_acc1 degree PGB_HSELL_TAIL_WARPS=5
run f_acc1_hub hub PGB_HSELL_TAIL_WARPS=5
run f_acc1_tw4 degree PGB_HSELL_TAIL_WARPS=4
run f_acc1_tw6 degree PGB_HSELL_TAIL_WARPS=6
run f_acc1_tw8 degree PGB_HSELL_TAIL_WARPS=8
run f_acc1_min16 degree PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_MIN_ENTRIES=16
run f_acc1_min8_hub hub PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_MIN_ENTRIES=8
run f_acc1_rc_hub hub PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4
run f_acc1_k96_hub hub PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_BLOCKS=96
run f_acc1_k128_rc_hub hub PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_BLOCKS=128 PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4
run f_acc1_skiphub degree PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_DEBUG_SKIP=1
run f_acc1_skiptail degree PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_DEBUG_SKIP=2
run f_acc1_skipboth degree PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_DEBUG_SKIP=3
DT=f64 run f_acc0_f64 degree PGB_HSELL_ACCUM=0
DT=f64 run f_acc1_f64 degree X=1
DT=f64 run f_acc1_f64_hub hub X=1
