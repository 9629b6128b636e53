#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; rel=$2; shift 2
  env "$@" timeout 300 python bench.py --kernel-only --relabel $rel ${DT:+--dtype $DT} ${SCALE:+--scale $SCALE} > gpurun_out/sw_$name.log 2>&1
  echo "$name rel=$rel $* :: $(tail -1 gpurun_out/sw_$name.log | cut -c1-120)"
}
run e_roll0 hub PGB_HSELL_ROLLING=0 PGB_HSELL_TAIL_WARPS=5
run e_roll1_tw4 hub PGB_HSELL_TAIL_WARPS=4
run e_roll1_tw5 hub PGB_HSELL_TAIL_WARPS=5
run e_roll1_tw6 hub PGB_HSELL_TAIL_WARPS=6
run e_roll1_tw8 hub PGB_HSELL_TAIL_WARPS=8
run e_roll1_tw10 hub PGB_HSELL_TAIL_WARPS=10
run e_roll1_upd2 hub PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_UPD_GROUP=2
run e_roll1_skiptail hub PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_DEBUG_SKIP=2
run e_roll0_skiptail hub PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_DEBUG_SKIP=2 PGB_HSELL_ROLLING=0
run e_roll1_b40k hub PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_BLOCK_COLS=40960
run e_roll1_k96 hub PGB_HSELL_TAIL_WARPS=5 PGB_HSELL_BLOCKS=96
run e_roll1_deg degree PGB_HSELL_TAIL_WARPS=5
DT=f64 run e_roll0_f64 hub PGB_HSELL_ROLLING=0
DT=f64 run e_roll1_f64 hub X=1
