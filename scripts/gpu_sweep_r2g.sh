#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; rel=$2; shift 2
  env "$@" timeout 300 python bench.py --kernel-only --relabel $rel ${DT:+--dtype $DT} ${SCALE:+--scale $SCALE} > gpurun_out/sw_$name.log 2>&1
  echo "$name rel=$rel $* :: $(tail -1 gpurun_out/sw_$name.log | cut -c1-100)"
}
run g_base hub X=1
run g_base2 hub X=1
run g_min16 hub PGB_HSELL_MIN_ENTRIES=16
run g_min24 hub PGB_HSELL_MIN_ENTRIES=24
run g_min48 hub PGB_HSELL_MIN_ENTRIES=48
run g_tw4 hub PGB_HSELL_TAIL_WARPS=4
run g_tw6 hub PGB_HSELL_TAIL_WARPS=6
run g_tw7 hub PGB_HSELL_TAIL_WARPS=7
run g_b40k hub PGB_HSELL_BLOCK_COLS=40960
run g_b36k hub PGB_HSELL_BLOCK_COLS=36864
run g_k80 hub PGB_HSELL_BLOCKS=80
run g_k48 hub PGB_HSELL_BLOCKS=48
run g_win2 hub PGB_HSELL_TAIL_WINDOWS=2
run g_win4 hub PGB_HSELL_TAIL_WINDOWS=4
run g_skipboth hub PGB_HSELL_DEBUG_SKIP=3
DT=f64 run g_f64 hub X=1
DT=f64 run g_f64_win2 hub PGB_HSELL_TAIL_WINDOWS=2
