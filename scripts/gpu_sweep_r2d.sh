#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; rel=$2; shift 2
  env "$@" timeout 300 python bench.py --kernel-only --relabel $rel ${DT:+--dtype $DT} ${SCALE:+--scale $SCALE} > gpurun_out/sw_$name.log 2>&1
  echo "$name rel=$rel $* :: $(tail -1 gpurun_out/sw_$name.log | cut -c1-120)"
}
run d_base hub PGB_HSELL_TAIL_WARPS=4
run d_upd2 hub PGB_HSELL_TAIL_WARPS=4 PGB_HSELL_UPD_GROUP=2
run d_tw2 hub PGB_HSELL_TAIL_WARPS=2
run d_tw3 hub PGB_HSELL_TAIL_WARPS=3
run d_tw5 hub PGB_HSELL_TAIL_WARPS=5
run d_l2win hub PGB_HSELL_TAIL_WARPS=4 PGB_HSELL_L2WIN=1
run d_l2win_notex hub PGB_HSELL_TAIL_WARPS=6 PGB_HSELL_L2WIN=1 PGB_HSELL_TEX=0
run d_b40k hub PGB_HSELL_TAIL_WARPS=4 PGB_HSELL_BLOCK_COLS=40960
run d_b48k hub PGB_HSELL_TAIL_WARPS=4 PGB_HSELL_BLOCK_COLS=49152
run d_b24k hub PGB_HSELL_TAIL_WARPS=4 PGB_HSELL_BLOCK_COLS=24576
run d_deg_tw4 degree PGB_HSELL_TAIL_WARPS=4
run d_nobank hub PGB_HSELL_TAIL_WARPS=4 PGB_HSELL_BANK_ORDER=0
SCALE=22 run d_s22 hub PGB_HSELL_TAIL_WARPS=4
SCALE=22 run d_s22_notex hub PGB_HSELL_TAIL_WARPS=6 PGB_HSELL_TEX=0
SCALE=25 run d_s25 hub PGB_HSELL_TAIL_WARPS=4
SCALE=25 run d_s25_notex degree PGB_HSELL_TAIL_WARPS=6 PGB_HSELL_TEX=0
