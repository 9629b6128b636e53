#!/bin/bash
# Round-2 sweep (b): L2 prefetch distance of the hsell gather kernel, RMAT-24 fp32 kernel-only
mkdir -p gpurun_out
run() {
  name=$1; rel=$2; shift 2
  env "$@" timeout 300 python bench.py --kernel-only --relabel $rel ${DT:+--dtype $DT} ${SCALE:+--scale $SCALE} > gpurun_out/sw_$name.log 2>&1
  echo "$name rel=$rel $* :: $(tail -1 gpurun_out/sw_$name.log | cut -c1-120)"
}
run pf0 hub PGB_HSELL_PF_DIST=0 PGB_HSELL_PF_TAIL=0
run pf0t hub PGB_HSELL_PF_DIST=0 PGB_HSELL_PF_TAIL=1
run pf16 hub PGB_HSELL_PF_DIST=16
run pf32 hub PGB_HSELL_PF_DIST=32
run pf64 hub PGB_HSELL_PF_DIST=64
run pf128 hub PGB_HSELL_PF_DIST=128
run pf32_nt hub PGB_HSELL_PF_DIST=32 PGB_HSELL_PF_TAIL=0
run pf32_deg degree PGB_HSELL_PF_DIST=32
run pf32_k96 hub PGB_HSELL_PF_DIST=32 PGB_HSELL_BLOCKS=96
run pf32_k128 hub PGB_HSELL_PF_DIST=32 PGB_HSELL_BLOCKS=128
run pf32_tw4 hub PGB_HSELL_PF_DIST=32 PGB_HSELL_TAIL_WARPS=4
run pf32_tw8 hub PGB_HSELL_PF_DIST=32 PGB_HSELL_TAIL_WARPS=8
run pf32_tw12 hub PGB_HSELL_PF_DIST=32 PGB_HSELL_TAIL_WARPS=12
run pf32_skiphub hub PGB_HSELL_PF_DIST=32 PGB_HSELL_DEBUG_SKIP=1
run pf32_skiptail hub PGB_HSELL_PF_DIST=32 PGB_HSELL_DEBUG_SKIP=2
run pf32_skipboth hub PGB_HSELL_PF_DIST=32 PGB_HSELL_DEBUG_SKIP=3
DT=f64 run pf32_f64 hub PGB_HSELL_PF_DIST=32
