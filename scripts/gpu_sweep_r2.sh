#!/bin/bash
# Round-2 sweep of the hsell shape knobs on RMAT-24 fp32 (kernel-only timing; one process per configuration).
mkdir -p gpurun_out
run() {
  name=$1; rel=$2; shift 2
  env "$@" timeout 300 python bench.py --kernel-only --relabel $rel ${DT:+--dtype $DT} ${SCALE:+--scale $SCALE} > gpurun_out/sw_$name.log 2>&1
  echo "$name rel=$rel $* :: $(tail -1 gpurun_out/sw_$name.log | cut -c1-400)"
}
run base degree X=1
run deg_rc degree PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4
run hub hub X=1
run hub_rc hub PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4
run hub_rc8 hub PGB_HSELL_MIN_ENTRIES=8 PGB_HSELL_ROUND_COST=4
run hub_k128 hub PGB_HSELL_BLOCKS=128
run hub_k128_rc hub PGB_HSELL_BLOCKS=128 PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4
run hub_k256_rc hub PGB_HSELL_BLOCKS=256 PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4
run hub_k512_rc hub PGB_HSELL_BLOCKS=512 PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4
run deg_k128_rc degree PGB_HSELL_BLOCKS=128 PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4
run hub_rc_skiphub hub PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4 PGB_HSELL_DEBUG_SKIP=1
run hub_rc_skiptail hub PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4 PGB_HSELL_DEBUG_SKIP=2
run hub_k256_rc_skiphub hub PGB_HSELL_BLOCKS=256 PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4 PGB_HSELL_DEBUG_SKIP=1
run hub_k256_rc_skiptail hub PGB_HSELL_BLOCKS=256 PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4 PGB_HSELL_DEBUG_SKIP=2
DT=f64 run base_f64 degree X=1
DT=f64 run hub_rc_f64 hub PGB_HSELL_MIN_ENTRIES=2 PGB_HSELL_ROUND_COST=5.4
