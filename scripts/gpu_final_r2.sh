#!/bin/bash
# round-2 closing run: the whole GPU suite, configs C1-C5, sanitizer over the small end-to-end pass, the default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/final_tests.log
timeout 900 python scripts/run_configs.py --tag r2 > gpurun_out/configs_r2.log 2>&1; grep -E "^C3|^C5" gpurun_out/configs_r2.log | cut -c1-900
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_mem.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_mem.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_race.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitize_race.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-700
