#!/bin/bash
mkdir -p gpurun_out
echo "== bench default (auto carveout)"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value=%.1f e2e=%.1f kernel_ms=%.3f frac=%.3f probe=%.3f cpu=%s'%(d['value'],d['e2e']['value'],r['kernel_ms'],r['frac'],r['gather_probe_ms'],d['cpu_baseline']['value']))"
echo "== configs"; timeout 1500 python scripts/run_configs.py --tag r1 > gpurun_out/configs_r1.log 2>&1; echo rc=$?; tail -5 gpurun_out/configs_r1.log | cut -c1-900
