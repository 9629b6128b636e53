#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/$name.log 2>&1; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.log').read().strip().splitlines()[-1]); r=d['roofline']
    print('$name', 'value=%.1f e2e=%.1f kernel_ms=%.3f frac=%.3f probe_ms=%.3f'%(d['value'],d['e2e']['value'],r['kernel_ms'],r['frac'],r['gather_probe_ms']))
except Exception as e: print('$name failed', e); print(open('gpurun_out/$name.log').read()[-1500:])
PY
}
L=$PWD/pygrank_b200/lib
run b_default env
run b_carve20 env PGB_SMEM_CARVEOUT=20
run b_carve28 env PGB_SMEM_CARVEOUT=28
run b_carve40 env PGB_SMEM_CARVEOUT=40
run b_mb4 env PGB_LIB=$L/libpgb200_mb4.so
run b_mb4_carve35 env PGB_LIB=$L/libpgb200_mb4.so PGB_SMEM_CARVEOUT=35
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/tests.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/tests.log | head -10
