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
for v in mb4 ipt7; do run b_$v env PGB_LIB=$L/libpgb200_$v.so; done
echo "== tests (default)"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/tests.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/tests.log | head -10
echo "== tests (ipt7)"; PGB_LIB=$L/libpgb200_ipt7.so timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/tests_ipt7.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/tests_ipt7.log | head -10
