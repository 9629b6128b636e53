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
for h in 0 2048 4096 8192 10240; do run b_hub$h env PGB_HUB_ENTRIES=$h; done
echo "== tests (default hub)"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/tests.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/tests.log | head -10
