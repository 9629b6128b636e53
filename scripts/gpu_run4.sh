#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/$name.log 2>&1; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.log').read().strip().splitlines()[-1]); r=d['roofline']
    print('$name', 'value=%.1f e2e=%.1f kernel_ms=%.3f frac=%.3f probe_ms=%.3f'%(d['value'],d['e2e']['value'],r['kernel_ms'],r['frac'],r['gather_probe_ms']))
except Exception as e: print('$name failed', e); print(open('gpurun_out/$name.log').read()[-800:])
PY
}
run b_default env
run b_carve0 env PGB_SMEM_CARVEOUT=0
run b_carve25 env PGB_SMEM_CARVEOUT=25
run b_carve50 env PGB_SMEM_CARVEOUT=50
run b_norelabel env PGB_DUMMY=1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --relabel none > gpurun_out/b_norelabel.log 2>&1; tail -1 gpurun_out/b_norelabel.log | cut -c 1-300; tail -1 gpurun_out/b_norelabel.log | grep -o '"roofline.*'
echo "== quick tests"; timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
