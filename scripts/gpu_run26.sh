#!/bin/bash
mkdir -p gpurun_out
echo "== hsell + maxdiff tests"; timeout 900 python -m pytest tests/test_hsell_gpu.py tests/test_gpu_parity.py -x -q -m gpu -k "hsell or lossless or forced or max_difference or fp32_mode" 2>&1 | tail -3
echo "== build time"; timeout 300 python - <<'PY'
import time, torch
import pygrank_b200 as pgb
from pygrank_b200 import device_synthetic
g = device_synthetic.rmat_graph_device(24, 16, seed=1, normalization="symmetric")
torch.cuda.synchronize()
for dt in (torch.float32, torch.float64):
    t0 = time.perf_counter(); f = g.in_view.hsell(dt); torch.cuda.synchronize(); print(dt, "hsell build s = %.3f" % (time.perf_counter() - t0), "bytes", f.nbytes())
PY
run() { echo "== $*"; env $1 $2 $3 timeout 300 python bench.py --kernel-only --steps 1 --warmup 1 $EXTRA 2>&1 | tail -1 | cut -c1-200; }
run PGB_X=0
