#!/bin/bash
# First GPU pass: smoke, parity tests, a small and the full bench, ncu launch list + one full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== tests" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -x --timeout=600 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -40 gpurun_out/tests.log
echo "== bench scale 20" ; timeout 600 python bench.py --scale 20 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench20.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench20.log
echo "== bench scale 24" ; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench24.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench24.log
echo "== bench scale 24 relabel none" ; timeout 900 python bench.py --steps 3 --warmup 3 --relabel none --no-cpu > gpurun_out/bench24_norelabel.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench24_norelabel.log
echo "== ncu launch list (scale 22)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --scale 22 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full (scale 24)"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 12 -c 2 -o gpurun_out/prof_r1 -f python bench.py --scale 24 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out
