#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_plugin_dropin.py tests/test_closed_form_options_gpu.py -x -q 2>&1 | tail -12
timeout 800 python bench.py --no-cpu --no-panel 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e_plugin'])"
