#!/bin/bash
mkdir -p gpurun_out
PGB_PANEL_TIMING=1 timeout 300 python scripts/panel_e2e.py 2>&1 | tail -12
PROBE_PANEL_CHUNK=4 timeout 300 python scripts/panel_e2e.py 2>&1 | tail -1
PROBE_PANEL_CHUNK=16 timeout 300 python scripts/panel_e2e.py 2>&1 | tail -1
PROBE_PANEL_GROUP=64 timeout 300 python scripts/panel_e2e.py 2>&1 | tail -1
PGB_PANEL=1 timeout 300 python scripts/panel_probe.py 2>&1 | tail -1 | cut -c1-60,330-420
