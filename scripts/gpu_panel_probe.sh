#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/panel_probe4.jsonl
: > $out
run() { env "$@" timeout 300 python scripts/panel_probe.py 2>&1 | tail -1 >> $out; }
run PGB_PANEL=1
run PGB_PANEL=1 PGB_HSELL_BLOCKS=448
run PGB_PANEL=1 PGB_HSELL_BLOCKS=512
run PGB_PANEL=1 PGB_HSELL_PANEL_TAIL_WARPS=6
run PGB_PANEL=1 PGB_HSELL_PANEL_TAIL_WARPS=10
python - <<'PY'
import json
for l in open('gpurun_out/panel_probe4.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad', l[:200]); continue
    e=d['env']; e.pop('PGB_PANEL',None)
    print(f"{d['ms_per_panel_step']:.3f} ms K {d['K']} pieces {d['pieces']/1e6:.2f}M hub_slots {d['hub_slots']/1e6:.0f}M tail_entries {d['tail_entries']/1e6:.0f}M", e)
PY
