#!/usr/bin/env python
"""Top stalled SASS instructions of an ncu report: python scripts/ncu_hot.py <file.ncu-rep> [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows))
stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[k] or 0) for r in rows) for k in stalls}
print("stall totals:", ", ".join(f"{k[6:]}={v} ({100*v/tot:.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:N]
for i in sorted(idx):
    r = rows[i]
    top = sorted(((int(r[k] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r['# Samples']):7d} {100*int(r['# Samples'])/tot:5.1f}%  exec={r['Instructions Executed']:>9}  {r['Source'].strip()[:70]:70s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
