#!/usr/bin/env python
"""gpurun_out/ ncu artefacts of scripts/gpu_profile_r2.sh -> profiles/{launches,step,kernel_*,traffic,sass}_<tag>.*"""
import csv, io, json, os, re, subprocess, sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_tex_mem_texture.sum",
        "l1tex__t_sectors_pipe_tex_mem_texture.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_red.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_active.min",
        "sm__cycles_active.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic"]


def raw(rep):
    """(header, units, values) of the one profiled launch: from the CSV exported on the GPU box (gpu_profile_r2.sh) or
    from the .ncu-rep itself."""
    csv_path = rep[:-len(".ncu-rep")] + ".raw.csv"
    if os.path.exists(csv_path):
        out = open(csv_path).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2]


def hot(rep):
    txt = rep[:-len(".ncu-rep")] + ".hot.txt"
    if os.path.exists(txt):
        return open(txt).read()
    return subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_hot.py"), rep, "14"], capture_output=True, text=True).stdout


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


traffic = {}
NOTES = {"gather": "RMAT scale 24 fp32, one launch of `bench.py --kernel-only`",
         "update": "RMAT scale 24 fp32, one launch of `bench.py --kernel-only`",
         "panel_gather": "panel path, RMAT scale 24, 4 x fp32 columns per node (f32x4 elements), one launch of `scripts/panel_probe.py`",
         "panel_update": "panel path, RMAT scale 24, 4 x fp32 columns, one launch of `scripts/panel_probe.py`"}
for suffix, dt in (("gather", "f32"), ("update", "f32"), ("gather_f64", "f64"), ("update_f64", "f64"),
                   ("panel_gather", "f32x4"), ("panel_update", "f32x4")):
    rep = os.path.join(GO, f"prof_{tag}_{suffix}.ncu-rep")
    if not os.path.exists(rep) and not os.path.exists(rep[:-len(".ncu-rep")] + ".raw.csv"):
        continue
    hdr, units, vals = raw(rep)
    name = vals[hdr.index("Kernel Name")]
    rd = to_bytes(vals[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
    wr = to_bytes(vals[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
    traffic.setdefault(dt, {"bytes_per_launch": {}})["bytes_per_launch"][name.split("(")[0]] = rd + wr
    if dt == "f64":
        continue
    with open(os.path.join(OUT, f"kernel_{tag}_{suffix}.md"), "w") as f:
        f.write(f"# `{name[:90]}` — ncu --set full --clock-control none ({tag}), {NOTES[suffix]}\n\n")
        f.write("Cold-cache, serialised profiler run: durations are indicative, counters are per launch.\n\n| metric | value | unit |\n|---|---:|---|\n")
        for i, h in enumerate(hdr):
            if h in WANT:
                f.write(f"| `{h}` | {vals[i]} | {units[i]} |\n")
        st = [(float(vals[i]), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        f.write("\nTop warp stall reasons (warps per issue-active cycle):\n\n| reason | value |\n|---|---:|\n")
        for v, h in sorted(st, reverse=True)[:7]:
            f.write("| %s | %.2f |\n" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
        if suffix in ("gather", "panel_gather"):
            f.write("\nMost-sampled SASS instructions (source page):\n\n```\n" + hot(rep) + "```\n")
with open(os.path.join(OUT, f"traffic_{tag}.json"), "w") as f:
    json.dump(dict(traffic, source="ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, bench.py --kernel-only "
                                   "(RMAT scale 24), gather + update kernels of one fused step"), f, indent=1)

path = os.path.join(GO, f"launches_{tag}.csv")
if os.path.exists(path):
    rows = list(csv.DictReader(io.StringIO("".join(l for l in open(path) if not l.startswith("==")))))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        sc = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r.get("Metric Unit", "ns"), 1e-6)
        a = agg[r["Kernel Name"].split("(")[0][:110]]
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", "")) * sc
    total = sum(v[1] for v in agg.values()) or 1.0
    with open(os.path.join(OUT, f"launches_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none -c 400` over `bench.py --steps 2 --warmup 3 --no-cpu --no-plugin`\n\n"
                "Per-launch times are cold-cache and serialised by the profiler: read the SHARES, not the absolutes.  The first 400 launches cover the graph build "
                "and the first solves.\n\n| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|\n")
        for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
            f.write(f"| `{name}` | {cnt} | {ms:.3f} | {100 * ms / total:.1f}% | {ms / cnt:.4f} |\n")
    step = {k: v for k, v in agg.items() if "hsell_gather_kernel" in k or "hsell_update" in k or "hsell_reduce" in k}
    tot = sum(ms / cnt for cnt, ms in step.values()) or 1.0
    with open(os.path.join(OUT, f"step_{tag}.md"), "w") as f:
        f.write(f"# Kernels of one fused PPR step ({tag}), RMAT scale 24 fp32 (from the launch list above)\n\nSerialised, cold-cache profiler times: read the SHARES.\n\n"
                "| kernel | launches | avg ms | share of the step |\n|---|---:|---:|---:|\n")
        for name, (cnt, ms) in sorted(step.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {cnt} | {ms / cnt:.4f} | {100 * (ms / cnt) / tot:.1f}% |\n")

# SASS opcode histogram of the step kernels (cuobjdump of the shipped library)
lib = os.path.join(ROOT, "pygrank_b200", "lib", "libpgb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
with open(os.path.join(OUT, f"sass_{tag}.md"), "w") as f:
    f.write(f"# SASS opcode histograms ({tag}): `cuobjdump -sass pygrank_b200/lib/libpgb200.so`, sm_100a\n\n")
    for want in ("hsell_gather_kernelIfLb1ELb1ELb0ELb0E", "hsell_gather_kernelINS_5f32x4ELb1ELb1ELb0ELb0E", "hsell_gather_kernelIfLb1ELb1ELb0ELb1E", "hsell_update_accum_kernelIfLi1ELb1E", "hsell_update_panel_kernelIfLi4ELb1E"):
        m = re.search(r"Function : (\S*" + want + r"\S*)(.*?)(?=Function : |\Z)", sass, re.S)
        if not m:
            continue
        ops = Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", m.group(2), re.M))
        fam = Counter()
        for op, c in ops.items():
            fam[op.split(".")[0]] += c
        f.write(f"## `{m.group(1)[:100]}`\n\n{sum(ops.values())} instructions.  Memory / special opcodes in full, the rest by family.\n\n| opcode | count |\n|---|---:|\n")
        for op, c in sorted(ops.items(), key=lambda kv: -kv[1]):
            if re.match(r"(LD|ST|TLD|TEX|RED|ATOM|SHFL|BAR|UBLK|UTMA|SYNCS|FENCE|CCTL|MEMBAR|LDS|STS|LDSM|ERRBAR)", op):
                f.write(f"| `{op}` | {c} |\n")
        f.write("\n| family | count |\n|---|---:|\n")
        for op, c in fam.most_common(14):
            f.write(f"| `{op}` | {c} |\n")
        f.write("\n")
print("done")
