#!/usr/bin/env python
"""Turns the ncu artefacts under gpurun_out/ into the small text summaries committed under profiles/.

    python scripts/summarize_profile.py <tag>      # e.g. r1
reads  gpurun_out/launches_<tag>.csv  (ncu --metrics gpu__time_duration.sum launch list of bench.py)
       gpurun_out/prof_<tag>.ncu-rep  (ncu --set full capture of the dominant kernel)
writes profiles/launches_<tag>.md, profiles/kernel_<tag>.md
"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def step_list():
    """Per-launch durations of the three kernels of one step (ncu -k regex:hsell_ over bench.py --kernel-only)."""
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}_step.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = defaultdict(list)
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r.get("Metric Unit", "ns"), 1e-6)
        agg[r["Kernel Name"].split("(")[0]].append(float(r["Metric Value"].replace(",", "")) * scale)
    tot = sum(sum(v) / len(v) for v in agg.values()) or 1.0
    with open(os.path.join(out_dir, f"step_{tag}.md"), "w") as f:
        f.write(f"# Kernels of one fused PPR step ({tag}), RMAT scale 24 fp32: `ncu --metrics gpu__time_duration.sum "
                "--clock-control none -k regex:hsell_` over `bench.py --kernel-only`\n\n")
        f.write("Serialised, cold-cache profiler times: read the SHARES.\n\n| kernel | launches | avg ms | share of the step |\n|---|---:|---:|---:|\n")
        for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{name}` | {len(v)} | {sum(v) / len(v):.4f} | {100 * (sum(v) / len(v)) / tot:.1f}% |\n")
    print("wrote step list")


def launch_list():
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += val * scale
    total = sum(v[1] for v in agg.values()) or 1.0
    with open(os.path.join(out_dir, f"launches_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py`\n\n")
        f.write("Per-launch times are cold-cache and serialised by the profiler: read the SHARES, not the absolutes.\n\n")
        f.write("| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|\n")
        for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {cnt} | {ms:.3f} | {100 * ms / total:.1f}% | {ms / cnt:.4f} |\n")
    print("wrote launches summary:", len(rows), "rows")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max"]


TRAFFIC = {}


def kernel_summary(rep, title, fname):
    path = os.path.join(ROOT, "gpurun_out", rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(out_dir, fname), "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` ({rep}); one launch.\n\n")
        for r in rows[2:3]:
            f.write(f"kernel: `{r[hdr.index('Kernel Name')]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"| {w} | {r[i]} | {units[i]} |\n")
            try:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tot = 0.0
                for w in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    i = hdr.index(w)
                    tot += float(r[i].replace(",", "")) * scale[units[i]]
                TRAFFIC[r[hdr.index("Kernel Name")].split("(")[0]] = tot
            except (ValueError, KeyError):
                pass
        src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        try:
            hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"][0]
            sh, data = srows[hi], srows[hi + 1:]
            col = {h: i for i, h in enumerate(sh)}
            stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
            agg = {s: 0 for s in stalls}
            for r in data:
                for s in stalls:
                    try:
                        agg[s] += int(r[col[s]] or 0)
                    except (ValueError, IndexError):
                        pass
            tot = sum(agg.values()) or 1
            f.write("\n## warp stall samples (all SASS instructions)\n\n| reason | samples | share |\n|---|---:|---:|\n")
            for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
                f.write(f"| {s} | {v} | {100 * v / tot:.1f}% |\n")
        except IndexError:
            pass
    print("wrote", fname)


launch_list()
step_list()
kernel_summary(f"prof_{tag}.ncu-rep", f"Dominant kernel ({tag}): fused PPR step on RMAT scale 24, fp32", f"kernel_{tag}.md")
kernel_summary(f"prof_{tag}_probe.ncu-rep", f"Gather probe ({tag}): index stream + gathers only, same graph", f"probe_{tag}.md")
kernel_summary(f"prof_{tag}_gather.ncu-rep",
               f"Dominant kernel ({tag}): hsell_gather_kernel<float> — one PPR step on RMAT scale 24, fp32", f"kernel_{tag}_gather.md")
kernel_summary(f"prof_{tag}_update.ncu-rep",
               f"Second kernel of the step ({tag}): hsell_update_kernel<float, AFFINE, SYMDEG>, same step", f"kernel_{tag}_update.md")

if TRAFFIC:
    import json
    with open(os.path.join(out_dir, f"traffic_{tag}.json"), "w") as f:
        json.dump({"source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                             "bench.py --kernel-only (RMAT scale 24, fp32)", "bytes_per_launch": TRAFFIC}, f, indent=1)
    print("wrote traffic")
