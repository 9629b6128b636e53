#!/usr/bin/env python
"""Quick look at ncu artefacts in gpurun_out/: python scripts/ncu_brief.py <tag> [kernel-suffix ...]"""
import csv, io, subprocess, sys, collections, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(path):
    rows = list(csv.DictReader(io.StringIO("".join(l for l in open(path) if not l.startswith("==")))))
    agg = collections.defaultdict(list)
    for r in rows:
        if r["Metric Name"] == "gpu__time_duration.sum":
            sc = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r["Metric Unit"]]
            agg[r["Kernel Name"].split("(")[0]].append(float(r["Metric Value"].replace(",", "")) * sc)
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:8]:
        v2 = sorted(v)
        print("%-58s n=%3d total=%8.3f med=%.4f max=%.4f" % (k[:58], len(v), sum(v), v2[len(v2) // 2], v2[-1]))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__cycles_active.max", "sm__cycles_active.min", "launch__grid_size", "launch__registers_per_thread",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
for suf in sys.argv[2:]:
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_{suf}.ncu-rep")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("==", suf, vals[hdr.index("Kernel Name")][:80])
    for i, h in enumerate(hdr):
        if h in WANT:
            print("  %-70s %12s %s" % (h, vals[i], units[i]))
    st = [(float(vals[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:6]:
        print("  stall %-60s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
