"""Summarise an .ncu-rep (read here, no GPU): key raw metrics, stall mix and the hottest
source lines per kernel.  Usage: python tools/ncu_summary.py file.ncu-rep [kernel-regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
sel = ["--kernel-name", "regex:" + kre] if kre else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in idx:
            print("%-75s %s %s" % (w, r[idx[w]][:90], units[idx[w]]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name",')
for blk in blocks[1:2]:
    lines = list(csv.reader(io.StringIO('"Kernel Name",' + blk)))
    h = lines[1]
    ix = {k: i for i, k in enumerate(h)}
    data = [l for l in lines[2:] if len(l) == len(h)]

    def g(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    tot = {}
    for r in data:
        for k in h:
            if k.startswith("stall_") and "Not Issued" not in k:
                tot[k] = tot.get(k, 0) + g(r, k)
    s = sum(tot.values()) or 1
    print("stall mix:", [(k, round(100 * v / s, 1)) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]])
    ts = sum(g(r, "# Samples") for r in data) or 1
    print("hottest instructions:")
    for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:22]:
        st = sorted(((k, g(r, k)) for k in h if k.startswith("stall_") and "Not Issued" not in k), key=lambda kv: -kv[1])[:2]
        print("  %5.1f%%  %-60s %s" % (100 * g(r, "# Samples") / ts, r[ix["Source"]].strip()[:60], st))
