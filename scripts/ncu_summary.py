#!/usr/bin/env python
"""Summarise an ncu capture of the score kernel: headline metrics + per-region stall mix + top stall lines.
usage: scripts/ncu_summary.py gpurun_out/prof_TAG   (expects _raw.csv and _src.csv next to it)"""
import csv, sys
base = sys.argv[1]
rows = list(csv.reader(open(base + "_raw.csv")))
d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed_op_shfl.sum"]
for k in keys:
    if k in d:
        print("{:75s} {:>12s} {}".format(k, d[k][0], d[k][1]))
rows = list(csv.reader(open(base + "_src.csv")))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
samp, src, ex = ci["# Samples"], ci["Source"], ci["Instructions Executed"]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
epi = [i for i, r in enumerate(data) if "USETMAXREG.TRY_ALLOC" in r[src]]
epi = epi[0] if epi else 0
for name, (lo, hi) in {"control warps": (0, epi), "epilogue warps": (epi, len(data))}.items():
    n = sum(int(r[samp]) for r in data[lo:hi]); e = sum(int(r[ex]) for r in data[lo:hi])
    print("\n[{}] sass lines {}-{}: samples {}  warp-instructions executed {}".format(name, lo, hi, n, e))
    agg = {k: sum(int(r[ci[k]] or 0) for r in data[lo:hi]) for k in reasons}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]:
        print("    {:28s} {:9d} {:5.1f}%".format(k, v, 100.0 * v / max(n, 1)))
print("\ntop epilogue lines by samples:")
for i in sorted(range(epi, len(data)), key=lambda i: -int(data[i][samp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    r = data[i]
    rs = sorted([(int(r[ci[k]] or 0), k) for k in reasons], reverse=True)[:2]
    print("  {:6d} {:8s} {:10s} {:60s} {}".format(i, r[samp], r[ex], r[src].strip()[:60], rs))
