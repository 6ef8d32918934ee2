#!/usr/bin/env python
"""Summarise an .ncu-rep (details + per-opcode + per-region) for one kernel.  Usage: ncu_summary.py rep [out.txt]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ["Duration", "Executed Ipc Active", "Issued Warp Per Scheduler", "Active Warps Per Scheduler", "Eligible Warps Per Scheduler",
        "Registers Per Thread", "Dynamic Shared Memory Per Block", "Theoretical Occupancy", "Achieved Occupancy", "Executed Instructions  ",
        "Avg. Active Threads Per Warp", "DRAM Throughput", "Memory Throughput", "L1/TEX Hit Rate", "L2 Hit Rate", "Compute (SM) Throughput",
        "SM Frequency", "Elapsed Cycles", "Block Limit", "Grid Size", "Block Size", "k_", "Local"]
out = []
for ln in det.splitlines():
    if any(k in ln for k in keys) and "OPT" not in ln:
        out.append(ln.rstrip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
if len(rows) >= 3:
    h = rows[0]
    for want in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__inst_executed_pipe_fp64.sum",
                 "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                 "smsp__warp_issue_stalled_wait_per_warp_active.pct", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"):
        for i, name in enumerate(h):
            if name == want:
                out.append(f"raw {want} = {rows[2][i]} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot, total_inst, total_samp = collections.Counter(), 0, 0
st = collections.Counter()
for r in data:
    try:
        n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except Exception:
        continue
    srcs = r[ix["Source"]].split()
    op = (srcs[1] if srcs[0].startswith("@") else srcs[0]).split(".")[0]
    tot[op] += n; total_inst += n; total_samp += s
    for hname in hdr:
        if hname.startswith("stall_") and "Not Issued" not in hname:
            try: st[hname] += int(r[ix[hname]])
            except Exception: pass
out.append(f"warp instructions executed: {total_inst}   stall samples: {total_samp}")
out.append("opcode mix (M warp-inst): " + ", ".join(f"{op} {n / 1e6:.1f}" for op, n in tot.most_common(28)))
out.append("stall samples: " + ", ".join(f"{k[6:]} {100 * v / max(1, sum(st.values())):.0f}%" for k, v in st.most_common(8)))
print("\n".join(out))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(out) + "\n")
