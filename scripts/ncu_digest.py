"""Digest of one `ncu --set full --import-source on` report: headline metrics, stall mix, and the hot SASS blocks
(consecutive instructions with a similar execution count = one loop), so a profile can be read without the GUI.
    python scripts/ncu_digest.py gpurun_out/x.ncu-rep [--sass]   > profiles/x.txt"""
import csv
import io
import subprocess
import sys


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    show_sass = "--sass" in sys.argv
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = {h: (units[i], vals[i]) for i, h in enumerate(hdr)}
    print("kernel:", d.get("Kernel Name", ("", "?"))[1])
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
            "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio"]
    for k in keys:
        if k in d:
            print(f"  {k:82s} {d[k][1]:>18s} {d[k][0]}")
    for h in hdr:
        if h.startswith("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts"):
            print(f"  {h:82s} {d[h][1]:>18s} (avg per SM)")
    print("stall reasons (warps stalled per issued instruction):")
    st = [(float(d[h][1]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h][1]]
    for v, h in sorted(st, reverse=True)[:9]:
        print(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:6.2f}")
    src = page(rep, "source", ["--print-source", "sass"])
    h2 = src[1]
    ix = {h: i for i, h in enumerate(h2)}
    rows = src[2:]
    nwarps = max(int(r[ix["Instructions Executed"]]) for r in rows[:4]) or 1
    tot = sum(int(r[ix["# Samples"]]) for r in rows) or 1
    print(f"hot SASS blocks (x = executions per launched warp; {nwarps} warps; {tot} samples):")
    blocks, start, prev = [], 0, None
    for k, r in enumerate(rows + [None]):
        x = int(r[ix["Instructions Executed"]]) / nwarps if r else -1
        if prev is None:
            prev = x
        if r is None or abs(x - prev) > max(0.6, 0.12 * prev):
            seg = rows[start:k]
            s = sum(int(q[ix["# Samples"]]) for q in seg)
            ins = sum(int(q[ix["Instructions Executed"]]) for q in seg) / nwarps
            shw = sum(int(q[ix["L1 Wavefronts Shared"]] or 0) for q in seg) / nwarps
            glq = sum(int(q[ix["L1 Tag Requests Global"]] or 0) for q in seg) / nwarps
            blocks.append((start, k - 1, prev, 100.0 * s / tot, ins, shw, glq, seg[0][ix["Source"]].strip()[:44]))
            start, prev = k, x
    for b in blocks:
        if b[3] >= 0.4:
            print(f"  sass {b[0]:5d}-{b[1]:5d}  x{b[2]:7.1f}  samples {b[3]:5.1f}%  inst/warp {b[4]:8.1f}  smem wavefronts/warp {b[5]:7.1f}  global tag requests/warp {b[6]:7.1f}  | {b[7]}")
    if show_sass:
        for k, r in enumerate(rows):
            s = int(r[ix["# Samples"]])
            print(f"{k:5d} {100.0 * s / tot:6.2f}% x{int(r[ix['Instructions Executed']]) / nwarps:8.1f} | {r[ix['Source']].strip()}")


if __name__ == "__main__":
    main()
