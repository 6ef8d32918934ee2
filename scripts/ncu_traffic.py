#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` captures: dram__bytes_read.sum + dram__bytes_write.sum per launch of the
dominant kernel, keyed by the MANGLED kernel symbol (what pst_kernel_name returns), so bench.py can only ever attach a
traffic figure to the instantiation it actually launched.
    python scripts/ncu_traffic.py <workload> <real> <report.ncu-rep> [<workload> <real> <report> ...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "traffic.json")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-kernel-base", "mangled"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    out = []
    for r in rows[2:]:
        d = {}
        for i, k in enumerate(h):
            v = r[i]
            if k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum") and units[i] in scale:
                v = float(v.replace(",", "")) * scale[units[i]]          # bytes / milliseconds
            d[k] = v
        out.append(d)
    return out


def _f(v):
    try:
        return float(str(v).replace(",", ""))
    except Exception:
        return None


def main():
    a = sys.argv[1:]
    entries = []
    if os.path.exists(OUT):
        old = json.load(open(OUT))
        entries = old.get("entries", [])
    for k in range(0, len(a), 3):
        workload, real, rep = a[k], a[k + 1], a[k + 2]
        for r in raw(rep):
            e = {"kernel_mangled": r["Kernel Name"], "workload": workload, "real": real,
                 "dram_bytes_per_launch": int(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"]),
                 "kernel_ms_under_ncu": r.get("gpu__time_duration.sum"),
                 "l1_data_pipe_pct_of_peak": _f(r.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")),
                 "fp64_pipe_pct_of_peak": _f(r.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")),
                 "dram_pct_of_peak": _f(r.get("dram__throughput.avg.pct_of_peak_sustained_elapsed")),
                 "l1_hit_rate_pct": _f(r.get("l1tex__t_sector_hit_rate.pct")), "l2_hit_rate_pct": _f(r.get("lts__t_sector_hit_rate.pct")),
                 "source": "ncu --set full --clock-control none, " + os.path.basename(rep)}
            entries = [x for x in entries if not (x["kernel_mangled"] == e["kernel_mangled"] and x["workload"] == workload and x["real"] == real)]
            entries.append(e)
    json.dump({"_comment": "dram bytes per launch of the dominant kernel from ncu captures; written by scripts/ncu_traffic.py; bench.py matches kernel_mangled + workload + real exactly",
               "entries": entries}, open(OUT, "w"), indent=1)
    print(f"{len(entries)} entries -> {OUT}")


if __name__ == "__main__":
    main()
