#!/bin/bash
# light ncu pass (a handful of metrics, one launch) of the pair kernel for each library build given: gpurun_out/r2_ncu_light_<name>.csv
M=gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,sm__inst_executed.avg.per_cycle_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
mkdir -p gpurun_out
for lib in "$@"; do
  name=$(basename $lib .so | sed 's/libprestige_b200_//')
  PRESTIGE_B200_LIB=$PWD/$lib timeout 600 ncu --metrics $M --clock-control none -k regex:k_wcsph_zrun -s 13 -c 1 --csv --log-file gpurun_out/r2_ncu_light_$name.csv python scripts/r2_ab.py --child 200,200,250 $R2_OPTS > gpurun_out/r2_ncu_light_$name.log 2>&1
  echo "== $name"; python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r2_ncu_light_$name.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
if hdr:
    h=rows[hdr[0]]; ix={k:i for i,k in enumerate(h)}
    for r in rows[hdr[0]+1:]:
        if len(r)>ix["Metric Value"]: print(f'{r[ix["Metric Name"]]:95s} {r[ix["Metric Value"]]:>20s} {r[ix["Metric Unit"]]}')
else:
    print(open("gpurun_out/r2_ncu_light_$name.csv").read()[-1500:])
PY
done
