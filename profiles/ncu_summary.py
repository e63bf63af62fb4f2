#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` dump into the few numbers the optimisation loop reads."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = [
 "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__cluster_size", "launch__grid_size",
 "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__cluster_max_active",
 "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
 "sm__inst_executed.sum", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.sum.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
 "smsp__thread_inst_executed_per_inst_executed.ratio",
 "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
 "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
 "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
 "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size"))
    for k in keys:
        if k in d: print(f"  {k:85s} {d[k]:>18s} {u[k]}")
    st = [(float(d[k].replace(',', '')), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and d[k]]
    for v, k in sorted(st, reverse=True)[:9]:
        print(f"  stall {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):40s} {v:.3f}")
