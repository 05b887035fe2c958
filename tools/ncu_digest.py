#!/usr/bin/env python3
"""Digest of an `ncu --set full` capture: the metrics DESIGN.md / profiles/*.md quote, one kernel per block.
usage: ncu -i rep --page raw --csv | tools/ncu_digest.py > profiles/<name>.md"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "memory_l1_wavefronts_shared",
    "memory_l1_wavefronts_shared_ideal", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    print(f"## {r[ik][:100]}\n")
    for m in WANT:
        if m in hdr:
            i = hdr.index(m)
            print(f"- {m} = {r[i]} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i].replace(",", "")), h[len(STALL):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("- stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st if v >= 0.2))
    print()
