#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel duration, DRAM bytes, FP64 pipe, occupancy, stalls.
usage: python tools/ncu_summary.py file.ncu-rep [more metrics...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_pipe_xu.sum', 'launch__grid_size', 'launch__block_size',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
        'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct',
        'derived__smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        ] + sys.argv[2:]
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print('==', r[ki][:110])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:80s} {r[i]:>18s} {units[i]}")
