#!/usr/bin/env python
"""Prints the metrics we track from an .ncu-rep (via `ncu -i rep --page raw --csv`). usage: ncu_keys.py file.ncu-rep"""
import csv, subprocess, sys, io
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_barrier.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
        'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_xu.sum',
        'smsp__inst_executed_pipe_fmaheavy.sum', 'smsp__inst_executed_pipe_fmalite.sum', 'smsp__inst_executed_pipe_uniform.sum', 'smsp__inst_executed_pipe_cbu.sum', 'smsp__inst_executed_pipe_adu.sum',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__cycles_active.avg']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:90s} {r[i]:>18s} {units[i]}")
    print()
