#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "bit_identical" > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu.log | tail -5
M="gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum"
for v in 2; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:k_forces_fluid -s 1 -c 1 --csv --log-file gpurun_out/ncu_forces_v$v.csv python tools/profile_forces.py 2000000 $v 20 > gpurun_out/prof_v$v.log 2>&1
  tail -1 gpurun_out/prof_v$v.log
done
python - <<'PY'
import csv
for v in (2,):
    try:
        rows=list(csv.reader(open(f"gpurun_out/ncu_forces_v{v}.csv")))
        hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
        print("== variant",v)
        for r in rows[hdr+1:]:
            print("  %-90s %s %s"%(r[-3],r[-1],r[-2]))
    except Exception as e:
        print("parse fail",v,e)
PY
