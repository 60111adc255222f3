#!/bin/bash
# ncu summary of one launch of each solve kernel (16 images x 100 copies): the metrics DESIGN.md quotes.
# usage: scripts/ncu_solve_metrics.sh <tag>   -> gpurun_out/<tag>_ncu_solve.txt (+ .ncu-rep)
tag=${1:-r02}
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,\
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,\
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,\
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,\
lts__t_sectors.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem,\
smsp__warps_eligible.avg.per_cycle_active,sm__cycles_elapsed.avg.per_second
ncu --metrics $M --clock-control none -k regex:'k_forward_residual|k_gradient_update' -s 4 -c 2 \
    python scripts/prof_solve.py 16 4 > gpurun_out/${tag}_ncu_solve.txt 2>&1
