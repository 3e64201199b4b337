#!/bin/bash
# Key metrics of every launch in a `--set full` .ncu-rep as "metric = value unit" lines.
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; u=r[1]
keys=['Kernel Name','launch__grid_size','launch__block_size','launch__registers_per_thread','gpu__time_duration.sum','sm__cycles_elapsed.avg.per_second','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','smsp__issue_active.avg.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_barrier','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_math_pipe_throttle','smsp__pcsamp_warps_issue_stalled_not_selected','smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_mio_throttle','smsp__pcsamp_warps_issue_stalled_lg_throttle','smsp__pcsamp_warps_issue_stalled_selected','smsp__pcsamp_warps_issue_stalled_membar','smsp__pcsamp_warps_issue_stalled_sleeping','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for row in r[2:]:
    print('-----')
    for k in keys:
        if k in h:
            i=h.index(k); print(k, '=', row[i][:160], u[i])
"
