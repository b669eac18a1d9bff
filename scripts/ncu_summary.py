#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the metrics DESIGN.md / bench.py cite (one column per launch).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv && python scripts/ncu_summary.py /tmp/raw.csv > profiles/NAME.csv
"""
import csv
import sys

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sass__inst_executed_shared_loads',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']


def main(path):
    rows = list(csv.reader(open(path)))
    head, units, data = rows[0], rows[1], rows[2:]
    w = csv.writer(sys.stdout)
    w.writerow(['metric', 'unit'] + [f'launch{i}' for i in range(len(data))])
    for k in KEEP:
        if k in head:
            i = head.index(k)
            w.writerow([k, units[i]] + [r[i] for r in data])


if __name__ == '__main__':
    main(sys.argv[1])
