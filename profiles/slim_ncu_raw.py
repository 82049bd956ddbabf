#!/usr/bin/env python
"""Keep the columns of an `ncu -i X.ncu-rep --page raw --csv` table that the docs quote (the full table has ~1500)."""
import csv
import sys
KEEP = ['ID', 'Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.avg.per_cycle_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def main(src, dst):
    rows = list(csv.reader(open(src)))
    idx = [rows[0].index(k) for k in KEEP if k in rows[0]]
    w = csv.writer(open(dst, "w"))
    for r in rows:
        w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
