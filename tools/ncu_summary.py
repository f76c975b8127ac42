#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py gpurun_out/x.ncu-rep [more keys...]"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit', 'launch__shared_mem_per_block ', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum ', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.avg.per_cycle_active',
        'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'smsp__average_warps_issue_stalled', 'sm__inst_executed_pipe_fma.avg.pct', 'sm__inst_executed_pipe_alu.avg.pct',
        'sm__inst_executed_pipe_fp64.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct', 'sm__pipe_fp64_cycles_active.avg.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'sm__cycles_active.avg ', 'sm__cycles_elapsed.avg ', 'smsp__cycles_active.avg ', 'sm__throughput.avg.pct',
        'l1tex__t_sector_hit_rate', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed_op_shared', 'local_load', 'local_store',
        'smsp__inst_executed_op_local']


def main():
    rep = sys.argv[1]
    keys = KEYS + sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('== kernel:', vals[hdr.index('Kernel Name')][:100])
        for h, u, v in zip(hdr, units, vals):
            hh = h + ' '
            if any(k in hh for k in keys) and '.min' not in h and '.max' not in h and v not in ('', '0'):
                print('  %-90s %-12s %s' % (h, u, v))


if __name__ == '__main__':
    main()
