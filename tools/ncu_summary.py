#!/usr/bin/env python
"""Prints the handful of ncu metrics that matter for this repo from a .ncu-rep (reads `ncu --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__cycles_elapsed.avg', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio', 'l1tex__data_bank_conflicts_pipe_lsu.sum']


def main(path, kernel_index=0):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print('== kernel:', name[:90])
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f'   {h:78s} {r[i]:>16s} {units[i]}')
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('per_issue_active.ratio') and r[i]]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f'   stall {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:30s} {v:8.2f} warps/issue')


if __name__ == '__main__':
    main(sys.argv[1])
