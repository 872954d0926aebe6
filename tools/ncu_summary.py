#!/usr/bin/env python
"""measurement aid: condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few numbers DESIGN.md quotes"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'sm__cycles_elapsed.max']


def main(path):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('###', r[hdr.index('Kernel Name')])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k); print('  %-72s %s %s' % (k, r[i], units[i]))
        st = []
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
                try:
                    st.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
                except ValueError:
                    pass
        print('  stalls per issue:', ', '.join('%s %.2f' % (n, v) for v, n in sorted(st, reverse=True)[:7]))
        print()


if __name__ == '__main__':
    main(sys.argv[1])
