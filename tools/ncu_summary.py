#!/usr/bin/env python
"""Key metrics per kernel from an .ncu-rep (`ncu -i rep --page raw --csv`)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'sm__cycles_elapsed.avg', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active' ]
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output = True, text = True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    stall = [x for x in h if 'smsp__average_warps_issue_stalled' in x and x.endswith('_per_issue_active.ratio')]
    for r in rows[2:]:
        print('=' * 20, r[h.index('Kernel Name')][:120])
        for w in WANT:
            if w in h:
                print(f'  {w:75s} {r[h.index(w)]:>16s} {units[h.index(w)]}')
        st = sorted(((float(r[h.index(x)] or 0), x) for x in stall), reverse = True)[:6]
        for v, x in st:
            print(f'  stall {x.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):40s} {v:8.2f}')
if __name__ == '__main__':
    main(sys.argv[1])
