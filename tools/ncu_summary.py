"""Selected metrics of `ncu --page raw --csv` exports, one column per capture (value and unit per cell).
    python tools/ncu_summary.py a_raw.csv b_raw.csv ... > profiles/rN_..._ncu_summary.txt"""
import csv
import os
import sys

WANT = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tma.sum',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def main():
    files = sys.argv[1:]
    cols = []
    for f in files:
        r = list(csv.reader(open(f)))
        cols.append(dict(zip(r[0], zip(r[1], r[2]))))
    names = [os.path.basename(f).replace('_raw.csv', '') for f in files]
    print('%-84s %s' % ('metric', ' '.join(n[-22:].rjust(22) for n in names)))
    for k in WANT:
        if not any(k in c for c in cols):
            continue
        cells = []
        for c in cols:
            u, v = c.get(k, ('', '-'))
            try:
                v = '%.4g' % float(v.replace(',', ''))
            except ValueError:
                pass
            cells.append((v + ' ' + u).strip().rjust(22))
        print('%-84s %s' % (k[:84], ' '.join(cells)))


if __name__ == '__main__':
    main()
