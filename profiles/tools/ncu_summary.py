"""Condense `ncu -i X.ncu-rep --page raw --csv` into one row per kernel with the metrics DESIGN.md quotes.
    python profiles/tools/ncu_summary.py raw.csv > summary.csv
Row 2 of the raw page holds the units; bytes are normalised to GB and durations to ms."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'lts__t_sector_hit_rate.pct']
SCALE = {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}


def main(path):
    rows = list(csv.reader(open(path)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {}
    for w in WANT:          # exact name first, else a section-prefixed copy that has a value in the first data row
        cands = [i for i, h in enumerate(head) if h == w] + [i for i, h in enumerate(head) if h.endswith('.' + w)]
        cands = [i for i in cands if data and data[0][i] != ''] or cands
        if cands:
            col[w] = cands[0]
    kn = head.index('Kernel Name')
    out = csv.writer(sys.stdout)
    out.writerow(['kernel'] + [w + ('[GB]' if 'bytes' in w else '[ms]' if 'duration' in w else '') for w in WANT])
    for r in data:
        vals = []
        for w in WANT:
            i = col.get(w)
            if i is None or r[i] == '':
                vals.append('')
                continue
            v = float(r[i].replace(',', ''))
            v *= SCALE.get(units[i], 1.0)
            vals.append('%.6g' % v)
        out.writerow([r[kn]] + vals)


if __name__ == '__main__':
    main(sys.argv[1])
