"""Helpers to read `ncu --page raw --csv` / `--page source --csv` exports (used for the summaries in profiles/)."""
import csv, sys
from collections import Counter

KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keys = list(KEYS) + [k for k in hdr if 'issue_stalled' in k and 'per_issue_active' in k and 'not_issued' not in k]
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', ''), '|',
                  ' | '.join(r[i][:44] for r in data), '|', units[i])


def source(path, nelem, top=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    data = []
    for r in rows[2:]:
        if r and r[0] == 'Address':
            break
        if len(r) == len(hdr):
            data.append(r)
    ix = {k: hdr.index(k) for k in ('Address', 'Source', '# Samples', 'Instructions Executed', 'stall_long_sb', 'stall_barrier', 'stall_short_sb',
                                    'stall_wait', 'stall_mio', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Excessive')}
    tot = sum(int(r[ix['# Samples']]) for r in data)
    ti = sum(int(r[ix['Instructions Executed']]) for r in data)
    tw = sum(int(r[ix['L1 Wavefronts Shared']] or 0) for r in data)
    print('samples', tot, 'warp-inst/elem %.0f' % (ti / nelem), 'smem wavefronts/elem %.0f' % (tw / nelem))
    for k in ('stall_long_sb', 'stall_barrier', 'stall_short_sb', 'stall_wait', 'stall_mio'):
        print(k, '%.1f%%' % (100 * sum(int(r[ix[k]]) for r in data) / tot))
    c = Counter()
    for r in data:
        op = [o for o in r[ix['Source']].split() if not o.startswith('@')]
        c[op[0].split('.')[0] if op else '?'] += int(r[ix['Instructions Executed']])
    print('mix/elem:', ', '.join('%s %.0f' % (k, v / nelem) for k, v in c.most_common(14)))
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top]:
        print(r[ix['Address']][-5:], r[ix['# Samples']].rjust(6), 'lsb', r[ix['stall_long_sb']].rjust(5), 'bar', r[ix['stall_barrier']].rjust(5),
              'ssb', r[ix['stall_short_sb']].rjust(5), 'wait', r[ix['stall_wait']].rjust(5), '|', r[ix['Source']][:80])


if __name__ == '__main__':
    if sys.argv[1] == 'raw':
        raw(sys.argv[2])
    else:
        source(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 40)
