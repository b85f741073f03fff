"""Turn ncu outputs (gpurun_out/) into the small text summaries kept under profiles/.
    python scripts/summarize_profile.py launches <launches.csv> <out.md>
    python scripts/summarize_profile.py kernel <prof.ncu-rep> <out.md>
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_active.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(',', '')))
        except ValueError:
            pass
    total = sum(sum(v) for v in agg.values())
    with open(out, 'w') as f:
        f.write('| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n')
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f'| `{k[:90]}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / 1e3:.1f} | {100 * sum(v) / total:.1f}% |\n')


def kernel(path, out):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    with open(out, 'w') as f:
        f.write('| metric | unit | ' + ' | '.join(f'launch {i}' for i in range(len(rows) - 2)) + ' |\n')
        f.write('|---|---|' + '---|' * (len(rows) - 2) + '\n')
        name = hdr.index('Kernel Name')
        f.write('| kernel | | ' + ' | '.join('`' + r[name][:60] + '`' for r in rows[2:]) + ' |\n')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f'| {k} | {rows[1][i]} | ' + ' | '.join(r[i] for r in rows[2:]) + ' |\n')


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
