"""profiles/r2_kernels_ncu_full.md from the outputs of scripts/gpu_call_r2_ncu.sh (gpurun_out/full_<section>.csv = `ncu --page raw
--csv` of a `--set full` capture, gpurun_out/zoo.jsonl = CUDA-event timings of the same workloads without ncu).
    python scripts/summarize_zoo.py gpurun_out profiles/r2_kernels_ncu_full.md
"""
import collections
import csv
import json
import os
import sys

src, out = sys.argv[1], sys.argv[2]
COLS = [('gpu__time_duration.sum', 'us'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__registers_per_thread', 'regs'), ('launch__waves_per_multiprocessor', 'waves'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM thr %'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM thr %'),
        ('dram__bytes_read.sum', 'DRAM rd'), ('dram__bytes_write.sum', 'DRAM wr'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long-sb'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier')]
BOUND = {
    'rot': 'rotated NMS, 32 images x 10 000 boxes (2 chunks of 16): pairwise ALU / latency; unit = IoU pair, 4.9995e7 algorithmic pairs per image',
    'rotbench': 'rotated NMS on bench.py\'s own workload (the best 10 000 of 64 512 RAPiD candidates @1024 per image, boxes of 18-250 px), 32 images',
    'rotclu': 'rotated NMS on detector-like clusters (40 objects x 250 mutually overlapping boxes per image), 32 images: the lazy narrow phase (rot_narrow = filter, rot_clip<0>, rot_clip<1>)',
    'atss': 'ATSS targets, batch 64 @640, 100 GT/image, 5 levels: HBM write of the dense target maps (86 floats + 2 mask bytes per cell) + ALU for 852 500 anchor-GT tests per image',
    'fcos': 'FCOS2 targets (central-region rule), same shapes as ATSS: HBM write of the target maps',
    'rowmax': 'IoU row-max, 64 x 25 575 boxes vs 100 GT: ALU (100 IoUs per 28 bytes moved)',
    'iou': 'pairwise AABB IoU 16 384 x 8 192 -> f32 matrix: HBM write (4 B per pair); second launch = the configs[3] shape 8 525 x 100',
    'iourot': 'pairwise rotated IoU 8 192 x 2 048 -> f64 matrix: ALU (float64 polygon clip per pair inside the circle cull)',
    'dense': 'dense scenes (configs[4]): decode + un-capped single-class NMS @704 (batch 64, 10 164 boxes) and @1536 (batch 8, 48 384 boxes): pairwise ALU / latency',
    'pre': 'image pre-processing 64 x 1080p -> 608 x 608 and 32 x 2048^2 -> 1024^2: HBM (3 B per source pixel + 12 B per output pixel)',
    'bench_decode_kernel': 'bench step, decode_kernel<FCOS, compact>, eager launch: HBM (200.8 MB algorithmic per launch)',
    'bench_postprocess_small_kernel': 'bench step, postprocess_small_kernel (48 registers), eager launch: latency (one CTA per image)',
}


def num(v):
    try:
        return float(v.replace(',', ''))
    except ValueError:
        return None


with open(out, 'w') as f:
    f.write('# ncu `--set full` of every kernel family, round 2\n\n'
            'Command per section: `ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:mydet '
            '-c N python scripts/kernel_zoo.py --once <section>` (scripts/gpu_call_r2_ncu.sh; one B200; each after the same command had '
            'exited 0 without ncu).  Times under `--set full` are serialised and cold-cache; the event-timed numbers of the same workloads '
            '(no profiler) are in the "timed" lines.  The .ncu-rep files are not kept (size); per-source-line stall samples of the hot '
            'kernels are in `r2_hot_lines_<section>.txt`.\n\n')
    timed = collections.defaultdict(list)
    zoo = os.path.join(src, 'zoo.jsonl')
    if os.path.exists(zoo):
        for line in open(zoo):
            d = json.loads(line)
            timed[d['section']].append(d)
    for sec, bound in BOUND.items():
        path = os.path.join(src, f'full_{sec}.csv')
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        name = hdr.index('Kernel Name')
        f.write(f'## {sec}\n\n{bound}\n\n')
        for d in timed.get(sec, []) + (timed.get('iou_cfg4', []) if sec == 'iou' else []):
            keys = [k for k in ('us_per_batch', 'us_per_image', 'us_per_call', 'us_per_launch', 'algorithmic_pairs_per_s', 'pairs_per_s',
                                'anchor_gt_pairs_per_s', 'frames_per_s', 'achieved_gbs', 'frac_of_hbm') if k in d]
            f.write('timed (CUDA events, no profiler): ' + d['workload'] + ': ' +
                    ', '.join(f'{k} {d[k]:.4g}' for k in keys) + '\n\n')
        agg = collections.OrderedDict()
        for r in rows[2:]:
            agg.setdefault(r[name].split('(')[0][:48], []).append(r)
        def head(col, label):
            u = units[hdr.index(col)] if col in hdr else ''
            return f'{label} [{u}]' if (label == 'us' and u) else label
        f.write('| kernel | launches | ' + ' | '.join(head(*c).replace('us [', 'time [') for c in COLS) + ' |\n|---|---|' + '---|' * len(COLS) + '\n')
        for k, rs in agg.items():
            cells = []
            for col, _ in COLS:
                if col not in hdr:
                    cells.append('')
                    continue
                i = hdr.index(col)
                vals = [num(r[i]) for r in rs if num(r[i]) is not None]
                if not vals:
                    cells.append('')
                    continue
                v = sum(vals) / len(vals)
                u = units[i]
                cells.append(f'{v:.3g} {u}' if col.startswith('dram__bytes') else f'{v:.4g}')
            f.write(f'| `{k}` | {len(rs)} | ' + ' | '.join(cells) + ' |\n')
        f.write('\n')
