"""Developer tool: per-source-line warp-stall samples of one kernel from an ncu report.
    python scripts/hot_lines.py <prof.ncu-rep> [file-substring] [top-N]
Reads `ncu --page source --print-source cuda,sass` (needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ''
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, out = None, None, []
for r in rows:
    if r and r[0] in ('File Name', 'File Path'):
        fname = r[1]
    elif r and r[0] == 'Line No':
        hdr = r
    elif hdr and fname and want in fname and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            smp = int(d['# Samples'])
        except (ValueError, KeyError):
            continue
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v)}
        out.append((int(r[0]), smp, int(d['Instructions Executed'] or 0), stalls, r[1].strip()[:90]))
total = sum(o[1] for o in out) or 1
print(f'total samples {total}')
for ln, smp, inst, st, src in sorted(out, key=lambda o: -o[1])[:top]:
    s = ' '.join(f'{k}:{v}' for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f'{ln:5d} {100 * smp / total:5.1f}% inst {inst:7d}  [{s}]  {src}')
