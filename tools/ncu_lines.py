"""Aggregates an .ncu-rep's source page by CUDA source line: samples, instructions, top stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = {}
fname = ''
for r in rows:
    if r and r[0] == 'File Path': fname = r[1].split('/')[-1]
    if r and r[0] == 'Line No': hdr = r; continue
    if not hdr or len(r) < len(hdr) - 2 or r[2] != '-': continue   # only the per-line summary rows
    try:
        line = int(r[0]); samples = int(r[hdr.index('# Samples')]); inst = int(r[hdr.index('Instructions Executed')])
    except ValueError:
        continue
    stalls = {}
    for i, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h and i < len(r):
            try: v = int(r[i])
            except ValueError: v = 0
            if v: stalls[h[6:]] = v
    key = (fname, line)
    a = agg.setdefault(key, [0, 0, {}, r[1]])
    a[0] += samples; a[1] += inst
    for k, v in stalls.items(): a[2][k] = a[2].get(k, 0) + v
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print(f'total samples {ts}  total warp-instructions {ti}')
for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ' '.join(f'{k}:{v}' for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3])
    print(f'{a[0]:7d} {100*a[0]/max(ts,1):5.1f}% inst {a[1]:9d} {f}:{line:<4} {a[3][:70]:70s} | {st}')
