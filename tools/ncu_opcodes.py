"""Aggregates an .ncu-rep's SASS page by opcode: executed warp-instructions and stall samples."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = collections.Counter(); samp = collections.Counter()
for r in rows:
    if r and r[0] in ('Address', '#'): hdr = r; continue
    if hdr is None and r and 'Source' in r: hdr = r; continue
    if not hdr or len(r) < len(hdr) - 2: continue
    try:
        src = r[hdr.index('Source')]; inst = int(r[hdr.index('Instructions Executed')]); s = int(r[hdr.index('# Samples')])
    except (ValueError, IndexError):
        continue
    t = src.split()
    if not t: continue
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    op = '.'.join(op.split('.')[:2]) if op.startswith(('FLO', 'IMAD', 'SHF', 'LOP3', 'ISETP', 'SEL', 'IADD3', 'SHFL', 'LD', 'ST', 'POPC', 'BRA', 'BSSY', 'BSYNC')) else op.split('.')[0]
    agg[op] += inst; samp[op] += s
tot = sum(agg.values()); ts = sum(samp.values())
print('total', tot, 'samples', ts)
for op, n in agg.most_common(40):
    print(f'{op:16s} {n:10d} {100*n/tot:5.1f}%  samples {100*samp[op]/max(1,ts):5.1f}%')
