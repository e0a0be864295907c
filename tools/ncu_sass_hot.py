"""Hottest SASS instructions of one kernel in an .ncu-rep (source page, SASS view): samples, executions, stalls."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, data, khdr = None, None, [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = r[1]; hdr = None; continue
    if r and r[0] == 'Address':
        hdr = r; continue
    if cur and pat in cur and hdr and len(r) >= 10:
        if khdr is None: khdr = hdr
        if hdr is khdr: data.append(r)   # first matching launch only
    if data and cur and pat not in cur:
        break
hdr = khdr
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
inst = sum(int(r[ix['Instructions Executed']]) for r in data)
print('kernel', pat, 'samples', tot, 'warp-instructions', inst, 'sass lines', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print('stall totals:', ' '.join(f'{k[6:]}:{v}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
if '--window' in sys.argv:   # print a contiguous window around the hottest instruction
    hot = max(range(len(data)), key=lambda i: int(data[i][ix['# Samples']]))
    lo, hi = max(0, hot - 120), min(len(data), hot + 120)
    for i in range(lo, hi):
        r = data[i]
        st = ' '.join(f'{s[6:]}:{r[ix[s]]}' for s in stalls if r[ix[s]] not in ('0', ''))
        print(f'{int(r[ix["# Samples"]]):6d} x{int(r[ix["Instructions Executed"]]):9d} {r[1][:60]:60s} {st}')
else:
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top]:
        st = ' '.join(f'{s[6:]}:{r[ix[s]]}' for s in stalls if r[ix[s]] not in ('0', ''))
        print(f'{int(r[ix["# Samples"]]):6d} x{int(r[ix["Instructions Executed"]]):9d} {r[1][:60]:60s} {st}')
