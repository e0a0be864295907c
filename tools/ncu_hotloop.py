"""Instructions of one kernel in an .ncu-rep whose execution count is at least `frac` of the maximum, in address order
(i.e. the hot loop), with samples and the top stall reason; then totals per opcode class.
usage: python tools/ncu_hotloop.py rep.ncu-rep "<kernel name substring>" [frac=0.5] [--list]"""
import collections, csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
frac = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 0.5
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, data = None, None, []
done = False
for r in rows:
    if r and r[0] == 'Kernel Name':
        if data: break
        cur = r[1]; hdr = None; continue
    if r and r[0] == 'Address':
        hdr = r; continue
    if cur and pat in cur and hdr and len(r) >= 10:
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
ex = [int(r[ix['Instructions Executed']]) for r in data]
mx = max(ex)
tot = sum(ex)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
hot = [(r, e) for r, e in zip(data, ex) if e >= frac * mx]
print(f'kernel {pat}: {len(data)} SASS lines, {tot} warp-instructions, max per line {mx}; {len(hot)} lines >= {frac} of max, '
      f'{sum(e for _, e in hot)} warp-instructions ({100 * sum(e for _, e in hot) / tot:.1f} %)')
cls = collections.Counter()
for r, e in hot:
    t = r[ix['Source']].split()
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    cls[op.split('.')[0]] += e
print('per opcode (executions / max line):', ' '.join(f'{k}:{v / mx:.1f}' for k, v in cls.most_common()))
if '--list' in sys.argv:
    for r, e in hot:
        st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        print(f'{int(r[ix["# Samples"]]):6d} x{e:9d} {r[ix["Source"]][:70]:70s} ' + ' '.join(f'{n}:{v}' for v, n in st if v))
