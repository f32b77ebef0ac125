"""Instruction mix and top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv`.
usage: python tools/ncu_source_hot.py src.csv [n_top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
H = rows[hdr]
ix = {h: i for i, h in enumerate(H)}
data = [r for r in rows[hdr + 1:] if len(r) >= len(H) and r[ix['# Samples']].isdigit()]
tot_s = sum(int(r[ix['# Samples']]) for r in data)
tot_i = sum(int(r[ix['Instructions Executed']]) for r in data)
print('kernel', rows[0][1][:80] if rows[0] else '', '| samples', tot_s, '| warp instructions', tot_i)
op, ops = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
    o = m.group(2).split('.')[0] if m else '?'
    op[o] += int(r[ix['Instructions Executed']])
    ops[o] += int(r[ix['# Samples']])
for o, c in op.most_common(24):
    print(f'{o:10s} inst {c:10d} {100 * c / tot_i:5.1f}%   samples {100 * ops[o] / max(tot_s, 1):5.1f}%')
print('--- top sampled instructions')
stalls = [h for h in H if h.startswith('stall_') and 'Not Issued' not in h]
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:3]
    print(r[ix['Source']].strip()[:50].ljust(50), r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), st)
agg = collections.Counter()
for r in data:
    for s in stalls:
        agg[s[6:]] += int(r[ix[s]])
print('--- stall totals', [(k, round(100 * v / max(tot_s, 1), 1)) for k, v in agg.most_common(10)])
