"""Top stalled SASS lines of one kernel from an ncu source-page CSV:  python tools/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
H = rows[hi]
si = H.index('Warp Stall Sampling (All Samples)')
ex = H.index('Instructions Executed')
reasons = [(i, h) for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) != len(H):
        continue
    try:
        int(r[si])
    except ValueError:
        continue
    data.append(r)
tot = sum(int(r[si]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {h: sum(int(r[i] or 0) for r in data) for i, h in reasons}
print('by reason:', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][si]))[:N]:
    why = max(reasons, key=lambda ih: int(r[ih[0]] or 0))[1]
    print(f'{int(r[si]):7d} {100*int(r[si])/tot:5.1f}%  #{idx:5d} exec={r[ex]:>8s} {why:18s} {r[1].strip()[:100]}')
