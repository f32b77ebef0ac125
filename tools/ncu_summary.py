"""Turn the ncu outputs of one GPU visit into a small tracked summary under profiles/.

    python tools/ncu_summary.py <gpurun_out/dir> <profiles/name.md> [--title "..."]

Reads <dir>/launches.csv (ncu --metrics gpu__time_duration.sum launch list) and, if present, <dir>/raw.csv
(`ncu -i prof.ncu-rep --page raw --csv`) and <dir>/bench.json.  Writes a markdown table per source.
"""
import collections
import csv
import json
import os
import sys

RAW_COLS = [
    ('gpu__time_duration.sum', 'us'),
    ('dram__bytes_read.sum', 'MB rd'),
    ('dram__bytes_write.sum', 'MB wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps %'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
]


def short(name):
    name = name.replace('void ', '')
    return name.split('(')[0][:44]


def launches_table(path):
    rows = list(csv.reader(open(path, errors='replace')))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    H = rows[hdr]
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(H):
            continue
        d = dict(zip(H, r))
        a = agg.setdefault(short(d['Kernel Name']), [0, 0.0, d['Grid Size'], d['Block Size']])
        a[0] += 1
        a[1] += float(d['Metric Value'].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    out = ['| kernel | launches | total us | share | grid | block |', '|---|---|---|---|---|---|']
    for k, v in agg.items():
        out.append(f'| {k} | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f}% | {v[2]} | {v[3]} |')
    out.append(f'| **all** | {sum(v[0] for v in agg.values())} | {tot / 1e3:.1f} | 100% | | |')
    return '\n'.join(out)


def raw_table(path):
    rows = list(csv.reader(open(path, errors='replace')))
    H = rows[0]
    idx = [(H.index(c), lab) for c, lab in RAW_COLS if c in H]
    kn = H.index('Kernel Name')
    out = ['| kernel | ' + ' | '.join(lab for _, lab in idx) + ' |', '|---|' + '---|' * len(idx)]
    for r in rows[2:]:
        if len(r) < len(H):
            continue
        vals = []
        for i, _ in idx:
            try:
                vals.append(f'{float(r[i].replace(",", "")):.1f}')
            except ValueError:
                vals.append(r[i])
        out.append(f'| {short(r[kn])} | ' + ' | '.join(vals) + ' |')
    return '\n'.join(out)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    title = sys.argv[sys.argv.index('--title') + 1] if '--title' in sys.argv else os.path.basename(dst)
    parts = [f'# {title}', '']
    b = os.path.join(src, 'bench.json')
    if os.path.exists(b):
        try:
            line = json.loads(open(b).read().strip().splitlines()[-1])
            parts += ['## bench.py line of the same build (not under ncu)', '', '```json', json.dumps(line), '```', '']
        except Exception:
            pass
    p = os.path.join(src, 'launches.csv')
    if os.path.exists(p):
        parts += ['## launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised)', '',
                  launches_table(p), '']
    p = os.path.join(src, 'raw.csv')
    if os.path.exists(p):
        parts += ['## `ncu --set full --clock-control none --import-source on`, raw page, one row per captured launch', '',
                  raw_table(p), '']
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, 'w').write('\n'.join(parts))
    print('wrote', dst)


if __name__ == '__main__':
    main()
