"""Per-kernel SASS evidence: counts of the Blackwell-native mnemonics in libaaconv_b200.so.

    python tools/sass_summary.py profiles/<name>.md

UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA load, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
MUFU.EX2 = ex2.approx, HMMA = legacy mma.sync (only the rank-dkh relative-position kernels use it, TF32).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'chexpert_b200', 'csrc', 'libaaconv_b200.so')
PAT = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'SYNCS', 'MUFU.EX2', 'HMMA', 'FFMA', 'LDG', 'STG',
       'LDS', 'STS', 'RED', 'ATOM']


def main():
    dst = sys.argv[1]
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kern = None
    counts = collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = kern.replace('(anonymous namespace)::', '').replace('aaconv::', '').replace('void ', '').split('(')[0]
            counts[kern] = collections.Counter(total=0)
            continue
        if kern is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        counts[kern]['total'] += 1
        for p in PAT:
            if op.startswith(p):
                counts[kern][p] += 1
    out = ['# SASS mnemonic counts per kernel (cuobjdump -sass libaaconv_b200.so, sm_100a)', '',
           'UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, '
           'SYNCS = mbarrier ops, MUFU.EX2 = ex2.approx; HMMA (legacy mma.sync, TF32) only in the rank-dkh relative-position kernels.', '',
           '| kernel | instrs | ' + ' | '.join(PAT) + ' |', '|---|---|' + '---|' * len(PAT)]
    order = sorted(counts.items(), key=lambda kc: -(kc[1]['UTCHMMA'] * 1000 + kc[1]['HMMA'] * 10 + kc[1]['UTMALDG']))
    for k, c in order:
        out.append(f'| {k[:60]} | {c["total"]} | ' + ' | '.join(str(c[p]) if c[p] else '' for p in PAT) + ' |')
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('wrote', dst, len(counts), 'kernels')


if __name__ == '__main__':
    main()
