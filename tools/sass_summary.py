"""Per-kernel SASS evidence: counts of the Blackwell-native mnemonics in libaaconv_b200.so, and (second argument) a SASS
LISTING of the tensor-core issue regions of the hot kernels: every line from 8 instructions before the first UTC*MMA to 8 after
the last one, per kernel.

    python tools/sass_summary.py profiles/<name>.md [profiles/<name>_listing.txt]

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
        m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
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
    if len(sys.argv) > 2:
        listing(sass, sys.argv[2])


HOT = ('pixel_gemm_tc_kernel', 'wgrad_tc_kernel', 'attn_fwd_cc_kernel<2, 1>', 'attn_bwd_dq_cc_kernel<2, 1, 112>',
       'attn_bwd_dkv_cc_kernel<2, 1>', 'attn_fwd_tc_kernel<1>')


def listing(sass, dst):
    """The MMA issue regions (SASS as cuobjdump prints it) of the hot kernels."""
    out, kern, lines = [], None, []

    def flush():
        if kern is None or not any(h in kern for h in HOT):
            return
        idx = [i for i, l in enumerate(lines) if re.search(r'UTC[A-Z]*MMA', l)]
        if not idx:
            return
        lo, hi = max(0, idx[0] - 8), min(len(lines), idx[-1] + 9, idx[0] + 240)
        out.append(f'==== {kern}   ({len(lines)} SASS instructions, {len(idx)} UTC*MMA between lines {idx[0]} and {idx[-1]}; '
                   f'lines {lo}..{hi} shown) ====')
        out.extend(lines[lo:hi])
        out.append('')
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            flush()
            kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = kern.replace('(anonymous namespace)::', '').replace('aaconv::', '').replace('void ', '').split('(')[0]
            lines = []
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,6}\*/', line):
            lines.append(re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', line.rstrip()))
    flush()
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('wrote', dst, sum(1 for l in out if l.startswith('====')), 'kernels')


if __name__ == '__main__':
    main()
