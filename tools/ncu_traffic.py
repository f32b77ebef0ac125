"""Per-kernel DRAM bytes per launch from an `ncu --set full` capture -> profiles/ncu_dram_bytes.json (what bench.py reports as
`roofline.traffic`; the number comes from the capture file of the build, not from a literal in bench.py).

    ncu -i gpurun_out/<dir>/prof.ncu-rep --page raw --csv > gpurun_out/<dir>/raw.csv
    python tools/ncu_traffic.py gpurun_out/<dir>/raw.csv "<what was captured>"
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# kernel function name -> the launch label bench.py / the library profile uses
LABEL = {'attn_fwd_cc_kernel': 'attn_fwd_cc', 'attn_bwd_dq_cc_kernel': 'attn_bwd_dq_cc', 'attn_bwd_dkv_cc_kernel': 'attn_bwd_dkv_cc',
         'aug_build_fwd_kernel': 'aug_build_fwd', 'rel_bwd_kernel': 'rel_bwd', 'wgrad_tc_kernel': 'conv_qkv_wgrad_tc',
         'pack_x_v4_kernel': 'pack_nhwc_bf16', 'out_bwd_patch_kernel': 'out_bwd_patch', 'out_proj_fwd_kernel': 'out_proj_fwd',
         'rel_bwd_reduce_kernel': 'rel_bwd_reduce', 'out_w_reduce_kernel': 'out_w_reduce', 'wgrad_reduce_kernel': 'wgrad_reduce',
         'pack_wf_kernel': 'pack_wf', 'pack_wd_kernel': 'pack_wd', 'attn_fwd_tc_kernel': 'attn_fwd_tc',
         'attn_bwd_dq_tc_kernel': 'attn_bwd_dq_tc', 'attn_bwd_dkv_tc_kernel': 'attn_bwd_dkv_tc',
         'aug_build_tc_kernel': 'aug_build_tc', 'rel_bwd_tc_kernel': 'rel_bwd_tc'}


def main():
    src, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
    rows = list(csv.reader(open(src, errors='replace')))
    H = rows[0]
    kn, rd, wr = H.index('Kernel Name'), H.index('dram__bytes_read.sum'), H.index('dram__bytes_write.sum')
    units = rows[1]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    acc = collections.OrderedDict()
    order = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(H):
            continue
        name = r[kn].replace('void ', '').split('(')[0].split('<')[0].replace('aaconv::', '').replace('(anonymous namespace)::', '').replace('<unnamed>::', '').replace('unnamed>::', '')
        label = LABEL.get(name, name)
        if name == 'pixel_gemm_tc_kernel':     # launched three ways per step, in this order: fprop, dgrad (wgrad has its own kernel)
            label = ('conv_qkv_fprop_tc', 'conv_qkv_dgrad_tc')[order[name] % 2]
            order[name] += 1
        b = float(r[rd].replace(',', '')) * scale.get(units[rd], 1.0) + float(r[wr].replace(',', '')) * scale.get(units[wr], 1.0)
        a = acc.setdefault(label, [0.0, 0])
        a[0] += b
        a[1] += 1
    out = {'source': f'ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch; {what}; {os.path.relpath(src, ROOT)}',
           'kernels': {k: v[0] / v[1] for k, v in acc.items()}, 'launches_captured': {k: v[1] for k, v in acc.items()}}
    dst = os.path.join(ROOT, 'profiles', 'ncu_dram_bytes.json')
    json.dump(out, open(dst, 'w'), indent=1)
    print('wrote', dst)
    for k, v in out['kernels'].items():
        print(f'  {k:28s} {v / 1e6:8.1f} MB/launch')


if __name__ == '__main__':
    main()
