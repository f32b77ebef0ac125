"""k-step / accumulator-width ablation of the three small-value-width attention kernels (results are wrong; timing only).
usage: AACONV_ABL_NKS=4 AACONV_ABL_NQ=64 python tools/attn_ablate2.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chexpert_b200 as cb  # noqa: E402
from chexpert_b200 import _lib  # noqa: E402
from bench import SHAPES    # noqa: E402
cin, hin, cout, dk, dv = SHAPES['T1']
H = hin // 2
torch.manual_seed(0)
m = cb.AAConv2d(cin, cout, 3, 2, dk, dv, 8, True, (H, H), precision='bf16').cuda()
x = torch.relu(torch.randn(16, cin, hin, hin, device='cuda')).requires_grad_(True)
dy = torch.randn(16, cout, H, H, device='cuda')
best = {}
for it in range(5):
    m.zero_grad(set_to_none=True); x.grad = None
    torch.cuda.synchronize()
    _lib.profile_begin(0)
    m(x).backward(dy)
    torch.cuda.synchronize()
    for k, v in _lib.profile_end():
        if it: best[k] = min(best.get(k, 1e9), v)
print('NKS', os.environ.get('AACONV_ABL_NKS'), 'NQ', os.environ.get('AACONV_ABL_NQ'), {k: round(v * 1e3, 1) for k, v in best.items() if k.startswith('attn')})
