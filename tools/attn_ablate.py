"""Ablation timing of attn_bwd_dq_cc_kernel (GPU box): which resource bounds the kernel?
modes: 0 normal | 1 no MUFU | 2 no global traffic after the first tiles | 4 no gradient MMAs | 8 no math | combinations"""
import ctypes, os, sys
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
lib = _lib.load()
for mode in (0, 1, 2, 4, 8, 3, 6, 7, 10, 12, 14, 15):
    lib.aaconv_debug_set_mode(mode)
    ts = []
    tf = []
    for it in range(4):
        m.zero_grad(set_to_none=True); x.grad = None
        torch.cuda.synchronize()
        _lib.profile_begin(torch.cuda.current_stream().cuda_stream)
        y = m(x)
        torch.cuda.synchronize()
        tf.append(dict(_lib.profile_end()).get('attn_fwd_cc', -1))
        _lib.profile_begin(torch.cuda.current_stream().cuda_stream)
        y.backward(dy)
        torch.cuda.synchronize()
        prof = dict(_lib.profile_end())
        ts.append(prof.get('attn_bwd_dq_cc', -1))
    print(f'mode {mode:2d}: attn_bwd_dq_cc {min(ts[1:])*1e3:8.1f} us   attn_fwd_cc {min(tf[1:])*1e3:8.1f} us')
lib.aaconv_debug_set_mode(0)
