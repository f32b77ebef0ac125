"""Dump the clock64 timeline CTA (0,0) of attn_bwd_dq_tc_kernel records through the debug hook (GPU box).

python tools/attn_timeline.py  -> per-tile stamps relative to the first one (cycles)
events: 0 MMA: K tile landed   1 MMA: S,dP issued   2 MMA: p_ready seen   3 WG: start waiting for S   4 WG: S ready
        5 WG: first TMEM load done   6 WG: arrived p_ready   7 [wg, ...] epilogue: loop end / final seen / stores done
"""
import ctypes
import os
import sys
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
buf = torch.zeros(12 * 64, dtype=torch.int64, device='cuda')
lib = _lib.load()
if len(sys.argv) > 1:
    lib.aaconv_debug_set_mode(int(sys.argv[1]))
for it in range(3):
    m.zero_grad(set_to_none=True)
    x.grad = None
    y = m(x)
    if it == 2:
        lib.aaconv_debug_set_timeline(ctypes.c_void_p(buf.data_ptr()))
    y.backward(dy)
torch.cuda.synchronize()
lib.aaconv_debug_set_timeline(None)
t = buf.cpu().reshape(12, 64)
t0 = int(t[t > 0].min())
names = ['S:top', 'S:K landed', 'S:slot free', 'S:issued', 'G:p_ready', 'wg:wait S', 'wg:S ready', 'wg:arrive']
order = [8, 0, 9, 1, 2, 3, 4, 6]
print('tile ' + ' '.join(f'{n:>13s}' for n in names))
for j in range(26):
    print(f'{j:4d} ' + ' '.join(f'{(int(t[e, j]) - t0) if t[e, j] > 0 else -1:13d}' for e in order))
print('CTA [entry, init done, stat landed, A in TMEM, loop end, final seen, stores done, after sync]:', [int(v) - t0 for v in t[10, :8]])
