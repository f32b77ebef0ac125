"""Run a few AAConv2d fwd+bwd steps at a Transition shape (ncu / sanitizer target; no timing).

python tools/one_step.py [--shape T1] [--steps 2] [--precision bf16] [--batch 16]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chexpert_b200 as cb  # noqa: E402
from bench import SHAPES    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--shape', default='T1')
ap.add_argument('--steps', type=int, default=2)
ap.add_argument('--precision', default='bf16')
ap.add_argument('--batch', type=int, default=16)
a = ap.parse_args()
cin, hin, cout, dk, dv = SHAPES[a.shape]
H = hin // 2
torch.manual_seed(0)
m = cb.AAConv2d(cin, cout, 3, 2, dk, dv, 8, True, (H, H), precision=a.precision)
for mod in m.modules():
    if isinstance(mod, torch.nn.Conv2d):
        torch.nn.init.kaiming_normal_(mod.weight)
m = m.cuda()
x = torch.relu(torch.randn(a.batch, cin, hin, hin, device='cuda')).requires_grad_(True)
dy = torch.randn(a.batch, cout, H, H, device='cuda')
for _ in range(a.steps):
    m.zero_grad(set_to_none=True)
    x.grad = None
    m(x).backward(dy)
torch.cuda.synchronize()
print('ok', float(x.grad.abs().sum()))
