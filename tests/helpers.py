"""Shared helpers for the parity tests (fixtures -> oracle shapes/params)."""
import glob
import os

import numpy as np
import torch

from oracle import aaconv_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
PARAM_NAMES = ('conv.weight', 'in_proj_qkv.weight', 'out_proj.weight', 'key_rel_h', 'key_rel_w')


def golden_cases(tag):
    return sorted(os.path.basename(f)[len('aaconv_'):-len(f'_{tag}.npz')]
                  for f in glob.glob(os.path.join(GOLDEN, f'aaconv_*_{tag}.npz')))


def load_case(name, tag):
    z = np.load(os.path.join(GOLDEN, f'aaconv_{name}_{tag}.npz'))
    B, Cin, Hin, Win, Cout, ks, st, dk, dv, nh, rel = (int(v) for v in z['cfg'])
    H, W = (Hin - 1) // st + 1, (Win - 1) // st + 1
    shape = O.AAConvShape(Cin, Cout, ks, st, dk, dv, nh, bool(rel), (H, W))
    params = {n: torch.from_numpy(z['p.' + n]) for n in PARAM_NAMES if 'p.' + n in z.files}
    grads = {n: torch.from_numpy(z['g.' + n]) for n in PARAM_NAMES if 'g.' + n in z.files}
    grads['x'] = torch.from_numpy(z['gx'])
    t = {k: torch.from_numpy(z[k]) for k in ('x', 'dy', 'y', 'weights')}
    return shape, params, grads, t


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
