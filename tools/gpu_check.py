"""Verbose parity report (GPU box): CUDA path vs oracle on the golden cases + seeded shapes.

python tools/gpu_check.py [fp32|bf16] [--big]
"""
import sys
import time

import torch

sys.path.insert(0, '.')
from oracle import aaconv_oracle as O          # noqa: E402  (checker only)
from tests.helpers import golden_cases, load_case, rel_err, PARAM_NAMES  # noqa: E402
import chexpert_b200 as cb                       # noqa: E402


def build_module(s, p, precision):
    m = cb.AAConv2d(s.in_channels, s.out_channels, s.kernel_size, s.stride, s.dk, s.dv, s.nh, s.relative,
                    s.input_dims, precision=precision)
    sd = {k: v.float() for k, v in p.items()}
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def run_case(name, s, p, g, t, precision):
    m = build_module(s, p, precision)
    x = t['x'].float().cuda().requires_grad_(True)
    y, w = m(x, return_attn=True)
    y.backward(t['dy'].float().cuda())
    torch.cuda.synchronize()
    out = {'y': rel_err(y.cpu(), t['y']), 'weights': rel_err(w.cpu(), t['weights']), 'gx': rel_err(x.grad.cpu(), g['x'])}
    for n, prm in m.named_parameters():
        out['g.' + n] = rel_err(prm.grad.cpu(), g[n])
    worst = max(out.values())
    print(f'{name:18s} worst={worst:.2e} ' + ' '.join(f'{k}={v:.1e}' for k, v in out.items()), flush=True)
    return worst


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith('-') else 'fp32'
    print(torch.cuda.get_device_name(0), 'precision', precision)
    worst = 0
    for name in golden_cases('f64'):
        s, p, g, t = load_case(name, 'f64')
        worst = max(worst, run_case(name, s, p, g, t, precision))
    if '--big' in sys.argv:
        for tag, shp, B, hin in (('T3', O.AAConvShape(1024, 512, 3, 2, 160, 48, 8, True, (10, 10)), 4, 20),
                                 ('T2', O.AAConvShape(512, 256, 3, 2, 160, 24, 8, True, (20, 20)), 2, 40),
                                 ('T1', O.AAConvShape(256, 128, 3, 2, 160, 8, 8, True, (40, 40)), 1, 80)):
            p = O.init_params(shp, seed=0)
            g0 = torch.Generator().manual_seed(1)
            x = torch.relu(torch.randn(B, shp.in_channels, hin, hin, generator=g0))
            dy = torch.randn(B, shp.out_channels, *shp.input_dims, generator=g0)
            t0 = time.time()
            y_ref, g_ref = O.aaconv_backward_closed(x.double(), {k: v.double() for k, v in p.items()}, shp, dy.double())
            t = {'x': x, 'dy': dy, 'y': y_ref, 'weights': O.aaconv_forward_closed(x.double(), {k: v.double() for k, v in p.items()}, shp, return_weights=True)[1]}
            print(f'  oracle {tag} {time.time()-t0:.1f}s')
            worst = max(worst, run_case(tag, shp, p, g_ref, t, precision))
    print('WORST', worst)


if __name__ == '__main__':
    main()
