"""A/B of one stage implementation selected by an environment variable (the library reads it once per process).

    python tools/stage_ab.py --env AACONV_AUG_BUILD --a legacy --b tc [--shapes T1,T2,T3,T1_512] [--batch 4]

Runs the same seeded AAConv2d fwd+bwd (bf16 mode) in two child processes, one per setting, dumps y, the attention map of a
small case, dx and the parameter gradients, compares them (max-abs error relative to the tensor's max-abs) and prints the
per-kernel device times of both runs.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(shape, batch, out_path):
    import torch
    sys.path.insert(0, ROOT)
    import chexpert_b200 as cb
    from chexpert_b200 import _lib
    from bench import SHAPES
    cin, hin, cout, dk, dv = SHAPES[shape]
    H = hin // 2
    torch.manual_seed(0)
    m = cb.AAConv2d(cin, cout, 3, 2, dk, dv, 8, True, (H, H), precision='bf16')
    for mod in m.modules():
        if isinstance(mod, torch.nn.Conv2d):
            torch.nn.init.kaiming_normal_(mod.weight)
    m = m.cuda()
    x = torch.relu(torch.randn(batch, cin, hin, hin, device='cuda')).requires_grad_(True)
    dy = torch.randn(batch, cout, H, H, device='cuda')
    res = {}
    want_map = H * H <= 1600 and batch <= 2
    for it in range(3):
        m.zero_grad(set_to_none=True)
        x.grad = None
        if it == 2:
            torch.cuda.synchronize()
            _lib.profile_begin(torch.cuda.current_stream().cuda_stream)
        if want_map and it == 0:
            y, w = m(x, return_attn=True)
            res['attn'] = w.detach().cpu()
        else:
            y = m(x)
        y.backward(dy)
    torch.cuda.synchronize()
    times = {}
    for name, t in _lib.profile_end():
        times[name] = times.get(name, 0.0) + t
    res['y'] = y.detach().cpu()
    res['dx'] = x.grad.cpu()
    for n, p in m.named_parameters():
        res['g.' + n] = p.grad.cpu()
    torch.save({'res': res, 'times': times}, out_path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--env', default='AACONV_AUG_BUILD')
    ap.add_argument('--a', default='legacy')
    ap.add_argument('--b', default='tc')
    ap.add_argument('--shapes', default='T1,T2,T3,T1_512')
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--child', nargs=3)
    a = ap.parse_args()
    if a.child:
        child(a.child[0], int(a.child[1]), a.child[2])
        return
    import torch
    worst = 0.0
    for shape in a.shapes.split(','):
        outs = []
        for setting in (a.a, a.b):
            f = tempfile.NamedTemporaryFile(suffix='.pt', delete=False).name
            env = dict(os.environ)
            env[a.env] = setting
            subprocess.run([sys.executable, os.path.abspath(__file__), '--child', shape, str(a.batch), f], env=env, check=True)
            outs.append(torch.load(f))
            os.unlink(f)
        ra, rb = outs[0]['res'], outs[1]['res']
        errs = {}
        for k in ra:
            d = (ra[k].double() - rb[k].double()).abs().max().item()
            errs[k] = d / (ra[k].double().abs().max().item() + 1e-30)
        worst = max(worst, max(errs.values()))
        ta, tb = outs[0]['times'], outs[1]['times']
        diff = {k: (round(ta.get(k, 0) * 1e3, 1), round(tb.get(k, 0) * 1e3, 1)) for k in sorted(set(ta) | set(tb))
                if abs(ta.get(k, 0) - tb.get(k, 0)) > 0.002 or k not in ta or k not in tb}
        print(json.dumps({'shape': shape, 'batch': a.batch, 'rel_max_err': {k: float('%.3g' % v) for k, v in errs.items()},
                          'us_changed(a,b)': diff, 'step_us(a,b)': (round(sum(ta.values()) * 1e3, 1), round(sum(tb.values()) * 1e3, 1))}))
    print('worst', worst)


if __name__ == '__main__':
    main()
