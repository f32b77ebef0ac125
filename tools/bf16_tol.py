"""Per-tensor statistics of the bf16 path against the fp64 oracle under the north_star tolerance (GPU box), next to
what PyTorch's own bf16 execution (torch.autocast) of the reference op sequence achieves on the same inputs.

    python tools/bf16_tol.py            # prints one line per tensor
"""
import sys
import torch
sys.path.insert(0, '.')
from oracle import aaconv_oracle as O          # noqa: E402
from tests.helpers import golden_cases, load_case  # noqa: E402
from tools.gpu_check import build_module       # noqa: E402


def autocast_reference(s, p, x, dy):
    """reference op order (oracle port of attn_aug_conv.py:65-97) under torch.autocast(bf16) on the GPU"""
    prm = {k: v.float().cuda().requires_grad_(True) for k, v in p.items()}
    xc = x.float().cuda().requires_grad_(True)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        y = O.aaconv_forward_sequential(xc, prm, s)
    y.float().backward(dy.float().cuda())
    out = {'y': y.float(), 'dx': xc.grad}
    out.update({k: v.grad for k, v in prm.items()})
    return out


def stats(name, got, want, ref):
    got, want, ref = got.double().cpu(), want.double(), ref.double().cpu()
    err, rerr = (got - want).abs(), (ref - want).abs()
    lim = 1e-2 + 2e-2 * want.abs()
    bad = (err > lim).double().mean().item()
    l2 = float((got - want).norm() / want.norm())
    rl2 = float((ref - want).norm() / want.norm())
    print(f'   {name:20s} max|want|={want.abs().max():.2e} rms={want.pow(2).mean().sqrt():.2e} | ours: maxerr={err.max():.2e} '
          f'relL2={l2:.2e} viol={bad*100:.2f}% | autocast: maxerr={rerr.max():.2e} relL2={rl2:.2e} | ratio max={float(err.max()/rerr.max()):.2f} '
          f'L2={l2/rl2:.2f}')


def run(tag, s, p, x, dy, y_ref, g_ref):
    print(tag)
    m = build_module(s, p, 'bf16')
    xc = x.float().cuda().requires_grad_(True)
    y = m(xc)
    y.backward(dy.float().cuda())
    ac = autocast_reference(s, p, x, dy)
    stats('y', y, y_ref, ac['y'])
    stats('dx', xc.grad, g_ref['x'], ac['dx'])
    for n, prm in m.named_parameters():
        stats(n, prm.grad, g_ref[n], ac[n])


for name in golden_cases('f64'):
    s, p, g, t = load_case(name, 'f64')
    run(name, s, p, t['x'], t['dy'], t['y'], g)
for tag, shp, B, hin in (('T3', O.AAConvShape(1024, 512, 3, 2, 160, 48, 8, True, (10, 10)), 4, 20),
                         ('T2', O.AAConvShape(512, 256, 3, 2, 160, 24, 8, True, (20, 20)), 2, 40),
                         ('T1', O.AAConvShape(256, 128, 3, 2, 160, 8, 8, True, (40, 40)), 1, 80)):
    p = O.init_params(shp, seed=0)
    g0 = torch.Generator().manual_seed(1)
    x = torch.relu(torch.randn(B, shp.in_channels, hin, hin, generator=g0))
    dy = torch.randn(B, shp.out_channels, *shp.input_dims, generator=g0)
    y_ref, g_ref = O.aaconv_backward_closed(x.double(), {k: v.double() for k, v in p.items()}, shp, dy.double())
    run(tag, shp, p, x, dy, y_ref, g_ref)
