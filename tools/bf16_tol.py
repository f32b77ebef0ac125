"""Per-tensor statistics of the bf16 path against the fixtures under the north_star tolerance (GPU box)."""
import sys
import torch
sys.path.insert(0, '.')
from oracle import aaconv_oracle as O          # noqa: E402
from tests.helpers import golden_cases, load_case  # noqa: E402
from tools.gpu_check import build_module       # noqa: E402


def stats(name, got, want):
    got, want = got.double().cpu(), want.double()
    err = (got - want).abs()
    lim = 1e-2 + 2e-2 * want.abs()
    bad = (err > lim).double().mean().item()
    print(f'   {name:22s} max|want|={want.abs().max():.3e} rms={want.pow(2).mean().sqrt():.3e} maxerr={err.max():.3e} '
          f'viol={bad*100:.3f}% worst_ratio={float((err/lim).max()):.2f}')


def run(tag, s, p, x, dy, y_ref, g_ref):
    print(tag)
    m = build_module(s, p, 'bf16')
    xc = x.float().cuda().requires_grad_(True)
    y = m(xc)
    y.backward(dy.float().cuda())
    stats('y', y, y_ref)
    stats('dx', xc.grad, g_ref['x'])
    for n, prm in m.named_parameters():
        stats(n, prm.grad, g_ref[n])


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
