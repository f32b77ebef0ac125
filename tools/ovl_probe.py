"""Probe (GPU box): does a pinned-host -> device upload on a side stream hide under the AAConv2d step?  (yes: 1.9 ms per step)"""
import os, sys, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chexpert_b200 as cb
from bench import SHAPES
cin, hin, cout, dk, dv = SHAPES['T1']; H = hin // 2; B = 16
dev = torch.device('cuda', 0)
m = cb.AAConv2d(cin, cout, 3, 2, dk, dv, 8, True, (H, H), precision='bf16').to(dev)
xh = torch.relu(torch.randn(B, cin, hin, hin)).pin_memory(); dyh = torch.randn(B, cout, H, H).pin_memory()
x = xh.to(dev); dy = dyh.to(dev)
def step(xin, dyin):
    for p in m.parameters(): p.grad = None
    y = m(xin.detach().requires_grad_(True)); y.backward(dyin); return y
def t(fn, n=5):
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
for _ in range(3): step(x, dy)
xb = torch.empty_like(x)
print('compute only', t(lambda: step(x, dy)))
print('h2d only (x)', t(lambda: xb.copy_(xh, non_blocking=True)))
up = torch.cuda.Stream()
def both():
    with torch.cuda.stream(up):
        xb.copy_(xh, non_blocking=True)
    step(x, dy)
def both_timed(n=5):
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): both()
    torch.cuda.current_stream().wait_stream(up); b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
print('h2d on side stream + compute', both_timed())
t0 = time.perf_counter()
for _ in range(5): step(x, dy)
print('cpu enqueue per step ms', (time.perf_counter() - t0) / 5 * 1e3); torch.cuda.synchronize()
