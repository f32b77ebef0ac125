"""Which results of repeated identical fwd+bwd passes differ from the first pass, and where (debugging aid for the bitwise test)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from oracle import aaconv_oracle as O
from test_gpu_parity import _module
hw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cin = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
shape = O.AAConvShape(cin, 128, 3, 2, 160, 8, 8, True, (hw, hw))
p = O.init_params(shape, seed=2)
m = _module(shape, p, 'bf16')
g0 = torch.Generator().manual_seed(4)
x = torch.relu(torch.randn(B, cin, 2 * hw, 2 * hw, generator=g0)).cuda().requires_grad_(True)
dy = torch.randn(B, 128, hw, hw, generator=g0).cuda()
names = ['y', 'dx'] + ['g.' + n for n, _ in m.named_parameters()]
first = None
for it in range(int(os.environ.get('ITERS', 60))):
    m.zero_grad(set_to_none=True); x.grad = None
    y = m(x); y.backward(dy)
    cur = [y.detach(), x.grad] + [q.grad for q in m.parameters()]
    torch.cuda.synchronize()
    if first is None:
        first = [t.clone() for t in cur]
        continue
    for n, a, b in zip(names, cur, first):
        if not torch.equal(a, b):
            d = (a.float() - b.float()).abs()
            idx = d.flatten().nonzero().flatten()
            print(f'it {it} {n}: {idx.numel()} of {d.numel()} differ, max {d.max().item():.3e} (ref max {b.abs().max().item():.3e}), first idx {idx[:6].tolist()} last {idx[-3:].tolist()} shape {tuple(a.shape)}')
print('done')
