# does cuDNN keep channels_last through a dense layer's convs, and what do strides look like in backward?
import torch, torch.nn as nn
torch.manual_seed(0)
c1 = nn.Conv2d(256, 128, 1, bias=False).cuda()
c2 = nn.Conv2d(128, 32, 3, padding=1, bias=False).cuda()
x = torch.randn(16, 256, 80, 80, device='cuda', dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
with torch.autocast('cuda', dtype=torch.bfloat16):
    y1 = c1(x); print('conv1 out strides', y1.stride(), y1.is_contiguous(memory_format=torch.channels_last))
    y1r = torch.relu(y1)
    y2 = c2(y1r); print('conv2 out strides', y2.stride())
g = torch.randn(16, 32, 80, 80, device='cuda', dtype=torch.bfloat16)   # NCHW grad, as from a narrow of the chain grad
def hook(name):
    def f(gr): print(name, 'grad strides', gr.stride()); return gr
    return f
y1.register_hook(hook('y1')); x.register_hook(hook('x'))
y2.backward(g)
import time
for fmt in (torch.contiguous_format, torch.channels_last):
    xx = torch.randn(16, 256, 80, 80, device='cuda', dtype=torch.bfloat16).contiguous(memory_format=fmt)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        for _ in range(3): c2(torch.relu(c1(xx)))
        torch.cuda.synchronize(); t=time.time()
        for _ in range(50): c2(torch.relu(c1(xx)))
        torch.cuda.synchronize(); print(fmt, (time.time()-t)/50*1e6, 'us per conv1+relu+conv2 fwd')
