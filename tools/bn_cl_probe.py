"""One forward + backward of the channels-last BN+ReLU ops at a dense-layer shape (ncu target).  usage: bn_cl_probe.py C HW_side"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chexpert_b200.fused_bn import tap_bn_relu_cl, bn_relu_cl
C = int(sys.argv[1]) if len(sys.argv) > 1 else 160
S = int(sys.argv[2]) if len(sys.argv) > 2 else 80
B, Ctot = 16, 256
buf = torch.randn(B, Ctot, S, S, device='cuda').bfloat16()
bn1, bn2 = torch.nn.BatchNorm2d(C).cuda(), torch.nn.BatchNorm2d(128).cuda()
stats = torch.empty(2 * B * Ctot, device='cuda')
gbuf = torch.randn(B, Ctot, S, S, device='cuda').bfloat16()
z = torch.randn(B, 128, S, S, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
for _ in range(3):
    x = buf[:, :C].detach().requires_grad_(True)
    f, y = tap_bn_relu_cl(bn1, x, (stats, 0))
    torch.autograd.backward([f, y], [gbuf[:, :C], torch.randn_like(y)])
    y2 = bn_relu_cl(bn2, z)
    y2.backward(torch.randn_like(y2))
torch.cuda.synchronize()
print('ok')
