"""BCE-with-logits on the competition classes, as one CUDA launch.

Replaces ``nn.BCEWithLogitsLoss(reduction='none')(out, target)`` + ``.sum(1).mean(0)``
(chexpert.py:530,160) for training and the element losses kept by evaluate() (chexpert.py:205).
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from .aaconv import _ptr, _stream

# dataset.py:20-25 -- positions of the 5 competition labels inside the 14 CSV label columns
COMPETITION_INDEX = (8, 2, 6, 5, 10)


def _launch(z, targets, cols, want_el, want_loss, want_dz, grad_scale=None):
    if not z.is_cuda:
        raise RuntimeError('chexpert_b200 loss runs on CUDA only; there is no CPU fallback')
    lib = _lib.load()
    z = z.detach().float().contiguous()
    t = targets.detach().to(device=z.device, dtype=torch.float32).contiguous()
    B, C = z.shape
    if cols is None and tuple(t.shape) != (B, C):
        raise RuntimeError(f'targets {tuple(t.shape)} do not match logits {(B, C)}')
    with torch.cuda.device(z.device):
        el = torch.empty_like(z) if want_el else None
        loss = torch.empty((), device=z.device, dtype=torch.float32) if want_loss else None
        dz = torch.empty_like(z) if want_dz else None
        _lib.check(lib.aaconv_bce_forward_backward(_ptr(z), _ptr(t), t.shape[1], _ptr(cols), B, C, _ptr(el), _ptr(loss),
                                                   _ptr(dz), _ptr(grad_scale), _stream()), 'aaconv_bce_forward_backward')
    return el, loss, dz


class _BCETrainLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, targets, cols):
        _, loss, dz = _launch(z, targets, cols, False, True, z.requires_grad)
        ctx.save_for_backward(dz)
        ctx.in_dtype = z.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return (dz * g).to(ctx.in_dtype), None, None


class _BCEElementLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, targets, cols):
        el, _, dz = _launch(z, targets, cols, True, False, z.requires_grad)
        ctx.save_for_backward(dz)
        ctx.B = z.shape[0]
        return el

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors            # d(sum_c mean_b el)/dz = (sigmoid(z)-t)/B
        return dz * ctx.B * g, None, None


class BCEWithLogitsLoss(nn.Module):
    """``reduction='none'`` -> (B,C) element losses (what chexpert.py:530 builds);
    ``reduction='train'`` -> scalar ``el.sum(1).mean(0)`` fused (chexpert.py:160).

    ``raw_labels=True`` accepts the 14-wide CheXpert label rows in {nan,-1,0,1} and applies the U-Ones
    policy and competition-column selection (dataset.py:25,139,142) inside the kernel.
    """

    def __init__(self, reduction='none', raw_labels=False):
        super().__init__()
        if reduction not in ('none', 'train'):
            raise ValueError("reduction must be 'none' or 'train'")
        self.reduction, self.raw_labels = reduction, raw_labels
        self.register_buffer('cols', torch.tensor(COMPETITION_INDEX, dtype=torch.int32), persistent=False)

    def forward(self, z, targets):
        if self.raw_labels:
            # the kernel reads targets[b, cols[c]] for c < C: check the shapes on the host (cols lives on the device)
            if z.dim() != 2 or z.shape[1] != len(COMPETITION_INDEX):
                raise RuntimeError(f'raw_labels=True needs (B, {len(COMPETITION_INDEX)}) logits (dataset.py:25), got {tuple(z.shape)}')
            if targets.dim() != 2 or targets.shape[0] != z.shape[0] or targets.shape[1] <= max(COMPETITION_INDEX):
                raise RuntimeError(f'raw_labels=True needs (B, >= {max(COMPETITION_INDEX) + 1}) label rows (dataset.py:20-23), '
                                   f'got {tuple(targets.shape)}')
        cols = self.cols.to(z.device) if self.raw_labels else None
        fn = _BCETrainLoss if self.reduction == 'train' else _BCEElementLoss
        return fn.apply(z, targets, cols)
