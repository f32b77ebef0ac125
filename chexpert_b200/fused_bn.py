"""BatchNorm2d (training mode) + ReLU as one op that reads its input through a batch stride -- the normalisation of a channel
slice of the pre-allocated DenseBlock feature buffer (torchvision densenet.py:36-41 under
models/attn_aug_conv.py:479-482).  Parameters, buffers and the arithmetic are nn.BatchNorm2d's (same state_dict, same running
statistics update); eval mode and CPU tensors go through the module's own forward.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib
from .aaconv import _ptr, _stream

_DT = {torch.float32: _lib.FP32, torch.bfloat16: _lib.BF16}


def _strided_ok(x):
    """(B, C, H, W) with dense (C, H, W) inside every sample: a contiguous tensor or a channel slice of the feature buffer."""
    B, C, H, W = x.shape
    return x.stride(3) == 1 and x.stride(2) == W and x.stride(1) == H * W and x.stride(0) >= C * H * W


class _BNReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, stats=None):
        lib = _lib.load()
        B, C, H, W = x.shape
        xs = x.detach()
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        with torch.cuda.device(x.device):
            y = torch.empty(B, C, H, W, device=x.device, dtype=x.dtype)
            saved = torch.empty(C, 2, device=x.device, dtype=torch.float32)
            # stats = (buffer, n_valid): plane statistics shared along a dense block -- channels [0, n_valid) of THIS input were
            # already reduced into `buffer` by the previous layer (channel-major (C_total, B) float2: a prefix is a valid (C, B) table)
            if stats is not None and stats[0].numel() * 4 >= 8 * B * C:
                ws, valid = stats[0], min(int(stats[1]), C)
            else:
                ws, valid = torch.empty(lib.aaconv_bn_relu_workspace_bytes(B, C) // 4, device=x.device, dtype=torch.float32), 0
            _lib.check(lib.aaconv_bn_relu_forward(_ptr(xs), _DT[x.dtype], B, C, H * W, xs.stride(0), _ptr(w), _ptr(b),
                                                  _ptr(running_mean), _ptr(running_var), float(momentum), float(eps), _ptr(y),
                                                  _ptr(saved), _ptr(ws), valid, _stream()), 'aaconv_bn_relu_forward')
        ctx.save_for_backward(xs, saved, w, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        xs, saved, w, b = ctx.saved_tensors
        B, C, H, W = xs.shape
        g = dy.detach().to(xs.dtype).contiguous()
        need = ctx.needs_input_grad
        with torch.cuda.device(xs.device):
            dx = torch.empty(B, C, H, W, device=xs.device, dtype=xs.dtype) if need[0] else None
            dw = torch.empty(C, device=xs.device, dtype=torch.float32) if need[1] else None
            db = torch.empty(C, device=xs.device, dtype=torch.float32) if need[2] else None
            ws = torch.empty(lib.aaconv_bn_relu_workspace_bytes(B, C), device=xs.device, dtype=torch.uint8)
            _lib.check(lib.aaconv_bn_relu_backward(_ptr(xs), _DT[xs.dtype], B, C, H * W, xs.stride(0), _ptr(g), _ptr(saved), _ptr(w),
                                                   _ptr(b), _ptr(dx), _ptr(dw), _ptr(db), _ptr(ws), _stream()), 'aaconv_bn_relu_backward')
        return dx, dw, db, None, None, None, None, None


def _fresh_alias(t):
    """A tensor over the same memory with its own autograd identity and version counter (later writes into other channels of the
    same feature buffer must not invalidate what this op saved)."""
    return torch.empty(0, dtype=t.dtype, device=t.device).set_(t.untyped_storage(), t.storage_offset(), t.shape, t.stride())


class _TapBNReLU(torch.autograd.Function):
    """(feats, relu(bn(feats))) for the concatenated features of a dense block: the first output IS the input (handed on to the
    next layer's concatenation), so that backward sees both consumers' gradients at once and ADDS this layer's dx into the
    gradient of the concatenation in place (aaconv_bn_relu_backward_acc) -- instead of a dense dx and autograd's own add."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, stats=None):
        y = _BNReLU.forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, stats)
        return _fresh_alias(x.detach()), y

    @staticmethod
    def backward(ctx, g_feats, dy):
        if dy is None:
            return g_feats, None, None, None, None, None, None, None
        xs, saved, w, b = ctx.saved_tensors
        B, C, H, W = xs.shape
        ok = (g_feats is not None and ctx.needs_input_grad[0] and g_feats.dtype == xs.dtype and tuple(g_feats.shape) == (B, C, H, W)
              and g_feats.stride(3) == 1 and g_feats.stride(2) == W and g_feats.stride(1) == H * W and g_feats.stride(0) >= C * H * W)
        if not ok:
            dx, dw, db = _BNReLU.backward(ctx, dy)[:3]
            if g_feats is not None and dx is not None:
                dx = dx + g_feats
            elif dx is None:
                dx = g_feats
            return dx, dw, db, None, None, None, None, None
        lib = _lib.load()
        g = dy.detach().to(xs.dtype).contiguous()
        need = ctx.needs_input_grad
        with torch.cuda.device(xs.device):
            dw = torch.empty(C, device=xs.device, dtype=torch.float32) if need[1] else None
            db = torch.empty(C, device=xs.device, dtype=torch.float32) if need[2] else None
            ws = torch.empty(lib.aaconv_bn_relu_workspace_bytes(B, C), device=xs.device, dtype=torch.uint8)
            _lib.check(lib.aaconv_bn_relu_backward_acc(_ptr(xs), _DT[xs.dtype], B, C, H * W, xs.stride(0), _ptr(g), _ptr(saved), _ptr(w),
                                                       _ptr(b), _ptr(g_feats), g_feats.stride(0), _ptr(dw), _ptr(db), _ptr(ws),
                                                       _stream()), 'aaconv_bn_relu_backward_acc')
        return g_feats, dw, db, None, None, None, None, None


def _fused_ok(bn, x):
    return (bn.training and x.is_cuda and x.dim() == 4 and x.dtype in _DT and bn.affine and bn.track_running_stats
            and bn.momentum is not None and _strided_ok(x) and x.shape[0] <= 65535)


def tap_bn_relu(bn, x, stats=None):
    """-> (x handed on, relu(bn(x))): bn_relu for an input that has a second consumer (the concatenation of a dense block)."""
    if not (_fused_ok(bn, x) and torch.is_grad_enabled() and x.requires_grad):
        return x, bn_relu(bn, x, stats)
    if bn.num_batches_tracked is not None and not getattr(bn, '_nbt_batched', False):
        bn.num_batches_tracked.add_(1)
    return _TapBNReLU.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, stats)


def bn_relu(bn, x, stats=None):
    """relu(bn(x)) for an nn.BatchNorm2d `bn`; the fused strided kernels when it is training on CUDA, the modules otherwise.
    ``stats = (float32 buffer of >= 2*B*C elements, n_valid_channels)`` shares the per-plane statistics between the layers of a
    dense block (see aaconv_bn_relu_forward)."""
    fused = (bn.training and x.is_cuda and x.dim() == 4 and x.dtype in _DT and bn.affine and bn.track_running_stats
             and bn.momentum is not None and _strided_ok(x) and x.shape[0] <= 65535)
    if not fused:
        return F.relu(bn(x), inplace=True)
    if bn.num_batches_tracked is not None and not getattr(bn, '_nbt_batched', False):   # TrainStep bumps all counters in one launch
        bn.num_batches_tracked.add_(1)
    return _BNReLU.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, stats)


# ------------------------------------------------------------------------------------------------
# channels-last side of a dense layer (csrc/bn_cl.cu): the convolutions see NHWC tensors, the feature buffer stays NCHW
# ------------------------------------------------------------------------------------------------
_CL = torch.channels_last


def _is_cl(t):
    return t.dim() == 4 and t.is_contiguous(memory_format=_CL)


def cl_ok(bn, x):
    """The channels-last kernels cover training-mode BatchNorm2d on CUDA with C and H*W multiples of 4."""
    B, C, H, W = x.shape
    return (bn.training and x.is_cuda and x.dtype in _DT and bn.affine and bn.track_running_stats and bn.momentum is not None
            and C % 4 == 0 and (H * W) % 4 == 0 and B <= 65535)


def _bump(bn):
    if bn.num_batches_tracked is not None and not getattr(bn, '_nbt_batched', False):
        bn.num_batches_tracked.add_(1)


class _TapBNReLUCL(torch.autograd.Function):
    """_TapBNReLU with an NHWC output: x = channel prefix of the NCHW feature buffer -> (x handed on, relu(bn(x)) channels-last);
    backward takes the NHWC gradient and adds dx into the NCHW gradient of the concatenation in place."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, stats):
        lib = _lib.load()
        B, C, H, W = x.shape
        xs = x.detach()
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        with torch.cuda.device(x.device):
            y = torch.empty((B, C, H, W), device=x.device, dtype=x.dtype, memory_format=_CL)
            saved = torch.empty(C, 2, device=x.device, dtype=torch.float32)
            if stats is not None and stats[0].numel() * 4 >= 8 * B * C:
                pst, valid = stats[0], min(int(stats[1]), C)
            else:
                pst, valid = torch.empty(lib.aaconv_bn_relu_workspace_bytes(B, C) // 4, device=x.device, dtype=torch.float32), 0
            ws = torch.empty(lib.aaconv_bn_relu_cl_workspace_bytes(B, C, H * W), device=x.device, dtype=torch.uint8)
            G, bpb = ctypes.c_int(0), ctypes.c_int(0)
            _lib.check(lib.aaconv_bn_stats_nchw(_ptr(xs), _DT[x.dtype], B, C, H * W, xs.stride(0), _ptr(pst), valid, ctypes.byref(G),
                                                ctypes.byref(bpb), _stream()), 'aaconv_bn_stats_nchw')
            _lib.check(lib.aaconv_bn_relu_cl_forward(_ptr(xs), _DT[x.dtype], B, C, H * W, 0, xs.stride(0), _ptr(w), _ptr(b),
                                                     _ptr(running_mean), _ptr(running_var), float(momentum), float(eps), _ptr(y), _ptr(saved),
                                                     _ptr(ws), _ptr(pst), G.value, bpb.value, _stream()), 'aaconv_bn_relu_cl_forward')
        ctx.save_for_backward(xs, saved, w, b)
        return _fresh_alias(xs), y

    @staticmethod
    def backward(ctx, g_feats, dy):
        xs, saved, w, b = ctx.saved_tensors
        if dy is None:
            return g_feats, None, None, None, None, None, None, None
        lib = _lib.load()
        B, C, H, W = xs.shape
        g = dy.detach().to(xs.dtype).contiguous(memory_format=_CL)
        need = ctx.needs_input_grad
        acc = (g_feats is not None and g_feats.dtype == xs.dtype and tuple(g_feats.shape) == (B, C, H, W) and g_feats.stride(3) == 1
               and g_feats.stride(2) == W and g_feats.stride(1) == H * W and g_feats.stride(0) >= C * H * W
               and g_feats.stride(0) % 4 == 0 and g_feats.data_ptr() % (4 * g_feats.element_size()) == 0)
        with torch.cuda.device(xs.device):
            dx = g_feats if acc else torch.empty(B, C, H, W, device=xs.device, dtype=xs.dtype)
            dw = torch.empty(C, device=xs.device, dtype=torch.float32) if need[1] else None
            db = torch.empty(C, device=xs.device, dtype=torch.float32) if need[2] else None
            ws = torch.empty(lib.aaconv_bn_relu_cl_workspace_bytes(B, C, H * W), device=xs.device, dtype=torch.uint8)
            _lib.check(lib.aaconv_bn_relu_cl_backward(_ptr(xs), _DT[xs.dtype], B, C, H * W, 0, xs.stride(0), _ptr(g), _ptr(saved), _ptr(w),
                                                      _ptr(b), _ptr(dx), dx.stride(0), int(acc), _ptr(dw), _ptr(db), _ptr(ws), _stream()),
                       'aaconv_bn_relu_cl_backward')
        if not acc and g_feats is not None:
            dx = dx + g_feats
        return dx, dw, db, None, None, None, None, None


class _BNReLUCL(torch.autograd.Function):
    """relu(bn(x)) for a channels-last x (the bottleneck between conv1 and conv2), channels-last out."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps):
        lib = _lib.load()
        B, C, H, W = x.shape
        xs = x.detach()
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        with torch.cuda.device(x.device):
            y = torch.empty((B, C, H, W), device=x.device, dtype=x.dtype, memory_format=_CL)
            saved = torch.empty(C, 2, device=x.device, dtype=torch.float32)
            ws = torch.empty(lib.aaconv_bn_relu_cl_workspace_bytes(B, C, H * W), device=x.device, dtype=torch.uint8)
            _lib.check(lib.aaconv_bn_relu_cl_forward(_ptr(xs), _DT[x.dtype], B, C, H * W, 1, C * H * W, _ptr(w), _ptr(b), _ptr(running_mean),
                                                     _ptr(running_var), float(momentum), float(eps), _ptr(y), _ptr(saved), _ptr(ws), None, 0, 0,
                                                     _stream()), 'aaconv_bn_relu_cl_forward')
        ctx.save_for_backward(xs, saved, w, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        xs, saved, w, b = ctx.saved_tensors
        B, C, H, W = xs.shape
        g = dy.detach().to(xs.dtype).contiguous(memory_format=_CL)
        need = ctx.needs_input_grad
        with torch.cuda.device(xs.device):
            dx = torch.empty((B, C, H, W), device=xs.device, dtype=xs.dtype, memory_format=_CL) if need[0] else None
            dw = torch.empty(C, device=xs.device, dtype=torch.float32) if need[1] else None
            db = torch.empty(C, device=xs.device, dtype=torch.float32) if need[2] else None
            ws = torch.empty(lib.aaconv_bn_relu_cl_workspace_bytes(B, C, H * W), device=xs.device, dtype=torch.uint8)
            _lib.check(lib.aaconv_bn_relu_cl_backward(_ptr(xs), _DT[xs.dtype], B, C, H * W, 1, C * H * W, _ptr(g), _ptr(saved), _ptr(w), _ptr(b),
                                                      _ptr(dx), C * H * W, 0, _ptr(dw), _ptr(db), _ptr(ws), _stream()),
                       'aaconv_bn_relu_cl_backward')
        return dx, dw, db, None, None, None, None


def tap_bn_relu_cl(bn, x, stats=None):
    _bump(bn)
    return _TapBNReLUCL.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, stats)


def bn_relu_cl(bn, x):
    _bump(bn)
    return _BNReLUCL.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps)


def slice_layout(src, dst, to_nchw):
    """Channel-slice layout change through the C ABI: src NHWC dense -> dst NCHW (batch stride) when to_nchw, else the reverse."""
    lib = _lib.load()
    nchw, cl = (dst, src) if to_nchw else (src, dst)
    B, C, H, W = nchw.shape
    with torch.cuda.device(src.device):
        _lib.check(lib.aaconv_slice_layout(_ptr(src), _ptr(dst), _DT[src.dtype], B, C, H * W, nchw.stride(0), int(to_nchw), _stream()),
                   'aaconv_slice_layout')
    return dst
