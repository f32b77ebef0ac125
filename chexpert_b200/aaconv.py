"""Drop-in AAConv2d backed by the sm_100a CUDA extension.

Mirrors the reference module contract (models/attn_aug_conv.py:19-100): same constructor, same
``forward(x)``, same state_dict keys (``conv.weight``, ``in_proj_qkv.weight``, ``out_proj.weight``,
``key_rel_h``, ``key_rel_w``), same attributes (``dk, dv, nh, relative``, ``conv`` may be ``None``).

Differences, all opt-in:
  * ``precision='fp32'|'bf16'`` keyword (default ``'fp32'`` or ``$AACONV_PRECISION``) selects the arithmetic.
  * the reference stores ``self.weights = softmax(logits)`` on EVERY call (attn_aug_conv.py:87, 1.3 GB at
    Transition1/B=16).  Here the map is only materialised when ``module.store_weights = True`` or
    ``forward(x, return_attn=True)``; the visualise path (chexpert.py:365,383-387) sets the flag.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Cfg:
    """Static geometry of one module instance + per-call dims builder."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, pad, dil, dk, dv, nh, relative, precision):
        self.args = (in_channels, out_channels, kernel_size, stride, pad, dil, dk, dv, nh, relative)
        self.precision = precision

    def dims(self, x):
        cin, cout, ks, st, pad, dil, dk, dv, nh, rel = self.args
        B, C, Hin, Win = x.shape
        if C != cin:
            raise RuntimeError(f'AAConv2d: expected {cin} input channels, got {C}')
        H, W = (Hin - 1) // st + 1, (Win - 1) // st + 1
        return _lib.Dims(B, cin, Hin, Win, cout, H, W, ks, st, pad, dil, dk, dv, nh, int(bool(rel)))


_IO_DTYPES = {torch.float32: _lib.FP32, torch.bfloat16: _lib.BF16}


class AAConvFunction(torch.autograd.Function):
    """Fused AAConv2d forward/backward (replaces autograd over attn_aug_conv.py:65-97).

    ``out_total``: if set, y is written as the first Cout channels of a freshly allocated (B, out_total, H, W) feature buffer
    (the next dense block's) and the returned tensor is that slice -- the block adopts the buffer
    instead of copying.  ``fused_in``: x is the input of the Transition's InstanceNorm2d + ReLU (attn_aug_conv.py:438-439),
    which the kernels then apply themselves (row f1); ``fused_in`` is the eps."""

    @staticmethod
    def forward(ctx, x, conv_w, qkv_w, out_w, key_rel_h, key_rel_w, cfg, want_weights, out_total=None, fused_in=None):
        if not x.is_cuda:
            raise RuntimeError('chexpert_b200.AAConv2d runs on CUDA (sm_100a) only; there is no CPU fallback')
        lib = _lib.load()
        prec = _lib.PRECISIONS[cfg.precision]
        in_dtype = x.dtype
        # boundary element type: bf16 activations (autocast) go through as they are in bf16 mode; anything else as fp32
        io_dtype = torch.bfloat16 if (x.dtype == torch.bfloat16 and prec == _lib.BF16) else torch.float32
        xf = x.detach().to(io_dtype).contiguous()
        params = [None if p is None else p.detach().float().contiguous()
                  for p in (conv_w, qkv_w, out_w, key_rel_h, key_rel_w)]
        d = cfg.dims(xf)
        with torch.cuda.device(xf.device):
            _lib.check(lib.aaconv_validate(ctypes.byref(d), prec), 'aaconv_validate')
            if d.relative and (tuple(key_rel_h.shape) != (d.dk // d.nh, 2 * d.H - 1)
                               or tuple(key_rel_w.shape) != (d.dk // d.nh, 2 * d.W - 1)):
                raise RuntimeError(f'AAConv2d: feature map {d.H}x{d.W} does not match input_dims of the relative tables')
            ctot = d.Cout if out_total is None else int(out_total)
            if ctot < d.Cout:
                raise RuntimeError(f'AAConv2d: out_total {ctot} < out_channels {d.Cout}')
            ybuf = torch.empty(d.B, ctot, d.H, d.W, device=xf.device, dtype=io_dtype)
            io = _lib.Io(_IO_DTYPES[io_dtype], _IO_DTYPES[io_dtype], ctot * d.H * d.W, int(fused_in is not None),
                         float(fused_in or 0.0))
            saved = torch.empty(lib.aaconv_saved_bytes_io(ctypes.byref(d), prec, ctypes.byref(io)), device=xf.device, dtype=torch.uint8)
            scratch = torch.empty(lib.aaconv_scratch_bytes_io(ctypes.byref(d), prec, ctypes.byref(io)), device=xf.device,
                                  dtype=torch.uint8)
            weights = (torch.empty(d.B, d.nh, d.H * d.W, d.H * d.W, device=xf.device, dtype=torch.float32)
                       if want_weights else None)
            pp = _lib.Params(*[_ptr(p) for p in params])
            _lib.check(lib.aaconv_forward_io(ctypes.byref(d), prec, ctypes.byref(io), _ptr(xf), ctypes.byref(pp), _ptr(ybuf),
                                             _ptr(weights), _ptr(saved), _ptr(scratch), _stream()), 'aaconv_forward')
        ctx.cfg, ctx.d, ctx.prec, ctx.in_dtype, ctx.io = cfg, d, prec, in_dtype, io
        ctx.save_for_backward(xf, saved, *[p for p in params if p is not None])
        ctx.present = [p is not None for p in params]
        y = ybuf if ctot == d.Cout else ybuf[:, :d.Cout]
        if y.dtype != in_dtype:
            y, ybuf = y.to(in_dtype), None                       # the buffer is only usable in the activations' own type
        if out_total is None:
            ybuf = None
        nd = [t for t in (weights, ybuf) if t is not None]
        if nd:
            ctx.mark_non_differentiable(*nd)
        return y, weights, ybuf

    @staticmethod
    def backward(ctx, dy, _dweights, _dbuf):
        lib = _lib.load()
        tensors = list(ctx.saved_tensors)
        xf, saved = tensors[0], tensors[1]
        it = iter(tensors[2:])
        params = [next(it) if present else None for present in ctx.present]
        d, prec, io = ctx.d, ctx.prec, ctx.io
        need = ctx.needs_input_grad
        dyf = dy.detach().float().contiguous()
        with torch.cuda.device(xf.device):
            dx = torch.empty_like(xf) if need[0] else None
            grads = [torch.empty_like(p) if (p is not None and need[i + 1]) else None for i, p in enumerate(params)]
            scratch = torch.empty(lib.aaconv_scratch_bytes_io(ctypes.byref(d), prec, ctypes.byref(io)), device=xf.device,
                                  dtype=torch.uint8)
            pp = _lib.Params(*[_ptr(p) for p in params])
            gg = _lib.ParamGrads(*[_ptr(g) for g in grads])
            _lib.check(lib.aaconv_backward_io(ctypes.byref(d), prec, ctypes.byref(io), _ptr(xf), ctypes.byref(pp), _ptr(dyf),
                                              _ptr(saved), _ptr(scratch), _ptr(dx), ctypes.byref(gg), _stream()), 'aaconv_backward')
        if dx is not None:
            dx = dx.to(ctx.in_dtype)
        return (dx, *grads, None, None, None, None)


class AAConv2d(nn.Module):
    """Attention-augmented convolution; constructor signature of models/attn_aug_conv.py:20."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, dk, dv, nh, relative, input_dims, **kwargs):
        super().__init__()
        self.dk, self.dv, self.nh, self.relative = dk, dv, nh, relative
        assert dk % nh == 0, 'nh must divide dk'
        assert dv % nh == 0, 'nh must divide dv'
        precision = kwargs.pop('precision', None) or os.environ.get('AACONV_PRECISION', 'fp32')
        if precision not in _lib.PRECISIONS:
            raise ValueError(f'precision must be one of {list(_lib.PRECISIONS)}')
        padding = kwargs.pop('padding', None)
        if not padding:
            padding = kernel_size // 2
        dilation = kwargs.pop('dilation', 1)
        if kwargs.pop('groups', 1) != 1:
            raise NotImplementedError('AAConv2d (B200): grouped conv branch is outside the supported hot path')
        if kwargs:
            raise TypeError(f'unsupported conv kwargs: {sorted(kwargs)}')
        # parameter holders with the reference's state_dict names; their own forward() is never called
        self.conv = (nn.Conv2d(in_channels, out_channels - dv, kernel_size, stride, padding, dilation=dilation, bias=False)
                     if out_channels > dv else None)
        self.in_proj_qkv = nn.Conv2d(in_channels, 2 * dk + dv, kernel_size=1, stride=stride, bias=False)
        self.out_proj = nn.Conv2d(dv, dv, kernel_size=1, bias=False)
        if relative:
            H, W = input_dims
            self.key_rel_h = nn.Parameter(dk ** -0.5 + torch.randn(dk // nh, 2 * H - 1))
            self.key_rel_w = nn.Parameter(dk ** -0.5 + torch.randn(dk // nh, 2 * W - 1))
        self.precision = precision
        self.store_weights = False
        self.weights = None
        self._geom = (in_channels, out_channels, kernel_size, stride, padding, dilation)

    def _cfg(self):
        cin, cout, ks, st, pad, dil = self._geom
        return _Cfg(cin, cout, ks, st, pad, dil, self.dk, self.dv, self.nh, self.relative, self.precision)

    def forward(self, x, return_attn=False, out_total=None, fused_in=None):
        """``forward(x)`` is the reference's (attn_aug_conv.py:65).  Keyword extras, all opt-in: ``return_attn`` also returns the
        softmax map; ``out_total`` writes y into the first channels of a (B, out_total, H, W) feature buffer and returns that
        slice; ``fused_in=eps`` applies the Transition's InstanceNorm2d + ReLU to x inside the kernels."""
        want = bool(return_attn or self.store_weights)
        y, w, buf = AAConvFunction.apply(
            x, self.conv.weight if self.conv is not None else None, self.in_proj_qkv.weight, self.out_proj.weight,
            self.key_rel_h if self.relative else None, self.key_rel_w if self.relative else None, self._cfg(), want,
            out_total, fused_in)
        if buf is not None:
            y.feature_buffer = buf       # (B, out_total, H, W); y == buf[:, :Cout].  BufferedDenseBlock adopts it.
        if self.store_weights:
            self.weights = w
        return (y, w) if return_attn else y

    def extra_repr(self):
        return 'dk={dk}, dv={dv}, nh={nh}, relative={relative}'.format(**self.__dict__)
