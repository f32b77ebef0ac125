"""Host-side wiring of the attention-augmented DenseNet121 around the B200 AAConv2d.

The wiring lives here: which dk/dv/input_dims each Transition gets, the
InstanceNorm -> ReLU -> AAConv2d(3x3, stride 2) sequence, and the weight initialisation, so that a checkpoint written
by the reference (``features.transition{1,2,3}.conv.{conv,in_proj_qkv,out_proj}.weight``, ``...key_rel_h/w``;
chexpert.py:90-123,504-518) loads strictly into this model and vice versa.  The dense blocks are torchvision's
(``_DenseBlock``), exactly as in the reference (models/attn_aug_conv.py:13,479); they are outside the hot path.

    reference                                   here
    models/attn_aug_conv.py:411-446 _Transition -> transition()
    models/attn_aug_conv.py:448-517 DenseNet    -> DenseNet
    chexpert.py:474-476 'aadensenet121'         -> aadensenet121()
"""
from collections import OrderedDict

import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.densenet import _DenseBlock, _DenseLayer

from .aaconv import AAConv2d
from .fused_bn import bn_relu, tap_bn_relu, tap_bn_relu_cl, bn_relu_cl, cl_ok, slice_layout, _is_cl


def transition_attn_dims(num_output_features, attn_params):
    """dk, dv, (H, W) of the AAConv2d in one Transition (models/attn_aug_conv.py:416-427)."""
    nh = attn_params['nh']
    dk = max(20 * nh, int((attn_params['k'] * num_output_features // nh) * nh))
    dv = int((attn_params['v'] * num_output_features // nh) * nh)
    dims = attn_params['input_dims'][0] // 2, attn_params['input_dims'][1] // 2   # attention runs on the strided map
    return dk, dv, dims


class AATransition(nn.Sequential):
    """InstanceNorm2d -> ReLU -> AAConv2d(3x3, stride 2) with the reference's child names ``norm`` / ``relu`` / ``conv``
    (models/attn_aug_conv.py:436-440), so state_dicts are interchangeable.

    ``fused_prologue=True``: on CUDA the InstanceNorm statistics, the normalisation and the ReLU
    run inside the AAConv2d kernels (one statistics pass + the operand pack forward; one fused adjoint pass backward) instead
    of as two separate torch modules; ``norm`` and ``relu`` stay as (parameter-free) children and define eps.
    ``forward(x, out_total=C)`` makes the AAConv2d epilogues write into the first channels of the next dense block's
    (B, C, H, W) feature buffer (row f3)."""

    def __init__(self, num_input_features, num_output_features, attn_params, precision=None, fused_prologue=False):
        super().__init__()
        dk, dv, dims = transition_attn_dims(num_output_features, attn_params)
        self.add_module('norm', nn.InstanceNorm2d(num_input_features))          # affine=False, no running stats (:438)
        self.add_module('relu', nn.ReLU(inplace=True))
        self.add_module('conv', AAConv2d(num_input_features, num_output_features, 3, 2, dk, dv, attn_params['nh'],
                                         attn_params['relative'], dims, precision=precision))
        self.fused_prologue = bool(fused_prologue)

    def forward(self, x, out_total=None):
        if self.fused_prologue and x.is_cuda and not self.norm.affine and not self.norm.track_running_stats:
            return self.conv(x, out_total=out_total, fused_in=self.norm.eps)
        return self.conv(self.relu(self.norm(x)), out_total=out_total)


def transition(num_input_features, num_output_features, attn_params=None, precision=None, fused_prologue=False):
    """Sequential with the reference's child names: norm, relu, conv (+ pool for the plain variant)."""
    if attn_params is None:   # stock DenseNet transition (models/attn_aug_conv.py:429-434)
        return nn.Sequential(OrderedDict([
            ('norm', nn.BatchNorm2d(num_input_features)),
            ('relu', nn.ReLU(inplace=True)),
            ('conv', nn.Conv2d(num_input_features, num_output_features, kernel_size=1, stride=1, bias=False)),
            ('pool', nn.AvgPool2d(kernel_size=2, stride=2)),
        ]))
    return AATransition(num_input_features, num_output_features, attn_params, precision, fused_prologue)


# ------------------------------------------------------------------------------------------------
# pre-allocated DenseBlock feature buffer
# ------------------------------------------------------------------------------------------------
def _alias(buf, c0, c1):
    """Tensor over channels [c0, c1) of the (B, C, H, W) buffer that shares its storage but NOT its autograd identity / version
    counter: later layers write channels >= c1 of the same storage, which must not invalidate what earlier layers saved."""
    B, C, H, W = buf.shape
    return torch.empty(0, dtype=buf.dtype, device=buf.device).set_(
        buf.untyped_storage(), buf.storage_offset() + c0 * H * W, (B, c1 - c0, H, W), (C * H * W, H * W, W, 1))


class _Adopt(torch.autograd.Function):
    """init_features -> the first channels of the block's buffer (copied unless they were produced there)."""

    @staticmethod
    def forward(ctx, init, buf):
        c0 = init.shape[1]
        view = _alias(buf, 0, c0)
        if init.data_ptr() != view.data_ptr() or init.stride() != view.stride():
            view.copy_(init)
        return view

    @staticmethod
    def backward(ctx, g):
        return g, None


class _Append(torch.autograd.Function):
    """[features so far | new_features]: the concatenation of torchvision's block (densenet.py:48,120-124) without the copy of
    everything that came before -- only the layer's own `growth_rate` channels are written, next to the rest."""

    @staticmethod
    def forward(ctx, prev, new, buf):
        c = prev.shape[1]
        ctx.c = c
        _alias(buf, c, c + new.shape[1]).copy_(new)
        return _alias(buf, 0, c + new.shape[1])

    @staticmethod
    def backward(ctx, g):
        return g[:, :ctx.c], g[:, ctx.c:], None


def _nchw_slice_ok(t):
    B, C, H, W = t.shape
    return (t.stride(3) == 1 and t.stride(2) == W and t.stride(1) == H * W and t.stride(0) % 4 == 0
            and t.data_ptr() % (4 * t.element_size()) == 0)


class _AppendCL(torch.autograd.Function):
    """_Append for channels-last new features: they are transposed into their channel slice of the NCHW buffer (one tiled kernel),
    and the slice of the buffer gradient is transposed back to NHWC for conv2's backward."""

    @staticmethod
    def forward(ctx, prev, new, buf):
        c, k = prev.shape[1], new.shape[1]
        ctx.c = c
        slice_layout(new.detach(), _alias(buf, c, c + k), to_nchw=True)
        return _alias(buf, 0, c + k)

    @staticmethod
    def backward(ctx, g):
        c = ctx.c
        gs = g[:, c:]
        B, k, H, W = gs.shape
        if _nchw_slice_ok(gs) and gs.is_cuda:
            gn = torch.empty((B, k, H, W), device=g.device, dtype=g.dtype, memory_format=torch.channels_last)
            slice_layout(gs, gn, to_nchw=False)
        else:
            gn = gs.contiguous(memory_format=torch.channels_last)
        return g[:, :c], gn, None


class BufferedDenseBlock(nn.ModuleDict):
    """torchvision's ``_DenseBlock`` (same ``denselayer%d`` children, parameters and arithmetic; the reference uses it at
    models/attn_aug_conv.py:479-482) over ONE pre-allocated (B, C_in + n * growth, H, W) feature buffer: layer i reads channels
    [0, c_i) of the buffer and writes its output behind them, so the per-layer ``torch.cat`` (O(layers^2) copy traffic and
    saved activations) disappears.  If ``init_features`` already lives in such a buffer (an AAConv2d called with ``out_total``),
    it is adopted without a copy.  The gradient of the concatenated features is accumulated along the chain
    (one add per layer over contiguous tensors) instead of per (producer, consumer) pair."""

    def __init__(self, num_layers, num_input_features, bn_size, growth_rate, drop_rate):
        super().__init__()
        for i in range(num_layers):
            self.add_module('denselayer%d' % (i + 1),
                            _DenseLayer(num_input_features + i * growth_rate, growth_rate=growth_rate, bn_size=bn_size,
                                        drop_rate=drop_rate))
        self.num_input_features, self.growth_rate, self.num_layers = num_input_features, growth_rate, num_layers
        self.out_channels = num_input_features + num_layers * growth_rate
        self.inner_channels_last = os.environ.get('AACONV_INNER_CL', '1') != '0'   # AACONV_INNER_CL=0: NCHW inside the layers (A/B, profiles/r02_b_scaling.md)

    def forward(self, init_features):
        B, C0, H, W = init_features.shape
        buf = getattr(init_features, 'feature_buffer', None)
        if (buf is None or tuple(buf.shape) != (B, self.out_channels, H, W) or buf.dtype != init_features.dtype
                or not buf.is_contiguous() or buf.data_ptr() != init_features.data_ptr()):
            buf = torch.empty(B, self.out_channels, H, W, dtype=init_features.dtype, device=init_features.device)
        feats = _Adopt.apply(init_features, buf)
        # per-(channel, sample) plane statistics of the buffer, shared by the norm1 of every layer: layer i only reduces the channels
        # layer i-1 has not seen (its own 32 new ones); the first layer reduces the block's input channels
        stats = torch.empty(2 * B * self.out_channels, device=buf.device, dtype=torch.float32) if buf.is_cuda else None
        valid = 0
        growth_ok = self.growth_rate % 4 == 0 and (H * W) % 4 == 0 and stats is not None
        for layer in self.values():
            # norm -> relu pairs: fused strided kernels in training on CUDA (chexpert_b200.fused_bn), the torch modules otherwise
            c = feats.shape[1]
            if (self.inner_channels_last and layer.drop_rate == 0 and torch.is_grad_enabled() and feats.requires_grad
                    and cl_ok(layer.norm1, feats) and _nchw_slice_ok(feats) and growth_ok):
                # The two convolutions of the layer run channels-last (cuDNN's tensor-core kernels are NHWC; with NCHW tensors it
                # transposes every operand of every fprop / dgrad / wgrad itself: 21 % of the step in the round-2 profile).  The
                # layout change rides on the BatchNorm + ReLU passes (csrc/bn_cl.cu); the buffer itself stays NCHW.
                feats, y1 = tap_bn_relu_cl(layer.norm1, feats, (stats, valid))
                z = layer.conv1(y1)
                if _is_cl(z) and cl_ok(layer.norm2, z):
                    new = layer.conv2(bn_relu_cl(layer.norm2, z))
                else:
                    new = layer.conv2(bn_relu(layer.norm2, z.contiguous()))
                valid = c
                new = new.to(buf.dtype)
                feats = (_AppendCL.apply(feats, new, buf) if _is_cl(new) else _Append.apply(feats, new.contiguous(), buf))
                continue
            # norm1 taps the concatenated features: backward adds its dx into the gradient of the concatenation in place
            feats, y1 = tap_bn_relu(layer.norm1, feats, None if stats is None else (stats, valid))
            new = layer.conv2(bn_relu(layer.norm2, layer.conv1(y1)))
            valid = c
            if layer.drop_rate > 0:
                new = F.dropout(new, p=layer.drop_rate, training=self.training)
            feats = _Append.apply(feats, new.to(buf.dtype), buf)
        return feats


class DenseNet(nn.Module):
    """DenseNet with attention-augmented transitions; constructor of models/attn_aug_conv.py:452-453 plus
    ``precision`` ('fp32' | 'bf16') for the AAConv2d kernels.  Unlike the reference, ``attn_params`` is not mutated."""

    def __init__(self, growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4, drop_rate=0,
                 num_classes=1000, attn_params=None, precision=None, feature_buffer=False, fused_prologue=False):
        super().__init__()
        self.feature_buffer = bool(feature_buffer)
        attn = dict(attn_params) if attn_params is not None else None
        if len(block_config) == 4:    # ImageNet stem: /4 before the first block (:460-468)
            stem = [('conv0', nn.Conv2d(3, num_init_features, kernel_size=7, stride=2, padding=3, bias=False)),
                    ('norm0', nn.BatchNorm2d(num_init_features)),
                    ('relu0', nn.ReLU(inplace=True)),
                    ('pool0', nn.MaxPool2d(kernel_size=3, stride=2, padding=1))]
            if attn is not None:
                attn['input_dims'] = attn['input_dims'][0] // 4, attn['input_dims'][1] // 4
        else:                         # CIFAR stem (:470-474)
            stem = [('conv0', nn.Conv2d(3, num_init_features, kernel_size=5, stride=1, padding=2, bias=False)),
                    ('norm0', nn.BatchNorm2d(num_init_features)),
                    ('relu0', nn.ReLU(inplace=True))]
        self.features = nn.Sequential(OrderedDict(stem))
        width = num_init_features
        for i, num_layers in enumerate(block_config):
            block = (BufferedDenseBlock(num_layers, width, bn_size, growth_rate, drop_rate) if feature_buffer else
                     _DenseBlock(num_layers=num_layers, num_input_features=width, bn_size=bn_size, growth_rate=growth_rate,
                                 drop_rate=drop_rate))
            self.features.add_module(f'denseblock{i + 1}', block)
            width += num_layers * growth_rate
            if i != len(block_config) - 1:
                self.features.add_module(f'transition{i + 1}', transition(width, width // 2, attn, precision, fused_prologue))
                width //= 2
            if attn is not None:      # every stage halves the map the next transition's attention sees (:491-493)
                attn['input_dims'] = attn['input_dims'][0] // 2, attn['input_dims'][1] // 2
        self.features.add_module('norm5', nn.BatchNorm2d(width))
        self.classifier = nn.Linear(width, num_classes)
        for m in self.modules():      # :503-510 -- includes the three Conv2d parameter holders of every AAConv2d
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.constant_(m.bias, 0)

    def _features(self, x):
        """-> relu(features(x)) (models/attn_aug_conv.py:513-514)."""
        if not self.feature_buffer:
            return F.relu(self.features(x), inplace=True)
        mods = list(self.features.named_children())
        skip = False
        for i, (name, m) in enumerate(mods):
            if skip:
                skip = False
                continue
            nxt = mods[i + 1][1] if i + 1 < len(mods) else None
            if isinstance(m, AATransition) and isinstance(nxt, BufferedDenseBlock):
                x = m(x, out_total=nxt.out_channels)      # the AAConv2d epilogues write into the next block's buffer
            elif isinstance(m, nn.BatchNorm2d) and isinstance(nxt, nn.ReLU):
                x = bn_relu(m, x)                         # stem: norm0 -> relu0 as one fused op (grid (C, B) instead of one CTA per channel)
                skip = True
            elif isinstance(m, nn.BatchNorm2d) and nxt is None:
                return bn_relu(m, x)                      # norm5 followed by forward()'s F.relu
            else:
                x = m(x)
        return F.relu(x, inplace=True)

    def forward(self, x):
        f = self._features(x)
        return self.classifier(F.adaptive_avg_pool2d(f, (1, 1)).flatten(1))

    def attn_layers(self):
        """The AAConv2d modules the visualise path hooks (chexpert.py:478)."""
        return [m for m in self.modules() if isinstance(m, AAConv2d)]


def aadensenet121(num_classes=5, input_dims=(320, 320), precision=None, feature_buffer=False, fused_prologue=False):
    """The model behind ``--model aadensenet121`` (chexpert.py:474-476); ``input_dims`` is the image size.
    ``feature_buffer`` / ``fused_prologue``: the B200 execution options of rows f3 / f1 (same parameters, same state_dict)."""
    return DenseNet(32, (6, 12, 24, 16), 64, num_classes=num_classes,
                    attn_params={'k': 0.2, 'v': 0.1, 'nh': 8, 'relative': True, 'input_dims': tuple(input_dims)},
                    precision=precision, feature_buffer=feature_buffer, fused_prologue=fused_prologue)
