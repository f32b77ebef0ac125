"""Host-side wiring of the attention-augmented DenseNet121 around the B200 AAConv2d.

Only the wiring lives here: which dk/dv/input_dims each Transition gets, the
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

import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.densenet import _DenseBlock

from .aaconv import AAConv2d


def transition_attn_dims(num_output_features, attn_params):
    """dk, dv, (H, W) of the AAConv2d in one Transition (models/attn_aug_conv.py:416-427)."""
    nh = attn_params['nh']
    dk = max(20 * nh, int((attn_params['k'] * num_output_features // nh) * nh))
    dv = int((attn_params['v'] * num_output_features // nh) * nh)
    dims = attn_params['input_dims'][0] // 2, attn_params['input_dims'][1] // 2   # attention runs on the strided map
    return dk, dv, dims


def transition(num_input_features, num_output_features, attn_params=None, precision=None):
    """Sequential with the reference's child names: norm, relu, conv (+ pool for the plain variant)."""
    if attn_params is None:   # stock DenseNet transition (models/attn_aug_conv.py:429-434)
        return nn.Sequential(OrderedDict([
            ('norm', nn.BatchNorm2d(num_input_features)),
            ('relu', nn.ReLU(inplace=True)),
            ('conv', nn.Conv2d(num_input_features, num_output_features, kernel_size=1, stride=1, bias=False)),
            ('pool', nn.AvgPool2d(kernel_size=2, stride=2)),
        ]))
    dk, dv, dims = transition_attn_dims(num_output_features, attn_params)
    return nn.Sequential(OrderedDict([
        ('norm', nn.InstanceNorm2d(num_input_features)),          # affine=False, no running stats (:438)
        ('relu', nn.ReLU(inplace=True)),
        ('conv', AAConv2d(num_input_features, num_output_features, 3, 2, dk, dv, attn_params['nh'],
                          attn_params['relative'], dims, precision=precision)),
    ]))


class DenseNet(nn.Module):
    """DenseNet with attention-augmented transitions; constructor of models/attn_aug_conv.py:452-453 plus
    ``precision`` ('fp32' | 'bf16') for the AAConv2d kernels.  Unlike the reference, ``attn_params`` is not mutated."""

    def __init__(self, growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4, drop_rate=0,
                 num_classes=1000, attn_params=None, precision=None):
        super().__init__()
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
            self.features.add_module(f'denseblock{i + 1}', _DenseBlock(num_layers=num_layers, num_input_features=width,
                                                                        bn_size=bn_size, growth_rate=growth_rate,
                                                                        drop_rate=drop_rate))
            width += num_layers * growth_rate
            if i != len(block_config) - 1:
                self.features.add_module(f'transition{i + 1}', transition(width, width // 2, attn, precision))
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

    def forward(self, x):
        f = F.relu(self.features(x), inplace=True)
        return self.classifier(F.adaptive_avg_pool2d(f, (1, 1)).flatten(1))

    def attn_layers(self):
        """The AAConv2d modules the visualise path hooks (chexpert.py:478)."""
        return [m for m in self.modules() if isinstance(m, AAConv2d)]


def aadensenet121(num_classes=5, input_dims=(320, 320), precision=None):
    """The model behind ``--model aadensenet121`` (chexpert.py:474-476); ``input_dims`` is the image size."""
    return DenseNet(32, (6, 12, 24, 16), 64, num_classes=num_classes,
                    attn_params={'k': 0.2, 'v': 0.1, 'nh': 8, 'relative': True, 'input_dims': tuple(input_dims)},
                    precision=precision)
