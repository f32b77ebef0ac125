"""Host wiring of aadensenet121: strict state_dict compatibility with the reference."""
import json
import os

import torch

from tests.helpers import GOLDEN


def test_state_dict_keys_and_shapes_match_reference():
    from chexpert_b200.densenet import aadensenet121
    meta = json.load(open(os.path.join(GOLDEN, 'aadensenet121_meta.json')))
    m = aadensenet121(5, (320, 320))
    sd = m.state_dict()
    assert set(sd) == set(meta['state_dict'])
    for k, shape in meta['state_dict'].items():
        assert list(sd[k].shape) == shape, k
    assert sum(p.numel() for p in m.parameters()) == meta['n_params'] == 12534381
    for i, layer in enumerate(m.attn_layers(), 1):
        t = meta['transitions'][str(i)]
        assert (layer.dk, layer.dv, layer.nh, layer.relative) == (t['dk'], t['dv'], t['nh'], t['relative'])
        assert layer.extra_repr() == t['repr']
        assert list(layer.key_rel_h.shape) == t['key_rel_h']


def test_attn_params_not_mutated_and_512_reachable():
    from chexpert_b200.densenet import DenseNet
    ap = {'k': 0.2, 'v': 0.1, 'nh': 8, 'relative': True, 'input_dims': (512, 512)}
    m = DenseNet(32, (6, 12, 24, 16), 64, num_classes=5, attn_params=ap)
    assert ap['input_dims'] == (512, 512)          # the reference mutates its argument (attn_aug_conv.py:468,493)
    assert list(m.features.transition1.conv.key_rel_w.shape) == [20, 127]   # L = 64x64 at Transition1


def test_synthetic_batch_is_radiograph_shaped():
    from chexpert_b200.train import synthetic_batch
    x, t = synthetic_batch(4, 320, seed=1)
    assert x.shape == (4, 3, 320, 320) and t.shape == (4, 14)
    assert torch.equal(x[:, 0], x[:, 1]) and -15.4 < float(x.min()) and float(x.max()) < 13.5
    vals = set(torch.nan_to_num(t, nan=7.0).unique().tolist())
    assert vals <= {7.0, -1.0, 0.0, 1.0}
