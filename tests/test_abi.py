"""CPU-side checks of the boundary: the shared object loads without a GPU and exports every symbol the
header declares; argument validation mirrors the reference's construct-time asserts."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from chexpert_b200 import _lib
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    from chexpert_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'aaconv_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(aaconv_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/aaconv_b200.h but not exported'
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))


def test_struct_layout_matches_header():
    from chexpert_b200 import _lib
    assert ctypes.sizeof(_lib.Dims) == 15 * 4
    assert ctypes.sizeof(_lib.Params) == 5 * ctypes.sizeof(ctypes.c_void_p)


def test_validate_rejects_bad_geometry(lib):
    from chexpert_b200 import _lib
    ok = _lib.Dims(2, 12, 10, 14, 24, 5, 7, 3, 2, 1, 1, 16, 8, 4, 1)
    assert lib.aaconv_validate(ctypes.byref(ok), 0) == 0
    assert lib.aaconv_saved_bytes(ctypes.byref(ok), 0) > 0
    assert lib.aaconv_saved_offset(ctypes.byref(ok), 0, b'lse') > 0
    assert lib.aaconv_saved_offset(ctypes.byref(ok), 0, b'nope') == -1
    bad = _lib.Dims(2, 12, 10, 14, 24, 5, 7, 3, 2, 1, 1, 18, 8, 4, 1)      # nh does not divide dk
    assert lib.aaconv_validate(ctypes.byref(bad), 0) != 0
    assert b'nh must divide dk' in lib.aaconv_last_error()
    bad = _lib.Dims(2, 12, 10, 14, 24, 4, 7, 3, 2, 1, 1, 16, 8, 4, 1)      # wrong output map
    assert lib.aaconv_validate(ctypes.byref(bad), 0) != 0


def test_module_contract_on_cpu():
    import chexpert_b200 as cb
    with pytest.raises(AssertionError):
        cb.AAConv2d(8, 16, 3, 2, 10, 4, 4, True, (4, 4))                    # attn_aug_conv.py:27
    m = cb.AAConv2d(12, 24, 3, 2, 16, 8, 4, True, (5, 7))
    assert set(m.state_dict()) == {'key_rel_h', 'key_rel_w', 'conv.weight', 'in_proj_qkv.weight', 'out_proj.weight'}
    assert m.state_dict()['key_rel_h'].shape == (4, 9) and m.state_dict()['key_rel_w'].shape == (4, 13)
    assert m.extra_repr() == 'dk=16, dv=8, nh=4, relative=True'
    assert cb.AAConv2d(6, 8, 3, 2, 8, 8, 2, True, (4, 5)).conv is None       # attn_aug_conv.py:34
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.randn(1, 12, 10, 14))


def test_missing_extension_fails_loudly(monkeypatch):
    from chexpert_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libaaconv_b200.so')
    with pytest.raises(RuntimeError, match='no CPU / eager fallback'):
        _lib.load()
