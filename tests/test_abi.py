"""The C-ABI shared library loads on a machine without a GPU and exports every symbol that
include/slvgpu.h declares; without a device the entry points fail loudly (no CPU path)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'slvgpu.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(slvgpu_[a-z_0-9]+)\s*\(', text)))


def test_header_symbols_exported():
    from slv_b200 import _lib
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(_lib.lib, n), 'symbol %s declared in include/slvgpu.h is not exported' % n
    assert b'sm_100a' in _lib.lib.slvgpu_version()


def test_struct_layout_matches_header():
    from slv_b200 import _lib
    assert ctypes.sizeof(_lib.PlanDesc) == 4 * 4 + 3 * 8 + 2 * 8 + 2 * 8 + 2 * 8 + 2 * 8 + 2 * 4 + 8 + 2 * 4
    assert _lib.NOUT == 20 and _lib.NMETA == 32 and _lib.NSOL == 16


def test_no_gpu_means_error_not_fallback(frags):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from slv_b200 import _lib
    from slv_b200.solefp import EFP
    efp = EFP(elect=True, freq=True, mode=8)
    efp.set(frags('nma').get_pos(), [0], [12], [frags('nma')], [None], [None])
    with pytest.raises(_lib.SlvGpuError):
        efp.eval(0)
    with pytest.raises(_lib.SlvGpuError):
        frags('water').copy().sup(frags('water').get_pos())


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'slv_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f
