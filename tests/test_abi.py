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
    assert _lib.NOUT == 24 and _lib.NMETA == 32 and _lib.NSOL == 16


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


def test_missing_parameter_tables_are_rejected_loudly(frags):
    """A fragment without the section a requested term reads (mecn / li+ / water2 carry no 'dpoli') must not reach the
    kernels with a -1 table offset: the reference raises KeyError there (solefp.py:1429); so does the plan, and the
    C ABI itself answers EINVAL before it even looks for a device."""
    from slv_b200 import _lib
    from slv_b200.plan import Plan
    nma = frags('nma')
    for bad in ('mecn', 'li+', 'water2'):
        b = frags(bad)
        with pytest.raises(KeyError):
            Plan([0, 1], [12, b.get_natoms()], [nma, b], [None, None], [None, None], 8,
                 dict(elect=True, pol=True, disp=True), 40.0, 17.0, 12.0)
    # the solute itself: water has no mode derivatives
    with pytest.raises((KeyError, AssertionError)):
        Plan([0, 1], [3, 12], [frags('water'), nma], [None, None], [None, None], 1, dict(elect=True, pol=True), 40.0, 17.0, 12.0)
    # straight through the C ABI with a table offset knocked out
    import ctypes as C
    meta = np.full((2, _lib.NMETA), -1, np.int32); sol = np.zeros(_lib.NSOL, np.int32)
    for t in (0, 1):
        meta[t][:6] = [3, 3, 3, 0, 0, 3]
        for k in ('IO_SUPL', 'IO_REORD', 'O_POS', 'O_RDMA', 'O_MOM', 'O_RPOL', 'O_POL', 'O_POLINV'):
            meta[t][_lib.M[k]] = 0
    meta[0][_lib.M['O_DPOLI']] = 0                      # the solvent type (1) has no dpoli
    itab = np.zeros(16, np.int32); dtab = np.zeros(512)
    d = _lib.PlanDesc()
    d.ntypes = 2; d.ninst = 2; d.natoms_total = 6; d.flags = _lib.FLAG['elect'] | _lib.FLAG['disp']
    d.ccut = d.pcut = d.ecut = 10.0
    d.type_meta = meta.ctypes.data; d.sol_meta = sol.ctypes.data; d.itab = itab.ctypes.data; d.n_itab = itab.size
    d.dtab = dtab.ctypes.data; d.n_dtab = dtab.size
    it = np.array([0, 1], np.int32); a0 = np.array([0, 3, 6], np.int32)
    d.inst_type = it.ctypes.data; d.inst_atom0 = a0.ctypes.data
    h = C.c_void_p()
    rc = _lib.lib.slvgpu_plan_create(C.byref(d), C.byref(h))
    assert rc == -1 and b'dpoli' in _lib.lib.slvgpu_last_error()
