"""Host logic: the .frg reader (scalings, triangular unpacking, 0-based mode list), the $Frag input
parser, SolCAMM contraction and the shift-dictionary assembly."""
import gzip
import os

import numpy as np
import pytest


def test_sizes_and_scalings(frags):
    p = frags('nma').get()
    assert (p['natoms'], p['ndma'], p['npol'], p['nmos'], p['nbasis'], p['nmodes']) == (12, 12, 20, 20, 164, 30)
    assert list(p['mode']) == [0, 7, 8]                       # file says 1,8,9 (1-based)
    assert p['dmao1'].shape == (30, 12, 10) and p['dpoli1'].shape == (30, 20, 12, 3, 3) and p['vecl1'].shape == (30, 20, 164)
    assert np.allclose(p['fock'], p['fock'].T) and np.allclose(p['fock1'], p['fock1'].transpose(0, 2, 1))
    g = p['gijk']
    assert np.allclose(g, g.transpose(1, 0, 2)) and np.allclose(g, g.transpose(0, 2, 1)) and np.allclose(g, g.transpose(2, 1, 0))
    w = frags('water').get()
    assert (w['natoms'], w['ndma'], w['npol'], w['nbasis']) == (3, 5, 5, 37) and 'dmac1' not in w


def test_first_derivatives_carry_sqrt_redmass(frags, tmp_path):
    """Re-read the raw numbers of one section and compare with the scaled array (slvpar.py:1278-1283)."""
    from slv_b200.slvpar import Frag
    f = frags('mescn'); p = f.get()
    opener = gzip.open if f._file.endswith('.gz') else open
    text = opener(f._file, 'rt').read()
    sec = [s for s in text.split('[') if s.startswith(' DMTP charges - first derivatives ]')][0]
    raw = np.array(' '.join(sec.split('\n')[1:]).split(), float)[:p['nmodes'] * p['ndma']].reshape(p['nmodes'], p['ndma'])
    assert np.allclose(p['dmac1'], np.sqrt(p['redmass'])[:, None] * raw, rtol=0, atol=1e-15)
    sec2 = [s for s in text.split('[') if s.startswith(' DMTP charges - second derivatives wrt mode]')][0]
    raw2 = np.array(' '.join(sec2.split('\n')[1:]).split(), float)[:len(p['mode']) * p['ndma']].reshape(len(p['mode']), p['ndma'])
    assert np.allclose(p['dmac2'], p['redmass'][p['mode']][:, None] * raw2, rtol=0, atol=1e-15)


def test_unknown_section_raises(tmp_path):
    from slv_b200.slvpar import Frag
    fn = tmp_path / 'bad.frg'
    fn.write_text(' [ molecule ]\n   name = X\n   natoms = 1\n\n [ Nonsense ]  N= 1\n 1.0\n')
    with pytest.raises(Exception):
        Frag(str(fn))
    with pytest.raises(IOError):
        Frag('water')._locate('no-such-fragment')


def test_text_to_list_and_mdinput():
    from slv_b200.slvpar import _text_to_list
    from slv_b200.md import MDInput
    assert list(_text_to_list('1-5')) == [1, 2, 3, 4, 5] and list(_text_to_list('1 2 9')) == [1, 2, 9]
    assert list(_text_to_list('1,3,5-7', delimiter=',')) == [1, 3, 5, 6, 7]
    inp = '''
$Frag
    NMA        nma
    reorder    1 3 5 2 7 10 11 12 4 6 8 9
    supl       1-5
    atoms      1-12                           1
$endFrag

$Frag
    water      water
    atoms      13-42                        10
    supl       1-3
$endFrag
'''
    args, idx = MDInput(inp).get()
    ind, nmol, bsm, supl, reord = args
    assert ind == [0] + [1] * 10 and [int(n) for n in nmol] == [12] + [3] * 10 and idx == list(range(42))
    assert list(supl[0]) == [0, 1, 2, 3, 4] and list(reord[0])[:4] == [0, 2, 4, 1] and list(reord[1]) == [0, 1, 2]
    assert bsm[0].get_natoms() == 12 and bsm[1].get_natoms() == 3


def test_solcamm_property_matches_plan_tables(frags):
    """The SolCAMM contraction (slvpar.py:1025-1073) equals the plan's s_mea + s_ea tables."""
    from slv_b200.plan import mode_weights, mom20
    f = frags('nma'); p = f.get()
    rdma, pos, c, d, q, o = f.get_property('solcamm', mode=8, ea=True, dma=False)
    g, den = mode_weights(p, 8)
    dma1 = mom20(p['dmac1'], p['dmad1'], p['dmaq1'], p['dmao1'])
    smea = -np.tensordot(g, dma1, (0, 0))
    m2 = list(p['mode']).index(7)
    sea = mom20(p['dmac2'][m2], p['dmad2'][m2], p['dmaq2'][m2], p['dmao2'][m2]) / den
    tot = mom20(c, d, q, o)
    assert np.abs(tot - (smea + sea)).max() < 1e-14


def test_shift_dictionary_assembly():
    from slv_b200.solefp import shifts_from_rows
    from slv_b200.units import HartreePerHbarToCmRec as c
    row = np.zeros((1, 24)); row[0, :8] = [1e-5, 2e-5, -3e-6, 4e-6, 5e-6, 6e-5, 7e-5, 8e-5]
    s = shifts_from_rows(row, True)
    assert abs(s['ele_tot'][0] - (1e-5 + 2e-5 - 3e-6 + 4e-6) * c) < 1e-9
    assert abs(s['total'][0] - (s['ele_tot'][0] + s['pol_mea'][0] + s['rep_mea'][0] + s['dis_mea'][0])) < 1e-9
    assert abs(s['total_ele'][0] - (s['total'][0] - s['rep_mea'][0])) < 1e-9 and s['pol_ea'][0] == 0.0
    assert abs(s['solcamm'][0] - 3e-5 * c) < 1e-9 and abs(s['dis_mea_iso'][0] - 8e-5 * c) < 1e-9


def test_exchange_repulsion_work_list_covers_every_ao_pair_once(frags):
    """slv_b200/plan.py subshell_work_list: the 23 (L_bra, L_ket, bra component) classes tile the AO-pair matrix of
    two fragments exactly once, items are sorted by primitive-pair count inside a class, and the sub-shells rebuild the
    basis (an SP shell is an S and a P sub-shell with the same exponents)."""
    import numpy as np
    from slv_b200.basis import BasisSet
    from slv_b200.plan import subshell_work_list, SUBSHELL_CLASSES
    pa, pb = frags('mescn').get(), frags('dmso').get()
    bA, bB = BasisSet(pa['atno'], pa['pos']), BasisSet(pb['atno'], pb['pos'])
    for b in (bA, bB):
        ncomp = np.array([1, 3, 6])[b.subshells[:, 2]]
        assert ncomp.sum() == len(b) and (b.powers.sum(1)[b.subshells[:, 1]] == b.subshells[:, 2]).all()
        assert (b.nprim[b.subshells[:, 1]] == b.subshells[:, 3]).all()
    offs, items = subshell_work_list(bA.subshells, bB.subshells)
    assert len(offs) == len(SUBSHELL_CLASSES) + 1 and offs[-1] == len(items)
    hits = np.zeros((len(bA), len(bB)), int)
    for cid, (LA, LB, IA) in enumerate(SUBSHELL_CLASSES):
        cost_prev = None
        for w in items[offs[cid]:offs[cid + 1]]:
            a, b = divmod(int(w), len(bB.subshells))
            sa, sb = bA.subshells[a], bB.subshells[b]
            assert sa[2] == LA and sb[2] == LB
            rows = range(sa[1], sa[1] + (1, 3, 6)[LA]) if IA < 0 else [sa[1] + IA]
            for k in rows:
                hits[k, sb[1]:sb[1] + (1, 3, 6)[LB]] += 1
            cost = sa[3] * sb[3]
            assert cost_prev is None or cost <= cost_prev
            cost_prev = cost
    assert (hits == 1).all()
