"""Oracle pinning: the NumPy restatement against the reference's own known-answer data
(examples/1_small-clusters/f and dat, committed as tests/golden/small_cluster.json)."""
import numpy as np
import pytest

from util import scan_geometries, SCAN_SUPL, SCAN_REORD


def _frame(frags, xyz, **kw):
    from oracle import solefp_oracle as O
    pars = [frags('nma').get(), frags('water').get()]
    return O.eval_frame(xyz, [0, 1], [12, 3], pars, SCAN_SUPL, SCAN_REORD, 8, **kw)


def test_golden_f_all_pinned_terms(frags, golden):
    from oracle import solefp_oracle as O
    r = _frame(frags, scan_geometries(golden, 1)[0], elect=True, ea=True, corr=True, pol=True, disp=True)
    c = O.HartreePerHbarToCmRec
    f = golden['f']
    for k, tol in (('ele_mea', 2e-9), ('ele_ea', 2e-9), ('cor_mea', 2e-9), ('cor_ea', 2e-9), ('pol_mea', 2e-9),
                   ('dis_mea', 2e-9), ('dis_mea_iso', 2e-9)):
        assert abs(r[k] * c - f[k]) < tol, (k, r[k] * c, f[k])
    ele_tot = (r['ele_mea'] + r['ele_ea'] + r['cor_mea'] + r['cor_ea']) * c
    assert abs(ele_tot - f['ele_tot']) < 5e-9
    assert r['rms_c'] < 1e-6 and r['rms_smax'] < 1e-6


def test_golden_scan_ele_tot(frags, golden):
    """dat column 2 at two decimals for all 21 rotated geometries (exercises Kabsch + tensor rotation)."""
    from oracle import solefp_oracle as O
    c = O.HartreePerHbarToCmRec
    for i, xyz in enumerate(scan_geometries(golden)):
        r = _frame(frags, xyz, elect=True, ea=True, corr=True)
        tot = (r['ele_mea'] + r['ele_ea'] + r['cor_mea'] + r['cor_ea']) * c
        assert abs(round(tot, 2) - golden['dat'][i][1]) < 0.011, (i, tot, golden['dat'][i][1])


def test_golden_scan_total_minus_rep(frags, golden):
    """dat column 1 is total = ele+pol+dis+rep; rep is the one term the tree's data cannot pin
    (DESIGN.md 'parity unpinned: rep'), so check total - rep_oracle stays below total and the pinned part
    reproduces point 0 exactly."""
    from oracle import solefp_oracle as O
    c = O.HartreePerHbarToCmRec
    r = _frame(frags, scan_geometries(golden, 1)[0], elect=True, ea=True, corr=True, pol=True, disp=True)
    pinned = (r['ele_mea'] + r['ele_ea'] + r['cor_mea'] + r['cor_ea'] + r['pol_mea'] + r['dis_mea']) * c
    assert abs(pinned - (golden['f']['total'] - golden['f']['rep_mea'])) < 1e-8


def test_mollst_strict_and_order(frags):
    from oracle import solefp_oracle as O
    rc = np.zeros((2, 3))
    mols = [np.array([[3.0, 0, 0]]), np.array([[5.0, 0, 0], [5.0, 0, 0]]), np.array([[0, 7.0, 0]])]
    ic, ip, ie = O.mollst(rc, mols, 5.0, 7.0, 3.0)
    assert list(ic) == [True, False, False]      # strict '<': the molecule at exactly 5.0 is outside ccut
    assert list(ip) == [True, True, False] and list(ie) == [False, False, False]


def test_one_atom_fragment_is_pure_shift():
    from oracle import solefp_oracle as O
    rot, tr, rms = O.kabsch(np.array([[1.0, 2.0, 3.0]]), np.array([[4.0, 4.0, 4.0]]))
    assert np.allclose(rot, np.eye(3)) and np.allclose(tr, [3.0, 2.0, 1.0]) and rms == 0.0


def test_wavefunction_fixtures_pin_basis_and_overlap_derivatives(frags):
    """The reference's own parameter files pin the integral conventions of the repulsion path: the LMO
    coefficients (`vecl`) must be orthonormal under the restated overlap, C S C^T = I, and their mode
    derivatives (`vecl1`, finite differences in the .frg) must satisfy d/dQ (C S C^T) = 0, i.e.
    C1 S C^T + C S C1^T + C dS C^T = 0 with dS built from the centre derivatives the way shftex.f:157-206
    contracts them.  This fixes basis data, function order, normalisation, and the sign/centre convention of
    getSA1B (solefp.py:1521-1524) without PyQuante."""
    from oracle import ints_oracle as IO
    from oracle import basis_oracle as basis
    for name in ('nma', 'water', 'meoh'):
        p = frags(name).get()
        b = basis.OracleBasis(p['atno'], p['pos'])
        S, T, S1, T1 = IO.one_electron(b, p['pos'], b, p['pos'])
        C = p['vecl']
        assert np.abs(C @ S @ C.T - np.eye(len(C))).max() < 1e-8, name
        if p.get('vecl1') is None:
            continue
        for m in range(len(p['vecl1'])):
            dS = np.einsum('ka,kla->kl', p['lvec'][m][b.center], S1)
            dS = dS + dS.T
            C1 = p['vecl1'][m]
            assert np.abs(C1 @ S @ C.T).max() > 1e-3          # the relation is not trivially satisfied
            assert np.abs(C1 @ S @ C.T + C @ S @ C1.T + C @ dS @ C.T).max() < 2e-6, (name, m)


@pytest.mark.xfail(strict=True, reason="golden rep_mea (examples/1_small-clusters/f:16) is NOT reproduced: the derivative integrals of the "
                   "golden run came from the un-vendored PyQuante 1.0.3-BBG (getSA1B/getTA1B); see DESIGN.md section 6 and "
                   "profiles/rep_gap_r02.txt -- the gap stays visible here until it is explained")
def test_golden_f_rep_mea(frags, golden):
    from oracle import solefp_oracle as O
    r = _frame(frags, scan_geometries(golden, 1)[0], elect=True, rep=True)
    assert abs(r['rep_mea'] * O.HartreePerHbarToCmRec - golden['f']['rep_mea']) < 1e-6


def test_rep_mea_value_is_stable(frags, golden):
    """The number the oracle (and the GPU path) give for the golden geometry: 9.34798 cm-1 against 11.54893 in `f`.
    Pinned here so that a change of the integrals or of the SHFTEX restatement cannot go unnoticed while the golden
    comparison above is an expected failure."""
    from oracle import solefp_oracle as O
    r = _frame(frags, scan_geometries(golden, 1)[0], elect=True, rep=True)
    assert abs(r['rep_mea'] * O.HartreePerHbarToCmRec - 9.34798) < 1e-4
