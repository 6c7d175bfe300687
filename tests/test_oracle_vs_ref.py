"""The NumPy restatement against the reference's own Fortran kernels (transpiled into oracle/_ref
by oracle/f77c.py): rotations, traceless forms, MOLLST, the polarisation shift SFTPLI, the
exchange-repulsion shift SHFTEX."""
import numpy as np
import pytest

from oracle import ref_natives as RN

pytestmark = pytest.mark.skipif(not RN.available(), reason='oracle/_ref/libslvref.so not built (needs /root/reference)')


def _cluster(frags, nw=5, seed=0, solute='nma'):
    from slv_b200 import synth
    from oracle import solefp_oracle as O
    sol, wat = frags(solute).get(), frags('water').get()
    X = synth.make_trajectory(sol['pos'], [wat['pos']], np.zeros(nw, int), 1, seed=seed)[0]
    n = sol['natoms']
    rot, tr, _ = O.kabsch(sol['pos'], X[:n]); parc = O.sup_params(sol, rot, tr)
    solv = []
    for k in range(nw):
        r, t, _ = O.kabsch(wat['pos'], X[n + 3 * k:n + 3 * k + 3]); solv.append(O.sup_params(wat, r, t))
    return sol, wat, parc, solv, rot


def test_rotations_and_traceless(frags):
    from oracle import solefp_oracle as O
    sol, wat, parc, solv, rot = _cluster(frags)
    d, q, o = RN.rotdma(sol['dmad'], sol['dmaq'], sol['dmao'], rot)
    assert np.abs(d - parc['dmad']).max() < 1e-14 and np.abs(q - parc['dmaq']).max() < 1e-14 and np.abs(o - parc['dmao']).max() < 1e-13
    d1, q1, o1 = RN.rotdm1(sol['dmad1'], sol['dmaq1'], sol['dmao1'], rot, 30, 12)
    assert np.abs(o1 - parc['dmao1'].ravel()).max() < 1e-12 and np.abs(q1 - parc['dmaq1'].ravel()).max() < 1e-13
    qt, ot = RN.tracls(parc['dmaq'], parc['dmao'])
    Th, Om = O.traceless_full(O.quad_full(parc['dmaq']), O.oct_full(parc['dmao']))
    assert np.abs(O.quad_reduced(Th) - qt).max() < 1e-14 and np.abs(O.oct_reduced(Om) - ot).max() < 1e-13
    from slv_b200.slvpar import traceless
    q2, o2 = traceless(parc['dmaq'], parc['dmao'])
    assert np.abs(q2 - qt).max() < 1e-15 and np.abs(o2 - ot).max() < 1e-14


def test_mollst(frags):
    from oracle import solefp_oracle as O
    rng = np.random.default_rng(3)
    rc = rng.normal(0, 2, (12, 3))
    mols = [rng.normal(0, 9, 3) + rng.normal(0, 1, (n, 3)) for n in (3, 3, 10, 1, 3, 7)]
    ic, ip, ie = O.mollst(rc, mols, 12.0, 9.0, 6.0)
    m = RN.mollst(rc, np.concatenate(mols), [len(x) for x in mols], 12.0, 9.0, 6.0)
    assert list(m[3]) == list(ic) and list(m[4]) == list(ip) and list(m[5]) == list(ie)


@pytest.mark.parametrize('nw,seed', [(1, 1), (6, 2)])
def test_sftpli(frags, nw, seed):
    from oracle import solefp_oracle as O
    sol, wat, parc, solv, rot = _cluster(frags, nw, seed)
    PAR = [parc] + solv
    QO = [RN.tracls(p['dmaq'], p['dmao']) for p in PAR]
    cat = lambda k: np.concatenate([p[k] for p in PAR])
    qadc1, octc1 = RN.tracl1(parc['dmaq1'], parc['dmao1'], 30, 12)
    pol = cat('dpol')
    mode = 8
    epol, spol, fi, avec, dipind, flds = RN.sftpli(
        cat('rdma'), cat('dmac'), cat('dmad'), np.concatenate([x[0] for x in QO]), np.concatenate([x[1] for x in QO]),
        parc['dmac1'], parc['dmad1'], qadc1, octc1, cat('rpol'), np.linalg.inv(pol), parc['redmass'], parc['freq'],
        parc['gijk'][:, mode - 1, mode - 1], parc['lmoc1'], parc['dpol1'], parc['lvec'],
        [p['ndma'] for p in PAR], [p['npol'] for p in PAR], mode, 12, 20)
    sh, fi2, dip2, av2 = O.pol_shift(parc, solv, mode)
    assert abs(sh - spol) < 1e-13 * max(1.0, abs(spol) * 1e3)
    assert np.abs(fi2 - fi).max() < 1e-12 * np.abs(fi).max()
    assert np.abs(dip2.ravel() - dipind).max() < 1e-12 * np.abs(dipind).max()
    assert np.abs(av2.ravel() - avec).max() < 1e-12 * np.abs(avec).max()
    sh3, _, _, _ = O.pol_shift(parc, solv, mode, literal_gemm=True)
    assert abs(sh3 - spol) < 1e-12 * max(1.0, abs(spol) * 1e3)
