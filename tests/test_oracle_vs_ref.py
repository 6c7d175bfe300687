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


def _rep_pair(frags, solute, partner, seed):
    """One (solute, partner) pair superimposed onto a synthetic geometry, with the wave-function arrays rotated the way
    Frag.sup does (slvpar.py:724-733), and the AO integrals of the oracle."""
    from slv_b200 import synth
    from oracle import solefp_oracle as O, basis_oracle as B, ints_oracle as IO
    sol, par = frags(solute).get(), frags(partner).get()
    X = synth.make_trajectory(sol['pos'], [par['pos']], np.zeros(1, int), 1, seed=seed, spacing=7.5)[0]
    n = sol['natoms']
    ra, ta, _ = O.kabsch(sol['pos'], X[:n]); pa = O.sup_params(sol, ra, ta)
    rb, tb, _ = O.kabsch(par['pos'], X[n:]); pb = O.sup_params(par, rb, tb)
    bA = B.OracleBasis(pa['atno'], pa['pos']); bB = B.OracleBasis(pb['atno'], pb['pos'])
    typA, typB = bA.powers.sum(1), bB.powers.sum(1)
    S, T, S1, T1 = IO.one_electron(bA, pa['pos'], bB, pb['pos'])
    return dict(pa=pa, pb=pb, ra=ra, rb=rb, bA=bA, bB=bB, typA=typA, typB=typB, S=S, T=T, S1=S1, T1=T1,
                CA=O.vecrot(pa['vecl'], ra, typA), CA1=O.vecrot(pa['vecl1'], ra, typA), CB=O.vecrot(pb['vecl'], rb, typB))


@pytest.mark.parametrize('solute,partner,mode,seed', [('nma', 'water', 8, 3), ('mescn', 'dmso', 4, 5), ('mescn', 'water', 4, 6)])
def test_shftex_restatement_matches_transpiled(frags, solute, partner, mode, seed):
    """oracle.rep_shift (einsum restatement of CALCIJ + CALCFI) against the reference's own shftex.f, transpiled, called
    with the argument list of solefp.py:1542-1545: SHFTMA and every fi_m to 1e-12 relative."""
    from oracle import solefp_oracle as O
    d = _rep_pair(frags, solute, partner, seed)
    pa, pb = d['pa'], d['pb']
    sh, fi = O.rep_shift(pa, [pb], mode, None, d['ra'], [d['rb']])
    ma, ea, fi_ref, sij, tij, smij, tmij = RN.shftex(
        pa['redmass'], pa['freq'], pa['gijk'][:, mode - 1, mode - 1], pa['lvec'], pa['lmoc'], pb['lmoc'], pa['pos'], pb['pos'],
        pa['lmoc1'], d['CA'], d['CB'], d['CA1'], d['S'], d['T'], d['S1'], d['T1'], pa['atno'], pb['atno'], len(d['bB']),
        d['bA'].center + 1, pa['fock'], pb['fock'], pa['fock1'], mode)
    assert ea == 0.0
    assert abs(sh - ma) < 1e-12 * abs(ma), (sh, ma)
    assert np.abs(fi - fi_ref).max() < 1e-12 * np.abs(fi_ref).max()
    assert np.abs(sij.reshape(len(d['CA']), -1) - d['CA'] @ d['S'] @ d['CB'].T).max() < 1e-13


def test_vecrot_vc1rot(frags):
    """VECROT / VC1ROT (efprot.f:22-283) against the oracle's coefficient rotation: p triples as vectors, d sextets as the
    symmetric matrix whose off-diagonals are the xy/xz/yz coefficients; random proper rotation."""
    from oracle import solefp_oracle as O, basis_oracle as B
    rng = np.random.default_rng(11)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    rot = q * np.sign(np.linalg.det(q))
    for name in ('nma', 'dmso', 'mescn'):
        p = frags(name).get()
        typ = B.OracleBasis(p['atno'], p['pos']).powers.sum(1)
        assert set(typ) == {0, 1, 2}
        v = RN.vecrot(p['vecl'], rot, typ)
        assert np.abs(v - O.vecrot(p['vecl'], rot, typ)).max() < 1e-14
        assert np.abs(v - p['vecl']).max() > 1e-2
        if 'vecl1' in p:
            v1 = RN.vc1rot(p['vecl1'], rot, typ)
            assert np.abs(v1 - O.vecrot(p['vecl1'], rot, typ)).max() < 1e-14


def test_product_basis_tables_equal_oracle_basis(frags):
    """The oracle parses the reference's dat/slvbas itself (oracle/basis_oracle.py); the product reads its own JSON
    conversion (slv_b200/basis.py).  Same functions, order, exponents and normalised coefficients."""
    from oracle import basis_oracle as B
    from slv_b200.basis import BasisSet
    for name in ('nma', 'water', 'dmso', 'mescn', 'na+', 'cl-', 'li+', 'so3--', 'phenol'):
        p = frags(name).get()
        a, b = B.OracleBasis(p['atno'], p['pos']), BasisSet(p['atno'], p['pos'])
        assert len(a) == len(b) == p['nbasis']
        assert (a.center == b.center).all() and (a.powers == b.powers).all() and (a.nprim == b.nprim).all()
        assert np.abs(a.exps - b.exps).max() == 0.0
        assert np.abs(a.coefs - b.coefs).max() < 1e-14 * np.abs(b.coefs).max()


@pytest.mark.parametrize('solute,partner,mode,seed', [('nma', 'water', 8, 3), ('mescn', 'water', 4, 6)])
def test_rep_force_is_the_derivative_of_the_reference_exrep_energy(frags, solute, partner, mode, seed):
    """fi_m of SHFTEX must be dE_exrep/dQ_m of the reference's own energy routine (exrep.f:140-260, transpiled): the
    solute is displaced along each normal mode -- atoms by lvec, LMO centroids by lmoc1, AO->LMO coefficients by vecl1,
    the Fock matrix by fock1, the basis functions ride on the atoms -- and the energy is differentiated by a 4-point
    central stencil.  This pins dT/dA (and dS/dA) against T and S: the derivative integrals the golden run took from
    the un-vendored PyQuante-mod are, in the oracle, the physical derivatives of the integrals EXREP consumes."""
    from oracle import solefp_oracle as O, basis_oracle as B, ints_oracle as IO
    d = _rep_pair(frags, solute, partner, seed)
    pa, pb = d['pa'], d['pb']
    _, fi = O.rep_shift(pa, [pb], mode, None, d['ra'], [d['rb']])

    def energy(m, h):
        pos = pa['pos'] + h * pa['lvec'][m]
        bA = B.OracleBasis(pa['atno'], pos)
        S, T, _, _ = IO.one_electron(bA, pos, d['bB'], pb['pos'])
        return RN.exrep(pa['lmoc'] + h * pa['lmoc1'][m], pb['lmoc'], pos, pb['pos'], pa['fock'] + h * pa['fock1'][m], pb['fock'],
                        d['CA'] + h * d['CA1'][m], d['CB'], S, T, pa['atno'], pb['atno'])

    h = 0.005
    nm = len(fi)
    modes = sorted(set(list(np.argsort(-np.abs(fi))[:4]) + [mode - 1, 0, nm - 1]))
    for m in modes:
        fd = (8.0 * (energy(m, h) - energy(m, -h)) - (energy(m, 2 * h) - energy(m, -2 * h))) / (12.0 * h)
        assert abs(fd - fi[m]) < 1e-6 * max(np.abs(fi).max(), abs(fi[m])), (m, fd, fi[m])


def test_solpol_energy(frags):
    """Energy mode (row N3): the oracle's polarisation energy -1/2 F . D^-1 F against the reference's own SOLPOL
    (solpol2.f:1153-1267, transpiled), called as at solefp.py:969."""
    from oracle import solefp_oracle as O
    sol, wat, parc, solv, rot = _cluster(frags, 7, 4)
    PAR = [parc] + solv
    QO = [RN.tracls(p['dmaq'], p['dmao']) for p in PAR]
    cat = lambda k: np.concatenate([p[k] for p in PAR])
    epol, flds, dipind = RN.solpol(cat('rdma'), cat('dmac'), cat('dmad'), np.concatenate([x[0] for x in QO]),
                                   np.concatenate([x[1] for x in QO]), cat('rpol'), np.linalg.inv(cat('dpol')),
                                   [p['ndma'] for p in PAR], [p['npol'] for p in PAR])
    F, D = O.pol_system(parc, solv)[:2]
    Fv = F.ravel()
    e = -0.5 * Fv @ np.linalg.solve(D, Fv)
    assert epol < 0 and abs(e - epol) < 1e-12 * abs(epol)
    assert np.abs(Fv - flds).max() < 1e-13 * np.abs(flds).max()
