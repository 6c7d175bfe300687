"""The __host__ __device__ math of slv_b200/csrc/slv_math.cuh, compiled for the host
(tests/libhostmath.so, built by __graft_entry__.build()), against the oracle."""
import ctypes
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LIBP = os.path.join(HERE, 'libhostmath.so')
pytestmark = pytest.mark.skipif(not os.path.isfile(LIBP), reason='tests/libhostmath.so not built')
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope='module')
def L():
    lib = ctypes.CDLL(LIBP)
    lib.hm_site_field.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
    lib.hm_kabsch_rot_fast.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
    lib.hm_site_dfield.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
    return lib


def _rows(par, sfx=''):
    from slv_b200.slvpar import traceless
    qt, ot = traceless(par['dmaq' + sfx], par['dmao' + sfx])
    return np.ascontiguousarray(np.concatenate([par['dmac' + sfx][..., None], par['dmad' + sfx], qt, ot], -1))


@pytest.mark.parametrize('name,sel', [('water', None), ('nma', [1, 7, 2, 4]), ('dmso', None)])
def test_kabsch_matches_svd(L, frags, name, sel):
    from slv_b200 import synth
    from oracle import solefp_oracle as O
    rng = np.random.default_rng(0)
    ref = frags(name).get()['pos']
    ref = ref if sel is None else ref[sel]
    worst = 0.0
    for t in range(100):
        Rt = synth.random_rotations(rng, 1)[0]
        tgt = ref @ Rt + rng.normal(0, 3, 3) + rng.normal(0, 0.03, ref.shape)
        rot, tr, rms = O.kabsch(ref, tgt)
        c = ref - ref.mean(0); r = tgt - tgt.mean(0)
        a = np.ascontiguousarray(c.T @ r); out = np.zeros(9)
        L.hm_kabsch_rot(P(a), P(out))
        worst = max(worst, np.abs(out.reshape(3, 3) - rot).max())
        L.hm_kabsch_rot_fast(P(a), float((c * c).sum() + (r * r).sum()), P(out))      # shifted inverse iteration path
        worst = max(worst, np.abs(out.reshape(3, 3) - rot).max())
    assert worst < 1e-12


def test_kabsch_large_rotation_and_reflection_guard(L):
    """180-degree rotations and mirror-image targets: the result is always a proper rotation."""
    from oracle import solefp_oracle as O
    ref = np.array([[0.0, 0.2, 0.0], [1.4, -0.85, 0.0], [-1.4, -0.85, 0.0], [0.3, 0.1, 0.9]])
    for Rt in (np.diag([-1.0, -1.0, 1.0]), np.diag([1.0, 1.0, -1.0])):
        tgt = ref @ Rt
        rot, tr, rms = O.kabsch(ref, tgt)
        c = ref - ref.mean(0); r = tgt - tgt.mean(0)
        out = np.zeros(9); L.hm_kabsch_rot_fast(P(np.ascontiguousarray(c.T @ r)), float((c * c).sum() + (r * r).sum()), P(out))
        out = out.reshape(3, 3)
        assert abs(np.linalg.det(out) - 1.0) < 1e-12
        assert np.abs(out - rot).max() < 1e-10


def test_moment_and_matrix_rotation(L, frags):
    from slv_b200 import synth
    from oracle import solefp_oracle as O
    par = frags('dmso').get()
    Rt = np.ascontiguousarray(synth.random_rotations(np.random.default_rng(1), 1)[0])
    p2 = O.sup_params(par, Rt, np.zeros(3))
    m, m2 = _rows(par), _rows(p2)
    for i in range(len(m)):
        o = np.zeros(20); L.hm_rot_mom20(P(Rt), P(np.ascontiguousarray(m[i])), P(o))
        assert np.abs(o - m2[i]).max() < 1e-13
    for i in range(par['npol']):
        o = np.zeros(9); L.hm_rot_mat(P(Rt), P(np.ascontiguousarray(par['dpol'][i].ravel())), P(o))
        assert np.abs(o.reshape(3, 3) - p2['dpol'][i]).max() < 1e-13


def test_pair_potentials_energy_and_charge_field(L, frags):
    from slv_b200 import synth
    from oracle import solefp_oracle as O
    A = frags('nma').get()
    Rt = synth.random_rotations(np.random.default_rng(2), 1)[0]
    B = O.sup_params(frags('water').get(), Rt, np.array([8.0, 1.0, -2.0]))
    qA, muA, ThA, OmA = O._lab_moments(A); qB, muB, ThB, OmB = O._lab_moments(B)
    e_ref = O.emtp_r4(A['rdma'], qA, muA, ThA, OmA, B['rdma'], qB, muB, ThB, OmB)
    sA, mB = _rows(A), _rows(B)
    e = 0.0; G = np.zeros((12, 3))
    for i in range(12):
        s = sA[i].copy(); L.hm_fold20(P(s))
        for j in range(5):
            Pv = np.zeros(23)
            L.hm_pair_potentials(P(np.ascontiguousarray(A['rdma'][i])), P(np.ascontiguousarray(B['rdma'][j])), P(np.ascontiguousarray(mB[j])), P(Pv))
            e += s @ Pv[:20]; G[i] += Pv[20:]
    assert abs(e - e_ref) < 1e-15 + 1e-12 * abs(e_ref)
    R = A['rdma'][:, None, :] - B['rdma'][None, :, :]
    Gref = -(qB[None, :, None] * R / ((R * R).sum(-1) ** 1.5)[..., None]).sum(1)
    assert np.abs(G - Gref).max() < 1e-14


def test_field_and_field_derivative(L, frags):
    from oracle import solefp_oracle as O
    par = frags('nma').get(); m = _rows(par)
    q, mu, Th, Om = O._lab_moments(par)
    rng = np.random.default_rng(1)
    for s in range(12):
        R = rng.normal(0, 4, 3); v = rng.normal(0, 1, 3)
        for fac in (5.0, 7.0):
            F = np.zeros(3); L.hm_site_field(P(R), P(np.ascontiguousarray(m[s])), fac, P(F))
            assert np.abs(F - O._field_traceless(R, q[s], mu[s], Th[s], Om[s], fac)).max() < 1e-15
        F = np.zeros(3); L.hm_site_dfield(P(R), P(v), P(np.ascontiguousarray(m[s])), 1.0, P(F))
        assert np.abs(F - O._dfield_traceless(R, v, q[s], mu[s], Th[s], Om[s])).max() < 1e-15


def test_lean_field_equals_generic(L):
    L.hm_site_field_acc.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    rng = np.random.default_rng(0)
    for t in range(20):
        R = rng.normal(0, 3, 3); m = rng.normal(0, 1, 20); F1 = np.zeros(3); F2 = np.zeros(3); F3 = np.ones(3)
        L.hm_site_field(P(R), P(m), 5.0, P(F1)); L.hm_site_field_acc(P(R), P(m), 5.0, 1.0, P(F2))
        assert np.abs(F1 - F2).max() < 1e-14 * np.abs(F1).max()
        L.hm_site_field_acc(P(np.array([1.0, 0.0, 0.0])), P(m), 5.0, 0.0, P(F3))      # masked: no contribution
        assert np.array_equal(F3, np.ones(3))


@pytest.mark.parametrize('na,nb', [('water', 'water'), ('mescn', 'dmso'), ('nma', 'meoh')])
def test_subshell_pair_integrals_match_oracle(L, frags, na, nb):
    """The sub-shell-pair specialisations k_rep runs (slv_ints.cuh subshell_pair_store, all 23 classes) against
    oracle/ints_oracle.py for two fragments in contact: S, T and the L-contracted centre derivatives."""
    from oracle import ints_oracle as IO
    from slv_b200.basis import BasisSet
    rng = np.random.default_rng(5)
    pa, pb = frags(na).get(), frags(nb).get()
    posA = np.ascontiguousarray(pa['pos'], np.float64)
    q = rng.normal(size=4); q /= np.linalg.norm(q); w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    posB = np.ascontiguousarray((pb['pos'] - pb['pos'].mean(0)) @ R + posA.mean(0) + np.array([5.5, 1.0, -0.7]))
    bA, bB = BasisSet(pa['atno'], posA), BasisSet(pb['atno'], posB)
    Lc = np.ascontiguousarray(rng.normal(size=(len(posA), 3)))
    S, T, S1, T1 = IO.one_electron(bA, posA, bB, posB)
    ref = np.stack([S, T, np.einsum('ka,kla->kl', Lc[bA.center], S1), np.einsum('ka,kla->kl', Lc[bA.center], T1)], -1)
    LP = len(bB) + 7                                   # records per row of the layout V[k][l][q] (any LP >= nbsB)
    Vp = np.full((len(bA), LP, 4), np.nan)
    eA, cA, eB, cB = [np.ascontiguousarray(v, np.float64) for v in (bA.exps, bA.coefs, bB.exps, bB.coefs)]
    sA, sB = np.ascontiguousarray(bA.subshells, np.int32), np.ascontiguousarray(bB.subshells, np.int32)
    L.hm_subshell_V(len(sA), P(sA), P(posA), P(eA), P(cA), P(Lc), len(sB), P(sB), P(posB), P(eB), P(cB), LP, P(Vp))
    V = np.ascontiguousarray(Vp[:, :len(bB)])
    assert np.isfinite(V).all()                       # every AO pair is written by exactly the classes walked
    assert np.isnan(Vp[:, len(bB):]).all()            # and nothing beyond the fragment's last basis function
    scale = np.abs(ref).max((0, 1))
    assert (np.abs(V - ref).max((0, 1)) < 1e-12 * np.maximum(scale, 1.0)).all(), np.abs(V - ref).max((0, 1))


@pytest.mark.parametrize('na,nb', [('mescn', 'dmso'), ('nma', 'water')])
def test_primitive_pair_records_are_a_sorted_screening_table(frags, na, nb):
    """The plan's tables behind the integral kernels (slv_b200/plan.py primitive_pair_records): every work item points at
    the records of its sub-shell pair; records carry (thr, mu, c_a c_b (pi/p)^1.5, 1/p, a, b), thresholds descend, and at
    any distance the primitive pairs the screening rule a b AB^2 <= 46 (a + b) keeps are exactly a prefix of the list."""
    from slv_b200.basis import BasisSet
    from slv_b200.plan import subshell_work_list, primitive_pair_records, SUBSHELL_CLASSES, PP_REC
    pa, pb = frags(na).get(), frags(nb).get()
    bA, bB = BasisSet(pa['atno'], pa['pos']), BasisSet(pb['atno'], pb['pos'])
    offs, items = subshell_work_list(bA.subshells, bB.subshells)
    irec, ptab = primitive_pair_records(bA, bB, items)
    assert irec.shape == (len(items), 8) and ptab.shape[1] == PP_REC == 6 and offs[-1] == len(items)
    rng = np.random.default_rng(9)
    nB = len(bB.subshells)
    for ci, (LA, LB, IA) in enumerate(SUBSHELL_CLASSES):
        for it in range(offs[ci], offs[ci + 1]):
            a, b = items[it] // nB, items[it] % nB
            sa, sb = bA.subshells[a], bB.subshells[b]
            assert (sa[2], sb[2]) == (LA, LB)
            assert tuple(irec[it, :4]) == (3 * sa[0], sa[1], 3 * sb[0], sb[1]) and irec[it, 5] == sa[3] * sb[3]
            r = ptab[irec[it, 4]:irec[it, 4] + irec[it, 5]]
            assert (np.diff(r[:, 0]) <= 0).all()
            ea, eb = bA.exps[sa[1], :sa[3]], bB.exps[sb[1], :sb[3]]
            ca, cb = bA.coefs[sa[1], :sa[3]], bB.coefs[sb[1], :sb[3]]
            A, B = np.repeat(ea, sb[3]), np.tile(eb, sa[3])
            want = np.stack([46.0 * (A + B) / (A * B), A * B / (A + B), np.repeat(ca, sb[3]) * np.tile(cb, sa[3]) * (np.pi / (A + B)) ** 1.5,
                             1.0 / (A + B), A, B], 1)
            o1, o2 = np.lexsort(r.T[::-1]), np.lexsort(want.T[::-1])
            assert np.allclose(r[o1], want[o2], rtol=1e-15, atol=0)            # the same set of primitive pairs
            for ab2 in rng.uniform(0.5, 200.0, 3):
                keep = r[:, 4] * r[:, 5] * ab2 <= 46.0 * (r[:, 4] + r[:, 5])
                n = int(keep.sum())
                assert keep[:n].all() and not keep[n:].any()


def test_d_component_norm_ratio_is_sqrt3_for_every_bundled_basis(frags):
    """slv_ints.cuh contracts a d sub-shell with the coefficients of its first component and rescales d_xy, d_xz, d_yz by
    the compile-time constant sqrt(3) (cart_norm_ratio).  That holds because primitive norms differ by exactly sqrt(3)
    between xx- and xy-type components while the contracted norms are equal: check it on the tables of every element the
    bundled basis set knows, for every primitive."""
    from slv_b200.basis import BasisSet
    for z in (1, 3, 6, 7, 8, 9, 11, 16, 17):
        b = BasisSet([z], np.zeros((1, 3)))
        for atom, f0, L, npr in b.subshells:
            if L != 2:
                continue
            c = b.coefs[f0:f0 + 6, :npr]
            assert np.allclose(c[1:3] / c[0], 1.0, rtol=1e-14, atol=0)
            assert np.allclose(c[3:6] / c[0], np.sqrt(3.0), rtol=1e-14, atol=0)
            assert (b.exps[f0:f0 + 6, :npr] == b.exps[f0, :npr]).all()
