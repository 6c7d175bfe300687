#!/usr/bin/env python
"""Second view of the golden rep_mea gap: contributions of the seventeen formula terms of CALCFI (shftex.f:145-287)
over the 21-point scan.  If the golden run used another version of one term (a later bug fix in shftex.f, a missing
factor), the residual need - sum is a CONSTANT multiple of that term's column."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from util import scan_geometries, SCAN_SUPL, SCAN_REORD
from oracle import solefp_oracle as O, basis_oracle as B, ints_oracle as IO
from slv_b200 import slvpar

NAMES = ['TS1', 'TS2', 'TS3', 'TS4', 'A1a', 'A1b', 'A1c', 'A1d', 'A1e', 'A2a', 'A2b', 'A2d', 'A2e', 'A3a', 'A3b', 'A4a', 'A4b']


def fi_terms(parc, pb, rotc, rb):
    bA = B.OracleBasis(parc['atno'], parc['pos']); bB = B.OracleBasis(pb['atno'], pb['pos'])
    typA = bA.powers.sum(1)
    CA = O.vecrot(parc['vecl'], rotc, typA); CA1 = O.vecrot(parc['vecl1'], rotc, typA)
    CB = O.vecrot(pb['vecl'], rb, bB.powers.sum(1))
    S, T, S1, T1 = IO.one_electron(bA, parc['pos'], bB, pb['pos'])
    L = parc['lvec']; atomA = bA.center
    Sij = CA @ S @ CB.T; Tij = CA @ T @ CB.T
    dS = np.einsum('mka,kla->mkl', L[:, atomA, :], S1); dT = np.einsum('mka,kla->mkl', L[:, atomA, :], T1)
    Sm = np.einsum('ik,mkl,jl->mij', CA, dS, CB) + np.einsum('mik,kl,jl->mij', CA1, S, CB)
    Tm = np.einsum('ik,mkl,jl->mij', CA, dT, CB) + np.einsum('mik,kl,jl->mij', CA1, T, CB)
    FA, FB, FA1 = parc['fock'], pb['fock'], parc['fock1']
    riA, riB, rnA, rnB, riA1 = parc['lmoc'], pb['lmoc'], parc['pos'], pb['pos'], parc['lmoc1']
    ZA = np.asarray(parc['atno'], float); ZB = np.asarray(pb['atno'], float)
    Rij = riA[:, None, :] - riB[None, :, :]; rij = np.linalg.norm(Rij, axis=-1)
    rm_ij = np.einsum('ija,mia->mij', Rij, riA1) / rij
    sl = np.log(np.abs(Sij)); aux = 2.0 * np.sqrt(-2.0 * sl / np.pi); sdr = Sij / rij
    Rxj = rnA[:, None, :] - riB[None, :, :]; rxj = np.linalg.norm(Rxj, axis=-1)
    rm_xj = np.einsum('xja,mxa->mxj', Rxj, L) / rxj
    Riy = riA[:, None, :] - rnB[None, :, :]; riy = np.linalg.norm(Riy, axis=-1)
    rm_iy = np.einsum('iya,mia->miy', Riy, riA1) / riy
    one = np.ones_like(Sm)
    t = [4.0 * sdr * Sm * (np.sqrt(-1.0 / (2.0 * np.pi * sl)) - aux - 1.0), 2.0 * sdr * sdr * rm_ij * (aux + 1.0),
         4.0 * Sm * Tij, 4.0 * Tm * Sij,
         -2.0 * Sm * (FA @ Sij), 8.0 * Sij * Sm * (1.0 / rij).sum(0)[None, :], -2.0 * Sij * np.einsum('mik,kj->mij', FA1, Sij),
         -2.0 * Sij * np.einsum('ik,mkj->mij', FA, Sm), -4.0 * Sij ** 2 * (rm_ij / rij ** 2).sum(1)[:, None, :] * one,
         -2.0 * Sm * (Sij @ FB.T), 8.0 * Sij * Sm * (1.0 / rij).sum(1)[:, None], -2.0 * Sij * np.einsum('jl,mil->mij', FB, Sm),
         -4.0 * Sij ** 2 * (rm_ij / rij ** 2).sum(2)[:, :, None] * one,
         2.0 * Sij ** 2 * (ZA[:, None] * rm_xj / rxj ** 2).sum(1)[:, None, :] * one, -4.0 * Sij * Sm * (ZA[:, None] / rxj).sum(0)[None, :],
         2.0 * Sij ** 2 * (ZB[None, :] * rm_iy / riy ** 2).sum(2)[:, :, None] * one, -4.0 * Sij * Sm * (ZB[None, :] / riy).sum(1)[:, None]]
    return np.array([x.sum((1, 2)) for x in t])


if __name__ == '__main__':
    golden = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'small_cluster.json')))
    pars = [slvpar.Frag('nma').get(), slvpar.Frag('water').get()]
    c = O.HartreePerHbarToCmRec; mode = 8
    w, den = O.mode_weights(pars[0], mode)
    A, need = [], []
    for i, xyz in enumerate(scan_geometries(golden)):
        r = O.eval_frame(xyz, [0, 1], [12, 3], pars, SCAN_SUPL, SCAN_REORD, mode, elect=True, ea=True, corr=True, pol=True, disp=True)
        pinned = (r['ele_mea'] + r['ele_ea'] + r['cor_mea'] + r['cor_ea'] + r['pol_mea'] + r['dis_mea']) * c
        ra, ta, _ = O.kabsch(pars[0]['pos'], xyz[:12], SCAN_SUPL[0]); pa = O.sup_params(pars[0], ra, ta)
        rb, tb, _ = O.kabsch(pars[1]['pos'], xyz[12:], SCAN_SUPL[1]); pb = O.sup_params(pars[1], rb, tb)
        A.append(-(w[None, :] * fi_terms(pa, pb, ra, rb)).sum(1) / den * c)
        need.append(golden['dat'][i][0] - pinned)
    A = np.array(A); need = np.array(need); need[0] = golden['f']['rep_mea']
    np.set_printoptions(precision=3, suppress=True, linewidth=220)
    print(NAMES)
    for i in range(21):
        print(i, A[i], '| %8.3f %8.3f' % (A[i].sum(), need[i]))
    d = need - A.sum(1)
    print('residual:', d)
    for k, n in enumerate(NAMES):
        ratio = d / A[:, k]
        print('%4s  ratio residual/term: min %8.3f max %8.3f  at point 0: %.6f' % (n, ratio.min(), ratio.max(), ratio[0]))
    np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'terms_table.npy'), np.column_stack([A, need]))
