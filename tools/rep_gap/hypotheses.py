#!/usr/bin/env python
"""Single-point hypothesis tests for the golden rep_mea (examples/1_small-clusters/f: 11.548932729375487 cm-1).
Each hypothesis is a concrete, plausible slip in what the golden run fed SHFTEX (atom map of the basis functions,
layout of the derivative integrals, rotation of the displacement vectors, ...); it must hit the full-precision
value, not fit a scan.  Output: the value each hypothesis gives."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from util import scan_geometries, SCAN_SUPL, SCAN_REORD
from oracle import solefp_oracle as O, basis_oracle as B, ints_oracle as IO
from slv_b200 import slvpar

GOLD = 11.548932729375487
c = O.HartreePerHbarToCmRec


def setup():
    golden = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'small_cluster.json')))
    pars = [slvpar.Frag('nma').get(), slvpar.Frag('water').get()]
    xyz = scan_geometries(golden)[0]
    ra, ta, _ = O.kabsch(pars[0]['pos'], xyz[:12], SCAN_SUPL[0]); pa = O.sup_params(pars[0], ra, ta)
    rb, tb, _ = O.kabsch(pars[1]['pos'], xyz[12:], SCAN_SUPL[1]); pb = O.sup_params(pars[1], rb, tb)
    return pars, pa, pb, ra, rb


def shftex(pa, pb, ra, rb, mode=8, *, L_ij=None, atom_map=None, S1=None, T1=None, CA1=None, FA1=None, riA1=None, L_nuc=None, ints=None):
    """SHFTEX with every mode-derivative input replaceable.  L_ij: displacement vectors used with the derivative
    integrals [nmodes, nbsA, 3]; L_nuc: those of the nuclear term [nmodes, natA, 3]."""
    bA = B.OracleBasis(pa['atno'], pa['pos']); bB = B.OracleBasis(pb['atno'], pb['pos'])
    typA = bA.powers.sum(1)
    CA = O.vecrot(pa['vecl'], ra, typA)
    CA1 = O.vecrot(pa['vecl1'], ra, typA) if CA1 is None else CA1
    CB = O.vecrot(pb['vecl'], rb, bB.powers.sum(1))
    S, T, S1_, T1_ = ints if ints is not None else IO.one_electron(bA, pa['pos'], bB, pb['pos'])
    S1 = S1_ if S1 is None else S1; T1 = T1_ if T1 is None else T1
    L = pa['lvec']
    amap = bA.center if atom_map is None else atom_map
    Lk = L[:, amap, :] if L_ij is None else L_ij
    Ln = L if L_nuc is None else L_nuc
    FA1 = pa['fock1'] if FA1 is None else FA1
    riA1 = pa['lmoc1'] if riA1 is None else riA1
    Sij = CA @ S @ CB.T; Tij = CA @ T @ CB.T
    dS = np.einsum('mka,kla->mkl', Lk, S1); dT = np.einsum('mka,kla->mkl', Lk, T1)
    Sm = np.einsum('ik,mkl,jl->mij', CA, dS, CB) + np.einsum('mik,kl,jl->mij', CA1, S, CB)
    Tm = np.einsum('ik,mkl,jl->mij', CA, dT, CB) + np.einsum('mik,kl,jl->mij', CA1, T, CB)
    FA, FB = pa['fock'], pb['fock']
    riA, riB, rnA, rnB = pa['lmoc'], pb['lmoc'], pa['pos'], pb['pos']
    ZA = np.asarray(pa['atno'], float); ZB = np.asarray(pb['atno'], float)
    Rij = riA[:, None, :] - riB[None, :, :]; rij = np.linalg.norm(Rij, axis=-1)
    rm_ij = np.einsum('ija,mia->mij', Rij, riA1) / rij
    sl = np.log(np.abs(Sij)); aux = 2.0 * np.sqrt(-2.0 * sl / np.pi); sdr = Sij / rij
    TS = (4.0 * sdr * Sm * (np.sqrt(-1.0 / (2.0 * np.pi * sl)) - aux - 1.0) + 2.0 * sdr * sdr * rm_ij * (aux + 1.0)
          + 4.0 * Sm * Tij + 4.0 * Tm * Sij)
    TA1 = (-2.0 * Sm * (FA @ Sij) + 8.0 * Sij * Sm * (1.0 / rij).sum(0)[None, :]
           - 2.0 * Sij * (np.einsum('mik,kj->mij', FA1, Sij) + np.einsum('ik,mkj->mij', FA, Sm))
           - 4.0 * Sij ** 2 * (rm_ij / rij ** 2).sum(1)[:, None, :])
    TA2 = (-2.0 * Sm * (Sij @ FB.T) + 8.0 * Sij * Sm * (1.0 / rij).sum(1)[:, None]
           - 2.0 * Sij * np.einsum('jl,mil->mij', FB, Sm) - 4.0 * Sij ** 2 * (rm_ij / rij ** 2).sum(2)[:, :, None])
    Rxj = rnA[:, None, :] - riB[None, :, :]; rxj = np.linalg.norm(Rxj, axis=-1)
    rm_xj = np.einsum('xja,mxa->mxj', Rxj, Ln) / rxj
    TA3 = 2.0 * Sij ** 2 * (ZA[:, None] * rm_xj / rxj ** 2).sum(1)[:, None, :] - 4.0 * Sij * Sm * (ZA[:, None] / rxj).sum(0)[None, :]
    Riy = riA[:, None, :] - rnB[None, :, :]; riy = np.linalg.norm(Riy, axis=-1)
    rm_iy = np.einsum('iya,mia->miy', Riy, riA1) / riy
    TA4 = 2.0 * Sij ** 2 * (ZB[None, :] * rm_iy / riy ** 2).sum(2)[:, :, None] - 4.0 * Sij * Sm * (ZB[None, :] / riy).sum(1)[:, None]
    fi = (TS + TA1 + TA2 + TA3 + TA4).sum((1, 2))
    w, den = O.mode_weights(pa, mode)
    return -(w * fi).sum() / den * c


if __name__ == '__main__':
    pars, pa, pb, ra, rb = setup()
    bA = B.OracleBasis(pa['atno'], pa['pos']); bB = B.OracleBasis(pb['atno'], pb['pos'])
    ints = IO.one_electron(bA, pa['pos'], bB, pb['pos'])
    S, T, S1, T1 = ints
    L = pa['lvec']; nm, nat = L.shape[0], L.shape[1]
    base = shftex(pa, pb, ra, rb, ints=ints)
    out = [('baseline (this repo)', base)]
    flat = np.concatenate([L.reshape(-1), np.zeros(3 * nat)])

    def shifted(d):                      # MLIST off by d: flat Fortran indexing into LVEC(NMODES*NATA*3)
        Lk = np.zeros((nm, len(bA.center), 3))
        for m in range(nm):
            for k, a in enumerate(bA.center):
                i0 = nat * 3 * m + 3 * (a + d)
                Lk[m, k] = flat[i0:i0 + 3] if i0 >= 0 else 0.0
        return Lk
    out.append(('MLIST one too large (get_bfsl already 1-based)', shftex(pa, pb, ra, rb, ints=ints, L_ij=shifted(1))))
    out.append(('MLIST one too small', shftex(pa, pb, ra, rb, ints=ints, L_ij=shifted(-1))))
    out.append(('derivative integrals w.r.t. the ket centre (sign)', shftex(pa, pb, ra, rb, ints=ints, S1=-S1, T1=-T1)))
    out.append(('kinetic derivative w.r.t. the ket centre only', shftex(pa, pb, ra, rb, ints=ints, T1=-T1)))
    out.append(('unrotated displacement vectors with the integrals', shftex(pa, pb, ra, rb, ints=ints, L_ij=pars[0]['lvec'][:, bA.center, :])))
    out.append(('unrotated displacement vectors everywhere', shftex(pa, pb, ra, rb, ints=ints, L_ij=pars[0]['lvec'][:, bA.center, :], L_nuc=pars[0]['lvec'])))
    out.append(('lvec laid out [atom][xyz][mode] (slv_check convention) read as [mode][atom][xyz]',
                shftex(pa, pb, ra, rb, ints=ints, L_ij=np.transpose(L, (1, 2, 0)).reshape(nm, nat, 3)[:, bA.center, :],
                       L_nuc=np.transpose(L, (1, 2, 0)).reshape(nm, nat, 3))))
    # derivative integrals stored [3][k][l] but read as [k][l][3]
    out.append(('S1/T1 stored [xyz][k][l], read [k][l][xyz]', shftex(pa, pb, ra, rb, ints=ints, S1=np.transpose(S1, (2, 0, 1)).reshape(S1.shape), T1=np.transpose(T1, (2, 0, 1)).reshape(T1.shape))))
    out.append(('S1/T1 stored [k][xyz][l], read [k][l][xyz]', shftex(pa, pb, ra, rb, ints=ints, S1=np.transpose(S1, (0, 2, 1)).reshape(S1.shape), T1=np.transpose(T1, (0, 2, 1)).reshape(T1.shape))))
    out.append(('vecl1 not rotated', shftex(pa, pb, ra, rb, ints=ints, CA1=pars[0]['vecl1'])))
    out.append(('no kinetic derivative integrals (T1 = 0)', shftex(pa, pb, ra, rb, ints=ints, T1=0 * T1)))
    out.append(('T1 := S1 (copy-paste slip)', shftex(pa, pb, ra, rb, ints=ints, T1=S1)))
    out.append(('kinetic derivative doubled', shftex(pa, pb, ra, rb, ints=ints, T1=2 * T1)))
    # deuterated parameter set
    pd = slvpar.Frag('nma-d7').get()
    xyz = scan_geometries(json.load(open(os.path.join(ROOT, 'tests', 'golden', 'small_cluster.json'))))[0]
    rd, td, _ = O.kabsch(pd['pos'], xyz[:12], SCAN_SUPL[0]); pdd = O.sup_params(pd, rd, td)
    out.append(('nma-d7 parameters', shftex(pdd, pb, rd, rb)))
    pw2 = slvpar.Frag('water2').get()
    try:
        r2, t2, _ = O.kabsch(pw2['pos'], xyz[12:], SCAN_SUPL[1]); pb2 = O.sup_params(pw2, r2, t2)
        out.append(('water2 as the solvent', shftex(pa, pb2, ra, r2)))
    except Exception as e:
        out.append(('water2 as the solvent: %s' % type(e).__name__, float('nan')))
    for name, v in out:
        print('%-85s %14.8f  %s' % (name, v, 'MATCH' if abs(v - GOLD) < 1e-5 else 'delta %+.5f' % (v - GOLD)))
