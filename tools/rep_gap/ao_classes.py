#!/usr/bin/env python
"""Third view of the golden rep_mea gap: the derivative integrals split by AO class.  dS/dA and dT/dA are
2a <l+1| - l <l-1| (raise / lower parts); rep_mea is linear in them, so the residual need - sum is regressed on
the contributions of {dS, dT} x {raise, lower} x {bra s, p, d} x {ket s, p, d}.  A convention difference in the
un-vendored getSA1B / getTA1B confined to one class would show up as a constant weight on its column."""
import itertools, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import scan_geometries, SCAN_SUPL, SCAN_REORD
from oracle import solefp_oracle as O, basis_oracle as B, ints_oracle as IO
from slv_b200 import slvpar
import decompose as D


def split_ints(bA, posA, bB, posB):
    """S, T and the raise / lower parts of the centre derivatives: S1r, S1l, T1r, T1l  [nA, nB, 3]."""
    nA, nB = len(bA.center), len(bB.center)
    A = posA[bA.center]; Bc = posB[bB.center]
    a = bA.exps[:, None, :, None]; b = bB.exps[None, :, None, :]
    cc = bA.coefs[:, None, :, None] * bB.coefs[None, :, None, :]
    la, lb = bA.powers, bB.powers
    shp = (nA, nB, a.shape[2], b.shape[3])
    sx = [IO._s1d(a, b, A[:, ax][:, None, None, None], Bc[:, ax][None, :, None, None], 3, 4) for ax in range(3)]

    def pick(ax, di, dj):
        out = np.zeros(shp)
        for i in range(3):
            for j in range(3):
                ii, jj = i + di, j + dj
                if ii < 0 or jj < 0:
                    continue
                mask = (la[:, ax] == i)[:, None] & (lb[:, ax] == j)[None, :]
                if mask.any():
                    out = out + np.where(mask[..., None, None], np.broadcast_to(sx[ax][ii][jj], shp), 0.0)
        return out
    lbf, laf = lb.astype(float), la.astype(float)

    def t_ax(ax, di=0):
        j = lbf[None, :, ax, None, None]
        return -0.5 * (j * (j - 1.0) * pick(ax, di, -2) - 2.0 * b * (2.0 * j + 1.0) * pick(ax, di, 0) + 4.0 * b * b * pick(ax, di, 2))
    s0 = [pick(ax, 0, 0) for ax in range(3)]; t0 = [t_ax(ax) for ax in range(3)]
    S = (cc * s0[0] * s0[1] * s0[2]).sum((2, 3))
    T = (cc * (t0[0] * s0[1] * s0[2] + s0[0] * t0[1] * s0[2] + s0[0] * s0[1] * t0[2])).sum((2, 3))
    out = [np.zeros((nA, nB, 3)) for _ in range(4)]
    for ax in range(3):
        i = laf[:, None, ax, None, None]
        o = [k for k in range(3) if k != ax]
        for part, (ds, dt) in enumerate(((2.0 * a * pick(ax, 1, 0), 2.0 * a * t_ax(ax, 1)), (-i * pick(ax, -1, 0), -i * t_ax(ax, -1)))):
            out[part][..., ax] = (cc * ds * s0[o[0]] * s0[o[1]]).sum((2, 3))
            out[2 + part][..., ax] = (cc * (dt * s0[o[0]] * s0[o[1]] + ds * t0[o[0]] * s0[o[1]] + ds * s0[o[0]] * t0[o[1]])).sum((2, 3))
    return S, T, out[0], out[1], out[2], out[3]


if __name__ == '__main__':
    golden = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'small_cluster.json')))
    pars = [slvpar.Frag('nma').get(), slvpar.Frag('water').get()]
    c = O.HartreePerHbarToCmRec; mode = 8
    w, den = O.mode_weights(pars[0], mode)
    cols, names, need, base = [], None, [], []
    for i, xyz in enumerate(scan_geometries(golden)):
        r = O.eval_frame(xyz, [0, 1], [12, 3], pars, SCAN_SUPL, SCAN_REORD, mode, elect=True, ea=True, corr=True, pol=True, disp=True)
        pinned = (r['ele_mea'] + r['ele_ea'] + r['cor_mea'] + r['cor_ea'] + r['pol_mea'] + r['dis_mea']) * c
        ra, ta, _ = O.kabsch(pars[0]['pos'], xyz[:12], SCAN_SUPL[0]); pa = O.sup_params(pars[0], ra, ta)
        rb, tb, _ = O.kabsch(pars[1]['pos'], xyz[12:], SCAN_SUPL[1]); pb = O.sup_params(pars[1], rb, tb)
        bA = B.OracleBasis(pa['atno'], pa['pos']); bB = B.OracleBasis(pb['atno'], pb['pos'])
        S, T, S1r, S1l, T1r, T1l = split_ints(bA, pa['pos'], bB, pb['pos'])
        LA, LB = bA.powers.sum(1), bB.powers.sum(1)
        full = D.fi_components(pa, pb, ra, rb, (S, T, S1r + S1l, T1r + T1l))
        tot = -(w[None, :] * full).sum(1) / den * c
        row, nm = [], []
        z = np.zeros_like(S1r)
        for q, arr in (('dSr', S1r), ('dSl', S1l), ('dTr', T1r), ('dTl', T1l)):
            for la_ in range(3):
                if q.endswith('l') and la_ == 0:
                    continue
                for lb_ in range(3):
                    m = ((LA == la_)[:, None] & (LB == lb_)[None, :])[..., None]
                    part = np.where(m, arr, 0.0)
                    comp = D.fi_components(pa, pb, ra, rb, (S, T, part if q[1] == 'S' else z, part if q[1] == 'T' else z))
                    k = 0 if q[1] == 'S' else 2
                    row.append(-(w * comp[k]).sum() / den * c); nm.append('%s%s%s' % (q, 'spd'[la_], 'spd'[lb_]))
        cols.append(row); names = nm
        need.append(golden['dat'][i][0] - pinned); base.append(tot.sum())
    G = np.array(cols); need = np.array(need); need[0] = golden['f']['rep_mea']; base = np.array(base)
    d = need - base
    np.set_printoptions(precision=3, suppress=True, linewidth=220)
    print('check: sum of dS + dT columns vs components', np.abs(G.sum(1)).max())
    print('point 0 columns:'); print(dict(zip(names, np.round(G[0], 4))))
    print('residual', d)
    best = []
    for k in (1, 2, 3):
        for cidx in itertools.combinations(range(len(names)), k):
            co, *_ = np.linalg.lstsq(G[:, cidx], d, rcond=None)
            best.append((np.abs(G[:, cidx] @ co - d).max(), [names[i] for i in cidx], np.round(1 + co, 3)))
    best.sort(key=lambda x: x[0])
    print('best 1-3 column re-weightings (max misfit over the 21 points, columns, weights):')
    for b in best[:12]:
        print('  %.3f' % b[0], b[1], b[2])
    # whole-quantity switches
    for nm_, sel in (('all dS', [i for i, n in enumerate(names) if n[1] == 'S']), ('all dT', [i for i, n in enumerate(names) if n[1] == 'T']),
                     ('all raise', [i for i, n in enumerate(names) if n[2] == 'r']), ('all lower', [i for i, n in enumerate(names) if n[2] == 'l'])):
        x = G[:, sel].sum(1)
        print('%-10s weight needed per point: min %.3f max %.3f' % (nm_, (1 + d / x).min(), (1 + d / x).max()))
