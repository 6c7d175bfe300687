"""NumPy oracle for the one-electron integrals of the exchange-repulsion path.
TEST INFRASTRUCTURE ONLY.

The reference obtains them from the un-vendored dependency PyQuante "1.0.3-BBG"
(github.com/globulion/pyq-mod; call sites solvshift/solefp.py:1523-1526):
  getSAB  -> S_kl  = <k|l>                        [nA,nB]
  getTAB  -> T_kl  = <k|-1/2 del^2|l>             [nA,nB]
  getSA1B -> dS_kl/dA_k (centre of k moves)       [nA,nB,3]
  getTA1B -> dT_kl/dA_k                           [nA,nB,3]
Restated here with Obara-Saika 1-D recursions over contracted Cartesian Gaussians built from
the basis tables of oracle/basis_oracle.py (center, powers, exps, coefs zero-padded).
"""
import numpy as np


def _s1d(a, b, A, B, imax, jmax):
    """1-D overlap table s[i][j] (unnormalised), arrays broadcast: shapes [...]."""
    p = a + b
    P = (a * A + b * B) / p
    PA = P - A; PB = P - B
    s = [[None] * (jmax + 1) for _ in range(imax + 1)]
    s[0][0] = np.sqrt(np.pi / p) * np.exp(-a * b / p * (A - B) ** 2)
    h = 0.5 / p
    for i in range(imax + 1):
        for j in range(jmax + 1):
            if i == 0 and j == 0:
                continue
            if i > 0:
                v = PA * s[i - 1][j]
                if i > 1:
                    v = v + h * (i - 1) * s[i - 2][j]
                if j > 0:
                    v = v + h * j * s[i - 1][j - 1]
            else:
                v = PB * s[i][j - 1]
                if j > 1:
                    v = v + h * (j - 1) * s[i][j - 2]
            s[i][j] = v
    return s


def one_electron(bA, posA, bB, posB):
    """S, T, dS/dA, dT/dA between basis tables bA (at posA) and bB (at posB)."""
    nA, nB = len(bA.center), len(bB.center)
    A = posA[bA.center]; B = posB[bB.center]
    a = bA.exps[:, None, :, None]; b = bB.exps[None, :, None, :]
    ca = bA.coefs[:, None, :, None]; cb = bB.coefs[None, :, None, :]
    la = bA.powers; lb = bB.powers
    S = np.zeros((nA, nB)); T = np.zeros((nA, nB)); S1 = np.zeros((nA, nB, 3)); T1 = np.zeros((nA, nB, 3))
    shp = (nA, nB, a.shape[2], b.shape[3])
    # per axis tables for all needed raised/lowered powers, then select by each function's powers
    sx = []
    for ax in range(3):
        tab = _s1d(a, b, A[:, ax][:, None, None, None], B[:, ax][None, :, None, None], 3, 4)
        sx.append(tab)

    def pick(ax, di, dj):
        """s(la+di, lb+dj) along axis ax, zero where a power would be negative."""
        out = np.zeros(shp)
        for i in range(3):
            for j in range(3):
                ii = i + di; jj = j + dj
                if ii < 0 or jj < 0:
                    continue
                mask = (la[:, ax] == i)[:, None] & (lb[:, ax] == j)[None, :]
                if mask.any():
                    out = out + np.where(mask[..., None, None], np.broadcast_to(sx[ax][ii][jj], shp), 0.0)
        return out

    lbf = lb.astype(float); laf = la.astype(float)

    def s_ax(ax, di=0):
        return pick(ax, di, 0)

    def t_ax(ax, di=0):
        j = lbf[None, :, ax, None, None]
        return -0.5 * (j * (j - 1.0) * pick(ax, di, -2) - 2.0 * b * (2.0 * j + 1.0) * pick(ax, di, 0)
                       + 4.0 * b * b * pick(ax, di, 2))

    s0 = [s_ax(ax) for ax in range(3)]
    t0 = [t_ax(ax) for ax in range(3)]
    cc = ca * cb
    S = (cc * s0[0] * s0[1] * s0[2]).sum((2, 3))
    T = (cc * (t0[0] * s0[1] * s0[2] + s0[0] * t0[1] * s0[2] + s0[0] * s0[1] * t0[2])).sum((2, 3))
    for ax in range(3):
        i = laf[:, None, ax, None, None]
        ds = 2.0 * a * s_ax(ax, +1) - i * s_ax(ax, -1)
        dt = 2.0 * a * t_ax(ax, +1) - i * t_ax(ax, -1)
        o = [k for k in range(3) if k != ax]
        S1[..., ax] = (cc * ds * s0[o[0]] * s0[o[1]]).sum((2, 3))
        T1[..., ax] = (cc * (dt * s0[o[0]] * s0[o[1]] + ds * t0[o[0]] * s0[o[1]] + ds * s0[o[0]] * t0[o[1]])).sum((2, 3))
    return S, T, S1, T1
