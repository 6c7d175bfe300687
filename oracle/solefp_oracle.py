"""NumPy restatement of the reference's SolEFP central-mode evaluation.

TEST INFRASTRUCTURE -- never imported by the product (slv_b200/).

Follows solvshift/solefp.py:1011-1629 (``EFP._eval_mode_central``), solvshift/slvpar.py:641-734
(``Frag.sup``), solvshift/src/efprot.f (rotations), solvshift/src/solpol2.f:3-115 (``MOLLST``).
The multipole electrostatics (``sdmtpm/sdmtpe/dmtcor``), the dispersion shift
(``sdisp6/tdisp6``) and the Kabsch superimposer live in the un-vendored dependency
``libbbg`` (>= 1.1.0, github.com/globulion/libbbg; no lockfile pins it): their published
algorithm is restated from SURVEY.md Appendix A and anchored on the reference call sites
solefp.py:1175-1190, 1434-1435, slvpar.py:646-652 and on the golden vector
examples/1_small-clusters/f.

Everything here works with full (unreduced) Cartesian tensors and einsum, i.e. a formulation
independent of the reduced-storage CUDA kernels it checks.
"""
import numpy as np

BohrToAngstrom = 0.529177249
HartreePerHbarToCmRec = 219474.63067

# reduced-storage index maps (efprot.f:312-320)
_Q_IDX = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
_O_IDX = [(0, 0, 0), (1, 1, 1), (2, 2, 2), (0, 0, 1), (0, 0, 2), (0, 1, 1), (1, 1, 2), (0, 2, 2), (1, 2, 2), (0, 1, 2)]


def quad_full(q):
    """[...,6] reduced -> [...,3,3] symmetric."""
    q = np.asarray(q)
    out = np.zeros(q.shape[:-1] + (3, 3))
    for k, (a, b) in enumerate(_Q_IDX):
        out[..., a, b] = q[..., k]
        out[..., b, a] = q[..., k]
    return out


def oct_full(o):
    """[...,10] reduced -> [...,3,3,3] fully symmetric."""
    import itertools
    o = np.asarray(o)
    out = np.zeros(o.shape[:-1] + (3, 3, 3))
    for k, abc in enumerate(_O_IDX):
        for perm in set(itertools.permutations(abc)):
            out[(Ellipsis,) + perm] = o[..., k]
    return out


def quad_reduced(Q):
    return np.stack([Q[..., a, b] for a, b in _Q_IDX], axis=-1)


def oct_reduced(O):
    return np.stack([O[..., a, b, c] for a, b, c in _O_IDX], axis=-1)


def traceless_full(Q, O):
    """Buckingham traceless forms (efprot.f:1256-1292)."""
    d = np.eye(3)
    trq = np.einsum('...aa->...', Q)
    Th = 1.5 * Q - 0.5 * trq[..., None, None] * d
    t = np.einsum('...abb->...a', O)
    Om = 2.5 * O - 0.5 * (np.einsum('...a,bc->...abc', t, d) + np.einsum('...b,ac->...abc', t, d)
                          + np.einsum('...c,ab->...abc', t, d))
    return Th, Om


# --------------------------------------------------------------------------- Kabsch
def kabsch(ref, target, supl=None):
    """Biopython-style SVDSuperimposer as used through libbbg (slvpar.py:641-679):
    rotate/translate the parameter geometry <ref> onto <target> using atoms <supl>.
    Row-vector convention x' = x.rot + transl.  One-atom fragments: identity + shift
    (slvpar.py:668-671).  Returns rot, transl, rms."""
    ref = np.asarray(ref, float); target = np.asarray(target, float)
    if len(target) == 1:
        return np.eye(3), (target - ref).reshape(3), 0.0
    if supl is not None:
        coords = ref[supl]; reference = target[supl]
    else:
        coords = ref; reference = target
    av1 = coords.mean(0); av2 = reference.mean(0)
    c = coords - av1; r = reference - av2
    a = c.T @ r
    u, d, vt = np.linalg.svd(a)
    rot = (vt.T @ u.T).T
    if np.linalg.det(rot) < 0:
        vt[2] = -vt[2]
        rot = (vt.T @ u.T).T
    tran = av2 - av1 @ rot
    diff = coords @ rot + tran - reference
    rms = np.sqrt((diff * diff).sum() / len(coords))
    return rot, tran, rms


def reorder_xyz(xyz, reord):
    """solefp.py:1633-1645: row i of the result is xyz[reord[i]], zeros where reord[i] < 0."""
    out = np.zeros((len(reord), 3))
    for i, I in enumerate(reord):
        if I > -1:
            out[i] = xyz[I]
    return out


def mollst(rc, rm_list, ccut, pcut, ecut):
    """MOLLST (solpol2.f:3-115): unweighted centre of the solute vs unweighted centre of each
    solvent molecule; in layer iff distance < cut (strict)."""
    c0 = rc.sum(0) / float(len(rc))
    icm, ipm, iem = [], [], []
    for xyz in rm_list:
        c = xyz.sum(0) / float(len(xyz))
        d = c0 - c
        rr = np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        icm.append(rr < ccut); ipm.append(rr < pcut); iem.append(rr < ecut)
    return np.array(icm, bool), np.array(ipm, bool), np.array(iem, bool)


# --------------------------------------------------------------------------- rotation of parameters
def sup_params(par, rot, transl):
    """Frag.sup body (slvpar.py:680-733): returns a new dict with every tensor transformed.
    Multipoles rotate as mu' = mu.R, Q' = R^T Q R, O'_ijk = O_abc R_ai R_bj R_ck (efprot.f:297-301);
    polarisabilities alpha' = R^T alpha R (slvpar.py:707-722)."""
    R = np.asarray(rot); t = np.asarray(transl)
    out = dict(par)
    for k in ('pos', 'lmoc', 'rpol', 'rdma'):
        if k in par:
            out[k] = par[k] @ R + t
    for k in ('lmoc1', 'lvec'):
        if k in par:
            out[k] = par[k] @ R
    for sfx in ('', '1', '2'):
        if 'dmad' + sfx in par:
            out['dmad' + sfx] = par['dmad' + sfx] @ R
            Q = quad_full(par['dmaq' + sfx]); O = oct_full(par['dmao' + sfx])
            out['dmaq' + sfx] = quad_reduced(np.einsum('...ab,ai,bj->...ij', Q, R, R))
            out['dmao' + sfx] = oct_reduced(np.einsum('...abc,ai,bj,ck->...ijk', O, R, R, R))
    for k in ('dpol', 'dpol1', 'dpoli', 'dpoli1'):
        if k in par:
            out[k] = np.einsum('ai,...ab,bj->...ij', R, par[k], R)
    return out


def _lab_moments(par, sfx=''):
    """q, mu, Theta(full, traceless), Omega(full, traceless) for suffix '', '1' or '2'."""
    q = par['dmac' + sfx]; mu = par['dmad' + sfx]
    Th, Om = traceless_full(quad_full(par['dmaq' + sfx]), oct_full(par['dmao' + sfx]))
    return q, mu, Th, Om


# --------------------------------------------------------------------------- electrostatics
def _T_tensors(R):
    """Interaction tensors of 1/R up to rank 3 for R[...,3]."""
    d = np.eye(3)
    r2 = (R * R).sum(-1); r = np.sqrt(r2)
    T0 = 1.0 / r
    T1 = -R / r[..., None] ** 3
    T2 = (3.0 * np.einsum('...a,...b->...ab', R, R) - r2[..., None, None] * d) / r[..., None, None] ** 5
    RRR = np.einsum('...a,...b,...c->...abc', R, R, R)
    Rd = (np.einsum('...a,bc->...abc', R, d) + np.einsum('...b,ac->...abc', R, d) + np.einsum('...c,ab->...abc', R, d))
    T3 = -(15.0 * RRR - 3.0 * r2[..., None, None, None] * Rd) / r[..., None, None, None] ** 7
    return T0, T1, T2, T3


def emtp_r4(rA, qA, muA, ThA, OmA, rB, qB, muB, ThB, OmB):
    """Multipole-multipole interaction energy between site sets A and B, orders R^-1..R^-4
    only (the truncation ``sdmtpm`` applies: q-q; q-mu; q-Theta + mu-mu; q-Omega + mu-Theta).
    Moment arrays of A may carry leading batch axes before the site axis."""
    R = rB[None, :, :] - rA[:, None, :]                      # [nA,nB,3]
    T0, T1, T2, T3 = _T_tensors(R)
    e = np.einsum('...i,ij,j->...', qA, T0, qB)
    e = e + np.einsum('...i,ija,ja->...', qA, T1, muB) - np.einsum('...ia,ija,j->...', muA, T1, qB)
    e = e + (np.einsum('...i,ijab,jab->...', qA, T2, ThB) + np.einsum('...iab,ijab,j->...', ThA, T2, qB)) / 3.0 \
          - np.einsum('...ia,ijab,jb->...', muA, T2, muB)
    e = e + (np.einsum('...i,ijabc,jabc->...', qA, T3, OmB) - np.einsum('...iabc,ijabc,j->...', OmA, T3, qB)) / 15.0 \
          - np.einsum('...ia,ijabc,jbc->...', muA, T3, ThB) / 3.0 + np.einsum('...iab,ijabc,jc->...', ThA, T3, muB) / 3.0
    return e


def mode_weights(parc, mode):
    """w_m = g_mjj/(M_m w_m^2), den = 2 M_j w_j  (mode is 1-based, as in the reference API)."""
    gijj = parc['gijk'][:, mode - 1, mode - 1]
    w = gijj / (parc['redmass'] * parc['freq'] ** 2)
    den = 2.0 * parc['redmass'][mode - 1] * parc['freq'][mode - 1]
    return w, den


def elect_shifts(parc, solv, mode, ea=True, corr=True):
    """ele_mea, ele_ea, cor_mea, cor_ea (a.u.) and fi_el[nmodes] for an already superimposed
    solute <parc> and list of superimposed solvent dicts <solv> (solefp.py:1167-1190)."""
    w, den = mode_weights(parc, mode)
    nmodes = len(w)
    if len(solv) == 0:
        return 0.0, 0.0, 0.0, 0.0, np.zeros(nmodes)
    rB = np.concatenate([s['rdma'] for s in solv])
    qB = np.concatenate([s['dmac'] for s in solv])
    muB = np.concatenate([s['dmad'] for s in solv])
    QB = np.concatenate([quad_full(s['dmaq']) for s in solv])
    OB = np.concatenate([oct_full(s['dmao']) for s in solv])
    ThB, OmB = traceless_full(QB, OB)
    rA = parc['rdma']
    q1, mu1, Th1, Om1 = _lab_moments(parc, '1')             # [nmodes, ndma, ...]
    fi = emtp_r4(rA, q1, mu1, Th1, Om1, rB, qB, muB, ThB, OmB)
    mea = -(w * fi).sum() / den
    eav = 0.0
    if ea:
        m2 = list(parc['mode']).index(mode - 1)
        q2 = parc['dmac2'][m2]; mu2 = parc['dmad2'][m2]
        Th2, Om2 = traceless_full(quad_full(parc['dmaq2'][m2]), oct_full(parc['dmao2'][m2]))
        eav = emtp_r4(rA, q2, mu2, Th2, Om2, rB, qB, muB, ThB, OmB) / den
    cmea = cea = 0.0
    if corr:
        # gradient of the potential of the solvent CHARGES at the solute sites (= atoms)
        R = rA[:, None, :] - rB[None, :, :]
        r3 = ((R * R).sum(-1)) ** 1.5
        G = -(qB[None, :, None] * R / r3[..., None]).sum(1)         # grad V at each solute site
        L = parc['lvec']                                            # [nmodes, natoms, 3]
        qA = parc['dmac']
        per_mode = np.einsum('x,mxa,xa->m', qA, L, G)
        cmea = -(w * per_mode).sum() / den
        cea = 2.0 * np.einsum('x,xa,xa->', parc['dmac1'][mode - 1], L[mode - 1], G) / den
    return mea, eav, cmea, cea, fi


# --------------------------------------------------------------------------- dispersion
_GL12 = np.array([
    [0.0471753363865118, -0.9815606342467192], [0.1069393259953184, -0.9041172563704749],
    [0.1600783285433462, -0.7699026741943047], [0.2031674267230659, -0.5873179542866175],
    [0.2334925365383548, -0.3678314989981802], [0.2491470458134028, -0.1252334085114689],
    [0.2491470458134028, 0.1252334085114689], [0.2334925365383548, 0.3678314989981802],
    [0.2031674267230659, 0.5873179542866175], [0.1600783285433462, 0.7699026741943047],
    [0.1069393259953184, 0.9041172563704749], [0.0471753363865118, 0.9815606342467192]])


def gl12_weights(v=0.3):
    """slvpar.py:989-1019."""
    return 2.0 * v * _GL12[:, 0] / (1.0 - _GL12[:, 1]) ** 2


def _disp_energy(ri, ai, rj, aj, W):
    """E_iso, E_ani for solute LMO set (ri [n,3], ai [n,12,3,3]) vs solvent (rj, aj).
    ani: -(1/2pi) sum W_k T_ab aj_cb T_cd ai_da ; iso: -(3/pi) sum W_k abar_i abar_j / R^6."""
    R = ri[:, None, :] - rj[None, :, :]
    r2 = (R * R).sum(-1)
    T = (3.0 * np.einsum('ija,ijb->ijab', R, R) - r2[..., None, None] * np.eye(3)) / r2[..., None, None] ** 2.5
    ani = -np.einsum('k,ijab,jkcb,ijcd,ikda->', W, T, aj, T, ai) / (2.0 * np.pi)
    abi = np.einsum('ikaa->ik', ai) / 3.0; abj = np.einsum('jkaa->jk', aj) / 3.0
    iso = -(3.0 / np.pi) * np.einsum('k,ik,jk,ij->', W, abi, abj, 1.0 / r2 ** 3)
    return iso, ani


def disp_shifts(parc, solv, mode, h=None):
    """dis_mea_iso, dis_mea (a.u.): -sum_m w_m dE/dQ_m / den with dE/dQ_m acting on the solute
    polarisabilities (dpoli1) and LMO centroids (lmoc1) (solefp.py:1427-1435).
    The centroid derivative is taken analytically through the linear pre-contraction
    u_i = sum_m w_m lmoc1_m and a complex-step-free central difference is NOT used: the
    directional derivative is evaluated exactly by differentiating T."""
    w, den = mode_weights(parc, mode)
    if len(solv) == 0:
        return 0.0, 0.0
    W = gl12_weights()
    ri = parc['rpol']; ai = parc['dpoli']
    rj = np.concatenate([s['rpol'] for s in solv]); aj = np.concatenate([s['dpoli'] for s in solv])
    A1 = np.tensordot(w, parc['dpoli1'], (0, 0))                 # [npol,12,3,3]
    u = np.tensordot(w, parc['lmoc1'], (0, 0))                   # [npol,3]
    # (1) polarisability-derivative part: E is linear in ai
    iso1, ani1 = _disp_energy(ri, A1, rj, aj, W)
    # (2) centroid part: d/de E(ri + e u) at e=0, analytic
    R = ri[:, None, :] - rj[None, :, :]
    r2 = (R * R).sum(-1); Ru = np.einsum('ija,ia->ij', R, u)
    d = np.eye(3)
    T = (3.0 * np.einsum('ija,ijb->ijab', R, R) - r2[..., None, None] * d) / r2[..., None, None] ** 2.5
    dT = (3.0 * (np.einsum('ia,ijb->ijab', u, R) + np.einsum('ija,ib->ijab', R, u)) - 2.0 * Ru[..., None, None] * d) \
        / r2[..., None, None] ** 2.5 - 5.0 * Ru[..., None, None] * T / r2[..., None, None]
    ani2 = -(np.einsum('k,ijab,jkcb,ijcd,ikda->', W, dT, aj, T, ai)
             + np.einsum('k,ijab,jkcb,ijcd,ikda->', W, T, aj, dT, ai)) / (2.0 * np.pi)
    abi = np.einsum('ikaa->ik', ai) / 3.0; abj = np.einsum('jkaa->jk', aj) / 3.0
    iso2 = -(3.0 / np.pi) * np.einsum('k,ik,jk,ij->', W, abi, abj, -6.0 * Ru / r2 ** 4)
    return -(iso1 + iso2) / den, -(ani1 + ani2) / den


# --------------------------------------------------------------------------- polarisation
def _field_traceless(R, q, mu, Th, Om, oct_fac):
    """Field at displacement R = P - r_site of a site with Buckingham-traceless moments:
    qR/R^3 + 3(mu.R)R/R^5 - mu/R^3 + 5(R Th R)R/R^7 - 2 Th R/R^5 + oct_fac (Om RRR) R/R^9 - 3 (Om RR)/R^7.
    oct_fac = 7 is the physical value; FFFFFF uses 5 (solpol2.f:1425).  Shapes: R [...,3],
    moments broadcastable against the leading axes of R."""
    r2 = (R * R).sum(-1); ri = 1.0 / np.sqrt(r2)
    r3 = ri ** 3; r5 = r3 * ri * ri; r7 = r5 * ri * ri; r9 = r7 * ri * ri
    mR = (mu * R).sum(-1)
    ThR = np.einsum('...ab,...b->...a', Th, R); RThR = (ThR * R).sum(-1)
    OmRR = np.einsum('...abc,...b,...c->...a', Om, R, R); OmRRR = (OmRR * R).sum(-1)
    return (q * r3)[..., None] * R + (3.0 * mR * r5)[..., None] * R - r3[..., None] * mu \
        + (5.0 * RThR * r7)[..., None] * R - 2.0 * r5[..., None] * ThR \
        + (oct_fac * OmRRR * r9)[..., None] * R - 3.0 * r7[..., None] * OmRR


def _dfield_traceless(R, v, q, mu, Th, Om):
    """Directional derivative along v (of the field point) of the PHYSICAL field (oct factor 7),
    with the reference's one slip reproduced: in the auxiliary sum SUM2 = Om_abc v_a R_b R_c the term
    O_xzz v_z R_x R_z is counted once instead of twice (solpol2.f:466-483 and 959-976)."""
    r2 = (R * R).sum(-1); ri2 = 1.0 / r2; ri = np.sqrt(ri2)
    r3 = ri ** 3; r5 = r3 * ri2; r7 = r5 * ri2; r9 = r7 * ri2
    Rv = (R * v).sum(-1)
    mR = (mu * R).sum(-1); mv = (mu * v).sum(-1)
    ThR = np.einsum('...ab,...b->...a', Th, R); Thv = np.einsum('...ab,...b->...a', Th, v)
    RThR = (ThR * R).sum(-1); vThR = (ThR * v).sum(-1)
    OmRR = np.einsum('...abc,...b,...c->...a', Om, R, R); OmRRR = (OmRR * R).sum(-1)
    OmvR = np.einsum('...abc,...b,...c->...a', Om, v, R)
    sum2 = (OmRR * v).sum(-1) - Om[..., 0, 2, 2] * v[..., 2] * R[..., 0] * R[..., 2]
    e = lambda s: s[..., None]
    out = e(q * r3) * (v - e(3.0 * Rv * ri2) * R)
    out = out + e(3.0 * r5) * (e(mR) * v + e(mv) * R - e(5.0 * ri2 * mR * Rv) * R + e(Rv) * mu)
    out = out + e(r5) * (e(5.0 * ri2) * (e(2.0 * Rv) * ThR + e(2.0 * vThR) * R + e(RThR) * v) - 2.0 * Thv
                         - e(35.0 * Rv * RThR * ri2 * ri2) * R)
    out = out + e(7.0 * r9) * (e(OmRRR) * v - e(9.0 * ri2 * OmRRR * Rv) * R + e(3.0 * sum2) * R + e(3.0 * Rv) * OmRR) \
        - e(6.0 * r7) * OmvR
    return out


def pol_system(parc, solv):
    """Fields FLDS and matrix D of FFFFFF (solpol2.f:1270-1511) for solute + solvent list.
    Returns F [N,3], D [3N,3N], centres [N,3], molecule id per centre, polinv [N,3,3]."""
    PAR = [parc] + list(solv)
    rp = np.concatenate([p['rpol'] for p in PAR]); molp = np.concatenate([[k] * p['npol'] for k, p in enumerate(PAR)])
    rd = np.concatenate([p['rdma'] for p in PAR]); mold = np.concatenate([[k] * p['ndma'] for k, p in enumerate(PAR)])
    q = np.concatenate([p['dmac'] for p in PAR]); mu = np.concatenate([p['dmad'] for p in PAR])
    Th, Om = traceless_full(np.concatenate([quad_full(p['dmaq']) for p in PAR]),
                            np.concatenate([oct_full(p['dmao']) for p in PAR]))
    pol = np.concatenate([p['dpol'] for p in PAR]); polinv = np.linalg.inv(pol)
    N = len(rp)
    R = rp[:, None, :] - rd[None, :, :]
    R = np.where((molp[:, None] != mold[None, :])[..., None], R, 1.0)       # own-molecule pairs are masked below
    Fij = _field_traceless(R, q[None], mu[None], Th[None], Om[None], 5.0)
    Fij = np.where((molp[:, None] != mold[None, :])[..., None], Fij, 0.0)
    F = Fij.sum(1)
    Rp = rp[:, None, :] - rp[None, :, :]
    r2 = (Rp * Rp).sum(-1); r2 = np.where(molp[:, None] != molp[None, :], r2, 1.0)
    T = np.eye(3) / r2[..., None, None] ** 1.5 - 3.0 * np.einsum('ija,ijb->ijab', Rp, Rp) / r2[..., None, None] ** 2.5
    T = np.where((molp[:, None] != molp[None, :])[..., None, None], T, 0.0)
    D = T.transpose(0, 2, 1, 3).reshape(3 * N, 3 * N).copy()
    for i in range(N):
        D[3 * i:3 * i + 3, 3 * i:3 * i + 3] = polinv[i]
    return F, D, rp, molp, polinv


def pol_shift(parc, solv, mode, literal_gemm=False):
    """pol_mea (a.u.), fi_pol[nmodes], dipind, avec -- SFTPLI + AAAAAA (solpol2.f:118-239, 242-1150).
    literal_gemm=True forms D^-1 dD_m D^-1 with two dense products per mode exactly like the
    reference (solpol2.f:1124-1129); otherwise the same vectors come from mat-vecs."""
    w, den = mode_weights(parc, mode)
    g = w / den
    nmodes = len(w)
    if len(solv) == 0:
        return 0.0, np.zeros(nmodes), None, None
    F, D, rp, molp, polinv = pol_system(parc, solv)
    N = len(rp); npc = parc['npol']; ndc = parc['ndma']
    assert ndc == parc['natoms'], 'AAAAAA indexes lvec/dma1 by DMTP site: needs ndma == natoms (solpol2.f:738-741, 873-876)'
    Dinv = np.linalg.inv(D)
    Fv = F.reshape(3 * N)
    dipind = Dinv @ Fv
    mat1 = Dinv + Dinv.T
    # solvent data
    rdB = np.concatenate([p['rdma'] for p in solv]); qB = np.concatenate([p['dmac'] for p in solv])
    muB = np.concatenate([p['dmad'] for p in solv])
    ThB, OmB = traceless_full(np.concatenate([quad_full(p['dmaq']) for p in solv]),
                              np.concatenate([oct_full(p['dmao']) for p in solv]))
    # solute data
    rdA = parc['rdma']; qA = parc['dmac']; muA = parc['dmad']
    ThA, OmA = traceless_full(quad_full(parc['dmaq']), oct_full(parc['dmao']))
    q1 = parc['dmac1']; mu1 = parc['dmad1']
    Th1, Om1 = traceless_full(quad_full(parc['dmaq1']), oct_full(parc['dmao1']))
    rpA = rp[:npc]; rpB = rp[npc:]
    R_AB = rpA[:, None, :] - rdB[None, :, :]                  # solute centres <- solvent sites
    R_pp = rpA[:, None, :] - rpB[None, :, :]                  # solute centres vs solvent centres
    r2pp = (R_pp * R_pp).sum(-1)
    R_BA = rpB[:, None, :] - rdA[None, :, :]                  # solvent centres <- solute sites
    avec = np.zeros(3 * N); fi = np.zeros(nmodes)
    for m in range(nmodes):
        dD = np.zeros((3 * N, 3 * N)); dF = np.zeros((N, 3))
        r1 = parc['lmoc1'][m]                                  # [npc,3]
        for i in range(npc):
            dD[3 * i:3 * i + 3, 3 * i:3 * i + 3] = -polinv[i] @ parc['dpol1'][m, i] @ polinv[i]
        # field derivative at solute centres (they move by r1)
        v = np.broadcast_to(r1[:, None, :], R_AB.shape)
        dF[:npc] = _dfield_traceless(R_AB, v, qB[None], muB[None], ThB[None], OmB[None]).sum(1)
        # D-matrix derivative, solute-solvent blocks: HALF of dT/dr_i . r1 (solpol2.f:699-706), mirrored
        R1R = np.einsum('ija,ia->ij', R_pp, r1)
        blk = R1R[..., None, None] * (np.eye(3) - 5.0 * np.einsum('ija,ijb->ijab', R_pp, R_pp) / r2pp[..., None, None]) \
            + np.einsum('ija,ib->ijab', R_pp, r1) + np.einsum('ia,ijb->ijab', r1, R_pp)
        blk = -blk * (1.5 / r2pp ** 2.5)[..., None, None]
        nB = N - npc
        off = blk.transpose(0, 2, 1, 3).reshape(3 * npc, 3 * nB)
        dD[:3 * npc, 3 * npc:] = off
        dD[3 * npc:, :3 * npc] = blk.transpose(1, 2, 0, 3).reshape(3 * nB, 3 * npc)
        # field derivative at solvent centres: solute sites move by lvec_m and carry dma1_m
        L = parc['lvec'][m]                                    # [natoms(=ndma),3]
        vL = np.broadcast_to(L[None, :, :], R_BA.shape)
        moving = _dfield_traceless(R_BA, vL, qA[None], muA[None], ThA[None], OmA[None])
        deriv = _field_traceless(R_BA, q1[m][None], mu1[m][None], Th1[m][None], Om1[m][None], 7.0)
        dF[npc:] = (deriv - moving).sum(1)
        dFv = dF.reshape(3 * N)
        if literal_gemm:
            mat3 = Dinv @ (dD @ Dinv)
            vec1 = mat3.T @ Fv
        else:
            vec1 = Dinv.T @ (dD.T @ (Dinv.T @ Fv))
        vec2 = mat1.T @ dFv
        mivec = vec1 - vec2
        avec += g[m] * mivec
        fi[m] = 0.5 * Fv @ mivec
    shift = -0.5 * Fv @ avec
    return shift, fi, dipind.reshape(N, 3), avec.reshape(N, 3)


# --------------------------------------------------------------------------- one frame, all terms
def eval_frame(pos, ind, nmol, pars, supl, reord, mode, ccut=1e10, pcut=1e10, ecut=1e10,
               elect=True, ea=True, corr=True, pol=False, disp=False, rep=False, make_basis=None, literal_gemm=False, removable=None):
    """EFP.set + EFP.eval(0) for one frame in central mode (solefp.py:223-277, 738-787, 1011-1629).
    pars: list of parameter dicts (Frag.get()); returns a dict of a.u. terms + rms values."""
    ind = np.asarray(ind, int); nmol = np.asarray(nmol, int)
    off = np.concatenate([[0], np.cumsum(nmol)])
    xyz = [pos[off[i]:off[i + 1]] for i in range(len(ind))]
    icm, ipm, iem = mollst(xyz[0], xyz[1:], ccut, pcut, ecut)

    def sup(i):
        t = ind[i]; p = pars[t]
        ro = reord[t] if reord is not None and reord[t] is not None else np.arange(p['natoms'])
        sl = supl[t] if supl is not None else None
        r, tr, rms = kabsch(p['pos'], reorder_xyz(xyz[i], ro), sl)
        rots[i] = r
        return sup_params(p, r, tr), rms

    rots = {}
    parc, rms_c = sup(0)
    out = dict(ele_mea=0.0, ele_ea=0.0, cor_mea=0.0, cor_ea=0.0, pol_mea=0.0, rep_mea=0.0, dis_mea=0.0, dis_mea_iso=0.0)
    lay_c = [sup(i + 1) for i in np.nonzero(icm)[0]]
    rms_s = [r for _, r in lay_c]
    out['rms_c'] = rms_c
    out['rms_smax'] = max(rms_s) if rms_s else 0.0
    out['rms_ave'] = (rms_c + sum(rms_s)) / (len(rms_s) + 1.0)
    out['n_elect'] = int(icm.sum()); out['n_pol'] = int(ipm.sum()); out['n_rep'] = int(iem.sum())
    if elect:
        m, e, cm, ce, fi = elect_shifts(parc, [p for p, _ in lay_c], mode, ea=ea, corr=corr)
        out.update(ele_mea=m, ele_ea=e, cor_mea=cm, cor_ea=ce, fi_el=fi)
        lay_p = None
        if pol or disp:
            lay_p = [sup(i + 1)[0] for i in np.nonzero(ipm)[0]]
        if pol:
            # remove_clashes (solefp.py:1224-1236, 1333-1404): fragments of removable TYPES leave the many-body
            # solve and each gets a two-body solve with the solute; the shifts add
            ipl = [i + 1 for i in np.nonzero(ipm)[0]]
            keep = [p_ for p_, i in zip(lay_p, ipl) if removable is None or ind[i] not in removable]
            add = [p_ for p_, i in zip(lay_p, ipl) if removable is not None and ind[i] in removable]
            sh, fip, dip, av = pol_shift(parc, keep, mode, literal_gemm=literal_gemm)
            sh_add = 0.0
            for p_ in add:
                s2, f2, _, _ = pol_shift(parc, [p_], mode)
                sh_add += s2; fip = fip + f2
            out.update(pol_mea=sh + sh_add, pol_add_mea=sh_add, fi_pol=fip, dipind=dip, avec=av)
        if disp:
            iso, ani = disp_shifts(parc, lay_p, mode)
            out.update(dis_mea=ani, dis_mea_iso=iso)
    if rep:
        ie = [i + 1 for i in np.nonzero(iem)[0]]
        lay_e = [sup(i)[0] for i in ie]
        if lay_e:
            out['rep_mea'], out['fi_rep'] = rep_shift(parc, lay_e, mode, make_basis, rots[0], [rots[i] for i in ie])
    return out


# --------------------------------------------------------------------------- exchange-repulsion
def vecrot(vec, rot, typ):
    """VECROT / VC1ROT (efprot.f:22-283): rotate AO->LMO coefficient rows [..., nbasis].  p triples as
    vectors; d sextets (xx yy zz xy xz yz) as the symmetric matrix whose off-diagonals ARE the xy/xz/yz
    coefficients (so they are counted twice in R^T M R -- the reference's convention, efprot.f:78-136)."""
    out = np.array(vec, float, copy=True)
    k = 0; n = len(typ)
    while k < n:
        if typ[k] == 1:
            out[..., k:k + 3] = vec[..., k:k + 3] @ rot
            k += 3
        elif typ[k] == 2:
            M = quad_full(vec[..., k:k + 6])
            out[..., k:k + 6] = quad_reduced(np.einsum('ai,...ab,bj->...ij', rot, M, rot))
            k += 6
        else:
            k += 1
    return out


def rep_shift(parc, solv, mode, make_basis=None, rotc=None, rots=None, par0c=None, par0s=None):
    """rep_mea (a.u.) and fi_rep[nmodes]: SHFTEX = CALCIJ + CALCFI (shftex.f:12-438) summed over the solvent
    molecules of the EXREP layer (solefp.py:1497-1560), integrals from oracle/ints_oracle.py.
    parc/solv are superimposed dicts that still carry UNROTATED wave-function arrays; rotc / rots are the
    rotations of their superimpositions (sup_params does not touch vecl/vecl1).  make_basis(atno, pos)
    returns the basis tables; the default is the oracle's own builder (oracle/basis_oracle.py, parsed from the
    reference's dat/slvbas) -- never the product's."""
    from . import ints_oracle as IO
    if make_basis is None:
        from .basis_oracle import make_basis
    w, den = mode_weights(parc, mode)
    nmodes = len(w)
    bA = make_basis(parc['atno'], parc['pos'])
    typA = bA.powers.sum(1)
    CA = vecrot(parc['vecl'], rotc, typA); CA1 = vecrot(parc['vecl1'], rotc, typA)
    FA, FA1 = parc['fock'], parc['fock1']
    riA, riA1, rnA, L = parc['lmoc'], parc['lmoc1'], parc['pos'], parc['lvec']
    ZA = np.asarray(parc['atno'], float)
    atomA = bA.center
    fi = np.zeros(nmodes)
    for pb, rb in zip(solv, rots):
        bB = make_basis(pb['atno'], pb['pos'])
        CB = vecrot(pb['vecl'], rb, bB.powers.sum(1))
        S, T, S1, T1 = IO.one_electron(bA, parc['pos'], bB, pb['pos'])
        FB = pb['fock']; riB = pb['lmoc']; rnB = pb['pos']; ZB = np.asarray(pb['atno'], float)
        Sij = CA @ S @ CB.T; Tij = CA @ T @ CB.T
        dS = np.einsum('mka,kla->mkl', L[:, atomA, :], S1); dT = np.einsum('mka,kla->mkl', L[:, atomA, :], T1)
        Sm = np.einsum('ik,mkl,jl->mij', CA, dS, CB) + np.einsum('mik,kl,jl->mij', CA1, S, CB)
        Tm = np.einsum('ik,mkl,jl->mij', CA, dT, CB) + np.einsum('mik,kl,jl->mij', CA1, T, CB)
        Rij = riA[:, None, :] - riB[None, :, :]; rij = np.linalg.norm(Rij, axis=-1)         # [i,j]
        rm_ij = np.einsum('ija,mia->mij', Rij, riA1) / rij
        sl = np.log(np.abs(Sij)); aux = 2.0 * np.sqrt(-2.0 * sl / np.pi)
        sdr = Sij / rij
        TS1 = 4.0 * sdr * Sm * (np.sqrt(-1.0 / (2.0 * np.pi * sl)) - aux - 1.0)
        TS2 = 2.0 * sdr * sdr * rm_ij * (aux + 1.0)
        TS3 = 4.0 * Sm * Tij; TS4 = 4.0 * Tm * Sij
        # TA1: sum over k in A (distances r_kj)
        TA1 = (-2.0 * Sm * (FA @ Sij)
               + 8.0 * Sij * Sm * (1.0 / rij).sum(0)[None, :]
               - 2.0 * Sij * (np.einsum('mik,kj->mij', FA1, Sij) + np.einsum('ik,mkj->mij', FA, Sm))
               - 4.0 * Sij ** 2 * (rm_ij / rij ** 2).sum(1)[:, None, :])
        # TA2: sum over l in B (distances r_il)
        TA2 = (-2.0 * Sm * (Sij @ FB.T)
               + 8.0 * Sij * Sm * (1.0 / rij).sum(1)[:, None]
               - 2.0 * Sij * np.einsum('jl,mil->mij', FB, Sm)
               - 4.0 * Sij ** 2 * (rm_ij / rij ** 2).sum(2)[:, :, None])
        Rxj = rnA[:, None, :] - riB[None, :, :]; rxj = np.linalg.norm(Rxj, axis=-1)           # [x,j]
        rm_xj = np.einsum('xja,mxa->mxj', Rxj, L) / rxj
        TA3 = 2.0 * Sij ** 2 * (ZA[:, None] * rm_xj / rxj ** 2).sum(1)[:, None, :] - 4.0 * Sij * Sm * (ZA[:, None] / rxj).sum(0)[None, :]
        Riy = riA[:, None, :] - rnB[None, :, :]; riy = np.linalg.norm(Riy, axis=-1)           # [i,y]
        rm_iy = np.einsum('iya,mia->miy', Riy, riA1) / riy
        TA4 = 2.0 * Sij ** 2 * (ZB[None, :] * rm_iy / riy ** 2).sum(2)[:, :, None] - 4.0 * Sij * Sm * (ZB[None, :] / riy).sum(1)[:, None]
        fi += (TS1 + TS2 + TS3 + TS4 + TA1 + TA2 + TA3 + TA4).sum((1, 2))
    return -(w * fi).sum() / den, fi


def eval_dma_frame(pos_c, parc0, supl_c, mode, env, ea=True, corr=True):
    """EFP.eval_dma (solefp.py:798-903): SolCAMM + corrections of the superimposed solute against an environment
    given as point charges env[n,4] = x,y,z,q; no cut-off, no superimposition of the environment."""
    r, tr, rms = kabsch(parc0['pos'], pos_c, supl_c)
    parc = sup_params(parc0, r, tr)
    n = len(env)
    pe = dict(rdma=env[:, :3], dmac=env[:, 3], dmad=np.zeros((n, 3)), dmaq=np.zeros((n, 6)), dmao=np.zeros((n, 10)))
    m, e, cm, ce, fi = elect_shifts(parc, [pe], mode, ea=ea, corr=corr)
    return dict(ele_env_mea=m, ele_env_ea=e, cor_env_mea=cm, cor_env_ea=ce)


# --------------------------------------------------------------------------- energy ("global") mode
def global_energies(pos, ind, nmol, pars, supl, reord, elect=True, pol=True):
    """EFP._eval_mode_global (solefp.py:905-1009): every molecule superimposed, no cut-offs; e_coul over all molecule pairs,
    e_pol = SOLPOL (solpol2.f:1153-1267) = -1/2 F . D^-1 F.  e_coul comes from libbbg.qm.clemtp.edmtpa, which is
    un-vendored and has no known-answer data in the tree: restated with the multipole series through R^-4 of the shift
    path (PARITY UNPINNED for this term).  Returns dict(ele, pol) in Hartree."""
    ind = np.asarray(ind, int); nmol = np.asarray(nmol, int)
    off = np.concatenate([[0], np.cumsum(nmol)])
    PAR = []
    for i in range(len(ind)):
        t = ind[i]; p = pars[t]
        ro = reord[t] if reord is not None and reord[t] is not None else np.arange(p['natoms'])
        r, tr, _ = kabsch(p['pos'], reorder_xyz(pos[off[i]:off[i + 1]], ro), supl[t] if supl is not None else None)
        PAR.append(sup_params(p, r, tr))
    out = dict(ele=0.0, pol=0.0)
    if elect:
        mom = [_lab_moments(p) for p in PAR]
        e = 0.0
        for a in range(len(PAR)):
            for b in range(a + 1, len(PAR)):
                qa, ma, Ta, Oa = mom[a]; qb, mb, Tb, Ob = mom[b]
                e += float(emtp_r4(PAR[a]['rdma'], qa, ma, Ta, Oa, PAR[b]['rdma'], qb, mb, Tb, Ob))
        out['ele'] = e
        if pol:
            F, D = pol_system(PAR[0], PAR[1:])[:2]
            Fv = F.ravel()
            out['pol'] = -0.5 * Fv @ np.linalg.solve(D, Fv)
    out['PAR'] = PAR
    return out
