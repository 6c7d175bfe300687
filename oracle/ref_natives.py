"""ctypes bindings to oracle/_ref/libslvref.so (the reference's Fortran kernels transpiled by
oracle/f77c.py) with the same call signatures f2py gives them in the reference
(solvshift/solefp.py:763, 1319-1322, 1542-1545; solvshift/slvpar.py:497-514, 689-733).
TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def available():
    return os.path.isfile(os.path.join(_HERE, '_ref', 'libslvref.so'))


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(os.path.join(_HERE, '_ref', 'libslvref.so'))
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, np.float64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i(v):
    return ctypes.byref(ctypes.c_int(int(v)))


def _f(v):
    return ctypes.byref(ctypes.c_double(float(v)))


def mollst(rc, rm, nat, ccut, pcut, ecut):
    """solpol2.mollst(rc, rm, ic, ip, ie, icm, ipm, iem, nat, ccut, pcut, ecut) -> masks."""
    rc = np.asfortranarray(rc, np.float64); rm = _d(rm).ravel()
    nat = np.ascontiguousarray(nat, np.int32)
    ncoord = rm.size // 3; nmols = nat.size
    m = [np.zeros(ncoord, np.int32) for _ in range(3)] + [np.zeros(nmols, np.int32) for _ in range(3)]
    lib().mollst_(_p(rc), _p(rm), *[_p(x) for x in m], _p(nat), _f(ccut), _f(pcut), _f(ecut),
                  _i(nmols), _i(ncoord), _i(rc.shape[0]))
    return [x.astype(bool) for x in m]


def tracls(qad, octu):
    q = np.asfortranarray(qad, np.float64).copy(order='F'); o = np.asfortranarray(octu, np.float64).copy(order='F')
    lib().tracls_(_p(q), _p(o), _i(q.shape[0]))
    return np.ascontiguousarray(q), np.ascontiguousarray(o)


def tracl1(qad1, oct1, nmodes, ndma):
    q = _d(qad1).ravel().copy(); o = _d(oct1).ravel().copy()
    lib().tracl1_(_p(q), _p(o), _i(nmodes), _i(ndma))
    return q, o


def rotdma(dip, qad, octu, rot):
    d = np.asfortranarray(dip, np.float64).copy(order='F'); q = np.asfortranarray(qad, np.float64).copy(order='F')
    o = np.asfortranarray(octu, np.float64).copy(order='F'); r = np.asfortranarray(rot, np.float64)
    lib().rotdma_(_p(d), _p(q), _p(o), _p(r), _i(d.shape[0]))
    return np.ascontiguousarray(d), np.ascontiguousarray(q), np.ascontiguousarray(o)


def rotdm1(dip1, qad1, oct1, rot, nmodes, ndma):
    d = _d(dip1).ravel().copy(); q = _d(qad1).ravel().copy(); o = _d(oct1).ravel().copy()
    r = np.asfortranarray(rot, np.float64)
    lib().rotdm1_(_p(d), _p(q), _p(o), _p(r), _i(nmodes), _i(ndma))
    return d, q, o


def vecrot(vec, rot, typs):
    v = np.asfortranarray(vec, np.float64).copy(order='F'); r = np.asfortranarray(rot, np.float64)
    t = np.ascontiguousarray(typs, np.int32)
    lib().vecrot_(_i(v.shape[0]), _i(v.shape[1]), _p(v), _p(r), _p(t))
    return np.ascontiguousarray(v)


def vc1rot(vec1, rot, typs):
    v = np.asfortranarray(vec1, np.float64).copy(order='F'); r = np.asfortranarray(rot, np.float64)
    t = np.ascontiguousarray(typs, np.int32)
    lib().vc1rot_(_i(v.shape[0]), _i(v.shape[1]), _i(v.shape[2]), _p(v), _p(r), _p(t))
    return np.ascontiguousarray(v)


def sftpli(rdma, chg, dip, qad, octu, chg1, dip1, qad1, oct1, rpol, polinv, redmss, freq, gijj,
           rpol1, pol1, lvec, ndma, npol, mode, ndmac, npolc):
    """solpol2.sftpli as called at solefp.py:1319-1322 (scratch arrays allocated here).
    Returns epol, shift, fi, avec, dipind, flds."""
    a = [_d(x).ravel() for x in (rdma, chg, dip, qad, octu, chg1, dip1, qad1, oct1, rpol, polinv)]
    redmss, freq, gijj, rpol1, pol1, lvec = [_d(x).ravel() for x in (redmss, freq, gijj, rpol1, pol1, lvec)]
    ndma = np.ascontiguousarray(ndma, np.int32); npol = np.ascontiguousarray(npol, np.int32)
    nmols = ndma.size; ndim = 3 * int(npol.sum()); ndmas = int(ndma.sum()); nmodes = redmss.size
    dmat = np.zeros((ndim, ndim), order='F'); dimat = np.zeros((ndim, ndim), order='F'); mat1 = np.zeros((ndim, ndim), order='F')
    flds = np.zeros(ndim); vec1 = np.zeros(ndim); vec2 = np.zeros(ndim); fivec = np.zeros(ndim); mivec = np.zeros(ndim)
    fi = np.zeros(nmodes); avec = np.zeros(ndim); dipind = np.zeros(ndim)
    epol = ctypes.c_double(0.0); shift = ctypes.c_double(0.0)
    lib().sftpli_(*[_p(x) for x in a], _p(mivec), _p(dmat), _p(flds), _p(dimat), _p(fivec), _p(vec1), _p(vec2),
                  _p(mat1), _p(redmss), _p(freq), _p(gijj), _p(rpol1), _p(pol1), _p(lvec),
                  _i(nmols), _p(ndma), _p(npol), _i(ndim), _i(ndmas), _i(mode), _i(nmodes), _i(ndmac), _i(npolc),
                  _i(a[2].size), _i(a[3].size), _i(a[4].size), _i(a[9].size), _i(a[10].size), _i(0),
                  ctypes.byref(epol), ctypes.byref(shift), _p(fi), _p(avec), _p(dipind))
    return epol.value, shift.value, fi, avec, dipind, flds


def shftex(redmss, freq, gijj, lvec, ria, rib, rna, rnb, ria1, cika, cikb, cika1, skm, tkm, sk1m, tk1m,
           za, zb, nbsb, mlist, faij, fbij, faij1, mode):
    """shftex.shftex as called at solefp.py:1542-1545.  Returns shftma, shftea, fi, sij, tij, smij, tmij."""
    redmss, freq, gijj, lvec, ria, rib, rna, rnb, ria1, cika, cikb, cika1, skm, tkm, sk1m, tk1m, faij, fbij, faij1 = \
        [_d(x).ravel() for x in (redmss, freq, gijj, lvec, ria, rib, rna, rnb, ria1, cika, cikb, cika1, skm, tkm,
                                  sk1m, tk1m, faij, fbij, faij1)]
    za = _d(za).ravel(); zb = _d(zb).ravel(); mlist = np.ascontiguousarray(mlist, np.int32)
    nmodes = redmss.size; nmosa = ria.size // 3; nmosb = rib.size // 3; nata = rna.size // 3; natb = rnb.size // 3
    nbsa = cika.size // nmosa
    sij = np.zeros(nmosa * nmosb); tij = np.zeros(nmosa * nmosb)
    smij = np.zeros(nmodes * nmosa * nmosb); tmij = np.zeros(nmodes * nmosa * nmosb); fi = np.zeros(nmodes)
    ma = ctypes.c_double(0.0); ea = ctypes.c_double(0.0)
    lib().shftex_(_p(redmss), _p(freq), _p(gijj), _p(lvec), _p(ria), _p(rib), _p(rna), _p(rnb), _p(ria1),
                  _p(cika), _p(cikb), _p(cika1), _p(skm), _p(tkm), _p(sk1m), _p(tk1m), _p(za), _p(zb),
                  _i(nbsa), _i(nbsb), _i(nmosa), _i(nmosb), _i(nata), _i(natb), _i(nmodes), _p(mlist),
                  _p(faij), _p(fbij), _p(faij1), _i(mode), _p(sij), _p(tij), _p(smij), _p(tmij), _p(fi),
                  ctypes.byref(ma), ctypes.byref(ea))
    return ma.value, ea.value, fi, sij, tij, smij, tmij


def gtcf(a, b, k):
    """genres.gtcf(a, b, k) -> tcf[k]  (util/lib/genres.f:1-23)."""
    a = _d(a).ravel(); b = _d(b).ravel()
    tcf = np.zeros(int(k))
    lib().gtcf_(_p(a), _p(b), _i(a.size), _i(k), _p(tcf))
    return tcf


def exrep(ria, rib, rna, rnb, faij, fbij, cika, cikb, skm, tkm, za, zb):
    """exrep.exrep(ria, rib, rna, rnb, faij, fbij, cika, cikb, skm, tkm, za, zb) -> eint (Hartree), the EFP2
    exchange-repulsion energy of one fragment pair (solvshift/src/exrep.f:140-168; call sites solefp.py:1549-1558
    (commented), 1781-1805).  2-D arguments are Fortran arrays: column-major copies are made here as f2py does."""
    F = lambda x: np.asfortranarray(x, np.float64)
    ria, rib, rna, rnb, faij, fbij, cika, cikb, skm, tkm = [F(x) for x in (ria, rib, rna, rnb, faij, fbij, cika, cikb, skm, tkm)]
    za = _d(za).ravel(); zb = _d(zb).ravel()
    assert ria.shape[0] <= 40 and rib.shape[0] <= 40, 'COMMON /INTIJ/ holds 40 x 40 LMO pairs (exrep.f:150)'
    e = ctypes.c_double(0.0)
    lib().exrep_(_p(ria), _p(rib), _p(rna), _p(rnb), _p(faij), _p(fbij), _p(cika), _p(cikb), _p(skm), _p(tkm), _p(za), _p(zb),
                 _i(cika.shape[1]), _i(cikb.shape[1]), _i(ria.shape[0]), _i(rib.shape[0]), _i(rna.shape[0]), _i(rnb.shape[0]),
                 ctypes.byref(e))
    return e.value


def solpol(rdma, chg, dip, qad, octu, rpol, polinv, ndma, npol):
    """solpol2.solpol as called at solefp.py:969 -> epol (Hartree), the EFP2 polarisation energy -1/2 F . D^-1 F
    (solvshift/src/solpol2.f:1153-1267); scratch arrays allocated here."""
    a = [_d(x).ravel() for x in (rdma, chg, dip, qad, octu, rpol, polinv)]
    ndma = np.ascontiguousarray(ndma, np.int32); npol = np.ascontiguousarray(npol, np.int32)
    nmols = ndma.size; ndim = 3 * int(npol.sum()); ndmas = int(ndma.sum())
    dmat = np.zeros((ndim, ndim), order='F'); flds = np.zeros(ndim); dipind = np.zeros(ndim)
    epol = ctypes.c_double(0.0)
    lib().solpol_(*[_p(x) for x in a], ctypes.byref(epol), _p(dmat), _p(flds), _p(dipind),
                  _i(nmols), _p(ndma), _p(npol), _i(ndim), _i(ndmas),
                  _i(a[2].size), _i(a[3].size), _i(a[4].size), _i(a[5].size), _i(a[6].size), _i(0))
    return epol.value, flds, dipind
