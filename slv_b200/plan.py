"""Static plan compilation (host, once per trajectory layout).

Turns the arguments of ``EFP.set`` (ind, nmol, bsm, supl, reord -- solvshift/solefp.py:223-277, produced
by ``MDInput.get``, solvshift/md.py:82-92) into the flat device tables of include/slvgpu.h and pre-contracts
every normal-mode sum of the reference: with g_m = g_mjj / (M_m w_m^2) / (2 M_j w_j),

  SolCAMM (mechanical)  s_mea = -sum_m g_m dma1_m              (slvpar.py:1049-1056)
  SolCAMM (electronic)  s_ea  = dma2_j / (2 M_j w_j)           (slvpar.py:1058-1064)
  corrections           c_mea = -q_x sum_m g_m L^m_x ,  c_ea = 2 q1^j_x L^j_x / (2 M_j w_j)
  dispersion            sum_m g_m dpoli1_m, sum_m g_m lmoc1_m
  polarisation          -alpha^-1 (sum_m g_m dpol1_m) alpha^-1, sum_m g_m lmoc1_m, sum_m g_m lvec_m, sum_m g_m dma1_m
  exchange-repulsion    sum_m g_m fock1_m, sum_m g_m vecl1_m, sum_m g_m lmoc1_m, sum_m g_m lvec_m

so that one pass over the solute-solvent pairs replaces the reference's (nmodes + 1) passes
(libbbg sdmtpm / solpol2.f AAAAAA / shftex.f CALCIJ+CALCFI all loop over the modes).  Pure data
re-indexing and parameter algebra on frame-invariant inputs; nothing per-frame happens here.
"""
import ctypes

import numpy

from . import _lib
from .slvpar import traceless


class _Tab(object):
    def __init__(self, dtype):
        self.chunks = []; self.n = 0; self.dtype = dtype

    def add(self, a):
        a = numpy.ascontiguousarray(a, self.dtype).ravel()
        off = self.n
        self.chunks.append(a); self.n += a.size
        return off

    def pad(self, k):
        """Make the next offset a multiple of k elements."""
        r = (-self.n) % k
        if r:
            self.add(numpy.zeros(r, self.dtype))

    def get(self):
        if not self.chunks:
            return numpy.zeros(1, self.dtype)
        return numpy.concatenate(self.chunks)


def mom20(q, d, Q, O):
    """[n,20] rows: charge, dipole, traceless quadrupole (6), traceless octupole (10)."""
    qt, ot = traceless(Q, O)
    return numpy.concatenate([numpy.asarray(q, numpy.float64)[..., None], d, qt, ot], axis=-1)


def mode_weights(par, mode):
    gijj = par['gijk'][:, mode - 1, mode - 1]
    den = 2.0 * par['redmass'][mode - 1] * par['freq'][mode - 1]
    return gijj / (par['redmass'] * par['freq'] ** 2) / den, den


# (L_bra, L_ket, bra component or -1 = all) in the order of SLV_SUBSHELL_CLASSES (slv_b200/csrc/slv_ints.cuh)
SUBSHELL_CLASSES = ([(la, 0, -1) for la in range(3)]
                    + [(la, lb, ia) for lb in (1, 2) for la in range(3) for ia in range((1, 3, 6)[la])])


def subshell_work_list(sA, sB):
    """Work list of the exchange-repulsion integral kernels for solute sub-shells sA against fragment sub-shells sB
    (rows: atom, first function, L, primitives): class offsets [len(SUBSHELL_CLASSES) + 1] and the items
    a * len(sB) + b, class by class, heaviest primitive-pair count first.  Every AO pair of the two fragments is
    produced by exactly one (class, item)."""
    cost = (sA[:, 3][:, None] * sB[:, 3][None, :])
    items, offs = [], [0]
    for LA, LB, IA in SUBSHELL_CLASSES:
        a, b = numpy.nonzero((sA[:, 2] == LA)[:, None] & (sB[:, 2] == LB)[None, :])
        o = numpy.argsort(-cost[a, b], kind='stable')
        items.append((a[o] * len(sB) + b[o]).astype(numpy.int32))
        offs.append(offs[-1] + len(o))
    return numpy.array(offs, numpy.int32), numpy.concatenate(items).astype(numpy.int32)


PP_REC = 6      # doubles per primitive-pair record (slv_ints.cuh PP_REC)


def primitive_pair_records(bA, bB, items):
    """Tables behind the integral kernels' work items (``items`` from subshell_work_list).

    Returns ``irec`` int32 [nitems, 8] = (3 * atom_A, first function A, 3 * atom_B, first function B, first record,
    records, 0, 0) and ``ptab`` float64 [nrec, PP_REC]: for every (solute sub-shell, fragment sub-shell) pair the
    records (thr, mu, pref0, 1/(a+b), a, b) of its primitive pairs -- thr = 46 (a+b)/(a b) is the squared distance beyond
    which exp(-mu AB^2) < exp(-46), mu = a b/(a+b), pref0 = c_a c_b (pi/(a+b))^1.5 with the coefficients of the first
    component of each sub-shell -- sorted by thr, largest first, so the surviving pairs at any distance are a prefix."""
    sA, sB = bA.subshells, bB.subshells
    nB = len(sB)
    first = numpy.zeros((len(sA), nB), numpy.int64)
    recs, n = [], 0
    for ia, (_, fa, _, npa) in enumerate(sA):
        ea, ca = bA.exps[fa, :npa], bA.coefs[fa, :npa]
        for ib, (_, fb, _, npb) in enumerate(sB):
            eb, cb = bB.exps[fb, :npb], bB.coefs[fb, :npb]
            a, b = numpy.repeat(ea, npb), numpy.tile(eb, npa)
            ip = 1.0 / (a + b)
            r = numpy.stack([46.0 * (a + b) / (a * b), a * b * ip, numpy.repeat(ca, npb) * numpy.tile(cb, npa) * (numpy.pi * ip) ** 1.5,
                             ip, a, b], 1)
            recs.append(r[numpy.argsort(-r[:, 0], kind='stable')])
            first[ia, ib] = n; n += len(r)
    ia, ib = items // nB, items % nB
    irec = numpy.zeros((len(items), 8), numpy.int32)
    irec[:, 0] = 3 * sA[ia, 0]; irec[:, 1] = sA[ia, 1]; irec[:, 2] = 3 * sB[ib, 0]; irec[:, 3] = sB[ib, 1]
    irec[:, 4] = first[ia, ib]; irec[:, 5] = sA[ia, 3] * sB[ib, 3]
    return irec, numpy.concatenate(recs).reshape(-1, PP_REC)


class Plan(object):
    """Device-resident evaluation plan (owns the slvgpu_plan handle)."""

    def __init__(self, ind, nmol, bsm, supl, reord, mode, flags, ccut, pcut, ecut,
                 max_chunk=4096, pol_maxit=1000, pol_tol=1e-10, pol_cent_cap=0, want_detail=False,
                 g_override=None, removable=None):
        ind = numpy.asarray(ind, int); nmol = numpy.asarray(nmol, int)
        assert len(ind) == len(nmol) and len(ind) >= 1
        self.ninst = len(ind)
        self.natoms_total = int(nmol.sum())
        self.flags = 0
        for k, v in flags.items():
            if v:
                self.flags |= _lib.FLAG[k]
        dt, it = _Tab(numpy.float64), _Tab(numpy.int32)
        ntypes = len(bsm)
        meta = numpy.full((ntypes, _lib.NMETA), -1, numpy.int32)
        M = _lib.M
        pars = [b.get() for b in bsm]
        need_rep = bool(self.flags & _lib.FLAG['rep'])
        self._bfs = {}
        for t, par in enumerate(pars):
            nat = int(par['natoms'])
            sl = supl[t] if supl is not None and supl[t] is not None else numpy.arange(nat)
            ro = reord[t] if reord is not None and reord[t] is not None else numpy.arange(nat)
            sl = numpy.asarray(sl, int); ro = numpy.asarray(ro, int)
            assert len(ro) == nat, 'reorder list of fragment %d must have natoms entries' % t
            m = meta[t]
            m[M['NATOMS']] = nat; m[M['NDMA']] = par.get('ndma', 0) or 0; m[M['NPOL']] = par.get('npol', 0) or 0
            m[M['NMOS']] = par.get('nmos', 0) or 0; m[M['NBASIS']] = par.get('nbasis', 0) or 0
            m[M['NSUPL']] = len(sl)
            # 2: part of the non-EFP electrostatic environment (EFP.eval_dma): point multipoles riding on frame atoms
            m[M['REMOVABLE']] = 2 if par.get('env') else (1 if (removable is not None and t in removable) else 0)
            m[M['IO_SUPL']] = it.add(sl); m[M['IO_REORD']] = it.add(ro)
            m[M['O_POS']] = dt.add(par['pos'])
            if 'rdma' in par:
                m[M['O_RDMA']] = dt.add(par['rdma'])
                m[M['O_MOM']] = dt.add(mom20(par['dmac'], par['dmad'], par['dmaq'], par['dmao']))
            if 'rpol' in par and 'dpol' in par:
                m[M['O_RPOL']] = dt.add(par['rpol'])
                m[M['O_POL']] = dt.add(par['dpol'])
                m[M['O_POLINV']] = dt.add(numpy.linalg.inv(par['dpol']))
            else:
                m[M['NPOL']] = 0
            if 'dpoli' in par:
                m[M['O_DPOLI']] = dt.add(par['dpoli'])
            if need_rep and 'vecl' in par:
                from .basis import BasisSet
                bfs = BasisSet(par['atno'], par['pos'])
                assert len(bfs) == par['nbasis'], 'basis size mismatch for fragment type %d' % t
                m[M['O_LMOC']] = dt.add(par['lmoc']); m[M['O_FOCK']] = dt.add(par['fock'])
                m[M['O_VECL']] = dt.add(par['vecl']); m[M['O_ZNUC']] = dt.add(numpy.asarray(par['atno'], float))
                m[M['IO_BFC']] = it.add(bfs.center); m[M['IO_BFL']] = it.add(bfs.powers); m[M['IO_BFNP']] = it.add(bfs.nprim)
                m[M['O_BFEXP']] = dt.add(bfs.exps); m[M['O_BFCOEF']] = dt.add(bfs.coefs)
                m[M['NSUBSH']] = len(bfs.subshells); m[M['IO_SUBSH']] = it.add(bfs.subshells)
                # rotation blocks of the AO->LMO coefficient rows (VECROT): (first function, angular momentum) per sub-shell
                it.pad(2)
                m[M['NBFBLK']] = len(bfs.subshells); m[M['IO_BFBLK']] = it.add(bfs.subshells[:, [1, 2]])
                self._bfs[t] = bfs
        # ---- every table a requested term reads must exist: the reference fails with KeyError on the missing parameter
        #      (solefp.py:1429 par['dpoli'], 1305 par['dpol'], 1531 par['fock']); here a missing table would be a bad address
        fl = self.flags
        for i in sorted(set(int(t) for t in ind[1:])):
            par, nm = pars[i], pars[i].get('name')
            if par.get('env'):
                continue
            if fl & _lib.FLAG['disp'] and meta[i][M['NPOL']] > 0 and 'dpoli' not in par:
                raise KeyError("'dpoli': dispersion needs the polarisabilities at imaginary frequencies of fragment %r" % nm)
            if fl & _lib.FLAG['rep'] and not ('vecl' in par and 'fock' in par and 'lmoc' in par):
                raise KeyError("'vecl'/'fock'/'lmoc': exchange repulsion needs the wave function of fragment %r" % nm)
            if fl & (_lib.FLAG['elect'] | _lib.FLAG['pol']) and 'rdma' not in par:
                raise KeyError("'rdma': fragment %r has no distributed multipoles" % nm)
        if len(ind) > 1 and not (fl & _lib.FLAG['glob']):
            pc0 = pars[int(ind[0])]
            for flag, keys in (('disp', ('dpoli', 'dpoli1', 'lmoc1')), ('pol', ('dpol', 'dpol1', 'lmoc1', 'dmac1', 'lvec')),
                               ('rep', ('vecl', 'vecl1', 'fock', 'fock1', 'lmoc', 'lmoc1', 'lvec')), ('elect', ('rdma', 'dmac1', 'gijk'))):
                for k in keys:
                    if fl & _lib.FLAG[flag] and k not in pc0:
                        raise KeyError("%r: the %s shift needs this section of the solute's parameter file (%r)" % (k, flag, pc0.get('name')))
        # ---- exchange-repulsion work list: for every fragment type B, the (solute sub-shell, B sub-shell) pairs
        #      grouped by the 23 (L_bra, L_ket, bra component) specialisations of slv_ints.cuh (ket s: all bra
        #      components in one item; ket p, d: one item per bra component) and sorted by primitive-pair count
        #      inside a class: the integral kernel is launched once per class over all (frame, fragment) pairs of a
        #      batch, so every warp runs one specialisation with near-equal loop lengths
        if need_rep and int(ind[0]) in self._bfs:
            sA = self._bfs[int(ind[0])].subshells
            for t, bB in self._bfs.items():
                offs, items = subshell_work_list(sA, bB.subshells)
                meta[t][M['IO_SHPAIR']] = it.add(numpy.concatenate([offs, items]))
                irec, ptab = primitive_pair_records(self._bfs[int(ind[0])], bB, items)
                it.pad(4); dt.pad(2)                       # 16-byte aligned records
                meta[t][M['IO_SHITEM']] = it.add(irec); meta[t][M['O_PPTAB']] = dt.add(ptab)
        # ---- solute: mode-contracted tables
        sol = numpy.full(_lib.NSOL, -1, numpy.int32)
        S = _lib.S
        pc = pars[int(ind[0])]
        self.mode = mode
        if mode is not None and 'gijk' in pc:
            g, den = mode_weights(pc, mode)
            if g_override is not None:          # one-hot mode weights: used by EFP.get_forces (fi_m = dE/dQ_m)
                g = numpy.asarray(g_override, numpy.float64)
                assert g.shape == (int(pc['nmodes']),)
            j = mode - 1
            dma1 = mom20(pc['dmac1'], pc['dmad1'], pc['dmaq1'], pc['dmao1'])          # [nmodes, ndma, 20]
            gd = numpy.tensordot(g, dma1, (0, 0))
            sol[S['SMEA']] = dt.add(-gd)
            sol[S['DMA1']] = dt.add(gd)
            if 'mode' in pc and (mode - 1) in list(pc['mode']):
                m2 = list(pc['mode']).index(mode - 1)
                sol[S['SEA']] = dt.add(mom20(pc['dmac2'][m2], pc['dmad2'][m2], pc['dmaq2'][m2], pc['dmao2'][m2]) / den)
            else:
                assert not (self.flags & _lib.FLAG['ea']) or not (self.flags & _lib.FLAG['elect']), \
                    " The second derivatives wrt mode %d are not evaluated for %s!" % (mode, pc.get('name'))
                sol[S['SEA']] = dt.add(numpy.zeros_like(gd))
            gL = numpy.tensordot(g, pc['lvec'], (0, 0))                                 # [natoms,3]
            nd = int(pc['ndma']); nat = int(pc['natoms']); n = min(nd, nat)
            cm = numpy.zeros((nd, 3)); ce = numpy.zeros((nd, 3))
            cm[:n] = -pc['dmac'][:n, None] * gL[:n]
            ce[:n] = 2.0 * pc['dmac1'][j][:n, None] * pc['lvec'][j][:n] / den
            sol[S['CMEA']] = dt.add(cm); sol[S['CEA']] = dt.add(ce)
            sol[S['LVEC']] = dt.add(gL)
            if 'lmoc1' in pc:
                sol[S['LMOC1']] = dt.add(numpy.tensordot(g, pc['lmoc1'], (0, 0)))
            if 'dpoli1' in pc:
                sol[S['DPOLI1']] = dt.add(numpy.tensordot(g, pc['dpoli1'], (0, 0)))
            if 'dpol1' in pc:
                ainv = numpy.linalg.inv(pc['dpol'])
                a1 = numpy.tensordot(g, pc['dpol1'], (0, 0))
                sol[S['DPOL1']] = dt.add(-numpy.einsum('iab,ibc,icd->iad', ainv, a1, ainv))
            if need_rep:
                sol[S['FOCK1']] = dt.add(numpy.tensordot(g, pc['fock1'], (0, 0)))
                sol[S['VECL1']] = dt.add(numpy.tensordot(g, pc['vecl1'], (0, 0)))
        self._keep = dict(meta=meta, sol=sol, itab=it.get(), dtab=dt.get(),
                          inst_type=numpy.ascontiguousarray(ind, numpy.int32),
                          inst_atom0=numpy.ascontiguousarray(numpy.concatenate([[0], numpy.cumsum(nmol)]), numpy.int32))
        k = self._keep
        d = _lib.PlanDesc()
        d.ntypes = ntypes; d.ninst = self.ninst; d.natoms_total = self.natoms_total; d.flags = self.flags
        d.ccut = float(ccut); d.pcut = float(pcut); d.ecut = float(ecut)
        d.type_meta = k['meta'].ctypes.data; d.sol_meta = k['sol'].ctypes.data
        d.itab = k['itab'].ctypes.data; d.n_itab = k['itab'].size
        d.dtab = k['dtab'].ctypes.data; d.n_dtab = k['dtab'].size
        d.inst_type = k['inst_type'].ctypes.data; d.inst_atom0 = k['inst_atom0'].ctypes.data
        d.max_chunk = int(max_chunk); d.pol_maxit = int(pol_maxit); d.pol_tol = float(pol_tol)
        d.pol_cent_cap = int(pol_cent_cap); d.want_detail = int(bool(want_detail))
        h = ctypes.c_void_p()
        _lib.check(_lib.lib.slvgpu_plan_create(ctypes.byref(d), ctypes.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, '_h', None)
        if h and _lib is not None and getattr(_lib, 'lib', None) is not None:
            _lib.lib.slvgpu_plan_destroy(h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def eval_device(self, coords, out=None, status=None, stream=None):
        """coords: CUDA torch tensor [nframes, natoms_total, 3] float64.  Returns (out[nframes,16], status)."""
        import torch
        assert coords.is_cuda and coords.dtype == torch.float64 and coords.is_contiguous()
        nf = coords.shape[0]
        assert coords.shape[1] == self.natoms_total and coords.shape[2] == 3
        if out is None:
            out = torch.empty((nf, _lib.NOUT), dtype=torch.float64, device=coords.device)
        if status is None:
            status = torch.empty((nf,), dtype=torch.int32, device=coords.device)
        st = torch.cuda.current_stream(coords.device).cuda_stream if stream is None else stream
        _lib.check(_lib.lib.slvgpu_eval_frames(self._h, coords.data_ptr(), nf, out.data_ptr(), status.data_ptr(), st))
        return out, status

    def eval_host(self, coords, out=None, status=None, scale=None, idx=None):
        """coords: numpy or pinned CPU torch tensor (host buffers; the H2D / D2H copies happen inside the call).
        float64 [nframes, natoms_total, 3] in Bohr, or float32 [nframes, natoms_src, 3] in the trajectory file's units:
        then `scale` converts to Bohr and `idx` (natoms_total record atoms, MDInput's idx) gathers the plan's atoms,
        both on the device."""
        is_t = hasattr(coords, 'data_ptr')
        f32 = (str(coords.dtype).endswith('float32'))
        if f32:
            if not is_t:
                coords = numpy.ascontiguousarray(coords, numpy.float32)
            else:
                assert coords.is_contiguous()
            nf, nsrc = int(coords.shape[0]), int(coords.shape[1])
            cptr = coords.data_ptr() if is_t else coords.ctypes.data
            if idx is not None:
                idx = numpy.ascontiguousarray(idx, numpy.int32)
                assert idx.shape == (self.natoms_total,), 'idx must name one trajectory atom per plan atom'
            else:
                assert nsrc == self.natoms_total
            if out is None:
                out = numpy.empty((nf, _lib.NOUT), numpy.float64)
            if status is None:
                status = numpy.empty((nf,), numpy.int32)
            optr = out.data_ptr() if hasattr(out, 'data_ptr') else out.ctypes.data
            sptr = status.data_ptr() if hasattr(status, 'data_ptr') else status.ctypes.data
            _lib.check(_lib.lib.slvgpu_eval_frames_host_f32(self._h, cptr, nf, nsrc, None if idx is None else idx.ctypes.data,
                                                           1.0 if scale is None else float(scale), optr, sptr))
            return out, status
        assert scale is None and idx is None, 'scale / idx belong to the float32 (trajectory-record) entry'
        if is_t:
            nf = coords.shape[0]; cptr = coords.data_ptr()
            assert coords.is_contiguous()
        else:
            coords = numpy.ascontiguousarray(coords, numpy.float64)
            nf = coords.shape[0]; cptr = coords.ctypes.data
        if out is None:
            out = numpy.empty((nf, _lib.NOUT), numpy.float64)
        if status is None:
            status = numpy.empty((nf,), numpy.int32)
        optr = out.data_ptr() if hasattr(out, 'data_ptr') else out.ctypes.data
        sptr = status.data_ptr() if hasattr(status, 'data_ptr') else status.ctypes.data
        _lib.check(_lib.lib.slvgpu_eval_frames_host(self._h, cptr, nf, optr, sptr))
        return out, status

    def set_timing(self, on=True):
        _lib.check(_lib.lib.slvgpu_set_timing(self._h, int(on)))

    def get_timing(self):
        """{kernel: (ms, launches)} accumulated since set_timing(True)."""
        ms = numpy.zeros(4); n = numpy.zeros(4, numpy.int64)
        _lib.check(_lib.lib.slvgpu_get_timing(self._h, ms.ctypes.data, n.ctypes.data))
        return dict((k, (float(ms[i]), int(n[i]))) for i, k in enumerate(('k_frame_elect', 'k_disp', 'k_pol', 'k_rep')))

    def last_launches(self):
        return int(_lib.lib.slvgpu_last_launches(self._h))

    def pol_detail(self, frame=0, max_centres=None):
        k = self._keep
        tot = int(sum(k['meta'][t][_lib.M['NPOL']] for t in k['inst_type']))
        n = min(max_centres or tot, tot)
        dip = numpy.zeros((n, 3)); fld = numpy.zeros((n, 3)); rp = numpy.zeros((n, 3)); av = numpy.zeros((n, 3))
        got = _lib.check(_lib.lib.slvgpu_get_pol_detail(self._h, frame, dip.ctypes.data, fld.ctypes.data, rp.ctypes.data,
                                                          av.ctypes.data, n))
        return dip[:got], fld[:got], rp[:got], av[:got]
