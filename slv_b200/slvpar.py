"""Fragment parameter container -- the ``solvshift.slvpar.Frag`` surface.

Reads the reference's ``*.frg`` format unchanged (solvshift/slvpar.py:302-321 and
1148-1535: sections ``[ name ] N= n`` followed by n numbers, five per line) and
applies the same scalings on read: first derivatives x sqrt(redmass), second
derivatives x redmass[mode] (slvpar.py:1278-1336, 1357-1370, 1425-1427, 1473-1474),
triangular unpacking of ``fock``/``fock1``/``gijk`` (1391-1408, 1429-1475); the
``mode`` list is 1-based in the file and 0-based in memory (1185-1187).

Superimposition and tensor rotation (``Frag.sup``, slvpar.py:641-734) run on the
GPU through the C ABI (``slvgpu_frag_sup``); there is no host arithmetic fallback.
"""
import copy as _copy
import gzip
import os
import re

import numpy

from .units import UNITS

__all__ = ['Frag', 'SEC_NAMES']

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')

# header -> key (slvpar.py:824-862).  The reference matches "<name> ]" as a substring of
# the header line, which makes the match exact on the section name.
SEC_NAMES = {
    'Atomic numbers': 'atno', 'Atomic masses': 'atms', 'Atomic coordinates': 'pos',
    'Harmonic frequencies': 'freq', 'Reduced masses': 'redmass',
    'Mass-weighted eigenvectors': 'lvec', 'Cubic anharmonic constants': 'gijk',
    'DMTP origins': 'origin', 'ESP charges': 'esp', 'ChelpG charges': 'chlpg',
    'DMTP centers': 'rdma', 'DMTP charges': 'dmac', 'DMTP dipoles': 'dmad',
    'DMTP quadrupoles': 'dmaq', 'DMTP octupoles': 'dmao',
    'DMTP charges - first derivatives': 'dmac1', 'DMTP dipoles - first derivatives': 'dmad1',
    'DMTP quadrupoles - first derivatives': 'dmaq1', 'DMTP octupoles - first derivatives': 'dmao1',
    'DMTP charges - second derivatives wrt mode': 'dmac2',
    'DMTP dipoles - second derivatives wrt mode': 'dmad2',
    'DMTP quadrupoles - second derivatives wrt mode': 'dmaq2',
    'DMTP octupoles - second derivatives wrt mode': 'dmao2',
    'Polarizable centers': 'rpol', 'Distributed polarizabilities': 'dpol',
    'Distributed polarizabilities - first derivatives': 'dpol1',
    'Distr. pol. wrt imaginary freq': 'dpoli',
    'Distr. pol. wrt imaginary freq - 1st derivatives': 'dpoli1',
    'LMO centroids': 'lmoc', 'LMO centroids - first derivatives': 'lmoc1',
    'Fock matrix': 'fock', 'Fock matrix - first derivatives': 'fock1',
    'Canonical Fock matrix': 'fckc', 'Canonical Fock matrix - first derivatives': 'fckc1',
    'AO->LMO matrix': 'vecl', 'AO->LMO matrix - first derivatives': 'vecl1',
    'AO->CMO matrix': 'vecc', 'AO->CMO matrix - first derivatives': 'vecc1',
}

_GL12 = numpy.array([
    [0.0471753363865118, -0.9815606342467192], [0.1069393259953184, -0.9041172563704749],
    [0.1600783285433462, -0.7699026741943047], [0.2031674267230659, -0.5873179542866175],
    [0.2334925365383548, -0.3678314989981802], [0.2491470458134028, -0.1252334085114689],
    [0.2491470458134028, 0.1252334085114689], [0.2334925365383548, 0.3678314989981802],
    [0.2031674267230659, 0.5873179542866175], [0.1600783285433462, 0.7699026741943047],
    [0.1069393259953184, 0.9041172563704749], [0.0471753363865118, 0.9815606342467192]])


def gauss_legendre_12_weights(v=0.3):
    """Weights W_k with int_0^inf f = sum_k W_k f_k, f tabulated at the 12 stored
    imaginary frequencies (slvpar.py:989-1019: W_k = 2 v w_k / (1 - t_k)^2)."""
    return 2.0 * v * _GL12[:, 0] / (1.0 - _GL12[:, 1]) ** 2


def _text_to_list(text, delimiter=None):
    """'1-5', '1 2 3', '1,2,5-7' -> int array (libbbg.utilities.text_to_list behaviour
    as used by slvpar.py:1186 and md.py:147-158)."""
    out = []
    parts = text.replace(',', ' ').split() if delimiter == ',' else text.split()
    for part in parts:
        part = part.strip()
        if not part:
            continue
        if '-' in part[1:]:
            a, b = part.split('-')
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return numpy.array(out, int)


def _tri_unpack(data, n):
    m = numpy.zeros((n, n), numpy.float64)
    il = numpy.tril_indices(n)
    m[il] = data
    m.T[il] = data
    return m


class Frag(UNITS):
    """Solvatochromic EFP fragment parameters (reference: solvshift/slvpar.py:24)."""

    params = ['water', 'water2', 'nma', 'nma-d7', 'mescn', 'meoac', 'meoh', 'dmso', 'dcm',
              'chcl3', 'na+', 'so3--', 'me-so3-', 'li+', 'et2coo', 'mecn', '4-me-imidazole',
              '4-me-phenol', 'chonh2', 'chonhme', 'comenh2', 'ethane', 'mecoo-', 'menh3+',
              'methane', 'n-propane', 'me-guanidinium+', 'ccl4', 'benzene', 'dmf', 'thf',
              'cyclohexane', 'etoh', 'mesh', 'menh2', 'phenol', 'cho-ch-nh2-ch3', 'dms',
              'imidazole', 'mecooh', 'me-s-co-ch-ch2', 'cl-', 'imidazolium+', 'me-s-cho',
              'phenolate-', 'me-s-13c-15n']

    def __init__(self, file=None):
        self._file = file
        self._dir = None
        self._mol = {}
        self._p = {}
        self._rot = None
        self._transl = None
        self._rms = None
        if file is not None:
            self(file)

    # ------------------------------------------------------------------ reading
    @staticmethod
    def _locate(name):
        """Built-in names resolve under $SLV_DATA/frg/<name>/<name>.frg (slvpar.py:304-311);
        without SLV_DATA the fixtures bundled with this package are used."""
        cands = []
        if 'SLV_DATA' in os.environ:
            cands.append(os.path.join(os.environ['SLV_DATA'], 'frg', name, name + '.frg'))
        cands.append(os.path.join(_DATA_DIR, 'frg', name + '.frg'))
        cands.append(os.path.join(_DATA_DIR, 'frg', name + '.frg.gz'))
        for c in cands:
            if os.path.isfile(c):
                return c
        raise IOError(" No such file: <%s>. There is no *.frg file for that molecule,\n"
                      " the name is misspelled or check if the $SLV_DATA variable is set properly\n"
                      " (echo $SLV_DATA)" % cands[0])

    def __call__(self, file):
        if file in self.params:
            self._file = self._locate(file)
            self._dir = os.path.dirname(self._file) + '/'
        else:
            self._file = file
        opener = gzip.open if self._file.endswith('.gz') else open
        with opener(self._file, 'rt') as fh:
            text = fh.read()
        sections = re.split(r'\[', text)[1:]
        for section in sections:
            self._read(section)
        return

    def _read(self, section):
        lines = section.split('\n')
        querry = lines[0]
        name = querry.split(']')[0].strip()
        if name == 'molecule':
            for line in lines[1:]:
                if '=' not in line:
                    continue
                key, arg = line.split('=', 1)
                key = key.strip(); arg = arg.strip()
                if key in ('name', 'shortname'):
                    self._mol[key] = arg
                elif key == 'atoms':
                    self._mol[key] = arg.split(',')
                elif key in ('natoms', 'nsites', 'nbasis', 'nmos', 'ncmos', 'nmodes', 'npol', 'ndma'):
                    self._mol[key] = int(arg)
                elif key == 'basis':
                    self._mol['method'], self._mol['basis'] = arg.split('/')
                elif key == 'mode':
                    self._mol['mode'] = _text_to_list(arg, delimiter=',') - 1
            return
        if name not in SEC_NAMES:
            raise Exception("Non-standard section name '%s' detected! Check the format of your file!" % name)
        key = SEC_NAMES[name]
        N = int(querry.split('N=')[1])
        nlines = N // 5 + bool(N % 5)
        data = numpy.array(' '.join(lines[1:1 + nlines]).split(), dtype=numpy.float64)
        assert data.size == N, 'section [ %s ]: expected %d numbers, found %d' % (name, N, data.size)
        m = self._mol
        p = self._p
        nat, ndma, npol = m.get('natoms'), m.get('ndma'), m.get('npol')
        nmodes, nmos, ncmos, nbasis = m.get('nmodes'), m.get('nmos'), m.get('ncmos'), m.get('nbasis')
        sder = m.get('mode')
        nsder = 0 if sder is None else len(sder)

        def s1(shape):      # first derivatives carry sqrt(reduced mass) of their mode
            assert 'redmass' in p, 'No reduced masses supplied!'
            d = data.reshape(shape)
            return numpy.sqrt(p['redmass']).reshape((nmodes,) + (1,) * (d.ndim - 1)) * d

        def s2(shape):      # second derivatives carry the reduced mass of their mode
            d = data.reshape(shape)
            return p['redmass'][sder].reshape((nsder,) + (1,) * (d.ndim - 1)) * d

        if key == 'pos':
            assert nat == N // 3, 'natoms in section [ molecule ] is not consistent with section [ Atomic coordinates ]!'
            p[key] = data.reshape(nat, 3)
        elif key == 'origin':
            p[key] = data.reshape(N // 3, 3)
        elif key == 'atno':
            assert nat == N
            p[key] = numpy.array(data, int)
        elif key in ('atms', 'esp', 'chlpg'):
            assert nat == N
            p[key] = data
        elif key == 'rdma':
            assert ndma == N // 3
            p[key] = data.reshape(ndma, 3)
        elif key == 'dmac':
            assert ndma == N
            p[key] = data
        elif key == 'dmad':
            p[key] = data.reshape(ndma, 3)
        elif key == 'dmaq':
            p[key] = data.reshape(ndma, 6)
        elif key == 'dmao':
            p[key] = data.reshape(ndma, 10)
        elif key == 'dmac1':
            p[key] = s1((nmodes, ndma))
        elif key == 'dmad1':
            p[key] = s1((nmodes, ndma, 3))
        elif key == 'dmaq1':
            p[key] = s1((nmodes, ndma, 6))
        elif key == 'dmao1':
            p[key] = s1((nmodes, ndma, 10))
        elif key == 'dmac2':
            p[key] = s2((nsder, ndma))
        elif key == 'dmad2':
            p[key] = s2((nsder, ndma, 3))
        elif key == 'dmaq2':
            p[key] = s2((nsder, ndma, 6))
        elif key == 'dmao2':
            p[key] = s2((nsder, ndma, 10))
        elif key == 'rpol':
            p[key] = data.reshape(npol, 3)
        elif key == 'dpol':
            assert npol == N // 9
            p[key] = data.reshape(npol, 3, 3)
        elif key == 'dpol1':
            p[key] = s1((nmodes, npol, 3, 3))
        elif key == 'dpoli':
            assert npol == N // 108
            p[key] = data.reshape(npol, 12, 3, 3)
        elif key == 'dpoli1':
            p[key] = s1((nmodes, npol, 12, 3, 3))
        elif key in ('freq', 'redmass'):
            assert nmodes == N
            p[key] = data
        elif key == 'lvec':
            p[key] = data.reshape(nmodes, nat, 3)
        elif key == 'gijk':
            g = numpy.zeros((nmodes, nmodes, nmodes), numpy.float64)
            K = 0
            for i in range(nmodes):
                for j in range(i + 1):
                    d = data[K:K + j + 1]
                    k = numpy.arange(j + 1)
                    g[i, j, k] = d; g[i, k, j] = d; g[j, i, k] = d
                    g[j, k, i] = d; g[k, i, j] = d; g[k, j, i] = d
                    K += j + 1
            p[key] = g
        elif key == 'lmoc':
            p[key] = data.reshape(nmos, 3)
        elif key == 'lmoc1':
            p[key] = s1((nmodes, nmos, 3))
        elif key == 'fock':
            p[key] = _tri_unpack(data, nmos)
        elif key == 'fckc':
            p[key] = _tri_unpack(data, ncmos)
        elif key in ('fock1', 'fckc1'):
            n = nmos if key == 'fock1' else ncmos
            d = data.reshape(nmodes, N // nmodes)
            f1 = numpy.array([_tri_unpack(d[i], n) for i in range(nmodes)])
            assert 'redmass' in p, 'No reduced masses supplied!'
            p[key] = numpy.sqrt(p['redmass'])[:, None, None] * f1
        elif key == 'vecl':
            p[key] = data.reshape(nmos, nbasis)
        elif key == 'vecc':
            p[key] = data.reshape(ncmos, nbasis)
        elif key == 'vecl1':
            p[key] = s1((nmodes, nmos, nbasis))
        elif key == 'vecc1':
            p[key] = s1((nmodes, ncmos, nbasis))
        return

    # ------------------------------------------------------------------ getters
    def get(self):
        """Parameter dictionary (slvpar.py:883-938)."""
        par = {}
        for k in ('ndma', 'npol', 'nmodes', 'natoms', 'name', 'basis', 'nbasis', 'nmos', 'mode'):
            if self._mol.get(k) is not None:
                par[k] = self._mol[k]
        par.update(self._p)
        return par

    def get_dir(self):
        return self._dir

    def get_name(self):
        return self._mol.get('name')

    def get_natoms(self):
        return self._mol.get('natoms')

    def get_pos(self):
        return self._p.get('pos')

    def get_atno(self):
        return self._p.get('atno')

    def get_atms(self):
        return self._p.get('atms')

    def get_rotranrms(self):
        """Rotation matrix, translation vector and rms of the last superimposition."""
        return self._rot, self._transl, self._rms

    def get_bfs(self):
        """Basis-set description of the fragment at its current position (replaces the
        PyQuante BasisSet of slvpar.py:458-462): see slv_b200.basis.BasisSet."""
        from .basis import BasisSet
        return BasisSet(self._p['atno'], self._p['pos'], self._mol.get('basis', '6-311++G**'))

    def get_traceless(self):
        """Traceless quadrupoles and octupoles (efprot.f:1214-1326 TRACLS)."""
        return traceless(self._p['dmaq'], self._p['dmao'])

    def get_traceless_1(self):
        """Flat traceless first derivatives (efprot.f:1329-1450 TRACL1)."""
        q, o = traceless(self._p['dmaq1'], self._p['dmao1'])
        return q.ravel(), o.ravel()

    def get_traceless_2(self, mode):
        """Traceless second derivatives wrt <mode> (1-based)."""
        m = list(self._mol['mode']).index(mode - 1)
        return traceless(self._p['dmaq2'][m], self._p['dmao2'][m])

    def get_property(self, property, **kwargs):
        if property.lower() == 'solcamm':
            return self._get_property_solcamm(**kwargs)
        raise NotImplementedError("property '%s' is outside the SolEFP frame-loop path" % property)

    def _get_property_solcamm(self, mode=None, ea=True, dma=False):
        """SolCAMM effective moments (slvpar.py:1025-1073); returns the dma=False tuple
        rdma, pos, solcamm_c, solcamm_d, solcamm_q, solcamm_o."""
        assert mode is not None, " You MUST specify <mode> to compute solvatochromic parameters (in normal numbers, helico)"
        if ea:
            assert (mode - 1) in list(self._mol['mode']), \
                " The second derivatives wrt mode %d are not evaluated for %s!" % (mode, self.get_name())
        p = self._p
        g_wi = p['gijk'][:, mode - 1, mode - 1] / (p['redmass'] * p['freq'] ** 2)
        g_wj = 2.0 * p['redmass'][mode - 1] * p['freq'][mode - 1]
        out = []
        for k1, k2 in (('dmac1', 'dmac2'), ('dmad1', 'dmad2'), ('dmaq1', 'dmaq2'), ('dmao1', 'dmao2')):
            s = -numpy.tensordot(g_wi, p[k1], (0, 0))
            if ea:
                s = s + p[k2][list(self._mol['mode']).index(mode - 1)]
            out.append(s / g_wj)
        return (p['rdma'], p['pos']) + tuple(out)

    # ------------------------------------------------------------------ actions
    def copy(self):
        return _copy.deepcopy(self)

    def sup(self, xyz, suplist=None, dxyz=None, rotran=None):
        """Superimpose the parameters onto <xyz> (Kabsch on the <suplist> atoms) and
        rotate every tensor; returns the rms in a.u. (slvpar.py:641-734).  Runs on the GPU."""
        from . import _lib
        xyz = numpy.ascontiguousarray(xyz, numpy.float64)
        if dxyz is not None:
            xyz = xyz - dxyz
        if rotran is None:
            rot, transl, rms = _lib.kabsch(self._p['pos'], xyz, suplist)
        else:
            rot, transl = rotran
            rms = 0.0
        self._rot, self._transl, self._rms = rot, transl, rms
        _lib.frag_transform(self, numpy.ascontiguousarray(rot, numpy.float64),
                            numpy.ascontiguousarray(transl, numpy.float64).reshape(3))
        return rms

    def rotate(self, rot):
        from . import _lib
        _lib.frag_transform(self, numpy.ascontiguousarray(rot, numpy.float64), numpy.zeros(3))

    def translate(self, transl):
        for k in ('pos', 'lmoc', 'rpol', 'rdma'):
            if k in self._p:
                self._p[k] = self._p[k] + numpy.asarray(transl, numpy.float64)

    def __repr__(self):
        m = self._mol
        return ' SLV SOLVATOCHROMIC EFP PARAMETERS: %s (natoms %s ndma %s npol %s nmos %s nmodes %s)' % (
            m.get('name'), m.get('natoms'), m.get('ndma'), m.get('npol'), m.get('nmos'), m.get('nmodes'))


def traceless(q, o):
    """Buckingham traceless forms in reduced storage (efprot.f:1256-1292):
    Theta = 3/2 Q - 1/2 tr(Q) delta ; Omega = 5/2 O - 1/2 (t_a d_bc + t_b d_ac + t_c d_ab),
    t_a = sum_b O_abb.  Q: XX YY ZZ XY XZ YZ; O: XXX YYY ZZZ XXY XXZ XYY YYZ XZZ YZZ XYZ.
    Pure re-indexing of parameters (host-side plan building)."""
    q = numpy.asarray(q, numpy.float64); o = numpy.asarray(o, numpy.float64)
    tr = q[..., 0] + q[..., 1] + q[..., 2]
    qt = 1.5 * q
    for k in range(3):
        qt[..., k] -= 0.5 * tr
    tx = o[..., 0] + o[..., 5] + o[..., 7]
    ty = o[..., 3] + o[..., 1] + o[..., 8]
    tz = o[..., 4] + o[..., 6] + o[..., 2]
    ot = 2.5 * o
    ot[..., 0] -= 1.5 * tx; ot[..., 1] -= 1.5 * ty; ot[..., 2] -= 1.5 * tz
    ot[..., 3] -= 0.5 * ty; ot[..., 4] -= 0.5 * tz; ot[..., 5] -= 0.5 * tx
    ot[..., 6] -= 0.5 * tz; ot[..., 7] -= 0.5 * tx; ot[..., 8] -= 0.5 * ty
    return qt, ot
