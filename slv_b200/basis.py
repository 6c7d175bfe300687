"""Contracted Cartesian Gaussian basis description (host side).

Replaces what the reference obtains from ``PyQuante.Ints.getbasis(mol, '6-311++G**')``
(solvshift/slvpar.py:458-462, solvshift/solefp.py:1497-1519): per atom, the shells of
``dat/slvbas`` in file order; an ``L`` shell contributes s then px,py,pz; d shells are the six
Cartesians xx,yy,zz,xy,xz,yz (efprot.f:17-20).  Every primitive carries its own normalisation
and every contracted Cartesian component is normalised to unity (PyQuante CGBF convention).
Only tables are built here -- the integrals themselves are evaluated on the GPU.
"""
import json
import math
import os

import numpy

_SYMBOL = {1: 'H', 3: 'Li', 6: 'C', 7: 'N', 8: 'O', 9: 'F', 11: 'Na', 16: 'S', 17: 'Cl'}
_POWERS = {'S': [(0, 0, 0)], 'P': [(1, 0, 0), (0, 1, 0), (0, 0, 1)],
           'D': [(2, 0, 0), (0, 2, 0), (0, 0, 2), (1, 1, 0), (1, 0, 1), (0, 1, 1)]}
_TABLE = None
MAXPRIM = 6


def _table():
    global _TABLE
    if _TABLE is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'slvbas.json')) as fh:
            _TABLE = json.load(fh)
    return _TABLE


def _dfact(n):
    return 1.0 if n <= 0 else n * _dfact(n - 2)


def prim_norm(alpha, l, m, n):
    """Normalisation constant of x^l y^m z^n exp(-alpha r^2)."""
    L = l + m + n
    return math.sqrt(2.0 ** (2 * L + 1.5) * alpha ** (L + 1.5)
                     / (_dfact(2 * l - 1) * _dfact(2 * m - 1) * _dfact(2 * n - 1) * math.pi ** 1.5))


def _self_overlap(exps, coefs, l, m, n):
    """<chi|chi> of a contracted Cartesian function with already primitive-normalised coefficients."""
    s = 0.0
    for a, ca in zip(exps, coefs):
        for b, cb in zip(exps, coefs):
            p = a + b
            s += ca * cb * (math.pi / p) ** 1.5 * _dfact(2 * l - 1) * _dfact(2 * m - 1) * _dfact(2 * n - 1) / (2.0 * p) ** (l + m + n)
    return s


class BasisSet(object):
    """Function tables of one fragment at given atomic positions.

    Attributes (numpy): center[nbf] atom index, powers[nbf,3], nprim[nbf], exps[nbf,MAXPRIM],
    coefs[nbf,MAXPRIM] (contraction coefficient x primitive norm x contracted norm; zero padded);
    shells[nshell,4] = atom, first function, number of functions (S 1, P 3, D 6, L = SP 4), primitives --
    all functions of a shell share the exponents of its first function; subshells[nsub,4] = atom, first function,
    angular momentum, primitives (L shells split into their S and P parts)."""

    def __init__(self, atno, pos, name='6-311++G**'):
        tab = _table()
        center, powers, nprim, exps, coefs = [], [], [], [], []
        shells = []            # (atom, first function, number of functions, number of primitives) per shell
        for ia, z in enumerate(numpy.asarray(atno, int)):
            for typ, prims in tab[_SYMBOL[int(z)]]:
                subs = [('S', 1), ('P', 2)] if typ == 'L' else [(typ, 1)]
                f0 = len(center)
                for st, col in subs:
                    for (l, m, n) in _POWERS[st]:
                        e = [p[0] for p in prims]
                        c = [p[col] * prim_norm(p[0], l, m, n) for p in prims]
                        nrm = 1.0 / math.sqrt(_self_overlap(e, c, l, m, n))
                        center.append(ia); powers.append((l, m, n)); nprim.append(len(e))
                        exps.append(e + [1.0] * (MAXPRIM - len(e)))
                        coefs.append([x * nrm for x in c] + [0.0] * (MAXPRIM - len(e)))
                shells.append((ia, f0, len(center) - f0, len(prims)))
        self.atno = numpy.asarray(atno, int)
        self.shells = numpy.array(shells, numpy.int32).reshape(-1, 4)
        # sub-shells: (atom, first function, angular momentum, primitives); an L shell is an S and a P sub-shell
        sub = []
        for ia, f0, nf, npr in shells:
            if nf == 4:
                sub += [(ia, f0, 0, npr), (ia, f0 + 1, 1, npr)]
            else:
                sub.append((ia, f0, {1: 0, 3: 1, 6: 2}[nf], npr))
        self.subshells = numpy.array(sub, numpy.int32).reshape(-1, 4)
        self.pos = numpy.asarray(pos, numpy.float64)
        self.center = numpy.array(center, numpy.int32)
        self.powers = numpy.array(powers, numpy.int32)
        self.nprim = numpy.array(nprim, numpy.int32)
        self.exps = numpy.array(exps, numpy.float64)
        self.coefs = numpy.array(coefs, numpy.float64)

    def __len__(self):
        return len(self.center)

    def get_bfsl(self):
        """0-based atom index of every basis function (PyQuante BasisSet.get_bfsl)."""
        return self.center.copy()

    def get_bfst(self):
        """Cartesian powers of every basis function (PyQuante BasisSet.get_bfst)."""
        return self.powers.copy()
