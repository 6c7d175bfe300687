"""The oracle's own basis-set builder.  TEST INFRASTRUCTURE ONLY -- never imported by the product.

The reference builds the AO basis of a fragment with ``PyQuante.Ints.getbasis(mol, par['basis'])``
(solvshift/solefp.py:1497-1498, 1518-1519; solvshift/slvpar.py:458-462) from the un-vendored PyQuante
"1.0.3-BBG"; the only in-tree source of the 6-311++G** tables is the GAMESS-format file ``dat/slvbas``
(copied verbatim as the fixture oracle/data/slvbas).  This module parses that text itself and builds the
function tables independently of ``slv_b200/basis.py`` (which reads a JSON conversion and uses closed-form
double-factorial norms): primitive and contracted norms come from Gamma-function moments of the 1-D Gaussians.  tests/test_oracle_vs_ref.py asserts that the two
builders agree table by table, so a wrong contraction, normalisation or function order in the product can no
longer hide behind an oracle that shares its tables.

Conventions (PyQuante CGBF): per atom the shells in file order; an ``L`` shell gives s then px, py, pz; d shells
are the six Cartesians xx, yy, zz, xy, xz, yz (solvshift/src/efprot.f:17-20); every primitive is normalised and
every contracted Cartesian component is normalised to unity.
"""
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SYMBOL = {1: 'H', 3: 'Li', 6: 'C', 7: 'N', 8: 'O', 9: 'F', 11: 'Na', 16: 'S', 17: 'Cl'}
CART = {'S': [(0, 0, 0)], 'P': [(1, 0, 0), (0, 1, 0), (0, 0, 1)],
        'D': [(2, 0, 0), (0, 2, 0), (0, 0, 2), (1, 1, 0), (1, 0, 1), (0, 1, 1)]}
MAXPRIM = 6
_CACHE = {}


def parse_gamess_basis(path=None):
    """{symbol: [(shell type, [(exponent, c_s(, c_p)), ...]), ...]} from a GAMESS $DATA-style basis listing:
    a header line '<symbol> <name>', then blocks '<S|P|D|L> <nprim>' followed by nprim numbered lines."""
    path = path or os.path.join(_HERE, 'data', 'slvbas')
    if path in _CACHE:
        return _CACHE[path]
    out, cur = {}, None
    with open(path) as fh:
        rows = [ln.split() for ln in fh]
    k = 0
    while k < len(rows):
        r = rows[k]
        if len(r) == 2 and r[0] in SYMBOL.values() and not r[1].isdigit():
            cur = r[0]; out[cur] = []
        elif len(r) == 2 and r[0] in ('S', 'P', 'D', 'L') and r[1].isdigit():
            n = int(r[1])
            ncol = 3 if r[0] == 'L' else 2                        # (the F d shell of dat/slvbas carries a spare column)
            prims = [tuple(float(x) for x in rows[k + 1 + q][1:1 + ncol]) for q in range(n)]
            assert all(len(p) == ncol for p in prims), (cur, r)
            out[cur].append((r[0], prims)); k += n
        k += 1
    _CACHE[path] = out
    return out


def _moment(a2, n):
    """integral of x^(2n) exp(-a2 x^2) over the real line."""
    return math.gamma(n + 0.5) / a2 ** (n + 0.5)


class OracleBasis(object):
    """center[nbf], powers[nbf,3], nprim[nbf], exps[nbf,MAXPRIM], coefs[nbf,MAXPRIM] (contraction coefficient x
    primitive norm x contracted norm, zero padded) -- the layout oracle/ints_oracle.py consumes."""

    def __init__(self, atno, pos, path=None):
        tab = parse_gamess_basis(path)
        center, powers, nprim, exps, coefs = [], [], [], [], []
        for ia, z in enumerate(np.asarray(atno, int)):
            for typ, prims in tab[SYMBOL[int(z)]]:
                for sub, col in ([('S', 1), ('P', 2)] if typ == 'L' else [(typ, 1)]):
                    for lmn in CART[sub]:
                        e = [p[0] for p in prims]
                        c = [p[col] / math.sqrt(_moment(2.0 * p[0], lmn[0]) * _moment(2.0 * p[0], lmn[1]) * _moment(2.0 * p[0], lmn[2]))
                             for p in prims]
                        center.append(ia); powers.append(lmn); nprim.append(len(e))
                        exps.append(e + [1.0] * (MAXPRIM - len(e))); coefs.append(c + [0.0] * (MAXPRIM - len(e)))
        self.atno = np.asarray(atno, int)
        self.pos = np.asarray(pos, np.float64)
        self.center = np.array(center, np.int32); self.powers = np.array(powers, np.int32)
        self.nprim = np.array(nprim, np.int32)
        self.exps = np.array(exps, np.float64); self.coefs = np.array(coefs, np.float64)
        # contracted norm: <chi|chi> = sum_pq c_p c_q prod_axis moment(a_p + a_q, l_axis) = 1
        for k in range(len(center)):
            l, m, n = powers[k]
            s = 0.0
            for a, ca in zip(self.exps[k], self.coefs[k]):
                for b, cb in zip(self.exps[k], self.coefs[k]):
                    s += ca * cb * _moment(a + b, l) * _moment(a + b, m) * _moment(a + b, n)
            self.coefs[k] /= math.sqrt(s)

    def __len__(self):
        return len(self.center)


def make_basis(atno, pos):
    return OracleBasis(atno, pos)
