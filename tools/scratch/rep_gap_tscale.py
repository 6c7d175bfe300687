"""Is the rep_mea gap a scale factor on the kinetic-energy integrals (or on any other single ingredient)?"""
import sys, json, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import scan_geometries, SCAN_SUPL, SCAN_REORD
from oracle import solefp_oracle as O, ints_oracle as IO
from slv_b200 import slvpar, basis
golden = json.load(open('tests/golden/small_cluster.json'))
pars = [slvpar.Frag('nma').get(), slvpar.Frag('water').get()]
c = O.HartreePerHbarToCmRec
geos = scan_geometries(golden)
need = {0: 11.550, 3: -4.497, 6: 24.872, 9: 101.167, 12: 143.069, 14: 129.892, 17: 67.717, 20: 16.301}
orig = IO.one_electron
def scaled(fs, ft, fs1, ft1):
    def f(bA, pA, bB, pB):
        S, T, S1, T1 = orig(bA, pA, bB, pB)
        return S * fs, T * ft, S1 * fs1, T1 * ft1
    return f
def rep(i, fac):
    IO.one_electron = scaled(*fac)
    r = O.eval_frame(geos[i], [0, 1], [12, 3], pars, SCAN_SUPL, SCAN_REORD, 8, elect=False, ea=False, corr=False, rep=True, make_basis=basis.BasisSet)
    IO.one_electron = orig
    return r['rep_mea'] * c
for name, f0, f1 in (('T and T1', (1, 0, 1, 0), (1, 1, 1, 1)), ('T1 only', (1, 1, 1, 0), (1, 1, 1, 1)), ('S1 and T1', (1, 1, 0, 0), (1, 1, 1, 1))):
    print(name)
    for i in sorted(need):
        a, b = rep(i, f0), rep(i, f1)          # b = a + part
        part = b - a
        lam = (need[i] - a) / part if abs(part) > 1e-12 else float('nan')
        print('  point %2d  without %9.3f  part %9.3f  need %9.3f  -> factor %.4f' % (i, a, part, need[i], lam))
