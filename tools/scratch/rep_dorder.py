import sys, json, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import scan_geometries, SCAN_SUPL, SCAN_REORD
from oracle import solefp_oracle as O
from slv_b200 import slvpar, basis
import slv_b200
golden = json.load(open('tests/golden/small_cluster.json'))
def frag(n): return slvpar.Frag(n)
pars = [frag('nma').get(), frag('water').get()]
c = O.HartreePerHbarToCmRec
class Swapped(basis.BasisSet):
    """integrals evaluated with PyQuante's native d order xx,yy,zz,xy,yz,xz"""
    def __init__(self, atno, pos):
        basis.BasisSet.__init__(self, atno, pos)
        pw = self.powers
        for i in range(len(pw)):
            if tuple(pw[i]) == (1,0,1) and tuple(pw[i+1]) == (0,1,1):
                self.powers[[i, i+1]] = self.powers[[i+1, i]]
need = {0:11.550,3:-4.497,6:24.872,9:101.167,12:143.069,14:129.892,17:67.717,20:16.301}
geos = scan_geometries(golden)
for variant, mb in (('std', basis.BasisSet), ('swap', Swapped)):
    for i in sorted(need):
        r = O.eval_frame(geos[i], [0,1],[12,3], pars, SCAN_SUPL, SCAN_REORD, 8, elect=False, ea=False, corr=False, rep=True, make_basis=mb)
        print(variant, i, need[i], r['rep_mea']*c)
