import sys; sys.path.insert(0,'/root/repo')
import warnings; warnings.simplefilter('ignore')
import numpy as np, bench
from slv_b200.solefp import EFP
w = bench.make_workload(4, 2000, 2, 2, 0)
X = bench.bohr_frames(w)
ind, nmol, bsm = np.array(w['ind']), np.array(w['nmol']), w['bsm']
off = np.concatenate([[0], np.cumsum(nmol)])
names = [b.get().get('name') for b in bsm]
def run(keep):
    keep = [0] + [i for i in keep if i != 0]
    cols = np.concatenate([np.arange(off[i], off[i + 1]) for i in keep])
    efp = EFP(elect=True, rep=True, freq=True, mode=4, ccut=45.0, pcut=16.0, ecut=13.0, cunit=True)
    rows, st = efp.eval_frames(np.ascontiguousarray(X[:, cols]), list(ind[keep]), list(nmol[keep]), bsm, w['supl'], w['reord'])
    return rows[0, 5], rows[0, 13], st[0]
allk = [i for i in range(len(ind)) if not bsm[ind[i]].get().get('env')]
print('all', run(allk))
for t in sorted(set(ind[1:])):
    if bsm[t].get().get('env'): continue
    k = [i for i in allk if ind[i] == t]
    r = run(k)
    print('type %2d %-28s inst %4d nmos %2d nbas %3d ->' % (t, names[t], len(k), bsm[t].get()['nmos'], bsm[t].get()['nbasis']), r, flush=True)
