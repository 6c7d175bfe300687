import sys, json, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import scan_geometries, SCAN_SUPL, SCAN_REORD
from oracle import solefp_oracle as O
from slv_b200 import slvpar, basis
golden = json.load(open('tests/golden/small_cluster.json'))
pars = [slvpar.Frag('nma').get(), slvpar.Frag('water').get()]
c = O.HartreePerHbarToCmRec
geos = scan_geometries(golden)
rows = []
for i, xyz in enumerate(geos):
    r = O.eval_frame(xyz, [0, 1], [12, 3], pars, SCAN_SUPL, SCAN_REORD, 8, elect=True, ea=True, corr=True, pol=True, disp=True, rep=True, make_basis=basis.BasisSet)
    tot = golden['dat'][i][0]
    t = {k: r[k] * c for k in ('ele_mea', 'ele_ea', 'cor_mea', 'cor_ea', 'pol_mea', 'dis_mea', 'rep_mea')}
    need = tot - sum(t[k] for k in ('ele_mea', 'ele_ea', 'cor_mea', 'cor_ea', 'pol_mea', 'dis_mea'))
    rows.append([i, tot, t['ele_mea'], t['ele_ea'], t['cor_mea'] + t['cor_ea'], t['pol_mea'], t['dis_mea'], t['rep_mea'], need, need - t['rep_mea']])
    print(' '.join('%9.3f' % v for v in rows[-1]))
R = np.array(rows)
d = R[:, 9]
names = ['tot', 'ele_mea', 'ele_ea', 'cor', 'pol', 'dis', 'rep', 'need']
for j, n in enumerate(names):
    x = R[:, 1 + j]
    print(n, 'corr %.4f' % np.corrcoef(x, d)[0, 1], 'slope %.4f' % (np.dot(x, d) / np.dot(x, x)))
A = np.stack([R[:, 7], R[:, 6], np.ones(len(R))], 1)
co, res, *_ = np.linalg.lstsq(A, d, rcond=None)
print('fit diff = a rep + b dis + c:', co, 'resid rms', np.sqrt(((A @ co - d) ** 2).mean()))
