import sys; sys.path.insert(0,'/root/repo')
import warnings; warnings.simplefilter('ignore')
import numpy as np, torch, bench
from slv_b200 import _lib
w = bench.make_workload(4, 2000, 10, 40, 0)
X = bench.bohr_frames(w)
efp = w['efp']; efp.max_chunk = 40
efp.set(X[0], w['ind'], w['nmol'], w['bsm'], w['supl'], w['reord'])
efp._removable = efp._removable_types('remove_by_name', False, False)
r1, s1 = efp.eval_frames(torch.from_numpy(X).cuda())
r1 = r1.cpu().numpy()
r2, s2 = efp.eval_frames(w['rec32'], scale=w['scale'], idx=w['idx'])
r3, s3 = efp.eval_frames(X)
names = sorted(_lib.OUT, key=_lib.OUT.get)
for k, n in enumerate(names):
    d12 = np.abs(r1[:, k] - r2[:, k]); d13 = np.abs(r1[:, k] - r3[:, k])
    print('%-12s f32-vs-dev %.3e  host64-vs-dev %.3e  nan %d  val %.6g' % (n, np.nanmax(d12), np.nanmax(d13), np.isnan(r1[:, k]).sum(), r1[0, k]))
r4, _ = efp.eval_frames(torch.from_numpy(X).cuda()); print('dev repeat identical', np.array_equal(r1, r4.cpu().numpy()))
