"""BASELINE config 3 shape: MeSCN in a 50/50 water/DMSO mix (400 fragments), full SolEFP (elect+pol+rep+disp)."""
import sys, json, time
import numpy as np, torch
sys.path.insert(0, '.')
from slv_b200 import synth
from slv_b200.slvpar import Frag
from slv_b200.solefp import EFP
nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 400
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
sol, wat, dmso = Frag('mescn'), Frag('water'), Frag('dmso')
types = np.random.default_rng(3).integers(0, 2, nfrag)
t = time.time()
X = synth.make_trajectory(sol.get_pos(), [wat.get_pos(), dmso.get_pos()], types, 32, seed=3, spacing=9.5, jit_lattice=0.3)
X = np.concatenate([X] * (nf // 32))
print('gen', time.time() - t, X.shape)
ind = [0] + [1 + int(t) for t in types]; nmol = [7] + [3 if t == 0 else 10 for t in types]
d = torch.from_numpy(X).cuda()
efp = EFP(elect=True, corr=True, pol=True, disp=True, rep=True, freq=True, mode=4, ccut=40.0, pcut=17.0, ecut=12.0, cunit=True)
rows, st = efp.eval_frames(d, ind, nmol, [sol, wat, dmso], [None] * 3, [None] * 3)
plan = efp._get_plan(); torch.cuda.synchronize()
plan.set_timing(True)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    plan.eval_device(d)
e1.record(); torch.cuda.synchronize()
tm = plan.get_timing()
r = rows.cpu().numpy()
s = efp.shifts_of(rows)
print(json.dumps(dict(frames=nf, ms=e0.elapsed_time(e1) / 3, fps=nf / (e0.elapsed_time(e1) / 3) * 1e3,
                      kernels_ms=dict((k, v[0] / 3) for k, v in tm.items()), n_elect=float(r[:, 11].mean()), n_pol=float(r[:, 12].mean()),
                      n_rep=float(r[:, 13].mean()), iters=float(r[:, 14].mean()), status=int(st.max()),
                      shifts0=dict((k, float(s[k][0])) for k in ('ele_tot', 'pol_mea', 'rep_mea', 'dis_mea', 'total')))))
