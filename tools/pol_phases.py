"""Where k_pol's time goes: run the bench workload with the BiCG iteration cap at 1, 2 and unlimited."""
import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
from slv_b200 import synth
from slv_b200.slvpar import Frag
from slv_b200.solefp import EFP
nw, nf = 500, 5888
sol, wat = Frag('nma-d7'), Frag('water')
X = synth.make_trajectory(sol.get_pos(), [wat.get_pos()], np.zeros(nw, int), 64, seed=2)
X = np.concatenate([X] * (nf // 64))
ind = [0] + [1] * nw; nmol = [12] + [3] * nw
d = torch.from_numpy(X).cuda()
for maxit in (1, 2, 3, 200):
    efp = EFP(elect=True, corr=True, pol=True, freq=True, mode=8, ccut=40.0, pcut=17.0, ecut=12.0, cunit=True)
    efp.pol_maxit = maxit
    rows, st = efp.eval_frames(d, ind, nmol, [sol, wat], [None, None], [None, None])
    plan = efp._get_plan(); torch.cuda.synchronize()
    plan.set_timing(True)
    for _ in range(3):
        plan.eval_device(d)
    t = plan.get_timing()
    print(json.dumps(dict(maxit=maxit, k_pol_ms=t['k_pol'][0] / 3, k_elect_ms=t['k_frame_elect'][0] / 3, iters=float(rows[:, 14].mean()))))
