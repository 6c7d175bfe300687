"""Quick device timing of the frame-loop kernels on a synthetic NMA-d7 + 500 waters batch."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
from slv_b200 import synth
from slv_b200.slvpar import Frag
from slv_b200.solefp import EFP
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 500
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 512
sol, wat = Frag('nma-d7'), Frag('water')
t = time.time()
X = synth.make_trajectory(sol.get_pos(), [wat.get_pos()], np.zeros(nw, int), 32, seed=2)
X = np.concatenate([X] * (nf // 32))
print('gen', time.time() - t)
ind = [0] + [1] * nw; nmol = [12] + [3] * nw
d = torch.from_numpy(X).cuda()
for flags in (dict(pol=False, disp=False), dict(pol=False, disp=True), dict(pol=True, disp=False)):
    efp = EFP(elect=True, corr=True, freq=True, mode=8, ccut=40.0, pcut=17.0, ecut=12.0, cunit=True, **flags)
    rows, st = efp.eval_frames(d, ind, nmol, [sol, wat], [None, None], [None, None])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        rows, st = efp.eval_frames(d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    r = rows.cpu().numpy()
    print(json.dumps(dict(flags=flags, frames=nf, ms=ms, fps=nf / ms * 1e3, n_elect=float(r[:, 11].mean()), n_pol=float(r[:, 12].mean()),
                          iters=float(r[:, 14].mean()), resid=float(r[:, 15].max()), status=int(st.max()),
                          ele=float(r[0, 0]), pol=float(r[0, 4]), dis=float(r[0, 6]))))
