"""Error of pol_mea (cm-1) against a tol=1e-14 solve as a function of the BiCG stopping tolerance, on the
bench workloads (config 2 and 3).  Decides how much margin the default tolerance has against the parity bar
(1e-6 cm-1 absolute)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from slv_b200.solefp import EFP
for cfg in (2, 3):
    wl = bench.WORKLOADS[cfg]
    bsm, ind, nmol, X = bench.make_frames(wl, wl['nsolv'], 250, 2000, 0)
    d = torch.from_numpy(X).cuda()
    ref = None
    for tol in (1e-14, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6):
        efp = EFP(freq=True, mode=wl['mode'], ccut=bench.CUTS[0], pcut=bench.CUTS[1], ecut=bench.CUTS[2], cunit=True, elect=True, corr=True, pol=True)
        efp.pol_tol = tol
        rows, st = efp.eval_frames(d, ind, nmol, bsm, [None] * len(bsm), [None] * len(bsm))
        rows = rows.cpu().numpy() if hasattr(rows, 'cpu') else np.asarray(rows)
        pol = rows[:, 4]
        if ref is None: ref = pol
        print(json.dumps(dict(config=cfg, tol=tol, iters=float(rows[:, 14].mean()), max_abs_err_cm1=float(np.abs(pol - ref).max()),
                              rms_err=float(np.sqrt(((pol - ref) ** 2).mean())), pol_mean_abs=float(np.abs(ref).mean()), status=int(np.asarray(st.cpu() if hasattr(st,'cpu') else st).max()))))
