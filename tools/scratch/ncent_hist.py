"""Distribution of polarisable centres per frame in the bench workloads (CPU, oracle mollst) and the
per-thread sweep length of k_pol for split factors S (rows x S lanes)."""
import sys, numpy as np
sys.path.insert(0, '.')
import bench
from oracle import solefp_oracle as O
for cfg in (2, 3, 5):
    wl = bench.WORKLOADS[cfg]
    bsm, ind, nmol, X = bench.make_frames(wl, wl['nsolv'], 100, 100, 0)
    pars = [b.get() for b in bsm]
    npol = [len(p['rpol']) for p in pars]
    offs = np.concatenate([[0], np.cumsum(nmol)])
    nc = []
    for f in range(len(X)):
        xc = X[f, :nmol[0]]
        mols = [X[f, offs[i]:offs[i + 1]] for i in range(1, len(nmol))]
        ic, ip, ie = O.mollst(xc, mols, *bench.CUTS)
        nc.append(npol[0] + sum(npol[ind[i + 1]] for i in np.nonzero(ip)[0]))
    nc = np.array(nc)
    print('config', cfg, 'npol', npol, 'ncent mean %.1f sd %.1f min %d max %d  frac>512 %.2f' % (nc.mean(), nc.std(), nc.min(), nc.max(), (nc > 512).mean()))
    ideal = (nc.astype(float) ** 2 / 256).sum()
    for S in (1, 2, 4, 8, 16):
        it = (np.ceil(nc * S / 256) * np.ceil(nc / S)).sum()
        print('   S=%2d  iterations/ideal = %.3f' % (S, it / ideal))
    best = np.min([np.ceil(nc * S / 256) * np.ceil(nc / S) for S in (1, 2, 4, 8, 16)], axis=0).sum()
    print('   best per frame      = %.3f' % (best / ideal))
