import sys; sys.path.insert(0,'/root/repo')
import warnings; warnings.simplefilter('ignore')
import numpy as np, glob, os
from slv_b200 import synth
from slv_b200.slvpar import Frag
from slv_b200.solefp import EFP
from oracle import solefp_oracle as O
sol = Frag('mescn')
names = sorted(os.path.basename(f)[:-7] for f in glob.glob('/root/repo/slv_b200/data/frg/*.frg.gz'))
c = O.HartreePerHbarToCmRec
for n in names:
    fr = Frag(n); p = fr.get()
    if 'vecl' not in p: continue
    X = synth.make_trajectory(sol.get_pos(), [fr.get_pos()], np.zeros(6, int), 2, seed=5, spacing=11.0, rmin=3.5)
    ind = [0] + [1] * 6; nmol = [7] + [fr.get_natoms()] * 6
    efp = EFP(elect=True, rep=True, freq=True, mode=4, ccut=1e10, pcut=1e10, ecut=16.0, cunit=True)
    try:
        rows, st = efp.eval_frames(X, ind, nmol, [sol, fr], [None, None], [None, None])
    except Exception as e:
        print('%-16s ERROR %s' % (n, str(e)[:100])); continue
    ref = O.eval_frame(X[0], ind, nmol, [sol.get(), p], [None, None], [None, None], 4, 1e10, 1e10, 16.0, elect=True, rep=True)
    print('%-16s nmos %2d nbas %3d  gpu %.8f  oracle %.8f  n_rep %d status %d' % (n, p['nmos'], p['nbasis'], rows[0, 5] * c, ref['rep_mea'] * c, rows[0, 13], st[0]), flush=True)
