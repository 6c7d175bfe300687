import sys, numpy as np
sys.path.insert(0, '.')
from oracle import ints_oracle as IO
from slv_b200 import slvpar, basis
for n in ('nma', 'water', 'meoh'):
    try: p = slvpar.Frag(n).get()
    except Exception as e: print(n, e); continue
    b = basis.BasisSet(p['atno'], p['pos'])
    S, T, S1, T1 = IO.one_electron(b, p['pos'], b, p['pos'])
    C = p['vecl']
    O = C @ S @ C.T
    print(n, C.shape, 'max|CSC^T - I| =', np.abs(O - np.eye(len(O))).max(), 'diag', np.diag(O)[:4])
    if p.get('vecl1') is not None:
        # d/dQ (C S C^T) = 0 :  C1 S C^T + C S C1^T + C dS C^T = 0
        L = p['lvec']; at = b.center
        # S1[k,l,a] = d<k|l>/dA_k (bra centre moves);  for same-molecule, ket centre also moves: d/dB = -(S1^T)
        m = 7
        dS = np.einsum('ka,kla->kl', L[m][at], S1)
        dS = dS + dS.T
        D = p['vecl1'][m] @ S @ C.T + C @ S @ p['vecl1'][m].T + C @ dS @ C.T
        print('   mode 8 orthonormality derivative residual', np.abs(D).max(), ' |C1 S C^T| max', np.abs(p['vecl1'][m] @ S @ C.T).max())
