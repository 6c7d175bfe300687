"""NumPy restatement of the shift-series post-processing.  TEST INFRASTRUCTURE ONLY.
cf: util/slv_calc-tcf:17-28; expres: util/slv_calc-expres:73-101 (the script's own NumPy routine -- the Fortran
EXPRES of util/lib/genres.f:25-79 is not what the script calls); gtcf: util/lib/genres.f:1-23."""
import numpy as np


def cf(y1, y2, norm=False):
    Y1 = y1 - np.mean(y1); Y2 = y2 - np.mean(y2)
    result = np.correlate(Y1, Y2, mode='full')
    f = result[result.size // 2:].copy()
    L = len(f)
    f /= (L - np.arange(L))
    f /= f[0]
    if not norm:
        f *= np.mean(Y1 * Y2)
    return f


def expres(data, dt, ndel=100):
    ReJ = np.zeros(ndel + 1); ReJ[0] = 1.0
    ImJ = np.zeros(ndel + 1)
    as_strided = np.lib.stride_tricks.as_strided
    S = np.zeros(len(data)); S[1:] = (data[1:] + data[0:-1]) * dt / 2.0
    N = len(data)
    for i in range(1, ndel + 1):
        neval = N - i
        s = as_strided(S.copy(), (neval + 1, i), (S.strides[0], S.strides[0])).sum(axis=1)
        ReJ[i] = np.cos(s).sum() / float(neval)
        ImJ[i] = -np.sin(s).sum() / float(neval)
    return ReJ, ImJ


def gtcf(a, b, k):
    n = len(a)
    return np.array([np.dot(a[:n - i], b[i:]) / float(n - i) for i in range(k)])
