"""NumPy restatement of the shift-series post-processing.  TEST INFRASTRUCTURE ONLY.
cf: util/slv_calc-tcf:17-28; expres: util/slv_calc-expres:73-101 (the script's own NumPy routine -- the Fortran
EXPRES of util/lib/genres.f:25-79 is not what the script calls); gtcf: util/lib/genres.f:1-23."""
import numpy as np

_trapz = getattr(np, 'trapezoid', None) or np.trapz


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


# ---- util/slv_calc-ftir (NumPy restatement of the script's own integrals; libbbg unit factor passed in) and util/slv_hist
def trig_transform(f, y0, dy, x, kind='cos'):
    """numpy.trapz(F, axis=0, dx=dy) with F[:, i] = trig(x[i] * y) * f  (slv_calc-ftir:181, 223-226, 399-407)."""
    y = y0 + dy * np.arange(len(f))
    F = (np.cos if kind == 'cos' else np.sin)(np.outer(y, x)) * np.asarray(f)[:, None]
    return _trapz(F, axis=0, dx=dy)


def spectral_density(ffcf, dtau, n_delay, w_min, w_max, n_w, n_pad, window, c):
    r1 = np.array(ffcf, float)[:n_delay] * c ** 2                       # :189
    T = np.arange(len(r1)) * dtau / 1e3                                  # :188
    TCF = r1
    n_t = len(T); dt = T[1] - T[0]
    if window:
        TCF = TCF * (0.54 + 0.46 * np.cos(np.pi * T / T[-1]))           # :193-195
    t_pad = dt * (n_t + n_pad - 1); n_t += n_pad                         # :199-204
    T = np.linspace(0.0, t_pad, n_t)
    TCF = np.concatenate([TCF, np.zeros(n_pad)])
    C0 = TCF[0]
    W = np.linspace(w_min, w_max * c, n_w)                               # :213-216
    F = np.zeros((n_t, n_w))
    for i in range(n_w):
        F[:, i] = 2 / C0 * np.cos(W[i] * T) * TCF                        # :223-225
    TCF_W = _trapz(F, axis=0, dx=dt)
    tau = _trapz(TCF / C0, axis=0, dx=dt)                              # :228
    return dict(T=T, TCF=TCF, W=W, TCF_W=TCF_W, C0=C0, tau=tau, t2star=1.0 / (tau * C0))


def ftir_classical(T, JR, JI, w0, w_min, w_max, T1, n_delay, n_pad, window, n_w, c):
    T = np.asarray(T, float); n_t = len(T); dt = T[1] - T[0]           # :372-376
    t_pad = dt * (n_t + n_pad - 1); n_t += n_pad                         # :378-384
    T = np.linspace(0.0, t_pad, n_t)
    JR = np.concatenate([JR, np.zeros(n_pad)])[:n_delay]; JI = np.concatenate([JI, np.zeros(n_pad)])[:n_delay]
    T = T[:n_delay]
    f = 0.54 + 0.46 * np.cos(np.pi * T / T[-1])                          # :390
    JR = JR * np.exp(-0.5 * T / T1); JI = JI * np.exp(-0.5 * T / T1)    # :391-392
    if window:
        JR = JR * f; JI = JI * f
    W = np.linspace(w_min * c, w_max * c, n_w); DW = W - w0 * c          # :396-401
    ReI = _trapz(np.cos(np.outer(T, DW)) * JR[:, None], axis=0, dx=dt)
    ImI = _trapz(-np.sin(np.outer(T, DW)) * JI[:, None], axis=0, dx=dt)
    return W / c, ReI + ImI, ReI, ImI


def hist(x, bins):
    """util/slv_hist:53-61 (pylab.hist(..., normed=True) is numpy.histogram(..., density=True))."""
    n, edges = np.histogram(x, bins, density=True)
    return float(np.average(x)), float(np.std(x)), n, edges
