"""Shift-series post-processing on the GPU.

  gtcf(a, b, k)         util/lib/genres.f:1-23 GTCF: TCF(i) = 1/(N-i) sum_j A(j) B(j+i), i = 0..k-1
  cf(y1, y2, norm)      util/slv_calc-tcf:17-28: correlation function of the mean-subtracted series, all lags
  expres(data, dt, ndel) util/slv_calc-expres:73-101: classical line-shape function Re/Im Jc(t)
  trig_transform(f, y0, dy, x, kind)  trapezoid cosine / sine transform: the integrals of util/slv_calc-ftir
  spectral_density(ffcf, ...)         util/slv_calc-ftir:187-231: cosine transform of the FFCF, tau, T2*
  ftir_classical(T, JR, JI, ...)      util/slv_calc-ftir:369-409: line shape from Jc(t)
  hist(x, bins)                       util/slv_hist:53-61: mean, std and the normalised histogram of a shift column

The unit factor of slv_calc-ftir, libbbg.units.UNITS.CmRecToHz / 1e12 "[rad/ps]", comes from the un-vendored libbbg;
CMREC_TO_RADPS = 2 pi c 1e-12 is the physical value of what its comment names and can be overridden per call.
"""
import numpy

_trapz = getattr(numpy, 'trapezoid', None) or numpy.trapz

from . import _lib


def gtcf(a, b, k, demean=False):
    a = numpy.ascontiguousarray(a, numpy.float64); b = numpy.ascontiguousarray(b, numpy.float64)
    assert a.ndim == 1 and a.shape == b.shape
    out = numpy.empty(int(k), numpy.float64)
    _lib.check(_lib.lib.slvgpu_tcf(a.ctypes.data, b.ctypes.data, a.size, int(k), int(bool(demean)), out.ctypes.data))
    return out


def cf(y1, y2, norm=False, kmax=None):
    """Correlation function between y1 and y2 (util/slv_calc-tcf:17-28): f[i] = <dY1(t+i) dY2(t)>, normalised to
    f[0] = 1 when norm is set.  kmax limits the number of lags (default: all, as the reference)."""
    n = len(y1)
    f = gtcf(y2, y1, n if kmax is None else kmax, demean=True)
    if norm:
        f = f / f[0]
    return f


def expres(data, dt, ndel=100):
    """Re[Jc], Im[Jc] at ndel+1 delay times (util/slv_calc-expres:73-101)."""
    d = numpy.ascontiguousarray(data, numpy.float64)
    re = numpy.empty(ndel + 1); im = numpy.empty(ndel + 1)
    _lib.check(_lib.lib.slvgpu_expres(d.ctypes.data, d.size, float(dt), int(ndel), re.ctypes.data, im.ctypes.data))
    return re, im


CMREC_TO_RADPS = 2.0 * numpy.pi * 2.99792458e10 * 1e-12       # rad/ps per cm^-1


def trig_transform(f, y0, dy, x, kind='cos'):
    """out[i] = trapz_y f(y) cos(x_i y)  (or sin) on the uniform grid y_j = y0 + j dy."""
    f = numpy.ascontiguousarray(f, numpy.float64); x = numpy.ascontiguousarray(x, numpy.float64)
    out = numpy.empty(x.size, numpy.float64)
    _lib.check(_lib.lib.slvgpu_trig_transform(f.ctypes.data, f.size, float(y0), float(dy), x.ctypes.data, x.size,
                                              {'cos': 0, 'sin': 1}[kind], out.ctypes.data))
    return out


def hamming(T):
    """f_window of util/slv_calc-ftir:193, 383."""
    T = numpy.asarray(T, numpy.float64)
    return 0.54 + 0.46 * numpy.cos(numpy.pi * T / T[-1])


def spectral_density(ffcf, dtau=1.0, n_delay=10000, w_min=1e-6, w_max=3000.0, n_w=4097, n_pad=0, window=True,
                     cmrec_to_radps=CMREC_TO_RADPS):
    """The cumulant model up to the spectral density (util/slv_calc-ftir:187-231).  ffcf: FFCF samples in cm^-2 at
    time points spaced dtau fs.  Returns dict: T [ps], TCF [rad^2/ps^2] (windowed, padded), W [rad/ps], W_cm1,
    TCF_W (2/C0 int cos(W t) TCF dt), C0, delta, tau, t2star."""
    tcf = numpy.array(ffcf, numpy.float64)[:n_delay] * cmrec_to_radps ** 2
    n_t = len(tcf)
    dt = dtau / 1e3
    T = numpy.arange(n_t) * dt
    if window:
        tcf = tcf * hamming(T)
    n_t += n_pad
    T = numpy.linspace(0.0, dt * (n_t - 1), n_t)
    tcf = numpy.concatenate([tcf, numpy.zeros(n_pad)])
    C0 = tcf[0]
    W = numpy.linspace(w_min, w_max * cmrec_to_radps, n_w)
    tcf_w = trig_transform(tcf, 0.0, dt, W, 'cos') * (2.0 / C0)
    tau = _trapz(tcf / C0, dx=dt)
    return dict(T=T, TCF=tcf, W=W, W_cm1=W / cmrec_to_radps, TCF_W=tcf_w, C0=C0, delta=numpy.sqrt(C0), tau=tau,
                t2star=1.0 / (tau * C0))


def ftir_classical(T, JR, JI, w0, w_min, w_max, T1=100.0, n_delay=None, n_pad=0, window=True, n_w=1024,
                   cmrec_to_radps=CMREC_TO_RADPS):
    """Classical-averaging line shape (util/slv_calc-ftir:369-409) from Jc(t) sampled at T [ps] (uniform grid starting
    at 0; expres() gives Re/Im Jc).  w0, w_min, w_max in cm^-1, T1 in ps.  Returns W_cm1, S, ReI, ImI."""
    T = numpy.asarray(T, numpy.float64); JR = numpy.asarray(JR, numpy.float64); JI = numpy.asarray(JI, numpy.float64)
    dt = T[1] - T[0]
    n_t = len(T) + n_pad
    Tp = numpy.linspace(0.0, dt * (n_t - 1), n_t)
    JR = numpy.concatenate([JR, numpy.zeros(n_pad)]); JI = numpy.concatenate([JI, numpy.zeros(n_pad)])
    if n_delay is not None:
        JR, JI, Tp = JR[:n_delay], JI[:n_delay], Tp[:n_delay]
    damp = numpy.exp(-0.5 * Tp / T1)
    JR = JR * damp; JI = JI * damp
    if window:
        f = hamming(Tp); JR = JR * f; JI = JI * f
    W = numpy.linspace(w_min * cmrec_to_radps, w_max * cmrec_to_radps, n_w)
    DW = W - w0 * cmrec_to_radps
    ReI = trig_transform(JR, 0.0, dt, DW, 'cos')
    ImI = -trig_transform(JI, 0.0, dt, DW, 'sin')
    return W / cmrec_to_radps, ReI + ImI, ReI, ImI


def hist(x, bins):
    """Mean, standard deviation and the normalised (density) histogram of a shift series with `bins` equal bins over
    [min, max] (util/slv_hist:53-61: average, std, pylab.hist(..., normed=True)).  Returns mean, std, density, edges."""
    x = numpy.ascontiguousarray(x, numpy.float64)
    lo, hi = float(x.min()), float(x.max())
    if hi == lo:
        lo, hi = lo - 0.5, hi + 0.5
    counts = numpy.zeros(int(bins), numpy.int64)
    _lib.check(_lib.lib.slvgpu_hist(x.ctypes.data, x.size, lo, hi, int(bins), counts.ctypes.data))
    edges = numpy.linspace(lo, hi, int(bins) + 1)
    density = counts / (counts.sum() * numpy.diff(edges))
    return float(x.mean()), float(x.std()), density, edges
