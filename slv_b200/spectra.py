"""Shift-series post-processing on the GPU.

  gtcf(a, b, k)         util/lib/genres.f:1-23 GTCF: TCF(i) = 1/(N-i) sum_j A(j) B(j+i), i = 0..k-1
  cf(y1, y2, norm)      util/slv_calc-tcf:17-28: correlation function of the mean-subtracted series, all lags
  expres(data, dt, ndel) util/slv_calc-expres:73-101: classical line-shape function Re/Im Jc(t)
"""
import numpy

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
