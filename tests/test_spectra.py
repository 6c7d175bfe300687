"""Shift-series post-processing (row N2): oracle vs the transpiled GTCF, and the CUDA kernels vs the oracle."""
import numpy as np
import pytest


def _series(n, seed):
    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.normal(0, 1, n)) * 0.05 + rng.normal(0, 1, n)       # correlated noise, like a shift trajectory
    return x


def test_oracle_gtcf_matches_reference_fortran():
    from oracle import ref_natives as RN, spectra_oracle as SO
    if not RN.available():
        pytest.skip('oracle/_ref not built')
    a, b = _series(700, 1), _series(700, 2)
    assert np.abs(RN.gtcf(a, b, 300) - SO.gtcf(a, b, 300)).max() < 1e-12
    # the script-level cf() is GTCF on the mean-subtracted series with the arguments swapped
    f = SO.cf(a, b)
    assert np.abs(f - RN.gtcf(b - b.mean(), a - a.mean(), 700)).max() < 1e-11 * np.abs(f).max()


@pytest.mark.gpu
@pytest.mark.parametrize('n,k', [(1, 1), (9, 9), (1000, 1000), (20011, 4097)])
def test_gpu_tcf(n, k):
    from slv_b200 import spectra
    from oracle import spectra_oracle as SO
    a, b = _series(n, 3), _series(n, 4)
    got = spectra.gtcf(a, b, k)
    ref = SO.gtcf(a, b, k)
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    if n > 8:
        f = spectra.cf(a, b, norm=True, kmax=k)
        assert np.abs(f - SO.cf(a, b, norm=True)[:k]).max() < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize('n,ndel', [(2, 1), (50, 49), (6000, 800)])
def test_gpu_expres(n, ndel):
    from slv_b200 import spectra
    from oracle import spectra_oracle as SO
    d = _series(n, 5); d -= d.mean()
    re, im = spectra.expres(d, 0.004, ndel)
    rre, rim = SO.expres(d, 0.004, ndel)
    assert np.abs(re - rre).max() < 1e-11 and np.abs(im - rim).max() < 1e-11 and re[0] == 1.0 and im[0] == 0.0


def test_oracle_ftir_classical_gives_the_lorentzian():
    """Known answer for the slv_calc-ftir restatement: Jc(t) = exp(-t / tau) with no window, long T1 and a fine, long
    time grid integrates to the Lorentzian tau / (1 + (dw tau)^2), centred at w0."""
    from oracle import spectra_oracle as SO
    c = 0.188365156731
    tau = 1.7
    T = np.arange(0, 60.0, 0.002)
    W, S, ReI, ImI = SO.ftir_classical(T, np.exp(-T / tau), np.zeros_like(T), 2000.0, 1990.0, 2010.0, 1e12, None, 0, False, 201, c)
    dw = (W - 2000.0) * c
    assert np.abs(S - tau / (1.0 + (dw * tau) ** 2)).max() < 2e-6
    assert abs(W[np.argmax(S)] - 2000.0) < 1e-9 and np.abs(ImI).max() == 0.0


def test_oracle_hist_normalisation():
    from oracle import spectra_oracle as SO
    x = _series(5000, 9)
    m, s, n, edges = SO.hist(x, 37)
    assert abs((n * np.diff(edges)).sum() - 1.0) < 1e-12 and abs(m - x.mean()) < 1e-14 and len(edges) == 38


@pytest.mark.gpu
@pytest.mark.parametrize('n,m,kind', [(2, 1, 'cos'), (777, 33, 'sin'), (8000, 1024, 'cos')])
def test_gpu_trig_transform(n, m, kind):
    from slv_b200 import spectra
    from oracle import spectra_oracle as SO
    f = _series(n, 11); x = np.linspace(-3.0, 7.0, m)
    got = spectra.trig_transform(f, 0.25, 0.01, x, kind)
    ref = SO.trig_transform(f, 0.25, 0.01, x, kind)
    assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
def test_gpu_ftir_models_and_hist():
    """FFCF -> spectral density, Jc(t) -> classical line shape, histogram: CUDA path against the NumPy restatement of
    util/slv_calc-ftir and util/slv_hist on a synthetic shift trajectory."""
    from slv_b200 import spectra
    from oracle import spectra_oracle as SO
    x = _series(20000, 5) * 3.0
    c = spectra.CMREC_TO_RADPS
    ffcf = spectra.cf(x, x, norm=False, kmax=3000)
    got = spectra.spectral_density(ffcf, dtau=2.0, n_delay=2500, w_max=400.0, n_w=513, n_pad=100)
    ref = SO.spectral_density(ffcf, 2.0, 2500, 1e-6, 400.0, 513, 100, True, c)
    assert np.abs(got['TCF_W'] - ref['TCF_W']).max() < 1e-10 * np.abs(ref['TCF_W']).max()
    assert abs(got['tau'] - ref['tau']) < 1e-12 * abs(ref['tau']) and abs(got['t2star'] - ref['t2star']) < 1e-12 * abs(ref['t2star'])
    re, im = spectra.expres(x * c, 0.002, 600)
    T = np.arange(601) * 0.002
    W, S, ReI, ImI = spectra.ftir_classical(T, re, im, 2100.0, 2050.0, 2150.0, T1=5.0, n_delay=550, n_pad=20, n_w=256)
    Wr, Sr, ReIr, ImIr = SO.ftir_classical(T, re, im, 2100.0, 2050.0, 2150.0, 5.0, 550, 20, True, 256, c)
    assert np.abs(W - Wr).max() < 1e-9 and np.abs(S - Sr).max() < 1e-11 * np.abs(Sr).max()
    assert np.abs(ReI - ReIr).max() < 1e-11 * np.abs(Sr).max() and np.abs(ImI - ImIr).max() < 1e-11 * np.abs(Sr).max()
    m, s, dens, edges = spectra.hist(x, 64)
    mr, sr, nr, er = SO.hist(x, 64)
    assert abs(m - mr) < 1e-12 and abs(s - sr) < 1e-12 and np.abs(edges - er).max() < 1e-12
    assert np.abs(dens - nr).max() < 1e-12 * nr.max()
