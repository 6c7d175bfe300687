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
