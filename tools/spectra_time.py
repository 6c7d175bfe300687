"""Timing of the shift-series post-processing kernels (row N2) next to the NumPy restatement."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from slv_b200 import spectra
from oracle import spectra_oracle as SO
rng = np.random.default_rng(0)
n, k, ndel = 1000000, 10000, 2000
x = np.cumsum(rng.normal(0, 1, n)) * 0.01 + rng.normal(0, 1, n); y = x[::-1].copy()
spectra.gtcf(x[:1000], y[:1000], 10)
t = time.time(); f = spectra.gtcf(x, y, k); tg = time.time() - t
t = time.time(); re, im = spectra.expres(x - x.mean(), 0.004, ndel); te = time.time() - t
ns = 100000
t = time.time(); SO.gtcf(x[:ns], y[:ns], 1000); tc = time.time() - t
t = time.time(); SO.expres((x - x.mean())[:ns], 0.004, 200); tce = time.time() - t
print(json.dumps(dict(tcf=dict(n=n, lags=k, gpu_s=tg, products_per_s=n * k / tg, cpu_numpy_products_per_s=ns * 1000 / tc),
                      expres=dict(n=n, ndel=ndel, gpu_s=te, sincos_per_s=n * ndel / te, cpu_numpy_per_s=ns * 200 / tce))))
