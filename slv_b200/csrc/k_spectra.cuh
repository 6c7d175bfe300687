// k_spectra.cuh -- post-processing of the per-frame shift series:
//   k_tcf    time-correlation function  TCF(i) = 1/(N-i) sum_j A(j) B(j+i)          (util/lib/genres.f:1-23 GTCF;
//            util/slv_calc-tcf:17-28 `cf` is the same sum on mean-subtracted series)
//   k_expres classical line-shape function Jc(t_i) = < exp(-i int_0^{t_i} dw) >       (util/slv_calc-expres:73-101)
//   k_trig_transform  cosine / sine trapezoid transforms of util/slv_calc-ftir (spectral density, line shape)
//   k_hist   histogram of a shift series (util/slv_hist)
// All are streaming sums over an L2-resident series.
#pragma once
#include "slv_common.cuh"

namespace slv {

constexpr int TCF_LAGS = 8;       // lags per CTA: a[j] is loaded once for 8 products
constexpr int SPEC_THREADS = 256;

__global__ void __launch_bounds__(SPEC_THREADS)
k_tcf(const double *__restrict__ a, const double *__restrict__ b, long long n, long long k, double *__restrict__ tcf) {
  __shared__ double s_red[32];
  const long long i0 = (long long)blockIdx.x * TCF_LAGS;
  double acc[TCF_LAGS];
#pragma unroll
  for (int l = 0; l < TCF_LAGS; ++l) acc[l] = 0.0;
  // every lag i0+l uses j < n - (i0+l); the common range first, the ragged tail after it
  const long long mfull = n - (i0 + TCF_LAGS - 1);
  for (long long j = threadIdx.x; j < mfull; j += blockDim.x) {
    const double aj = a[j];
#pragma unroll
    for (int l = 0; l < TCF_LAGS; ++l) acc[l] = fma(aj, b[j + i0 + l], acc[l]);
  }
  for (long long j = (mfull > 0 ? mfull : 0) + threadIdx.x; j < n - i0; j += blockDim.x) {
    const double aj = a[j];
#pragma unroll
    for (int l = 0; l < TCF_LAGS; ++l)
      if (j + i0 + l < n) acc[l] = fma(aj, b[j + i0 + l], acc[l]);
  }
#pragma unroll
  for (int l = 0; l < TCF_LAGS; ++l) {
    const double s = block_sum(acc[l], s_red);
    if (threadIdx.x == 0 && i0 + l < k) tcf[i0 + l] = s / (double)(n - (i0 + l));
  }
}

// mean of a series (one CTA, deterministic)
__global__ void __launch_bounds__(1024) k_mean(const double *__restrict__ x, long long n, double *__restrict__ out) {
  __shared__ double s_red[32];
  double s = 0.0;
  for (long long j = threadIdx.x; j < n; j += blockDim.x) s += x[j];
  s = block_sum(s, s_red);
  if (threadIdx.x == 0) out[0] = s / (double)n;
}

__global__ void k_shift(double *__restrict__ x, long long n, const double *__restrict__ mean) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) x[j] -= mean[0];
}

// C[m] = sum_{t<m} S[t], S[0] = 0, S[t] = (f[t] + f[t-1]) dt/2 ; m = 0..n   (one CTA: chunked serial sums + scan of chunk totals)
__global__ void __launch_bounds__(1024) k_trapz_prefix(const double *__restrict__ f, long long n, double dt, double *__restrict__ C) {
  __shared__ double s_tot[1024];
  const int t = threadIdx.x, nt = blockDim.x;
  const long long per = (n + nt - 1) / nt, lo = (long long)t * per, hi = lo + per < n ? lo + per : n;
  double s = 0.0;
  for (long long m = lo; m < hi; ++m) s += (m > 0 ? (f[m] + f[m - 1]) * dt * 0.5 : 0.0);
  s_tot[t] = s;
  __syncthreads();
  if (t == 0) {
    double run = 0.0;
    for (int i = 0; i < nt; ++i) { const double v = s_tot[i]; s_tot[i] = run; run += v; }
  }
  __syncthreads();
  double run = s_tot[t];
  for (long long m = lo; m < hi; ++m) {
    C[m] = run;
    run += (m > 0 ? (f[m] + f[m - 1]) * dt * 0.5 : 0.0);
  }
  if (hi == n && lo < n) C[n] = run;
  if (n == 0 && t == 0) C[0] = 0.0;
}

// Re/Im J(i) = +/- sum_{n=0}^{N-i} cos/sin(C[n+i] - C[n]) / (N - i),  i = 1..ndel  (J(0) = 1)
__global__ void __launch_bounds__(SPEC_THREADS)
k_expres(const double *__restrict__ C, long long n, long long ndel, double *__restrict__ re, double *__restrict__ im) {
  __shared__ double s_red[32];
  const long long i = (long long)blockIdx.x;
  if (i == 0) { if (threadIdx.x == 0) { re[0] = 1.0; im[0] = 0.0; } return; }
  double sc = 0.0, ss = 0.0;
  for (long long m = threadIdx.x; m <= n - i; m += blockDim.x) {
    double sv, cv;
    sincos(C[m + i] - C[m], &sv, &cv);
    sc += cv; ss += sv;
  }
  sc = block_sum(sc, s_red); ss = block_sum(ss, s_red);
  if (threadIdx.x == 0) { re[i] = sc / (double)(n - i); im[i] = -ss / (double)(n - i); }
}

// out[i] = trapezoid integral over y of f(y) trig(x_i y) on the uniform grid y_j = y0 + j dy, j < n; trig = cos (kind 0)
// or sin (kind 1).  This is every frequency/time integral of util/slv_calc-ftir: the spectral density
// (:223-226), the line-broadening functions (:252-259) and the classical line shape (:399-407).  One CTA per x_i.
__global__ void __launch_bounds__(SPEC_THREADS)
k_trig_transform(const double *__restrict__ f, long long n, double y0, double dy, const double *__restrict__ x, int kind,
                 double *__restrict__ out) {
  __shared__ double s_red[32];
  const double xi = x[blockIdx.x];
  double acc = 0.0;
  for (long long j = threadIdx.x; j < n; j += blockDim.x) {
    const double w = (j == 0 || j == n - 1) ? 0.5 : 1.0;
    const double a = xi * (y0 + (double)j * dy);
    acc = fma(w * f[j], kind ? sin(a) : cos(a), acc);
  }
  acc = block_sum(acc, s_red);
  if (threadIdx.x == 0) out[blockIdx.x] = acc * dy;
}

// counts of x in `bins` equal-width bins over [lo, hi] with numpy.histogram's edge rule (the last bin is closed);
// per-CTA shared-memory histograms merged with integer atomics (exact, order-independent)
__global__ void __launch_bounds__(SPEC_THREADS)
k_hist(const double *__restrict__ x, long long n, double lo, double hi, int bins, unsigned long long *__restrict__ counts) {
  extern __shared__ unsigned int s_h[];
  for (int b = threadIdx.x; b < bins; b += blockDim.x) s_h[b] = 0u;
  __syncthreads();
  const double sc = (double)bins / (hi - lo);
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const double v = x[j];
    if (v >= lo && v <= hi) {
      int b = (int)((v - lo) * sc);
      if (b >= bins) b = bins - 1;
      atomicAdd(&s_h[b], 1u);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += blockDim.x)
    if (s_h[b]) atomicAdd(&counts[b], (unsigned long long)s_h[b]);
}

}  // namespace slv
