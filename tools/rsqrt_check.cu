// Accuracy of slv::rsqrt_fast (hardware seed + one cubic step) against 1/sqrt in FP64 over the range of squared
// inter-site distances (1e-2 .. 1e6 Bohr^2): prints the maximum relative error.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../slv_b200/csrc/slv_math.cuh"
__global__ void k(double *err, double *seed_err, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = exp(log(1e-2) + (log(1e6) - log(1e-2)) * (i + 0.37) / n);
  double y = slv::rsqrt_fast(x), ref = 1.0 / sqrt(x), s;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(x));
  err[i] = fabs(y - ref) / ref; seed_err[i] = fabs(s - ref) / ref;
}
int main() {
  const int n = 1 << 22;
  double *e, *s; cudaMallocManaged(&e, n * sizeof(double)); cudaMallocManaged(&s, n * sizeof(double));
  k<<<n / 256, 256>>>(e, s, n); cudaDeviceSynchronize();
  double m = 0, ms = 0; for (int i = 0; i < n; ++i) { if (e[i] > m) m = e[i]; if (s[i] > ms) ms = s[i]; }
  printf("{\"rsqrt_fast_max_rel_err\": %.3e, \"hardware_seed_max_rel_err\": %.3e, \"samples\": %d}\n", m, ms, n);
  return 0;
}
