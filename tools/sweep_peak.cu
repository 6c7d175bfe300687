// Practical ceiling of the k_pol sweep body: the inner loop of slv::k_pol (slv::sweep_pair over a resident
// 512-centre shared-memory tile) run alone -- no BiCG phases, no barriers -- at the kernel's launch shape
// (2 x 256 threads per SM, 128 registers), next to a pure DFMA loop at the same occupancy.  Prints the FP64
// instruction issue rate of each as a fraction of the DFMA rate, i.e. what the instruction mix itself allows.
#include <cstdio>
#include <cuda_runtime.h>
#include "../slv_b200/csrc/slv_math.cuh"
namespace slv {
__device__ __forceinline__ void sweep_pair(const double rc[3], const bool same, const double tt[9], double ax[3], double ay[3]) {
  const double Rx = rc[0] - tt[0], Ry = rc[1] - tt[1], Rz = rc[2] - tt[2];
  const double r2 = same ? 1.0 : (Rx * Rx + Ry * Ry + Rz * Rz);
  const double u = rsqrt_fast(r2), u2 = u * u;
  const double u3 = same ? 0.0 : u2 * u, t5 = 3.0 * u3 * u2;
  const double Rxj = fma(Rz, tt[5], fma(Ry, tt[4], Rx * tt[3]));
  const double Ryj = fma(Rz, tt[8], fma(Ry, tt[7], Rx * tt[6]));
  const double fx = -t5 * Rxj, fy = -t5 * Ryj;
  ax[0] = fma(fx, Rx, fma(u3, tt[3], ax[0])); ax[1] = fma(fx, Ry, fma(u3, tt[4], ax[1])); ax[2] = fma(fx, Rz, fma(u3, tt[5], ax[2]));
  ay[0] = fma(fy, Rx, fma(u3, tt[6], ay[0])); ay[1] = fma(fy, Ry, fma(u3, tt[7], ay[1])); ay[2] = fma(fy, Rz, fma(u3, tt[8], ay[2]));
}
}
constexpr int TILE = 512;
template <int ROWS>
__global__ void __launch_bounds__(256, 2) k_sweep(double *out, int reps) {
  __shared__ double s_tile[TILE * 10];
  __shared__ int s_mol[TILE];
  for (int j = threadIdx.x; j < TILE; j += blockDim.x) {
    for (int d = 0; d < 10; ++d) s_tile[j * 10 + d] = 0.37 * ((j * 7 + d * 13) % 101) + 0.01 * d;
    s_mol[j] = j / 5;
  }
  __syncthreads();
  double r[ROWS][3], ax[ROWS][3], ay[ROWS][3]; int m[ROWS];
  for (int k = 0; k < ROWS; ++k) {
    const int c = threadIdx.x + k * 256;
    for (int d = 0; d < 3; ++d) { r[k][d] = s_tile[c * 10 + d] + 0.5; ax[k][d] = 0; ay[k][d] = 0; }
    m[k] = s_mol[c];
  }
  for (int it = 0; it < reps; ++it) {
#pragma unroll 2
    for (int jj = 0; jj < TILE; ++jj) {
      const double2 *t2 = reinterpret_cast<const double2 *>(s_tile + jj * 10);
      const double2 q0 = t2[0], q1 = t2[1], q2 = t2[2], q3 = t2[3], q4 = t2[4];
      const double tt[9] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y, q4.x};
      const int mj = s_mol[jj];
#pragma unroll
      for (int k = 0; k < ROWS; ++k) slv::sweep_pair(r[k], mj == m[k], tt, ax[k], ay[k]);
    }
  }
  double s = 0;
  for (int k = 0; k < ROWS; ++k) for (int d = 0; d < 3; ++d) s += ax[k][d] + ay[k][d];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256, 2) k_dfma(double *out, int iters, double a, double b) {
  double x[12];
  for (int i = 0; i < 12; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int k = 0; k < 12; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0; for (int k = 0; k < 12; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> static float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 2;
  double *out; cudaMalloc(&out, sizeof(double) * blocks * 256);
  const int iters = 8192, reps = 40;
  const float t0 = timeit([&] { k_dfma<<<blocks, 256>>>(out, iters, 0.999999, 1e-7); });
  const double dfma_rate = 12.0 * 8 * iters * (double)blocks * 256 / (t0 * 1e-3);          // thread-DFMA per second
  const float t1 = timeit([&] { k_sweep<1><<<blocks, 256>>>(out, reps); });
  const float t2 = timeit([&] { k_sweep<2><<<blocks, 256>>>(out, reps); });
  const double p1 = (double)reps * TILE * blocks * 256 / (t1 * 1e-3), p2 = 2.0 * reps * TILE * blocks * 256 / (t2 * 1e-3);
  printf("{\"dfma_tflops_16warps\": %.3f, \"pairs_per_s_1row\": %.4e, \"pairs_per_s_2row\": %.4e, "
         "\"fp64_instr_frac_of_dfma_1row\": %.3f, \"fp64_instr_frac_of_dfma_2row\": %.3f, \"alg_tflops_1row\": %.3f, \"alg_tflops_2row\": %.3f}\n",
         2 * dfma_rate / 1e12, p1, p2, 35.0 * p1 / dfma_rate, 35.0 * p2 / dfma_rate, 49.0 * p1 / 1e12, 49.0 * p2 / 1e12);
  return 0;
}
