// FP64 tensor-core micro-benchmark: mma.sync.aligned.m8n8k4.row.col.f64 (DMMA) throughput on this GPU, next to the DFMA
// figure of tools/fp64_peak.cu.  NCH independent accumulator chains per warp, W warps per CTA, 2 CTAs per SM.
// Prints TFLOP/s (one DMMA = 8 x 8 x 4 x 2 = 512 flop per warp) for several (warps, chains) shapes.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NCH>
__global__ void k(double *out, int iters, double a, double b) {
  double d0[NCH], d1[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) { d0[c] = threadIdx.x * 1e-3 + c; d1[c] = c * 0.5; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) dmma(d0[c], d1[c], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < NCH; ++c) s += d0[c] + d1[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCH>
static double run(int sms, int warps, double *out) {
  const int blocks = sms * 2, threads = warps * 32, iters = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    k<NCH><<<blocks, threads>>>(out, iters, 1e-3, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 512.0 * 8 * NCH * (double)iters * blocks * warps;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep >= 2 && tf > best) best = tf;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 2 * 1024);
  printf("{\"sms\": %d, \"dmma_m8n8k4_tflops\": {", p.multiProcessorCount);
  const int ws[4] = {2, 4, 8, 16};
  double best = 0;
  for (int wi = 0; wi < 4; ++wi) {
    const int w = ws[wi];
    const double t1 = run<1>(p.multiProcessorCount, w, out), t2 = run<2>(p.multiProcessorCount, w, out),
                 t4 = run<4>(p.multiProcessorCount, w, out), t8 = run<8>(p.multiProcessorCount, w, out);
    printf("%s\"warps_per_cta_%d\": {\"chains1\": %.2f, \"chains2\": %.2f, \"chains4\": %.2f, \"chains8\": %.2f}", wi ? ", " : "", w, t1, t2, t4, t8);
    if (t8 > best) best = t8; if (t4 > best) best = t4; if (t2 > best) best = t2;
  }
  printf("}, \"dmma_peak_tflops\": %.2f}\n", best);
  return 0;
}
