// FP64 DFMA peak micro-benchmark (MEASURED_PEAKS.json has no FP64 figure): 8 independent FMA chains per
// thread, 1024 threads x (4 x SMs) blocks, CUDA-event timed.  Prints TFLOP/s (FMA = 2 flops).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 4, threads = 512, iters = 4096;
  double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 8; ++rep) {
    cudaEventRecord(e0);
    k<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep >= 2 && tf > best) best = tf;
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"fp64_dfma_tflops\": %.3f, \"sms\": %d, \"clock_khz_attr\": %d}\n", best, p.multiProcessorCount, clk);
  return 0;
}
