// FP64 pipe peak of the GPU: a DFMA microbenchmark (SURVEY.md 8d asks for a
// measured number, MEASURED_PEAKS.json has none).  8 independent DFMA chains
// per thread, 1024 threads per CTA, 2 CTAs per SM; prints TFLOP/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int ctas = p.multiProcessorCount * 2, thr = 1024, iters = 4096;
  double* d;
  cudaMalloc(&d, (size_t)ctas * thr * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    k_dfma<<<ctas, thr>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 8 * 16 * (double)iters * ctas * thr;
    const double tf = flop / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %.0f, \"fp64_dfma_tflops\": %.2f, "
         "\"dfma_per_clk_per_sm\": %.1f}\n",
         p.name, p.multiProcessorCount, clk / 1e3, best,
         best * 1e12 / 2.0 / (p.multiProcessorCount * (clk * 1e3)));
  return cudaGetLastError() != cudaSuccess;
}
