// Measured fp64 (DFMA) issue rate of the SMs: the denominator for the device-evaluation roofline
// (BASELINE.json asks for achieved FP64 FLOP/s against the B200 FP64 peak; MEASURED_PEAKS.json holds
// only HBM and bf16 numbers).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CHAINS>
static void run(int blocks_per_sm, int threads, int sms, double* out) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma<CHAINS><<<sms * blocks_per_sm, threads>>>(out, 100, 0.999, 1e-3);
  cudaEventRecord(e0);
  k_dfma<CHAINS><<<sms * blocks_per_sm, threads>>>(out, iters, 0.999, 1e-3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)sms * blocks_per_sm * threads * iters * CHAINS;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("chains=%d blocks/SM=%d threads=%d: %.3f ms, %.2f TFLOP/s, %.1f DFMA lanes/clk/SM at %.0f MHz nominal\n", CHAINS,
         blocks_per_sm, threads, ms, 2 * fma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / sms / (clk * 1e3), clk / 1e3);
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  run<8>(1, 128, sms, out);
  run<8>(1, 512, sms, out);
  run<8>(2, 1024, sms, out);
  run<4>(2, 1024, sms, out);
  run<1>(2, 1024, sms, out);
  run<1>(1, 32, sms, out);   // latency: one warp, one chain
  return 0;
}
