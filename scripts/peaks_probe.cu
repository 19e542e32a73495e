// Micro-benchmarks of the two per-SM rates SURVEY 8(d) builds its rooflines on: FP32 FMA and MUFU
// (lg2 / ex2 .approx) throughput of one B200, under the clock the device actually holds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/peaks_probe scripts/peaks_probe.cu
// Prints one JSON line.
#include <cuda_runtime.h>
#include <cstdio>

template <int OP>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = seed + 0.001f * (threadIdx.x + j);
  const float m = 1.0000001f, c = 1e-7f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) a[j] = fmaf(a[j], m, c);
      if (OP == 1) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
      if (OP == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
    }
    if (OP == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fabsf(a[j]) + 1.5f;     // keep lg2's argument positive (1 ALU op, other pipe)
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  if (s == 12345.678f) out[0] = s;
}

template <int OP>
static double run(int blocks, int iters) {
  float* d;
  cudaMalloc(&d, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<OP><<<blocks, 256>>>(d, iters / 4, 1.1f);
  cudaDeviceSynchronize();
  double best = 1e30;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    probe<OP><<<blocks, 256>>>(d, iters, 1.1f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaFree(d);
  return (double)blocks * 256.0 * 8.0 * iters / (best * 1e-3);     // ops per second
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, iters = 1 << 16;
  const double fma = run<0>(blocks, iters), lg2 = run<1>(blocks, iters), ex2 = run<3>(blocks, iters);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f, \"fp32_fma_tflops\": %.2f, "
         "\"fp32_fma_per_clk_per_sm_at_max_clock\": %.1f, \"mufu_lg2_gops\": %.1f, "
         "\"mufu_ex2_gops\": %.1f, \"mufu_lg2_per_clk_per_sm_at_max_clock\": %.2f}\n",
         p.name, p.multiProcessorCount, clk / 1e3, 2.0 * fma / 1e12, fma / (clk * 1e3) / p.multiProcessorCount,
         lg2 / 1e9, ex2 / 1e9, lg2 / (clk * 1e3) / p.multiProcessorCount);
  return 0;
}
