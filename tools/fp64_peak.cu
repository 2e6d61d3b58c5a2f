// Measured FP64 (DFMA) issue peak of the device: the denominator of the Dyson k-sum's FP64-pipe fraction.
// Independent fused multiply-adds on distinct registers, no memory traffic; timed with CUDA events over ~50 ms per shape;
// prints one JSON object (bench.py reads profiles/fp64_peak.json, a copy of it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k_fma(double *out, int iters, double s) {
  double a[ILP], b[ILP], c[ILP];
  for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x + i; b[i] = 1.0 - 1e-9 * (i + 1) * s; c[i] = s * (i + 2); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b[i], c[i]);
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], b[(i + 1) % ILP], a[(i + 3) % ILP]);
  }
  double r = 0;
  for (int i = 0; i < ILP; ++i) r += a[i] + c[i];
  if (r == 12345.678) out[0] = r;
}
int main() {
  double *out;
  cudaMalloc(&out, 64);
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  int best_nt = 0;
  const int nts[4] = {128, 256, 512, 1024};
  printf("{\"shapes\": [");
  for (int k = 0; k < 4; ++k) {
    const int nt = nts[k], iters = 40000;
    k_fma<8><<<sms, nt>>>(out, 1000, 0.5);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_fma<8><<<sms, nt>>>(out, iters, 0.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double lane_instr = (double)iters * 16 * nt * sms; // 2 x ILP fma per thread per iteration
    const double T = lane_instr / (ms * 1e-3) / 1e12;
    printf("%s{\"threads_per_sm\": %d, \"ms\": %.3f, \"T_lane_instr_per_s\": %.3f}", k ? ", " : "", nt, ms, T);
    if (T > best) { best = T; best_nt = nt; }
  }
  printf("], \"dfma_T_instr_per_s\": %.3f, \"best_threads_per_sm\": %d, \"sm_count\": %d, \"max_sm_khz\": %d, "
         "\"dfma_per_clk_per_sm_at_max_clock\": %.2f, \"fp64_tflops\": %.2f}\n",
         best, best_nt, sms, khz, best * 1e12 / ((double)sms * khz * 1e3), 2 * best);
  return 0;
}
