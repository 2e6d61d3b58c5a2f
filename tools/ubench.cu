// Developer micro-benchmarks (not part of the product): FP64 pipe rates, FP64+INT co-issue, LDS.128 bandwidth on one B200.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int OP, int ILP> __global__ void __launch_bounds__(512) k_fp64(double *out, double s) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double b = s, c = 1.0 - s * 1e-9;
  int ii = threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) a[i] = fma(a[i], c, b);
      if (OP == 1) a[i] = a[i] + b;
      if (OP == 2) a[i] = a[i] * c;
      if (OP == 3) { a[i] = fma(a[i], c, b); ii = ii * 3 + it; }               // 1 int per fp64
      if (OP == 4) { a[i] = fma(a[i], c, b); ii = ii * 3 + it; ii ^= (ii >> 3); } // 2 int per fp64
    }
  }
  double r = 0;
  for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 12345.678 || ii == 123456789) out[0] = r;
}
__global__ void __launch_bounds__(512) k_lds(double *out, int iters) {
  extern __shared__ double2 sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = make_double2(i, 1);
  __syncthreads();
  double2 acc = make_double2(0, 0);
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      double2 v = sm[(idx + u * 512) & 8191];
      acc.x += v.x; acc.y += v.y;
    }
    idx = (idx + 37) & 8191;
  }
  if (acc.x == 1.2345) out[0] = acc.y;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double *out; cudaMalloc(&out, 64);
  int nsm = 148; const int blocks = nsm * 4;
  const char *names[] = {"DFMA", "DADD", "DMUL", "DFMA+1int", "DFMA+3int"};
#define RUN(OP, ILP, NT) { float ms = timeit([&] { k_fp64<OP, ILP><<<blocks, NT>>>(out, 0.5); }); \
    double ops = (double)blocks * NT * ILP * ITER; printf("%-10s ILP=%d NT=%d : %.3f ms  %.2f T instr-lanes/s  (%.1f lanes/clk/SM @1.965GHz)\n", names[OP], ILP, NT, ms, ops / ms / 1e9, ops / ms / 1e3 / 148 / 1.965e6); }
  RUN(0, 8, 512) RUN(0, 4, 512) RUN(0, 2, 512) RUN(0, 1, 512) RUN(0, 8, 128) RUN(0, 4, 128) RUN(0, 2, 128) RUN(0, 1, 128)
  RUN(1, 8, 512) RUN(2, 8, 512) RUN(3, 8, 512) RUN(4, 8, 512)
  {
    cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
    int iters = 2000;
    float ms = timeit([&] { k_lds<<<nsm, 512, 8192 * 16>>>(out, iters); });
    double bytes = (double)nsm * 512 * iters * 8 * 16;
    printf("LDS.128 : %.3f ms  %.1f TB/s total  %.1f B/clk/SM @1.965GHz\n", ms, bytes / ms / 1e9, bytes / ms / 1e3 / 148 / 1.965e6);
  }
  return 0;
}
