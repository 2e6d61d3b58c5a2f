// Developer micro-benchmark: FP64 dependent-issue latency and single-SMSP throughput vs. warps x ILP (one block on one SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP> __global__ void k_lat(double *out, long long *cyc, double s) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double b = s, c = 1.0 - s * 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < 512; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) a[i] = fma(a[i], c, b);
      if (OP == 1) a[i] = a[i] + b;
      if (OP == 2) a[i] = a[i] * c;
    }
  }
  long long t1 = clock64();
  double r = 0;
  for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 12345.678) out[0] = r;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP, int ILP> void run(int nthreads, double *out, long long *cyc) {
  k_lat<OP, ILP><<<1, nthreads>>>(out, cyc, 0.5);
  cudaDeviceSynchronize();
  k_lat<OP, ILP><<<1, nthreads>>>(out, cyc, 0.5);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const char *nm[] = {"DFMA", "DADD", "DMUL"};
  int warps = nthreads / 32;
  double per_smsp_warps = warps / 4.0;
  printf("%s warps/SM=%2d ILP=%d : %7lld cycles for 512 iters -> %.2f cyc per dependent step; %.3f warp-instr/clk/SMSP\n", nm[OP], warps, ILP, c,
         c / 512.0, (per_smsp_warps < 1 ? 1 : per_smsp_warps) * ILP * 512.0 / c);
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
  run<0, 1>(32, out, cyc); run<1, 1>(32, out, cyc); run<2, 1>(32, out, cyc);
  run<0, 2>(32, out, cyc); run<0, 4>(32, out, cyc); run<0, 8>(32, out, cyc); run<0, 16>(32, out, cyc);
  run<0, 1>(128, out, cyc); run<0, 2>(128, out, cyc); run<0, 4>(128, out, cyc); run<0, 8>(128, out, cyc); run<0, 16>(128, out, cyc);
  run<0, 1>(256, out, cyc); run<0, 2>(256, out, cyc); run<0, 4>(256, out, cyc); run<0, 8>(256, out, cyc);
  run<0, 1>(512, out, cyc); run<0, 2>(512, out, cyc); run<0, 4>(512, out, cyc); run<0, 8>(512, out, cyc);
  run<0, 1>(1024, out, cyc); run<0, 2>(1024, out, cyc); run<0, 4>(1024, out, cyc);
  run<1, 4>(128, out, cyc); run<1, 8>(128, out, cyc); run<1, 4>(512, out, cyc);
  return 0;
}
