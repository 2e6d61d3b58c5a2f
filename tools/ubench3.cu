// Developer micro-benchmark: sustained FP64 rate of the register-resident butterflies themselves (no memory traffic).
#include <cstdio>
#include "../triqs_b200/csrc/tq_fft2.cuh"
template <int R, int UNR> __global__ void k_bfly(cd *out, long long *cyc, int iters, double s) {
  cd a[R];
  for (int i = 0; i < R; ++i) a[i] = cmk(threadIdx.x * 1e-3 + i, s * i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      Dft<R, 1>::run(a);
#pragma unroll
      for (int i = 0; i < R; ++i) a[i] = cscale(0.25, a[i]);
    }
  }
  long long t1 = clock64();
  cd r = cmk(0, 0);
  for (int i = 0; i < R; ++i) r = cadd(r, a[i]);
  if (r.x == 12345.678) out[0] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// distinct-register FMA stream: a[i] = fma(b[i], c[i], a[i])
template <int ILP> __global__ void k_fma3(double *out, long long *cyc, int iters, double s) {
  double a[ILP], b[ILP], c[ILP];
  for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x + i; b[i] = 1.0 - 1e-9 * (i + 1) * s; c[i] = s * (i + 2); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b[i], c[i]);
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], b[(i + 1) % ILP], a[(i + 3) % ILP]);
  }
  long long t1 = clock64();
  double r = 0;
  for (int i = 0; i < ILP; ++i) r += a[i] + c[i];
  if (r == 12345.678) out[0] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  cd *out; long long *cyc;
  cudaMalloc(&out, 64); cudaMalloc(&cyc, 8 * 1024);
  long long h;
#define RUNB(R, UNR, NT, FP64_PER) { int iters = 200; k_bfly<R, UNR><<<148, NT>>>(out, cyc, iters, 0.5); cudaDeviceSynchronize(); k_bfly<R, UNR><<<148, NT>>>(out, cyc, iters, 0.5); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); double instr = (double)iters * UNR * (FP64_PER + 2 * R) * (NT / 32) / 4.0; \
    printf("radix-%-2d warps/SM=%2d : %8lld cycles, FP64 warp-instr/clk/SMSP = %.3f (peak 0.5)\n", R, NT / 32, h, instr / h); }
  RUNB(5, 4, 128, 36) RUNB(5, 4, 256, 36) RUNB(5, 4, 384, 36) RUNB(5, 4, 512, 36)
  RUNB(10, 2, 128, 92) RUNB(10, 2, 256, 92) RUNB(10, 2, 384, 92) RUNB(10, 2, 512, 92)
  RUNB(20, 1, 128, 224) RUNB(20, 1, 256, 224) RUNB(20, 1, 384, 224)
  RUNB(25, 1, 128, 424) RUNB(25, 1, 256, 424) RUNB(25, 1, 384, 424)
  RUNB(16, 1, 128, 188) RUNB(16, 1, 384, 188)
#define RUNF(ILP, NT) { int iters = 2000; k_fma3<ILP><<<148, NT>>>((double *)out, cyc, iters, 0.5); cudaDeviceSynchronize(); k_fma3<ILP><<<148, NT>>>((double *)out, cyc, iters, 0.5); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); double instr = (double)iters * 2 * ILP * (NT / 32) / 4.0; \
    printf("fma3 ILP=%d warps/SM=%2d : %8lld cycles, FP64 warp-instr/clk/SMSP = %.3f\n", ILP, NT / 32, h, instr / h); }
  RUNF(8, 128) RUNF(8, 384) RUNF(4, 384)
  return 0;
}
