// Developer micro-benchmark: SM-side cost of one tile FFT in shared memory, two-pass large-radix engine (tq_fft2.cuh, what the
// pipelined kernels use: 11-12 warps at 168 registers) against the three-pass small-radix engine (tq_fft3.cuh: 16-32 warps).
// Lock-step passes (__syncthreads), results stored to an L2-resident global buffer like the pass-B epilogue does.
#include <cstdio>
#include <cstring>
#include "tq_fft3.cuh"

template <int STEP> struct PostGlobal {
  cd *out;
  int tcw, c, klow;
  TQ_HD void begin(int k, int cc) { klow = k; c = cc; }
  template <int T> TQ_HD void store(cd v) { out[(size_t)(klow + STEP * T) * tcw + c] = v; }
};

template <int N, int RA, int RB, int NT> __global__ void __launch_bounds__(NT, 1) k_two(cd *out, const cd *twg, long long *cyc, int iters, int tcw, int pitch) {
  extern __shared__ __align__(16) unsigned char smem[];
  cd *tile = reinterpret_cast<cd *>(smem);
  cd *tw = tile + (size_t)N * pitch;
  for (int i = threadIdx.x; i < N * pitch; i += NT) tile[i] = cmk(1e-150 * (i % 7), 1e-150);
  for (int i = threadIdx.x; i < N; i += NT) tw[i] = twg[i];
  __syncthreads();
  PreIdentity pre;
  PostGlobal<RA> post{out + (size_t)blockIdx.x * N * tcw, tcw, 0, 0};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    fft2_pass_a<N, RA, RB, 1, false>(tile, pitch, tcw, tw, threadIdx.x, NT, pre);
    __syncthreads();
    fft2_pass_b<N, RA, RB, 1>(tile, pitch, tcw, threadIdx.x, NT, post);
    __syncthreads();
  }
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

template <int N, int R1, int R2, int R3, int NT> __global__ void __launch_bounds__(NT, 1) k_three(cd *out, const cd *tw1g, const cd *tw2g, long long *cyc, int iters, int tcw) {
  extern __shared__ __align__(16) unsigned char smem[];
  cd *tile = reinterpret_cast<cd *>(smem);
  cd *tw1 = tile + (size_t)N * tcw;
  cd *tw2 = tw1 + N;
  for (int i = threadIdx.x; i < N * tcw; i += NT) tile[i] = cmk(1e-150 * (i % 7), 1e-150);
  for (int i = threadIdx.x; i < N; i += NT) tw1[i] = tw1g[i];
  for (int i = threadIdx.x; i < R2 * R3; i += NT) tw2[i] = tw2g[i];
  __syncthreads();
  PreIdentity pre;
  PostGlobal<R1 * R2> post{out + (size_t)blockIdx.x * N * tcw, tcw, 0, 0};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    fft3_pass1<N, R1, R2, R3, 1>(tile, tcw, tcw, tw1, threadIdx.x, NT, pre);
    __syncthreads();
    fft3_pass2<N, R1, R2, R3, 1>(tile, tcw, tcw, tw2, threadIdx.x, NT);
    __syncthreads();
    fft3_pass3<N, R1, R2, R3, 1>(tile, tcw, tcw, threadIdx.x, NT, post);
    __syncthreads();
  }
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main(int argc, char **argv) {
  cd *out, *tw;
  long long *cyc;
  cudaMalloc(&out, sizeof(cd) * 148 * 400 * 16);
  cudaMalloc(&tw, sizeof(cd) * 2048);
  cudaMemset(tw, 0, sizeof(cd) * 2048);
  cudaMalloc(&cyc, 8 * 1024);
  long long h;
  const int iters = 40;
#define RUN2(N, RA, RB, NT, TCW) RUN2P(N, RA, RB, NT, TCW, TCW)
#define RUN2P(N, RA, RB, NT, TCW, PITCH)                                                                                          \
  {                                                                                                                              \
    auto k = k_two<N, RA, RB, NT>;                                                                                               \
    size_t sm = sizeof(cd) * ((size_t)N * PITCH + N);                                                                            \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                                               \
    for (int rep = 0; rep < 2; ++rep) {                                                                                          \
      k<<<148, NT, sm>>>(out, tw, cyc, iters, TCW, PITCH);                                                                              \
      cudaDeviceSynchronize();                                                                                                   \
    }                                                                                                                            \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                                                              \
    cudaFuncAttributes fa;                                                                                                       \
    cudaFuncGetAttributes(&fa, k);                                                                                               \
    printf("two-pass   N=%d = %2d x %2d      threads %4d regs %3d spill %3zu  tcw %2d pitch %2d : %7.0f cycles/tile  %.3f cycles/element  (%s)\n", N, RA, RB, NT,  \
           fa.numRegs, fa.localSizeBytes, TCW, PITCH, (double)h / iters, (double)h / iters / (N * TCW), cudaGetErrorString(cudaGetLastError()));                      \
  }
#define RUN3(N, R1, R2, R3, NT, TCW)                                                                                             \
  {                                                                                                                              \
    auto k = k_three<N, R1, R2, R3, NT>;                                                                                         \
    size_t sm = sizeof(cd) * ((size_t)N * TCW + N + R2 * R3);                                                                    \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                                               \
    for (int rep = 0; rep < 2; ++rep) {                                                                                          \
      k<<<148, NT, sm>>>(out, tw, tw, cyc, iters, TCW);                                                                          \
      cudaDeviceSynchronize();                                                                                                   \
    }                                                                                                                            \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                                                              \
    cudaFuncAttributes fa;                                                                                                       \
    cudaFuncGetAttributes(&fa, k);                                                                                               \
    printf("three-pass N=%d = %2d x %2d x %2d threads %4d regs %3d spill %3zu  tcw %2d : %7.0f cycles/tile  %.3f cycles/element  (%s)\n", N, R1, R2, R3, \
           NT, fa.numRegs, fa.localSizeBytes, TCW, (double)h / iters, (double)h / iters / (N * TCW), cudaGetErrorString(cudaGetLastError()));                  \
  }
  const char *sel = argc > 1 ? argv[1] : "all";
  if (!strcmp(sel, "all") || !strcmp(sel, "base")) {
    RUN2(250, 10, 25, 352, 13) RUN2(250, 10, 25, 384, 13) RUN2(250, 10, 25, 352, 9)
    RUN3(250, 10, 5, 5, 512, 13) RUN3(250, 10, 5, 5, 768, 13) RUN3(250, 10, 5, 5, 1024, 13) RUN3(250, 5, 10, 5, 1024, 13) RUN3(250, 10, 5, 5, 1024, 9)
    RUN2(400, 20, 20, 352, 9) RUN2(400, 20, 20, 384, 9) RUN2(400, 20, 20, 352, 8)
    RUN3(400, 10, 8, 5, 512, 9) RUN3(400, 10, 8, 5, 768, 9) RUN3(400, 10, 8, 5, 1024, 9) RUN3(400, 8, 10, 5, 1024, 9) RUN3(400, 5, 8, 10, 1024, 9) RUN3(400, 10, 8, 5, 1024, 8)
  }
  if (!strcmp(sel, "all") || !strcmp(sel, "var")) {
    // factorisation, warp count and row pitch of the two-pass engine
    RUN2P(250, 10, 25, 352, 13, 13) RUN2P(250, 10, 25, 352, 13, 14) RUN2P(250, 10, 25, 352, 13, 15) RUN2P(250, 10, 25, 352, 13, 16)
    RUN2P(250, 25, 10, 352, 13, 13) RUN2P(250, 25, 10, 352, 13, 14) RUN2P(250, 10, 25, 256, 13, 13) RUN2P(250, 25, 10, 256, 13, 13)
    RUN2P(250, 10, 25, 352, 9, 9) RUN2P(250, 10, 25, 352, 9, 10) RUN2P(250, 10, 25, 352, 9, 11) RUN2P(250, 25, 10, 352, 9, 9)
    RUN2P(400, 20, 20, 352, 9, 9) RUN2P(400, 20, 20, 352, 9, 10) RUN2P(400, 20, 20, 352, 9, 11) RUN2P(400, 20, 20, 352, 9, 12)
    RUN2P(400, 20, 20, 352, 8, 8) RUN2P(400, 20, 20, 352, 8, 9) RUN2P(400, 16, 25, 352, 9, 9) RUN2P(400, 25, 16, 352, 9, 9)
    RUN2P(400, 20, 20, 256, 9, 9) RUN2P(400, 20, 20, 288, 9, 9)
  }
  return 0;
}
