// Developer micro-benchmark: LSU cost of 16-byte global stores in the scratch-scatter pattern of the pass-1 kernels versus
// contiguous stores and shared-memory stores (cycles per warp instruction, 12 warps per SM, all SMs).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(384) k_st(double2 *g, long long *cyc, int iters, long stride_el) {
  extern __shared__ double2 sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2 v = make_double2(threadIdx.x, blockIdx.x);
  // lane -> (segment, column): 13 columns per segment, segments `stride_el` elements apart (pattern of PostScratch)
  const int seg = lane / 13, col = lane - seg * 13;
  double2 *base = g + ((long)blockIdx.x * 12 + warp) * 4096 + (long)seg * stride_el + col;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 25; ++u) {
      if (MODE == 0) base[(long)u * 10 * stride_el / 16 + (it & 7) * 16] = v;          // scatter: 25 rows, far apart
      if (MODE == 1) g[((long)blockIdx.x * 12 + warp) * 4096 + u * 32 + lane + (it & 7) * 800] = v; // contiguous 512 B
      if (MODE == 2) sm[warp * 1024 + u * 32 + lane] = v;                              // shared
      v.x += 1.0;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double2 *g; long long *cyc;
  size_t n = (size_t)1 << 28; // 4 GB
  cudaMalloc(&g, n * 16); cudaMalloc(&cyc, 8 * 256);
  cudaFuncSetAttribute(k_st<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 1024 * 16);
  long long h;
  const char *nm[] = {"STG.128 scatter (3 x 208 B runs)", "STG.128 contiguous 512 B", "STS.128"};
  for (int mode = 0; mode < 3; ++mode) {
    int iters = 200;
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k_st<0><<<148, 384, 0>>>(g, cyc, iters, 10000);
      if (mode == 1) k_st<1><<<148, 384, 0>>>(g, cyc, iters, 10000);
      if (mode == 2) k_st<2><<<148, 384, 12 * 1024 * 16>>>(g, cyc, iters, 10000);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-36s: %8lld cycles for %d warp-stores per warp x 12 warps -> %.1f SM cycles per warp instruction (%s)\n", nm[mode], h, iters * 25,
           (double)h / (iters * 25.0 * 12), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
