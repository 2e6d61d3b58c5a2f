// Batched orbital-sized complex matrix inversion and the fused lattice Dyson k-sum.
//
// Restates  invert_in_place            c++/triqs/gfs/functions/functions2.hpp:197-201
//           _gf_invert_data_in_place   c++/triqs/gfs/gf/gf_expr.hpp:219-222
//           SumkDiscrete.__call__      python/triqs/sumk/sumk_discrete.py:140-166
// nda::inverse_in_place (LAPACK zgetrf + zgetri for n > 3) is external to the reference tree; here every
// thread inverts one n x n matrix entirely in registers by Gauss-Jordan elimination with partial (row)
// pivoting, fully unrolled, with predicated register swaps instead of indexed accesses.  In the k-sum the
// inverse is consumed immediately: G(k, w) (214 GB at 64^3 x 2048 x 5x5) never exists in memory.
#include "tq_common.cuh"
#include "tq_linalg.cuh"

#include <algorithm>

// -------------------------------------------------------------------------------------------------
// tq_invert_batched: a[n_pts][n][n] in place.  Coalesced staging through shared memory.
// -------------------------------------------------------------------------------------------------
template <int N> __global__ void __launch_bounds__(128) k_invert_batched(cd *a, long n_pts) {
  constexpr int NN = N * N;
  constexpr int STR = (NN % 2 == 0) ? NN + 1 : NN; // odd number of complex per matrix: conflict-free 16-byte accesses
  extern __shared__ double2 sm[];                  // [128][STR]
  const long base = (long)blockIdx.x * 128;
  const int cnt = (int)min(128L, n_pts - base);
  cd *g = a + (size_t)base * NN;
  for (int i = threadIdx.x; i < cnt * NN; i += 128) {
    int m = i / NN, e = i - m * NN;
    sm[m * STR + e] = g[i];
  }
  __syncthreads();
  if ((int)threadIdx.x < cnt) {
    cd M[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) M[i][j] = sm[threadIdx.x * STR + i * N + j];
    if (!tq_invert_inplace_nopivot<N>(M)) { // natural pivots rejected: redo with partial pivoting
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) M[i][j] = sm[threadIdx.x * STR + i * N + j];
      tq_invert_inplace<N>(M);
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) sm[threadIdx.x * STR + i * N + j] = M[i][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cnt * NN; i += 128) {
    int m = i / NN, e = i - m * NN;
    g[i] = sm[m * STR + e];
  }
}


// -------------------------------------------------------------------------------------------------
// n > 8 (nda::inverse_in_place has no size limit, functions2.hpp:197-201): one WARP per matrix, the matrix in shared memory,
// the same Gauss-Jordan elimination with partial (row) pivoting as the register version, lanes over the rows / the
// (row, column) pairs of each step.  `form`: the matrix is built on the fly as A(w) - eps_k (Dyson) instead of read.
// -------------------------------------------------------------------------------------------------
constexpr int TQ_GENERIC_MATRIX_MAX = 64;
struct WarpInvForm {
  const cd *eps, *sigma; // eps [n_k][n][n], sigma [n_w][n][n]
  long k0;               // first k-point of this launch
  int n_w, first, stat;
  double beta, mu;
};
template <bool FORM> __global__ void k_invert_warp(cd *a, long n_mat, int n, WarpInvForm F) {
  extern __shared__ double2 sm[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = n | 1; // odd pitch: conflict-free column access
  cd *A = sm + (size_t)warp * ((size_t)n * ld + n);
  cd *col = A + (size_t)n * ld;
  int *piv = reinterpret_cast<int *>(sm + (size_t)warps * ((size_t)n * ld + n)) + warp * n;
  const int nn = n * n;
  for (long m = (long)blockIdx.x * warps + warp; m < n_mat; m += (long)gridDim.x * warps) {
    cd *g = a + (size_t)m * nn;
    if (FORM) { // matrix m = (k - k0) * n_w + w:  (i w_n + mu) 1 - Sigma(w) - eps_k
      const long kk = m / F.n_w;
      const int w = (int)(m - kk * F.n_w);
      const double om = M_PI * (double)(2 * (long)(F.first + w) + F.stat) / F.beta;
      const cd *e = F.eps + (size_t)(F.k0 + kk) * nn, *sg = F.sigma + (size_t)w * nn;
      for (int i = lane; i < nn; i += 32) {
        const int r = i / n, c = i - r * n;
        cd v = cmk(-sg[i].x - e[i].x, -sg[i].y - e[i].y);
        if (r == c) v = cmk(v.x + F.mu, v.y + om);
        A[r * ld + c] = v;
      }
    } else {
      for (int i = lane; i < nn; i += 32) A[(i / n) * ld + (i % n)] = g[i];
    }
    __syncwarp();
    for (int p = 0; p < n; ++p) {
      // pivot search in column p, rows p .. n-1 (first maximum wins, like the serial loop)
      double best = -1.0;
      int r = p;
      for (int i = p + lane; i < n; i += 32) {
        const double v = cabs2(A[i * ld + p]);
        if (v > best) {
          best = v;
          r = i;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int orr = __shfl_xor_sync(0xffffffffu, r, o);
        if (ob > best || (ob == best && orr < r)) {
          best = ob;
          r = orr;
        }
      }
      if (lane == 0) piv[p] = r;
      if (r != p)
        for (int j = lane; j < n; j += 32) {
          const cd t = A[p * ld + j];
          A[p * ld + j] = A[r * ld + j];
          A[r * ld + j] = t;
        }
      __syncwarp();
      const cd d = A[p * ld + p];
      const double inv = 1.0 / cabs2(d);
      const cd ip = cmk(d.x * inv, -d.y * inv);
      for (int i = lane; i < n; i += 32) col[i] = A[i * ld + p];
      __syncwarp();
      for (int j = lane; j < n; j += 32) A[p * ld + j] = cmul(j == p ? cmk(1.0, 0.0) : A[p * ld + j], ip);
      for (int i = lane; i < n; i += 32)
        if (i != p) A[i * ld + p] = cmk(0.0, 0.0);
      __syncwarp();
      for (int e = lane; e < nn; e += 32) {
        const int i = e / n, j = e - i * n;
        if (i != p) A[i * ld + j] = cfma(cneg(col[i]), A[p * ld + j], A[i * ld + j]);
      }
      __syncwarp();
    }
    // undo the row interchanges as column interchanges, in reverse order
    for (int p = n - 1; p >= 0; --p) {
      const int c = piv[p];
      if (c != p)
        for (int i = lane; i < n; i += 32) {
          const cd t = A[i * ld + p];
          A[i * ld + p] = A[i * ld + c];
          A[i * ld + c] = t;
        }
      __syncwarp();
    }
    for (int i = lane; i < nn; i += 32) g[i] = A[(i / n) * ld + (i % n)];
    __syncwarp();
  }
}
static size_t warp_inv_smem(int n, int warps) { return (sizeof(cd) * ((size_t)n * (n | 1) + n) + sizeof(int) * n) * warps + 16; }
template <bool FORM> static tq_status launch_invert_warp(tq_ctx *ctx, cd *d, long n_mat, int n, const WarpInvForm &F) {
  int warps = 8;
  while (warps > 1 && warp_inv_smem(n, warps) > (size_t)ctx->max_smem_optin / 2) warps >>= 1;
  const size_t smem = warp_inv_smem(n, warps);
  if (smem > (size_t)ctx->max_smem_optin) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "matrix too large for one SM's shared memory");
  auto k = k_invert_warp<FORM>;
  TQ_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long blocks = std::min<long>((n_mat + warps - 1) / warps, 16L * ctx->sm_count);
  {
    KernelScope ks(ctx, FORM ? "dyson_invert_warp" : "invert_batched_warp");
    k<<<(unsigned)blocks, warps * 32, smem, ctx->stream>>>(d, n_mat, n, F);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  return TQ_OK;
}
// sum over the k-points of a slab: out[w][e] (+)= sum_k weight_k G[k][w][e], fixed order (deterministic)
__global__ void k_dyson_slab_sum(const cd *__restrict__ g, long nk, long n_elem /* n_w n n */, const double *__restrict__ weights, long k0,
                                 double uniform_weight, int accumulate, cd *__restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_elem) return;
  cd s = accumulate ? out[i] : cmk(0.0, 0.0);
  for (long k = 0; k < nk; ++k) s = caxpy(weights ? __ldg(&weights[k0 + k]) : uniform_weight, g[(size_t)k * n_elem + i], s);
  out[i] = s;
}

extern "C" tq_status tq_invert_batched(tq_ctx *ctx, int n, long n_pts, tq_complex *a) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!a || n < 1 || n_pts < 0) return tq_fail(ctx, TQ_ERR_INVALID, "tq_invert_batched: invalid argument");
  if (n > TQ_MAX_MATRIX_DIM) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tq_invert_batched: n > " + std::to_string(TQ_MAX_MATRIX_DIM));
  static_assert(TQ_MAX_MATRIX_DIM == TQ_GENERIC_MATRIX_MAX, "header and kernel limits agree");
  if (n_pts == 0) return TQ_OK;
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = sizeof(cd) * (size_t)n_pts * n * n;
  const bool dev = tq_is_device_ptr(a);
  cd *d = (cd *)a;
  if (!dev) {
    TQ_TRY(tq_reserve(ctx, ctx->stage_in, bytes));
    d = (cd *)ctx->stage_in.p;
    TQ_CUDA(ctx, cudaMemcpyAsync(d, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  const long nb = (n_pts + 127) / 128;
  if (nb > 0x7fffffffL) return tq_fail(ctx, TQ_ERR_INVALID, "tq_invert_batched: too many points");
  if (n > 8) { // warp-per-matrix kernel
    TQ_TRY(launch_invert_warp<false>(ctx, d, n_pts, n, WarpInvForm{}));
    if (!dev) {
      TQ_CUDA(ctx, cudaMemcpyAsync(a, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
      TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return TQ_OK;
  }
#define TQ_INV_CASE(NV)                                                                                                          \
  case NV: {                                                                                                                     \
    constexpr int NN = NV * NV, STR = (NN % 2 == 0) ? NN + 1 : NN;                                                               \
    const size_t smem = sizeof(cd) * 128 * STR;                                                                                  \
    auto k = k_invert_batched<NV>;                                                                                               \
    TQ_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                               \
    k<<<(unsigned)nb, 128, smem, ctx->stream>>>(d, n_pts);                                                                       \
  } break;
  KernelScope ks(ctx, "invert_batched");
  switch (n) {
    TQ_INV_CASE(1)
    TQ_INV_CASE(2)
    TQ_INV_CASE(3)
    TQ_INV_CASE(4)
    TQ_INV_CASE(5)
    TQ_INV_CASE(6)
    TQ_INV_CASE(7)
    TQ_INV_CASE(8)
  }
#undef TQ_INV_CASE
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  if (!dev) {
    TQ_CUDA(ctx, cudaMemcpyAsync(a, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return TQ_OK;
}

// -------------------------------------------------------------------------------------------------
// fused k-sum.  CTA = 128 frequencies x one chunk of k-points; thread = one frequency.
//   A(w) = (i w_n + mu) 1 - Sigma(w)  lives in shared memory ([element][thread], conflict free)
//   per k: M = A - eps_k (eps_k read with warp-uniform loads), M <- M^-1 in registers, acc += weight * M
//   partial[chunk][w][n][n] is reduced by a second tiny kernel in fixed order (deterministic result).
// -------------------------------------------------------------------------------------------------
#ifndef TQ_DYSON_MINB
#define TQ_DYSON_MINB 1
#endif
template <int N> __global__ void __launch_bounds__(128, TQ_DYSON_MINB) k_dyson_partial(const cd *__restrict__ eps, const double *__restrict__ weights,
                                                                       double uniform_weight, const cd *__restrict__ sigma, int n_w, long n_k,
                                                                       long k_per_chunk, double beta, int stat, int first, double mu,
                                                                       cd *__restrict__ partial) {
  constexpr int NN = N * N;
  extern __shared__ double As_[]; // [2 NN][128]
  double(*As)[128] = reinterpret_cast<double(*)[128]>(As_);
  const int w = blockIdx.x * 128 + threadIdx.x;
  const bool active = w < n_w;
  const int wc = active ? w : n_w - 1;
  {
    const double om = M_PI * (double)(2 * (long)(first + wc) + stat) / beta; // matsubara.hpp:55
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        cd s = sigma[(size_t)wc * NN + i * N + j];
        double re = -s.x, im = -s.y;
        if (i == j) {
          re += mu;
          im += om;
        }
        As[2 * (i * N + j)][threadIdx.x] = re;
        As[2 * (i * N + j) + 1][threadIdx.x] = im;
      }
  }
  cd acc[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) acc[i][j] = cmk(0, 0);
  const long k0 = (long)blockIdx.y * k_per_chunk;
  const long k1 = min(n_k, k0 + k_per_chunk);
  for (long k = k0; k < k1; ++k) {
    const cd *e = eps + (size_t)k * NN;
    cd M[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        cd ev = __ldg(&e[i * N + j]);
        M[i][j] = cmk(As[2 * (i * N + j)][threadIdx.x] - ev.x, As[2 * (i * N + j) + 1][threadIdx.x] - ev.y);
      }
    if (!tq_invert_inplace_nopivot<N>(M)) { // natural pivots rejected (rare): redo this matrix with partial pivoting
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
          cd ev = __ldg(&e[i * N + j]);
          M[i][j] = cmk(As[2 * (i * N + j)][threadIdx.x] - ev.x, As[2 * (i * N + j) + 1][threadIdx.x] - ev.y);
        }
      tq_invert_inplace<N>(M);
    }
    const double wk = weights ? __ldg(&weights[k]) : uniform_weight;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) acc[i][j] = caxpy(wk, M[i][j], acc[i][j]);
  }
  if (active) {
    cd *p = partial + ((size_t)blockIdx.y * n_w + w) * NN;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) p[i * N + j] = acc[i][j];
  }
}

__global__ void k_dyson_reduce(const cd *__restrict__ partial, int n_chunks, long n_elem, cd *__restrict__ out) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_elem) return;
  cd s = cmk(0, 0);
  for (int c = 0; c < n_chunks; ++c) s = cadd(s, partial[(size_t)c * n_elem + i]);
  out[i] = s;
}

extern "C" tq_status tq_dyson_ksum(tq_ctx *ctx, int n, long n_k, long n_iw, double beta, int statistic, double mu, const tq_complex *eps_k,
                                   const double *weights, double uniform_weight, const tq_complex *sigma, tq_complex *out, int all_reduce) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!eps_k || !sigma || !out || n < 1 || n_k < 0 || n_iw < 1 || !(beta > 0) || (statistic != TQ_FERMION && statistic != TQ_BOSON))
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_dyson_ksum: invalid argument");
  if (n > TQ_MAX_MATRIX_DIM) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tq_dyson_ksum: n > " + std::to_string(TQ_MAX_MATRIX_DIM));
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long last = n_iw - 1;
  const long first = -(last + (statistic == TQ_FERMION ? 1 : 0));
  const long n_w = last - first + 1;
  const long nn = (long)n * n;
  const void *d_eps = nullptr, *d_sig = nullptr, *d_wts = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, eps_k, sizeof(cd) * (size_t)n_k * nn, ctx->stage_in, &d_eps));
  TQ_TRY(tq_stage_in(ctx, sigma, sizeof(cd) * (size_t)n_w * nn, ctx->stage_in2, &d_sig));
  if (weights) TQ_TRY(tq_stage_in(ctx, weights, sizeof(double) * (size_t)n_k, ctx->stage_in3, &d_wts));
  TQ_TRY(tq_stage_out_begin(ctx, out, sizeof(cd) * (size_t)n_w * nn, ctx->stage_out, &d_out, &out_host));
  if (n > 8) {
    // large targets: slabs of k-points; G(k, w) of a slab is formed and inverted in a scratch of <= 256 MB (warp per matrix), then
    // summed over k with the weights in fixed order.  The register path below never materialises G(k, w).
    const long n_elem = n_w * nn;
    const long slab = std::max<long>(1, std::min<long>(std::max<long>(n_k, 1), (256L << 20) / (long)(sizeof(cd) * n_elem)));
    TQ_TRY(tq_reserve(ctx, ctx->scratch, sizeof(cd) * (size_t)slab * n_elem));
    cd *d_g = (cd *)ctx->scratch.p;
    if (n_k == 0) TQ_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(cd) * (size_t)n_elem, ctx->stream));
    for (long k0 = 0; k0 < n_k; k0 += slab) {
      const long nk = std::min(slab, n_k - k0);
      WarpInvForm F{(const cd *)d_eps, (const cd *)d_sig, k0, (int)n_w, (int)first, statistic, beta, mu};
      TQ_TRY(launch_invert_warp<true>(ctx, d_g, nk * n_w, n, F));
      KernelScope ks(ctx, "dyson_slab_sum");
      k_dyson_slab_sum<<<(unsigned)((n_elem + 127) / 128), 128, 0, ctx->stream>>>(d_g, nk, n_elem, (const double *)d_wts, k0, uniform_weight, k0 > 0,
                                                                                 (cd *)d_out);
      tq_count_launch(ctx);
      TQ_CUDA(ctx, cudaGetLastError());
    }
    if (all_reduce) TQ_TRY(tq_all_reduce_sum(ctx, (double *)d_out, (size_t)2 * n_elem)); // sumk_discrete.py:163
    TQ_TRY(tq_stage_out_finish(ctx, out, sizeof(cd) * (size_t)n_w * nn, d_out, out_host));
    return TQ_OK;
  }
  const int w_tiles = (int)((n_w + 127) / 128);
  // enough CTAs for ~4 waves of 4 CTAs/SM, but at least 64 k-points per chunk
  long want_chunks = std::max<long>(1, (4L * 4 * ctx->sm_count + w_tiles - 1) / w_tiles);
  long k_per_chunk = std::max<long>(64, (n_k + want_chunks - 1) / want_chunks);
  int n_chunks = (int)std::max<long>(1, (n_k + k_per_chunk - 1) / k_per_chunk);
  if (n_chunks > 65535) {
    n_chunks = 65535;
    k_per_chunk = (n_k + n_chunks - 1) / n_chunks;
    n_chunks = (int)((n_k + k_per_chunk - 1) / k_per_chunk);
  }
  TQ_TRY(tq_reserve(ctx, ctx->scratch, sizeof(cd) * (size_t)n_chunks * n_w * nn));
  cd *d_part = (cd *)ctx->scratch.p;
  dim3 grid(w_tiles, n_chunks);
#define TQ_DYSON_CASE(NV)                                                                                                        \
  case NV: {                                                                                                                     \
    const size_t smem = sizeof(double) * 2 * NV * NV * 128;                                                                      \
    auto k = k_dyson_partial<NV>;                                                                                                \
    TQ_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                               \
    k<<<grid, 128, smem, ctx->stream>>>((const cd *)d_eps, (const double *)d_wts, uniform_weight, (const cd *)d_sig, (int)n_w, n_k,     \
                                        k_per_chunk, beta, statistic, (int)first, mu, d_part);                                   \
  } break;
  {
  KernelScope ks(ctx, "dyson_partial");
  switch (n) {
    TQ_DYSON_CASE(1)
    TQ_DYSON_CASE(2)
    TQ_DYSON_CASE(3)
    TQ_DYSON_CASE(4)
    TQ_DYSON_CASE(5)
    TQ_DYSON_CASE(6)
    TQ_DYSON_CASE(7)
    TQ_DYSON_CASE(8)
  }
  }
#undef TQ_DYSON_CASE
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  const long n_elem = n_w * nn;
  {
    KernelScope ks(ctx, "dyson_reduce");
    k_dyson_reduce<<<(unsigned)((n_elem + 255) / 256), 256, 0, ctx->stream>>>(d_part, n_chunks, n_elem, (cd *)d_out);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  if (all_reduce) TQ_TRY(tq_all_reduce_sum(ctx, (double *)d_out, (size_t)2 * n_elem)); // sumk_discrete.py:163
  TQ_TRY(tq_stage_out_finish(ctx, out, sizeof(cd) * (size_t)n_w * nn, d_out, out_host));
  return TQ_OK;
}
