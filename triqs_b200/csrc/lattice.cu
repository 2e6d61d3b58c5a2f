// Plain batched multi-dimensional DFT (the _fourier_base seam) and the cyclat <-> brzone transforms.
//
// Restates c++/triqs/gfs/transform/fourier_common.cpp:25-45 (fftw_plan_many_dft, stride = howmany, dist = 1,
// unnormalised, BACKWARD = e^{+i...}) and c++/triqs/gfs/transform/fourier_lattice.cpp:30-59.
// The data is [prod(dims)][howmany] with the batch index innermost, so a transform along mesh axis d sees
// the array as [outer][N_d][inner] with `inner` contiguous: a CTA stages a tile [N_d][TC] of TC adjacent
// inner columns in shared memory (coalesced TC*16-byte runs), transforms it in place and writes it back.
#include "tq_fft.cuh"

#include <algorithm>
#include <map>
#include <memory>
#include <vector>

struct PlainArgs {
  FftPlan P;
  const cd *in;
  cd *out;
  long inner;     // contiguous elements per (outer, n)
  int TC, tiles;  // tiles = ceil(inner / TC)
  int direct_tw, smem_tw;
  double scale;   // applied on store (1 = none)
  // WRAP: x_r = (in_r - a1 h1_r - a2 h2_r) pin_r  ->  DFT  ->  out_k = pout_k X_k + a1 g1_k + a2 g2_k   (fourier_real.cu)
  const cd *rin;  // [N][3] = h1, h2, pin
  const cd *rout; // [N][3] = g1, g2, pout
  const cd *coef; // [outer][inner][2] = a1, a2
};

template <int SGN, int NT, int MINB, bool WRAP = false> __global__ void __launch_bounds__(NT, MINB) plain_tile_kernel(const PlainArgs A) {
  extern __shared__ double2 smem[];
  const long o = blockIdx.x / A.tiles;
  const int tile_id = blockIdx.x - (int)(o * A.tiles);
  const long c0 = (long)tile_id * A.TC;
  const int tc = (int)min((long)A.TC, A.inner - c0);
  const int N = A.P.N;
  cd *tile = smem;
  cd *twS = tile + (size_t)N * A.TC;                       // [N] twiddles (if A.smem_tw)
  int *posS = reinterpret_cast<int *>(twS + (A.smem_tw ? N : 0)); // [N] output positions (if A.smem_tw)
  const cd *src = A.in + (size_t)o * N * A.inner + c0;
  cd *dst = A.out + (size_t)o * N * A.inner + c0;
  const TileThread T = tile_thread(threadIdx.x, NT, tc);
  if (A.smem_tw)
    for (int i = threadIdx.x; i < N; i += NT) {
      twS[i] = __ldg(&A.P.tw[i]);
      posS[i] = __ldg(&A.P.pos[i]);
    }
  constexpr int LU = 8; // independent loads in flight per thread
  if (T.nbt > 0) {
    for (int c = T.ct; c < tc; c += T.nct) {
      const cd *gp = src + c;
      cd *tp = tile + c;
      cd a1 = cmk(0.0, 0.0), a2 = a1;
      if (WRAP) {
        a1 = __ldg(&A.coef[2 * ((size_t)o * A.inner + c0 + c)]);
        a2 = __ldg(&A.coef[2 * ((size_t)o * A.inner + c0 + c) + 1]);
      }
      for (int r0 = T.bt; r0 < N; r0 += T.nbt * LU) {
        cd g[LU];
#pragma unroll
        for (int u = 0; u < LU; ++u) {
          const int r = r0 + u * T.nbt;
          if (r < N) {
            g[u] = gp[(size_t)r * A.inner];
            if (WRAP) {
              const cd *e = A.rin + 3 * (size_t)r;
              g[u] = cmul(csub(csub(g[u], cmul(a1, __ldg(e))), cmul(a2, __ldg(e + 1))), __ldg(e + 2));
            }
          }
        }
#pragma unroll
        for (int u = 0; u < LU; ++u) {
          const int r = r0 + u * T.nbt;
          if (r < N) tp[r * tc] = g[u];
        }
      }
    }
  }
  __syncthreads();
  fft_tile<SGN, false>(A.P, tile, tc, tc, A.smem_tw ? twS : A.P.tw, A.direct_tw != 0);
  const bool scaled = A.scale != 1.0;
  if (T.nbt > 0) {
    for (int c = T.ct; c < tc; c += T.nct) {
      cd a1 = cmk(0.0, 0.0), a2 = a1;
      if (WRAP) {
        a1 = __ldg(&A.coef[2 * ((size_t)o * A.inner + c0 + c)]);
        a2 = __ldg(&A.coef[2 * ((size_t)o * A.inner + c0 + c) + 1]);
      }
      for (int k = T.bt; k < N; k += T.nbt) {
        const int pos = A.smem_tw ? posS[k] : __ldg(&A.P.pos[k]);
        cd v = tile[pos * tc + c];
        if (scaled) v = cscale(A.scale, v);
        if (WRAP) {
          const cd *e = A.rout + 3 * (size_t)k;
          v = cadd(cadd(cmul(__ldg(e + 2), v), cmul(a1, __ldg(e))), cmul(a2, __ldg(e + 1)));
        }
        dst[(size_t)k * A.inner + c] = v;
      }
    }
  }
}

// Direct O(N^2) evaluation for axis lengths the tile FFT cannot take (a prime factor > 128, or longer than one SM's shared
// memory): FFTW accepts any length (fourier_common.cpp:30), so these must work, if slowly.  One thread per output row k and
// DIRECT_CW columns; the input rows are warp-uniform (broadcast) loads, w^{nk} comes from the N-entry table by the running
// index (n k) mod N.  The affine wrap of fourier_real.cu is applied on the fly.
constexpr int DIRECT_CW = 8;
template <bool WRAP> __global__ void __launch_bounds__(128) k_direct_axis(const cd *__restrict__ in, cd *__restrict__ out, const cd *__restrict__ w,
                                                                         int N, long inner, int sign, double scale, const cd *__restrict__ rin,
                                                                         const cd *__restrict__ rout, const cd *__restrict__ coef) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const long c0 = (long)blockIdx.y * DIRECT_CW;
  const long o = blockIdx.z;
  const int wd = (int)min((long)DIRECT_CW, inner - c0);
  const cd *src = in + (size_t)o * N * inner + c0;
  cd a1[DIRECT_CW], a2[DIRECT_CW], acc[DIRECT_CW];
#pragma unroll
  for (int c = 0; c < DIRECT_CW; ++c) {
    acc[c] = cmk(0.0, 0.0);
    a1[c] = a2[c] = cmk(0.0, 0.0);
    if (WRAP && c < wd) {
      a1[c] = __ldg(&coef[2 * ((size_t)o * inner + c0 + c)]);
      a2[c] = __ldg(&coef[2 * ((size_t)o * inner + c0 + c) + 1]);
    }
  }
  const int kk = k < N ? k : 0;
  int m = 0;
  for (int n = 0; n < N; ++n) {
    cd tw = __ldg(&w[m]);
    if (sign < 0) tw = cconj(tw);
    m += kk;
    if (m >= N) m -= N;
    cd h1 = cmk(0.0, 0.0), h2 = h1, pin = cmk(1.0, 0.0);
    if (WRAP) {
      h1 = __ldg(&rin[3 * (size_t)n]);
      h2 = __ldg(&rin[3 * (size_t)n + 1]);
      pin = __ldg(&rin[3 * (size_t)n + 2]);
    }
#pragma unroll
    for (int c = 0; c < DIRECT_CW; ++c)
      if (c < wd) {
        cd x = __ldg(src + (size_t)n * inner + c);
        if (WRAP) x = cmul(csub(csub(x, cmul(a1[c], h1)), cmul(a2[c], h2)), pin);
        acc[c] = cfma(tw, x, acc[c]);
      }
  }
  if (k >= N) return;
  cd *dst = out + (size_t)o * N * inner + (size_t)k * inner + c0;
  for (int c = 0; c < wd; ++c) {
    cd v = acc[c];
    if (scale != 1.0) v = cscale(scale, v);
    if (WRAP) {
      const cd *e = rout + 3 * (size_t)k;
      v = cadd(cadd(cmul(__ldg(e + 2), v), cmul(a1[c], __ldg(e))), cmul(a2[c], __ldg(e + 1)));
    }
    dst[c] = v;
  }
}

struct DirectTwiddles {
  std::map<long, cd *> w;
  ~DirectTwiddles() {
    for (auto &kv : w) cudaFree(kv.second);
  }
};

static tq_status launch_direct(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, double scale, const cd *rin,
                               const cd *rout, const cd *coef) {
  if (N > 65536) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tq_fft_many: axis length " + std::to_string(N) + " is neither factorisable nor <= 65536");
  if (outer > 65535) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tq_fft_many: too many outer slices for the direct transform");
  if (!ctx->direct_tw) ctx->direct_tw = std::make_shared<DirectTwiddles>();
  cd *&w = ctx->direct_tw->w[N];
  if (!w) {
    std::vector<cd> h;
    tq_fft_twiddles_host((int)N, h);
    TQ_CUDA(ctx, cudaMalloc(&w, sizeof(cd) * N));
    TQ_CUDA(ctx, cudaMemcpyAsync(w, h.data(), sizeof(cd) * N, cudaMemcpyHostToDevice, ctx->stream));
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  cd *dst = out;
  const size_t bytes = sizeof(cd) * (size_t)outer * N * inner;
  if (in == out) { // every output needs every input: go through a temporary
    TQ_TRY(tq_reserve(ctx, ctx->scratch, bytes));
    dst = (cd *)ctx->scratch.p;
  }
  const dim3 grid((unsigned)((N + 127) / 128), (unsigned)((inner + DIRECT_CW - 1) / DIRECT_CW), (unsigned)outer);
  if (grid.y > 65535) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tq_fft_many: too many columns for the direct transform");
  {
    KernelScope ks(ctx, "fft_axis_direct");
    if (rin)
      k_direct_axis<true><<<grid, 128, 0, ctx->stream>>>(in, dst, w, (int)N, inner, sign, scale, rin, rout, coef);
    else
      k_direct_axis<false><<<grid, 128, 0, ctx->stream>>>(in, dst, w, (int)N, inner, sign, scale, nullptr, nullptr, nullptr);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  if (dst != out) TQ_CUDA(ctx, cudaMemcpyAsync(out, dst, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return TQ_OK;
}

// fourier_matsubara.cu: Bluestein convolution for axis lengths without a plan
bool tq_bluestein_axis_possible(tq_ctx *ctx, long N);
tq_status tq_bluestein_plain_axis(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, double scale, const cd *rin,
                                  const cd *rout, const cd *coef);
// lattice_pipe.cu: TMA-pipelined fast path for axis lengths with a compile-time two-pass plan
tq_status tq_lattice_pipe_axis(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, double scale, bool *handled);

static tq_status launch_plain(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, double scale, const cd *rin = nullptr,
                              const cd *rout = nullptr, const cd *coef = nullptr) {
  const bool wrap = rin != nullptr;
  if (!wrap) {
    bool handled = false;
    TQ_TRY(tq_lattice_pipe_axis(ctx, outer, N, inner, in, out, sign, scale, &handled));
    if (handled) return TQ_OK;
  }
  const size_t budget = (size_t)ctx->max_smem_optin;
  {
    FftPlan probe;
    if (!tq_fft_factorize((int)N, probe) || sizeof(cd) * (size_t)N > budget) {
      // a prime factor > 127 (or a column longer than one SM's shared memory): Bluestein convolution through the pipelined axis kernel
      // (fourier_matsubara.cu; the affine wrap of the real-time transforms rides on its load / store kernels), the direct O(N^2) sum for short axes
      if (tq_bluestein_axis_possible(ctx, N)) return tq_bluestein_plain_axis(ctx, outer, N, inner, in, out, sign, scale, rin, rout, coef);
      return launch_direct(ctx, outer, N, inner, in, out, sign, scale, rin, rout, coef);
    }
  }
  std::shared_ptr<FftPlanHost> fp;
  TQ_TRY(tq_get_fft_plan(ctx, N, &fp));
  size_t half = budget / 2 - 1024;
  if (const char *e = TQ_DEV_ENV("TQ_TILE_BYTES")) half = std::min<size_t>((size_t)atol(e), budget);
  if (sizeof(cd) * (size_t)N > budget)
    return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tq_fft_many: axis length " + std::to_string(N) + " does not fit one SM's shared memory");
  // per row: tc columns + (twiddle + position when tabulated in shared memory)
  auto fits = [&](long tc, size_t lim) { return sizeof(cd) * (size_t)N * tc + (tc >= 4 ? (sizeof(cd) + sizeof(int)) * (size_t)N : 0) <= lim; };
  long tc_half = (long)(half / (sizeof(cd) * N)), tc_full = (long)(budget / (sizeof(cd) * N));
  while (tc_half > 0 && !fits(tc_half, half)) --tc_half;
  while (tc_full > 1 && !fits(tc_full, budget)) --tc_full;
  long tcmax = tc_half >= 8 ? tc_half : tc_full;
  tcmax = std::max<long>(1, std::min(tcmax, inner));
  long tc = tcmax;
  // prefer a tile width that divides the contiguous run and keeps 64-byte granularity
  for (long t = tcmax; t >= std::max<long>(4, tcmax / 2); --t)
    if (inner % t == 0 && t % 4 == 0) { tc = t; break; }
  PlainArgs A{};
  A.P = fp->plan;
  A.in = in;
  A.out = out;
  A.inner = inner;
  A.TC = (int)tc;
  A.tiles = (int)((inner + tc - 1) / tc);
  A.direct_tw = A.smem_tw = tc >= 4;
  A.scale = scale;
  A.rin = rin;
  A.rout = rout;
  A.coef = coef;
  const size_t smem = sizeof(cd) * (size_t)N * tc + (A.smem_tw ? (sizeof(cd) + sizeof(int)) * (size_t)N : 0);
  const long nblocks = outer * A.tiles;
  if (nblocks > 0x7fffffffL) return tq_fail(ctx, TQ_ERR_INVALID, "tq_fft_many: too many tiles");
  const bool big = smem > budget / 2 - 1024;
#define TQ_LAUNCH_PLAIN(SGN, NT, MINB)                                                                                           \
  do {                                                                                                                           \
    auto k = wrap ? plain_tile_kernel<SGN, NT, MINB, true> : plain_tile_kernel<SGN, NT, MINB, false>;                            \
    TQ_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                               \
    k<<<(unsigned)nblocks, NT, smem, ctx->stream>>>(A);                                                                          \
  } while (0)
  KernelScope ks(ctx, wrap ? "fft_axis_wrapped" : "fft_axis");
  if (sign > 0) {
    if (big) TQ_LAUNCH_PLAIN(1, 512, 1); else TQ_LAUNCH_PLAIN(1, 256, 2);
  } else {
    if (big) TQ_LAUNCH_PLAIN(-1, 512, 1); else TQ_LAUNCH_PLAIN(-1, 256, 2);
  }
#undef TQ_LAUNCH_PLAIN
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  return TQ_OK;
}

// fourier_real.cu: one axis transform with the affine pre/post wrap fused into the load and the store of the tile
tq_status tq_wrapped_fft_axis(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, const cd *rin, const cd *rout,
                              const cd *coef) {
  return launch_plain(ctx, outer, N, inner, in, out, sign, 1.0, rin, rout, coef);
}

// fourier_matsubara.cu (Bluestein): one plain unnormalised axis transform of [outer][N][inner] on device pointers, in place allowed
tq_status tq_fft_axis_device(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign) {
  return launch_plain(ctx, outer, N, inner, in, out, sign, 1.0);
}

// ... the same with a pointwise product fused into the stores where the pipelined kernel takes the length:
//   out(o, k, i) *= tab[(o % omod) * os + k * rs + i / cdiv]  (conjugated if conj);  *fused = false: transformed, NOT multiplied
tq_status tq_lattice_pipe_axis_mul(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, const cd *tab, int omod, int os,
                                   int rs, int cdiv, int conj, bool *handled);
tq_status tq_fft_axis_mul_device(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, const cd *tab, int omod, int os, int rs,
                                 int cdiv, int conj, bool *fused) {
  TQ_TRY(tq_lattice_pipe_axis_mul(ctx, outer, N, inner, in, out, sign, tab, omod, os, rs, cdiv, conj, fused));
  if (*fused) return TQ_OK;
  return launch_plain(ctx, outer, N, inner, in, out, sign, 1.0);
}

static tq_status fft_many_device(tq_ctx *ctx, int rank, const long *dims, long howmany, const cd *d_in, cd *d_out, int sign, double scale) {
  long total = 1;
  for (int d = 0; d < rank; ++d) total *= dims[d];
  bool first = true;
  int last_axis = -1;
  for (int d = 0; d < rank; ++d)
    if (dims[d] > 1) last_axis = d;
  if (last_axis < 0) { // all lengths 1: identity (times scale)
    if (d_in != d_out) TQ_CUDA(ctx, cudaMemcpyAsync(d_out, d_in, sizeof(cd) * (size_t)total * howmany, cudaMemcpyDeviceToDevice, ctx->stream));
    return TQ_OK;
  }
  // innermost axis first (largest contiguous runs while the input is still cold)
  for (int d = rank - 1; d >= 0; --d) {
    if (dims[d] == 1) continue;
    long outer = 1, inner = howmany;
    for (int e = 0; e < d; ++e) outer *= dims[e];
    for (int e = d + 1; e < rank; ++e) inner *= dims[e];
    bool is_last = true;
    for (int e = d - 1; e >= 0; --e)
      if (dims[e] > 1) is_last = false;
    TQ_TRY(launch_plain(ctx, outer, dims[d], inner, first ? d_in : d_out, d_out, sign, is_last ? scale : 1.0));
    first = false;
  }
  return TQ_OK;
}

extern "C" tq_status tq_fft_many(tq_ctx *ctx, int rank, const int *dims, long howmany, const tq_complex *in, tq_complex *out, int sign) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!in || !out || rank < 1 || rank > 3 || !dims || howmany < 1 || (sign != 1 && sign != -1))
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_fft_many: invalid argument");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  long ld[3] = {1, 1, 1}, total = 1;
  for (int d = 0; d < rank; ++d) {
    if (dims[d] < 1) return tq_fail(ctx, TQ_ERR_INVALID, "tq_fft_many: dims must be >= 1");
    ld[d] = dims[d];
    total *= dims[d];
  }
  const size_t bytes = sizeof(cd) * (size_t)total * howmany;
  const void *d_in = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, in, bytes, ctx->stage_in, &d_in));
  TQ_TRY(tq_stage_out_begin(ctx, out, bytes, ctx->stage_out, &d_out, &out_host));
  TQ_TRY(fft_many_device(ctx, rank, ld, howmany, (const cd *)d_in, (cd *)d_out, sign, 1.0));
  TQ_TRY(tq_stage_out_finish(ctx, out, bytes, d_out, out_host));
  return TQ_OK;
}

extern "C" tq_status tq_fourier_lattice(tq_ctx *ctx, const long dims[3], int to_k, long n_others, const tq_complex *in, tq_complex *out) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!in || !out || !dims || n_others < 1 || dims[0] < 1 || dims[1] < 1 || dims[2] < 1)
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_fourier_lattice: invalid argument");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long nk = dims[0] * dims[1] * dims[2];
  const size_t bytes = sizeof(cd) * (size_t)nk * n_others;
  const void *d_in = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, in, bytes, ctx->stage_in, &d_in));
  TQ_TRY(tq_stage_out_begin(ctx, out, bytes, ctx->stage_out, &d_out, &out_host));
  // fourier_lattice.cpp:59 r -> k: FFTW_BACKWARD, no scale;  :51-55 k -> r: FFTW_FORWARD then /= N_k
  TQ_TRY(fft_many_device(ctx, 3, dims, n_others, (const cd *)d_in, (cd *)d_out, to_k ? 1 : -1, to_k ? 1.0 : 1.0 / (double)nk));
  TQ_TRY(tq_stage_out_finish(ctx, out, bytes, d_out, out_host));
  return TQ_OK;
}
