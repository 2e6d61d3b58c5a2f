// Legendre <-> Matsubara transforms.
//
// Restates c++/triqs/gfs/transform/legendre_matsubara.hpp:37-129 and c++/triqs/utility/legendre.{hpp,cpp}:
//   legendre -> imfreq : G(i w_n) = sum_l T_{nl} G_l,   T_{nl} = sqrt(2l+1) e^{i (n + 1/2) pi} i^l j_l((n + 1/2) pi), T_{-n-1,l} = conj T_{nl}
//                        (legendre.cpp:24-42; sqrt(pi/2x) J_{l+1/2}(x) = j_l(x) absorbs the 1/sqrt(2n+1))
//   legendre -> imtime : G(tau) = sum_l sqrt(2l+1)/beta P_l(2 tau/beta - 1) G_l                    (:56-70)
//   imtime -> legendre : G_l = delta sum_t c_t sqrt(2l+1) P_l(x_t) G(tau_t), c = 1/2 at both ends  (:97-118)
//   imfreq -> legendre : through G(tau) on 50000 points                                            (:76-93)
// P_l by the three-term recurrence of legendre_generator (legendre.hpp:39-58), evaluated in the same order.
// All three kernels are skinny matrix products with the matrix generated on the fly; the reduction over tau is
// chunked with a fixed-order second stage (deterministic).
#include "tq_common.cuh"

#include <cmath>
#include <map>
#include <memory>
#include <vector>

struct LegendrePlanCache {
  std::map<std::tuple<long, int, int>, cd *> T; // (n_iw, stat, n_l) -> device [n_w][n_l]
  ~LegendrePlanCache() {
    for (auto &kv : T) cudaFree(kv.second);
  }
};

// gw[b][w][c] = sum_l T[w][l] gl[b][l][c]
__global__ void k_leg_to_iw(const cd *__restrict__ T, const cd *__restrict__ gl, long n_w, int n_l, long n_cols, long total, cd *__restrict__ gw) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long c = i % n_cols, w = (i / n_cols) % n_w, b = i / (n_cols * n_w);
  const cd *g = gl + (size_t)b * n_l * n_cols + c;
  const cd *t = T + (size_t)w * n_l;
  cd acc = cmk(0.0, 0.0);
  for (int l = 0; l < n_l; ++l) acc = cfma(__ldg(t + l), __ldg(g + (size_t)l * n_cols), acc);
  gw[i] = acc;
}

// legendre_generator::next (legendre.hpp:44-55)
struct LegGen {
  double x, p0, p1;
  int n;
  TQ_D void reset(double xx) {
    x = xx;
    n = 0;
    p0 = 1.0;
    p1 = xx;
  }
  TQ_D double next() {
    double r;
    if (n > 1) {
      r = ((2 * n - 1) * x * p1 - (n - 1) * p0) / n;
      p0 = p1;
      p1 = r;
    } else
      r = n == 0 ? p0 : p1;
    ++n;
    return r;
  }
};

// gt[b][t][c] = sum_l sqrt(2l+1)/beta P_l(x_t) gl[b][l][c]
__global__ void k_leg_to_tau(const cd *__restrict__ gl, long n_tau, int n_l, long n_cols, long total, double beta, cd *__restrict__ gt) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long c = i % n_cols, t = (i / n_cols) % n_tau, b = i / (n_cols * n_tau);
  const double wr = (double)t / (double)(n_tau - 1);
  const double tau = 0.0 * (1 - wr) + beta * wr; // linear.hpp:154-160
  LegGen L;
  L.reset(2 * tau / beta - 1);
  const cd *g = gl + (size_t)b * n_l * n_cols + c;
  cd acc = cmk(0.0, 0.0);
  for (int l = 0; l < n_l; ++l) {
    const double f = sqrt((double)(2 * l + 1)) / beta * L.next();
    acc = caxpy(f, __ldg(g + (size_t)l * n_cols), acc);
  }
  gt[i] = acc;
}

// partial[b][chunk][l][c] = sum_{t in chunk} c_t sqrt(2l+1) P_l(x_t) gt[b][t][c]
constexpr int LEG_ROWS = 64; // tau rows per CTA
constexpr int LEG_LMAX = 128;
__global__ void __launch_bounds__(256) k_tau_to_leg_partial(const cd *__restrict__ gt, long n_tau, int n_l, long n_cols, int n_chunks, double beta,
                                                            cd *__restrict__ partial) {
  extern __shared__ double P[]; // [n_l][LEG_ROWS]
  const long b = blockIdx.x / n_chunks;
  const int chunk = blockIdx.x - (int)(b * n_chunks);
  const long t0 = (long)chunk * LEG_ROWS;
  const int rows = (int)min((long)LEG_ROWS, n_tau - t0);
  for (int r = threadIdx.x; r < LEG_ROWS; r += blockDim.x) {
    const long t = t0 + r;
    double coef = 0.0, x = 0.0;
    if (r < rows) {
      const double wr = (double)t / (double)(n_tau - 1);
      const double tau = 0.0 * (1 - wr) + beta * wr;
      x = 2 * tau / beta - 1;
      coef = (t == 0 || t == n_tau - 1) ? 0.5 : 1.0; // :108-111
    }
    LegGen L;
    L.reset(x);
    for (int l = 0; l < n_l; ++l) P[l * LEG_ROWS + r] = coef * sqrt((double)(2 * l + 1)) * L.next();
  }
  __syncthreads();
  const cd *g = gt + ((size_t)b * n_tau + t0) * n_cols;
  cd *out = partial + ((size_t)b * n_chunks + chunk) * n_l * n_cols;
  for (long e = threadIdx.x; e < (long)n_l * n_cols; e += blockDim.x) {
    const int l = (int)(e / n_cols);
    const long c = e - (long)l * n_cols;
    cd acc = cmk(0.0, 0.0);
    for (int r = 0; r < rows; ++r) acc = caxpy(P[l * LEG_ROWS + r], __ldg(g + (size_t)r * n_cols + c), acc);
    out[e] = acc;
  }
}
__global__ void k_tau_to_leg_final(const cd *__restrict__ partial, int n_chunks, long n_lc, long total, double delta, cd *__restrict__ gl) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long b = i / n_lc, e = i - b * n_lc;
  cd acc = cmk(0.0, 0.0);
  for (int k = 0; k < n_chunks; ++k) acc = cadd(acc, partial[((size_t)b * n_chunks + k) * n_lc + e]);
  gl[i] = cscale(delta, acc); // :117
}

static tq_status leg_check(tq_ctx *ctx, const char *name, const void *a, const void *b, long n_l, long n_pts, long n_others, long batch) {
  if (!a || !b || n_l < 1 || n_pts < 1 || n_others < 1 || batch < 1) return tq_fail(ctx, TQ_ERR_INVALID, std::string(name) + ": invalid argument");
  if (n_l > LEG_LMAX) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, std::string(name) + ": more than 128 Legendre coefficients");
  return TQ_OK;
}

extern "C" tq_status tq_legendre_to_imfreq(tq_ctx *ctx, int statistic, long n_l, long n_iw, long n_others, long batch, const tq_complex *gl,
                                           tq_complex *gw) {
  if (!ctx) return TQ_ERR_INVALID;
  TQ_TRY(leg_check(ctx, "tq_legendre_to_imfreq", gl, gw, n_l, n_iw, n_others, batch));
  if (statistic != TQ_FERMION && statistic != TQ_BOSON) return tq_fail(ctx, TQ_ERR_INVALID, "tq_legendre_to_imfreq: invalid statistic");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long last = n_iw - 1, first = -(last + (statistic == TQ_FERMION ? 1 : 0)), n_w = last - first + 1; // imfreq.hpp:59-60
  if (!ctx->leg_cache) ctx->leg_cache = std::make_shared<LegendrePlanCache>();
  cd *&T = ctx->leg_cache->T[std::make_tuple(n_iw, statistic, (int)n_l)];
  if (!T) {
    std::vector<cd> h((size_t)n_w * n_l);
    for (long w = 0; w < n_w; ++w) {
      long n = first + w; // legendre_T(om.index(), l): the reference uses the Matsubara index whatever the statistic (:47-48)
      const bool neg = n < 0;
      if (neg) n = std::labs(n + 1);
      const double x = ((double)n + 0.5) * M_PI;
      const double ph_re = std::cos(x), ph_im = std::sin(x); // e^{i (n + 1/2) pi}
      for (long l = 0; l < n_l; ++l) {
        const double f = std::sqrt((double)(2 * l + 1)) * std::sph_bessel((unsigned)l, x);
        cd il = (l & 3) == 0 ? cmk(1, 0) : (l & 3) == 1 ? cmk(0, 1) : (l & 3) == 2 ? cmk(-1, 0) : cmk(0, -1); // i^l
        cd r = cscale(f, cmul(cmk(ph_re, ph_im), il));
        h[(size_t)w * n_l + l] = neg ? cconj(r) : r;
      }
    }
    TQ_CUDA(ctx, cudaMalloc(&T, sizeof(cd) * h.size()));
    TQ_CUDA(ctx, cudaMemcpyAsync(T, h.data(), sizeof(cd) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  const size_t in_bytes = sizeof(cd) * (size_t)batch * n_l * n_others, out_bytes = sizeof(cd) * (size_t)batch * n_w * n_others;
  const void *d_in = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, gl, in_bytes, ctx->stage_in, &d_in));
  TQ_TRY(tq_stage_out_begin(ctx, gw, out_bytes, ctx->stage_out, &d_out, &out_host));
  const long total = batch * n_w * n_others;
  {
    KernelScope ks(ctx, "legendre_to_imfreq");
    k_leg_to_iw<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(T, (const cd *)d_in, n_w, (int)n_l, n_others, total, (cd *)d_out);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(tq_stage_out_finish(ctx, gw, out_bytes, d_out, out_host));
  return TQ_OK;
}

extern "C" tq_status tq_legendre_to_imtime(tq_ctx *ctx, double beta, long n_l, long n_tau, long n_others, long batch, const tq_complex *gl,
                                           tq_complex *gt) {
  if (!ctx) return TQ_ERR_INVALID;
  TQ_TRY(leg_check(ctx, "tq_legendre_to_imtime", gl, gt, n_l, n_tau, n_others, batch));
  if (!(beta > 0) || n_tau < 2) return tq_fail(ctx, TQ_ERR_INVALID, "tq_legendre_to_imtime: invalid mesh");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t in_bytes = sizeof(cd) * (size_t)batch * n_l * n_others, out_bytes = sizeof(cd) * (size_t)batch * n_tau * n_others;
  const void *d_in = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, gl, in_bytes, ctx->stage_in, &d_in));
  TQ_TRY(tq_stage_out_begin(ctx, gt, out_bytes, ctx->stage_out, &d_out, &out_host));
  const long total = batch * n_tau * n_others;
  {
    KernelScope ks(ctx, "legendre_to_imtime");
    k_leg_to_tau<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_in, n_tau, (int)n_l, n_others, total, beta, (cd *)d_out);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(tq_stage_out_finish(ctx, gt, out_bytes, d_out, out_host));
  return TQ_OK;
}

// device-resident imtime -> legendre
static tq_status tau_to_leg_device(tq_ctx *ctx, double beta, long n_tau, long n_l, long n_others, long batch, const cd *d_gt, cd *d_gl) {
  const int n_chunks = (int)((n_tau + LEG_ROWS - 1) / LEG_ROWS);
  const size_t part_bytes = sizeof(cd) * (size_t)batch * n_chunks * n_l * n_others;
  TQ_TRY(tq_reserve(ctx, ctx->scratch, part_bytes));
  const size_t smem = sizeof(double) * (size_t)n_l * LEG_ROWS;
  TQ_CUDA(ctx, cudaFuncSetAttribute(k_tau_to_leg_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    KernelScope ks(ctx, "imtime_to_legendre");
    k_tau_to_leg_partial<<<(unsigned)(batch * n_chunks), 256, smem, ctx->stream>>>(d_gt, n_tau, (int)n_l, n_others, n_chunks, beta,
                                                                                    (cd *)ctx->scratch.p);
  }
  const long total = batch * n_l * n_others;
  {
    KernelScope ks(ctx, "imtime_to_legendre_final");
    k_tau_to_leg_final<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)ctx->scratch.p, n_chunks, n_l * n_others, total,
                                                                                  beta / (double)(n_tau - 1), d_gl);
  }
  tq_count_launch(ctx, 2);
  TQ_CUDA(ctx, cudaGetLastError());
  return TQ_OK;
}

extern "C" tq_status tq_imtime_to_legendre(tq_ctx *ctx, double beta, long n_tau, long n_l, long n_others, long batch, const tq_complex *gt,
                                           tq_complex *gl) {
  if (!ctx) return TQ_ERR_INVALID;
  TQ_TRY(leg_check(ctx, "tq_imtime_to_legendre", gt, gl, n_l, n_tau, n_others, batch));
  if (!(beta > 0) || n_tau < 2) return tq_fail(ctx, TQ_ERR_INVALID, "tq_imtime_to_legendre: invalid mesh");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t in_bytes = sizeof(cd) * (size_t)batch * n_tau * n_others, out_bytes = sizeof(cd) * (size_t)batch * n_l * n_others;
  const void *d_in = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, gt, in_bytes, ctx->stage_in, &d_in));
  TQ_TRY(tq_stage_out_begin(ctx, gl, out_bytes, ctx->stage_out, &d_out, &out_host));
  TQ_TRY(tau_to_leg_device(ctx, beta, n_tau, n_l, n_others, batch, (const cd *)d_in, (cd *)d_out));
  TQ_TRY(tq_stage_out_finish(ctx, gl, out_bytes, d_out, out_host));
  return TQ_OK;
}

extern "C" tq_status tq_fourier_iw_to_tau(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_tau, long n_others, long batch,
                                          const tq_complex *gw, const tq_complex *moments, int n_moments, tq_complex *gt, int *warn,
                                          double *tail_err);

extern "C" tq_status tq_imfreq_to_legendre(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_l, long n_others, long batch,
                                           const tq_complex *gw, tq_complex *gl) {
  if (!ctx) return TQ_ERR_INVALID;
  TQ_TRY(leg_check(ctx, "tq_imfreq_to_legendre", gw, gl, n_l, n_iw, n_others, batch));
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long Nt = 50000; // :85
  const size_t gt_bytes = sizeof(cd) * (size_t)batch * Nt * n_others, out_bytes = sizeof(cd) * (size_t)batch * n_l * n_others;
  TQ_TRY(tq_reserve(ctx, ctx->leg_tmp, gt_bytes));
  // gt() = fourier(gw): the tail is fitted by the transform itself (:89); gw may be a host pointer, the transform stages it
  std::vector<int> warn((size_t)batch);
  std::vector<double> err((size_t)batch);
  TQ_TRY(tq_fourier_iw_to_tau(ctx, beta, statistic, n_iw, Nt, n_others, batch, gw, nullptr, 0, (tq_complex *)ctx->leg_tmp.p, warn.data(), err.data()));
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_out_begin(ctx, gl, out_bytes, ctx->stage_out, &d_out, &out_host));
  TQ_TRY(tau_to_leg_device(ctx, beta, Nt, n_l, n_others, batch, (const cd *)ctx->leg_tmp.p, (cd *)d_out));
  TQ_TRY(tq_stage_out_finish(ctx, gl, out_bytes, d_out, out_host));
  return TQ_OK;
}
