// retime <-> refreq transforms.
//
// Restates c++/triqs/gfs/transform/fourier_real.cpp:35-78 (direct) and :82-131 (inverse).  Both have the form
//     x_r = (in_r - a1 h1_r - a2 h2_r) pin_r   ->   unnormalised DFT of length L   ->   out_k = pout_k X_k + a1 g1_k + a2 g2_k
// with per-column pole weights a1,2 = (m1 +- i m2 / a) / 2  (:66, :114), a = delta_w sqrt(L) (:60, :106), and per-row factors
// that depend only on the two meshes:
//     direct  : h = th_expo(t, a), th_expo_neg(t, a);  pin = e^{i t w_min};   g = 1/(w + i a), 1/(w - i a);  pout = delta_t e^{i (w - w_min) t_min}
//     inverse : h = 1/(w + i a), 1/(w - i a);  pin = e^{-i w t_min};  g = th_expo, th_expo_neg;  pout = e^{i w_min (t_min - t)} / (delta_t L)
// The row factors are tabulated once per mesh pair (host, cached in the context) and the wrap is fused into the load and the
// store of the tile FFT (lattice.cu, plain_tile_kernel<.., WRAP = true>): G is read once and written once.
#include "tq_common.cuh"

#include <cmath>
#include <map>
#include <memory>
#include <tuple>
#include <vector>

tq_status tq_wrapped_fft_axis(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, const cd *rin, const cd *rout,
                              const cd *coef);

struct RealPlanCache {
  struct Tables {
    cd *rin = nullptr, *rout = nullptr;
  };
  std::map<std::tuple<double, double, double, double, long, int>, Tables> tabs;
  ~RealPlanCache() {
    for (auto &kv : tabs) {
      cudaFree(kv.second.rin);
      cudaFree(kv.second.rout);
    }
  }
};

// mesh/bases/linear.hpp:154-160
static double linear_value(double xmin, double xmax, long n, long i) {
  if (n == 1) return xmin;
  const double wr = (double)i / (double)(n - 1);
  return xmin * (1 - wr) + xmax * wr;
}
static cd th_expo(double t, double a) { // fourier_real.cpp:27
  return t > 0 ? cmk(0.0, -std::exp(-a * t)) : (t < 0 ? cmk(0.0, 0.0) : cmk(0.0, -0.5 * std::exp(-a * t)));
}
static cd th_expo_neg(double t, double a) { // :28
  return t < 0 ? cmk(0.0, std::exp(a * t)) : (t > 0 ? cmk(0.0, 0.0) : cmk(0.0, 0.5 * std::exp(a * t)));
}
static cd inv_c(double re, double im) { // 1 / (re + i im)
  const double d = re * re + im * im;
  return cmk(re / d, -im / d);
}
static cd expi_d(double x) { return cmk(std::cos(x), std::sin(x)); }

// pole weights a1, a2 per column and max |m0|    (:45-47, :66)
__global__ void k_real_coefs(const cd *__restrict__ mom, int n_mom, long n_cols, long total, double a, cd *__restrict__ coef,
                             unsigned long long *max_m0_bits) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long gf = i / n_cols, c = i - gf * n_cols;
  const cd *m = mom + (size_t)gf * n_mom * n_cols + c;
  const cd m0 = m[0], m1 = m[n_cols], m2 = m[2 * n_cols];
  const cd im2 = cmk(-m2.y / a, m2.x / a); // i m2 / a
  coef[2 * i] = cscale(0.5, cadd(m1, im2));
  coef[2 * i + 1] = cscale(0.5, csub(m1, im2));
  const double a0 = hypot(m0.x, m0.y);
  atomicMax(max_m0_bits, (unsigned long long)__double_as_longlong(a0)); // non-negative doubles order like their bit patterns
}

static tq_status real_transform(tq_ctx *ctx, bool direct, double t_min, double t_max, double w_min, double w_max, long L, long n_others, long batch,
                                const tq_complex *in, const tq_complex *moments, int n_moments, tq_complex *out) {
  const char *name = direct ? "tq_fourier_t_to_w" : "tq_fourier_w_to_t";
  if (!ctx) return TQ_ERR_INVALID;
  if (!in || !out || L < 2 || n_others < 1 || batch < 1) return tq_fail(ctx, TQ_ERR_INVALID, std::string(name) + ": invalid argument");
  const bool have_moments = moments != nullptr && n_moments > 0;
  if (have_moments && n_moments < 3) // :45, :92
    return tq_fail(ctx, TQ_ERR_RUNTIME, direct ? " Direct RealTime Fourier transform requires known moments up to order 2."
                                               : " Direct Real-Frequency Fourier transform requires known moments up to order 2.");
  const double dt = (t_max - t_min) / (double)(L - 1), dw = (w_max - w_min) / (double)(L - 1);
  if (std::abs(dt * dw * (double)L / (2 * M_PI) - 1) > 1.e-10) return tq_fail(ctx, TQ_ERR_RUNTIME, "Meshes are not compatible"); // :50-51, :100-101
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const double a = dw * std::sqrt((double)L);

  if (!ctx->real_cache) ctx->real_cache = std::make_shared<RealPlanCache>();
  RealPlanCache::Tables &tb = ctx->real_cache->tabs[std::make_tuple(t_min, t_max, w_min, w_max, L, direct ? 1 : 0)];
  if (!tb.rin) {
    std::vector<cd> rin(3 * (size_t)L), rout(3 * (size_t)L);
    const double corr = 1.0 / (dt * (double)L);
    for (long i = 0; i < L; ++i) {
      const double t = linear_value(t_min, t_max, L, i), w = linear_value(w_min, w_max, L, i);
      cd *tt = direct ? &rin[3 * i] : &rout[3 * i], *ww = direct ? &rout[3 * i] : &rin[3 * i];
      tt[0] = th_expo(t, a);
      tt[1] = th_expo_neg(t, a);
      ww[0] = inv_c(w, a);
      ww[1] = inv_c(w, -a);
      if (direct) {
        tt[2] = expi_d(t * w_min);                          // :68
        ww[2] = cscale(dt, expi_d((w - w_min) * t_min));    // :75
      } else {
        ww[2] = expi_d(-w * t_min);                         // :116
        tt[2] = cscale(corr, expi_d(w_min * (t_min - t))); // :123-124
      }
    }
    TQ_CUDA(ctx, cudaMalloc(&tb.rin, sizeof(cd) * 3 * L));
    TQ_CUDA(ctx, cudaMalloc(&tb.rout, sizeof(cd) * 3 * L));
    TQ_CUDA(ctx, cudaMemcpyAsync(tb.rin, rin.data(), sizeof(cd) * 3 * L, cudaMemcpyHostToDevice, ctx->stream));
    TQ_CUDA(ctx, cudaMemcpyAsync(tb.rout, rout.data(), sizeof(cd) * 3 * L, cudaMemcpyHostToDevice, ctx->stream));
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }

  const size_t bytes = sizeof(cd) * (size_t)batch * L * n_others;
  const void *d_in = nullptr, *d_known = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, in, bytes, ctx->stage_in, &d_in));
  if (have_moments) TQ_TRY(tq_stage_in(ctx, moments, sizeof(cd) * (size_t)batch * n_moments * n_others, ctx->stage_in2, &d_known));
  TQ_TRY(tq_stage_out_begin(ctx, out, bytes, ctx->stage_out, &d_out, &out_host));
  const long total = batch * n_others;
  const size_t st_off = (sizeof(cd) * 2 * (size_t)total + 255) & ~size_t(255);
  TQ_TRY(tq_reserve(ctx, ctx->small, st_off + 8));
  cd *d_coef = (cd *)ctx->small.p;
  unsigned long long *d_st = (unsigned long long *)((char *)ctx->small.p + st_off);
  TQ_CUDA(ctx, cudaMemsetAsync(d_coef, 0, st_off + 8, ctx->stream)); // no moments: zero tail, the raw data is transformed (:41-43)
  if (have_moments) {
    KernelScope ks(ctx, "real_coefs");
    k_real_coefs<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_known, n_moments, n_others, total, a, d_coef, d_st);
    tq_count_launch(ctx);
    TQ_CUDA(ctx, cudaGetLastError());
  }
  // direct: FFTW_BACKWARD (:71);  inverse: FFTW_FORWARD (:119)
  TQ_TRY(tq_wrapped_fft_axis(ctx, batch, L, n_others, (const cd *)d_in, (cd *)d_out, direct ? 1 : -1, tb.rin, tb.rout, d_coef));
  if (have_moments) {
    TQ_TRY(tq_reserve_pinned(ctx, 8));
    TQ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, d_st, 8, cudaMemcpyDeviceToHost, ctx->stream));
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double a0;
    memcpy(&a0, ctx->pinned, sizeof(double));
    if (!(a0 < 1e-8)) // :46-47, :94-95
      return tq_fail(ctx, TQ_ERR_RUNTIME,
                     std::string(direct ? "ERROR: Direct" : "ERROR: Inverse") + " Fourier implementation requires vanishing 0th moment\n  error is :" +
                        std::to_string(a0));
  }
  TQ_TRY(tq_stage_out_finish(ctx, out, bytes, d_out, out_host));
  return TQ_OK;
}

extern "C" tq_status tq_fourier_t_to_w(tq_ctx *ctx, double t_min, double t_max, double w_min, double w_max, long n_pts, long n_others, long batch,
                                       const tq_complex *gt, const tq_complex *moments, int n_moments, tq_complex *gw) {
  return real_transform(ctx, true, t_min, t_max, w_min, w_max, n_pts, n_others, batch, gt, moments, n_moments, gw);
}

extern "C" tq_status tq_fourier_w_to_t(tq_ctx *ctx, double t_min, double t_max, double w_min, double w_max, long n_pts, long n_others, long batch,
                                       const tq_complex *gw, const tq_complex *moments, int n_moments, tq_complex *gt) {
  return real_transform(ctx, false, t_min, t_max, w_min, w_max, n_pts, n_others, batch, gw, moments, n_moments, gt);
}
