// Elementwise companions of the Matsubara path: make_hermitian, is_gf_hermitian, replace_by_tail.
//
// Restates c++/triqs/gfs/functions/imfreq.hpp:81-124 (is_gf_hermitian), :175-215 (make_hermitian), :258-261 (replace_by_tail)
// and tail_eval c++/triqs/mesh/tail_fitter.hpp:49-64.  Data is [batch][n_pts][d][d] (d = 1 scalar, d = n matrix, d = n1 n2 for a
// rank-4 target viewed as a (ij),(kl) matrix); on a full imfreq mesh the point -i w_n sits at the mirrored data index
// n_pts - 1 - i for both statistics (imfreq.hpp:59-60), on imtime it is the point itself.  All three are HBM streams.
#include "tq_common.cuh"

#include <cmath>

// out[p](I,K) = 0.5 (g[p](I,K) + conj(g[mirror(p)](K,I)))
__global__ void k_make_hermitian(const cd *__restrict__ g, cd *__restrict__ out, long n_pts, int d, long total, int mirror) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int dd = d * d;
  const long pt = i / dd; // global point index over the batch
  const int e = (int)(i - pt * dd);
  const int I = e / d, K = e - I * d;
  const long gf = pt / n_pts, p = pt - gf * n_pts;
  const long pm = mirror ? n_pts - 1 - p : p;
  const cd a = g[i], b = g[((gf * n_pts + pm) * dd) + K * d + I];
  out[i] = cmk(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
}

// flag[0] = 1 as soon as |conj(g[mirror(p)](K,I)) - g[p](I,K)| > tolerance somewhere
__global__ void k_is_hermitian(const cd *__restrict__ g, long n_pts, int d, long total, int mirror, double tol, int *flag) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int dd = d * d;
  const long pt = i / dd;
  const int e = (int)(i - pt * dd);
  const int I = e / d, K = e - I * d;
  const long gf = pt / n_pts, p = pt - gf * n_pts;
  const long pm = mirror ? n_pts - 1 - p : p;
  const cd a = g[i], b = g[((gf * n_pts + pm) * dd) + K * d + I];
  const double dx = b.x - a.x, dy = -b.y - a.y;
  if (!(hypot(dx, dy) <= tol)) atomicExch(flag, 1);
}

// g[iw] = sum_n tail[n] / (i w)^n  for iw.n >= n_min or iw.n < -n_min     (imfreq.hpp:258-261, tail_fitter.hpp:49-64)
__global__ void k_replace_by_tail(cd *__restrict__ g, const cd *__restrict__ tail, int n_moments, long n_w, int n_cols, long total, double beta,
                                  int stat, int first, int n_min) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long pt = i / n_cols;
  const int c = (int)(i - pt * n_cols);
  const long gf = pt / n_w, p = pt - gf * n_w;
  const long n = first + p;
  if (!(n >= n_min || n < -n_min)) return;
  const double w = M_PI * (double)(2 * n + stat) / beta; // matsubara.hpp:55; om = i w
  const cd *t = tail + (size_t)gf * n_moments * n_cols + c;
  cd z = cmk(1.0, 0.0), res = cmk(0.0, 0.0);
  for (int m = 0; m < n_moments; ++m) {
    res = cfma(t[(size_t)m * n_cols], z, res);
    z = cmk(z.y / w, -z.x / w); // z / (i w)
  }
  g[i] = res;
}

static tq_status herm_args(tq_ctx *ctx, int mesh_kind, long n_pts, int d, long batch, const void *g) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!g || n_pts < 1 || d < 1 || batch < 1 || (mesh_kind != 0 && mesh_kind != 1)) return tq_fail(ctx, TQ_ERR_INVALID, "hermiticity helper: invalid argument");
  return TQ_OK;
}

extern "C" tq_status tq_make_hermitian(tq_ctx *ctx, int mesh_kind, long n_pts, int d, long batch, const tq_complex *g, tq_complex *out) {
  TQ_TRY(herm_args(ctx, mesh_kind, n_pts, d, batch, g));
  if (!out) return tq_fail(ctx, TQ_ERR_INVALID, "tq_make_hermitian: invalid argument");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long total = batch * n_pts * d * d;
  const size_t bytes = sizeof(cd) * (size_t)total;
  const void *d_in = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, g, bytes, ctx->stage_in, &d_in));
  TQ_TRY(tq_stage_out_begin(ctx, out, bytes, ctx->stage_out, &d_out, &out_host));
  if (d_in == d_out) return tq_fail(ctx, TQ_ERR_INVALID, "tq_make_hermitian: in-place operation is not supported");
  {
    KernelScope ks(ctx, "make_hermitian");
    k_make_hermitian<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_in, (cd *)d_out, n_pts, d, total, mesh_kind == 0);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(tq_stage_out_finish(ctx, out, bytes, d_out, out_host));
  return TQ_OK;
}

extern "C" tq_status tq_is_gf_hermitian(tq_ctx *ctx, int mesh_kind, long n_pts, int d, long batch, const tq_complex *g, double tolerance, int *result) {
  TQ_TRY(herm_args(ctx, mesh_kind, n_pts, d, batch, g));
  if (!result) return tq_fail(ctx, TQ_ERR_INVALID, "tq_is_gf_hermitian: invalid argument");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long total = batch * n_pts * d * d;
  const void *d_in = nullptr;
  TQ_TRY(tq_stage_in(ctx, g, sizeof(cd) * (size_t)total, ctx->stage_in, &d_in));
  TQ_TRY(tq_reserve(ctx, ctx->small, 256));
  int *d_flag = (int *)ctx->small.p;
  TQ_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
  {
    KernelScope ks(ctx, "is_gf_hermitian");
    k_is_hermitian<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_in, n_pts, d, total, mesh_kind == 0, tolerance, d_flag);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(tq_reserve_pinned(ctx, sizeof(int)));
  TQ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *result = *(int *)ctx->pinned ? 0 : 1;
  return TQ_OK;
}

extern "C" tq_status tq_replace_by_tail(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_cols, long batch, tq_complex *g,
                                        const tq_complex *tail, int n_moments, int n_min) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!g || !tail || n_iw < 1 || n_cols < 1 || batch < 1 || n_moments < 1 || !(beta > 0) || (statistic != TQ_FERMION && statistic != TQ_BOSON))
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_replace_by_tail: invalid argument");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long last = n_iw - 1, first = -(last + (statistic == TQ_FERMION ? 1 : 0));
  const long n_w = last - first + 1;
  const long total = batch * n_w * n_cols;
  const size_t bytes = sizeof(cd) * (size_t)total;
  const void *d_tail = nullptr;
  TQ_TRY(tq_stage_in(ctx, tail, sizeof(cd) * (size_t)batch * n_moments * n_cols, ctx->stage_in2, &d_tail));
  const bool dev = tq_is_device_ptr(g);
  cd *d_g = (cd *)g;
  if (!dev) {
    TQ_TRY(tq_reserve(ctx, ctx->stage_in, bytes));
    d_g = (cd *)ctx->stage_in.p;
    TQ_CUDA(ctx, cudaMemcpyAsync(d_g, g, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  {
    KernelScope ks(ctx, "replace_by_tail");
    k_replace_by_tail<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_g, (const cd *)d_tail, n_moments, n_w, (int)n_cols, total, beta, statistic,
                                                                               (int)first, n_min);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  if (!dev) {
    TQ_CUDA(ctx, cudaMemcpyAsync(g, d_g, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return TQ_OK;
}


// -------------------------------------------------------------------------------------------------
// gf_evaluator<imfreq>   c++/triqs/gfs/evaluator.hpp:47-75:  g(i w_n) for ANY Matsubara index n -- the mesh value inside the window,
// the fitted high-frequency tail  sum_p A_p / (i w_n)^p  outside (fit_tail_no_normalize with x = |w_max| / i w is the same sum).
// Batched: one fit per gf (the reference refits on every call), then one thread per (point, column).
// -------------------------------------------------------------------------------------------------
tq_status tq_fit_tail_device(tq_ctx *ctx, double beta, int statistic, long n_iw, double tail_fraction, int n_tail_max, int expansion_order,
                             int hermitian, int inner_dim, long n_cols, long batch, const cd *d_g, const cd *d_known, int n_known,
                             cd *d_out_moments, int *n_moments, double *d_err);

__global__ void k_eval_imfreq(const cd *__restrict__ g, const cd *__restrict__ tail, int n_moments, long n_w, int n_cols, long n_pts,
                              const long *__restrict__ ns, long total, double beta, int stat, long first, cd *__restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long q = i / n_cols;
  const int c = (int)(i - q * n_cols);
  const long gf = q / n_pts, p = q - gf * n_pts;
  const long n = ns[p];
  if (n >= first && n < first + n_w) {
    out[i] = g[((size_t)gf * n_w + (n - first)) * n_cols + c];
    return;
  }
  const double w = M_PI * (double)(2 * n + stat) / beta; // matsubara.hpp:55
  const cd *t = tail + (size_t)gf * n_moments * n_cols + c; // (tq_fit_tail_device: dense [batch][n_moments][n_cols])
  cd z = cmk(1.0, 0.0), res = cmk(0.0, 0.0);
  for (int m = 0; m < n_moments; ++m) {
    res = cfma(t[(size_t)m * n_cols], z, res);
    z = cmk(z.y / w, -z.x / w); // z / (i w)
  }
  out[i] = res;
}

extern "C" tq_status tq_evaluate_imfreq(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_cols, long batch, const tq_complex *g, long n_pts,
                                        const long *ns, tq_complex *out, double *tail_err) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!g || !out || !ns || n_iw < 1 || n_cols < 1 || batch < 1 || n_pts < 1 || !(beta > 0) || (statistic != TQ_FERMION && statistic != TQ_BOSON))
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_evaluate_imfreq: invalid argument");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long last = n_iw - 1, first = -(last + (statistic == TQ_FERMION ? 1 : 0)), n_w = last - first + 1;
  const void *d_g = nullptr, *d_ns = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, g, sizeof(cd) * (size_t)batch * n_w * n_cols, ctx->stage_in, &d_g));
  TQ_TRY(tq_stage_in(ctx, ns, sizeof(long) * (size_t)n_pts, ctx->stage_in2, &d_ns));
  TQ_TRY(tq_stage_out_begin(ctx, out, sizeof(cd) * (size_t)batch * n_pts * n_cols, ctx->stage_out, &d_out, &out_host));
  bool outside = true; // host-visible indices: fit only when a point lies outside the mesh
  if (!tq_is_device_ptr(ns)) {
    outside = false;
    for (long p = 0; p < n_pts; ++p) outside |= (ns[p] < first || ns[p] > last);
  }
  const size_t mom_bytes = sizeof(cd) * (size_t)batch * TQ_MAX_MOMENTS * n_cols;
  const size_t err_off = (mom_bytes + 255) & ~size_t(255);
  TQ_TRY(tq_reserve(ctx, ctx->leg_tmp, err_off + sizeof(double) * batch));
  cd *d_mom = (cd *)ctx->leg_tmp.p;
  double *d_err = (double *)((char *)ctx->leg_tmp.p + err_off);
  int n_moments = 0;
  if (outside) {
    TQ_CUDA(ctx, cudaMemsetAsync(d_err, 0, sizeof(double) * batch, ctx->stream));
    TQ_TRY(tq_fit_tail_device(ctx, beta, statistic, n_iw, 0.2, 30, -1, 0, 1, n_cols, batch, (const cd *)d_g, nullptr, 0, d_mom, &n_moments, d_err));
  }
  const long total = batch * n_pts * n_cols;
  {
    KernelScope ks(ctx, "evaluate_imfreq");
    k_eval_imfreq<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_g, d_mom, n_moments, n_w, (int)n_cols, n_pts, (const long *)d_ns,
                                                                           total, beta, statistic, first, (cd *)d_out);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  if (tail_err) {
    if (outside) {
      TQ_CUDA(ctx, cudaMemcpyAsync(tail_err, d_err, sizeof(double) * batch, cudaMemcpyDeviceToHost, ctx->stream));
      TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
      for (long b = 0; b < batch; ++b) tail_err[b] = 0.0;
    }
  }
  TQ_TRY(tq_stage_out_finish(ctx, out, sizeof(cd) * (size_t)batch * n_pts * n_cols, d_out, out_host));
  return TQ_OK;
}
