// Batched off-mesh evaluation G(tau) by linear interpolation.
//
// Restates  gf_evaluator<imtime>::operator()   c++/triqs/gfs/evaluator.hpp:35-43
//           linear::evaluate                    c++/triqs/mesh/bases/linear.hpp:221-228
// (the reference evaluates one point per call and allocates a matrix for each, benchmarks/GfInterp.cpp:37-39).
// One thread per output element: the table (n_tau x n_elem, a few MB) stays L2/L1 resident, the 16-byte
// results are written with streaming stores because they are never re-read.
#include "tq_common.cuh"

__global__ void __launch_bounds__(256) k_interp_tau(const cd *__restrict__ table, long n_tau, int n_elem, double xmin, double delta_inv,
                                                    const double *__restrict__ taus, long n_pts, cd *__restrict__ out) {
  const long total = n_pts * n_elem;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long pt = t / n_elem;
    const int e = (int)(t - pt * n_elem);
    double x = fmax(__ldg(&taus[pt]), xmin);
    double a = (x - xmin) * delta_inv;
    long i = min((long)a, n_tau - 2);
    double w = fmin(a - (double)i, 1.0);
    cd g0 = __ldg(&table[(size_t)i * n_elem + e]), g1 = __ldg(&table[(size_t)(i + 1) * n_elem + e]);
    double u = 1 - w;
    // (1 - w) * f(i) + w * f(i + 1), products rounded separately like the reference's expression
    cd r = cmk(__dadd_rn(__dmul_rn(u, g0.x), __dmul_rn(w, g1.x)), __dadd_rn(__dmul_rn(u, g0.y), __dmul_rn(w, g1.y)));
    __stcs(&out[t], r);
  }
}

extern "C" tq_status tq_interp_tau(tq_ctx *ctx, double beta, long n_tau, long n_elem, const tq_complex *table, long n_pts, const double *taus,
                                   tq_complex *out) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!table || !taus || !out || n_tau < 2 || n_elem < 1 || n_pts < 0 || !(beta > 0))
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_interp_tau: invalid argument");
  if (n_pts == 0) return TQ_OK;
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const void *d_tab = nullptr, *d_tau = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  const size_t out_bytes = sizeof(cd) * (size_t)n_pts * n_elem;
  TQ_TRY(tq_stage_in(ctx, table, sizeof(cd) * (size_t)n_tau * n_elem, ctx->stage_in, &d_tab));
  TQ_TRY(tq_stage_in(ctx, taus, sizeof(double) * (size_t)n_pts, ctx->stage_in2, &d_tau));
  TQ_TRY(tq_stage_out_begin(ctx, out, out_bytes, ctx->stage_out, &d_out, &out_host));
  const double delta = (beta - 0.0) / (double)(n_tau - 1); // linear.hpp:51
  const double delta_inv = 1.0 / delta;                    // linear.hpp:52
  const long total = n_pts * n_elem;
  long nb = (total + 255) / 256;
  nb = std::min<long>(nb, (long)ctx->sm_count * 8 * 64);
  KernelScope ks(ctx, "interp_tau");
  k_interp_tau<<<(unsigned)nb, 256, 0, ctx->stream>>>((const cd *)d_tab, n_tau, (int)n_elem, 0.0, delta_inv, (const double *)d_tau, n_pts,
                                                      (cd *)d_out);
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(tq_stage_out_finish(ctx, out, out_bytes, d_out, out_host));
  return TQ_OK;
}
