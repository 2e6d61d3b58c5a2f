// Batched off-mesh evaluation G(tau) by linear interpolation.
//
// Restates  gf_evaluator<imtime>::operator()   c++/triqs/gfs/evaluator.hpp:35-43
//           linear::evaluate                    c++/triqs/mesh/bases/linear.hpp:221-228
// (the reference evaluates one point per call and allocates a matrix for each, benchmarks/GfInterp.cpp:37-39).
// One thread per output element: the table (n_tau x n_elem, a few MB) stays L2/L1 resident, the 16-byte
// results are written with streaming stores because they are never re-read.
#include "tq_common.cuh"

// A CTA takes INTERP_PB points at a time: the mesh cell and weight of every point are computed ONCE (one thread per point, into
// shared memory) instead of once per output element, and the element loop needs no 64-bit division (the old kernel spent most of its
// issue slots on index arithmetic: ncu r02, 67 % issue-slot utilisation at 52 % of DRAM peak).  Same arithmetic per element.
constexpr int INTERP_PB = 256;
__global__ void __launch_bounds__(256) k_interp_tau(const cd *__restrict__ table, long n_tau, int n_elem, double xmin, double delta_inv,
                                                    const double *__restrict__ taus, long n_pts, cd *__restrict__ out) {
  __shared__ int s_i[INTERP_PB];
  __shared__ double s_w[INTERP_PB];
  // t / n_elem by multiplication: exact while t * n_elem < 2^32, i.e. INTERP_PB * n_elem^2 < 2^32
  const bool fast = n_elem > 1 && n_elem < 4000;
  const unsigned magic = fast ? (unsigned)((0x100000000ull + n_elem - 1) / n_elem) : 0u;
  const long n_chunks = (n_pts + INTERP_PB - 1) / INTERP_PB;
  for (long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const long p0 = ch * INTERP_PB;
    const int np = (int)min((long)INTERP_PB, n_pts - p0);
    if ((int)threadIdx.x < np) {
      double x = fmax(__ldg(&taus[p0 + threadIdx.x]), xmin);
      double a = (x - xmin) * delta_inv;
      long i = min((long)a, n_tau - 2);
      s_i[threadIdx.x] = (int)i;
      s_w[threadIdx.x] = fmin(a - (double)i, 1.0);
    }
    __syncthreads();
    const int tot = np * n_elem; // < 2^31: INTERP_PB * n_elem
    cd *o = out + (size_t)p0 * n_elem;
    for (int t = threadIdx.x; t < tot; t += 256) {
      const int pl = fast ? (int)__umulhi((unsigned)t, magic) : (n_elem > 1 ? t / n_elem : t);
      const int e = t - pl * n_elem;
      const int i = s_i[pl];
      const double w = s_w[pl];
      const cd g0 = __ldg(&table[(size_t)i * n_elem + e]), g1 = __ldg(&table[(size_t)(i + 1) * n_elem + e]);
      const double u = 1 - w;
      // (1 - w) * f(i) + w * f(i + 1), products rounded separately like the reference's expression
      const cd r = cmk(__dadd_rn(__dmul_rn(u, g0.x), __dmul_rn(w, g1.x)), __dadd_rn(__dmul_rn(u, g0.y), __dmul_rn(w, g1.y)));
      __stcs(&o[t], r);
    }
    __syncthreads();
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
  if (n_elem > (1 << 22) || n_tau >= (1L << 31)) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tq_interp_tau: table too large");
  long nb = (n_pts + INTERP_PB - 1) / INTERP_PB;
  nb = std::min<long>(nb, (long)ctx->sm_count * 8 * 16);
  KernelScope ks(ctx, "interp_tau");
  k_interp_tau<<<(unsigned)nb, 256, 0, ctx->stream>>>((const cd *)d_tab, n_tau, (int)n_elem, 0.0, delta_inv, (const double *)d_tau, n_pts,
                                                      (cd *)d_out);
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(tq_stage_out_finish(ctx, out, out_bytes, d_out, out_host));
  return TQ_OK;
}
