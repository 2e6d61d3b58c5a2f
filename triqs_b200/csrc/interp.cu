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

// -------------------------------------------------------------------------------------------------
// Off-mesh evaluation on a PRODUCT mesh: multi-linear interpolation, one factor per component mesh.
//
// Restates  evaluate(std::tuple<M...>, f, x1, x...)   c++/triqs/mesh/evaluate.hpp:97-114  (the recursion: the sum over the FIRST
//           component is the innermost one), with the per-component rules
//             linear::evaluate          c++/triqs/mesh/bases/linear.hpp:221-228   (imtime, retime, refreq: clamped cell, weight w)
//             evaluate(brzone1d, f, vi) c++/triqs/mesh/brzone.hpp:355-359         (periodic: floor, positive_modulo; dim == 1 -> 0)
//             evaluate(m, f, index)     c++/triqs/mesh/evaluate.hpp:73            (an index passes through: imfreq n, cyclat index)
// One thread per output element; the cell and weight of every (point, component) are computed once per point in shared memory.
// Every two-point combination is (1 - w) * A + w * B with the products rounded separately, like the reference's expression.
constexpr int PROD_PB = 128;
struct ProdArgs {
  int n_dims;
  int kind[TQ_PROD_MAX_DIMS];
  long dim[TQ_PROD_MAX_DIMS];
  long stride[TQ_PROD_MAX_DIMS]; // in elements of n_elem complex numbers
  double xmin[TQ_PROD_MAX_DIMS], delta_inv[TQ_PROD_MAX_DIMS];
};
TQ_D long prod_pmod(long i, long d) {
  const long r = i % d;
  return r < 0 ? r + d : r;
}
__global__ void __launch_bounds__(256) k_evaluate_prod(const __grid_constant__ ProdArgs P, const cd *__restrict__ table, int n_elem,
                                                       const double *__restrict__ coords, long n_pts, cd *__restrict__ out) {
  __shared__ long s_lo[TQ_PROD_MAX_DIMS][PROD_PB], s_hi[TQ_PROD_MAX_DIMS][PROD_PB];
  __shared__ double s_w[TQ_PROD_MAX_DIMS][PROD_PB];
  const int nd = P.n_dims;
  const long n_chunks = (n_pts + PROD_PB - 1) / PROD_PB;
  for (long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const long p0 = ch * PROD_PB;
    const int np = (int)min((long)PROD_PB, n_pts - p0);
    for (int t = threadIdx.x; t < np * nd; t += blockDim.x) {
      const int pl = t / nd, d = t - pl * nd;
      const double x = __ldg(&coords[(p0 + pl) * nd + d]);
      long lo, hi;
      double w;
      if (P.kind[d] == TQ_AXIS_LINEAR) { // linear.hpp:221-228
        const double xc = fmax(x, P.xmin[d]);
        const double a = (xc - P.xmin[d]) * P.delta_inv[d];
        lo = min((long)a, P.dim[d] - 2);
        hi = lo + 1;
        w = fmin(a - (double)lo, 1.0);
      } else if (P.kind[d] == TQ_AXIS_PERIODIC) { // brzone.hpp:355-359
        const long i = (long)floor(x);
        w = x - (double)i;
        lo = prod_pmod(i, P.dim[d]);
        hi = P.dim[d] == 1 ? 0 : prod_pmod(i + 1, P.dim[d]);
      } else { // an index: pass through (evaluate.hpp:73)
        lo = hi = (long)llrint(x);
        w = -1.0; // marks "no interpolation along this component"
      }
      s_lo[d][pl] = lo * P.stride[d];
      s_hi[d][pl] = hi * P.stride[d];
      s_w[d][pl] = w;
    }
    __syncthreads();
    const long tot = (long)np * n_elem;
    cd *o = out + (size_t)p0 * n_elem;
    for (long t = threadIdx.x; t < tot; t += blockDim.x) {
      const int pl = (int)(t / n_elem);
      const int e = (int)(t - (long)pl * n_elem);
      // corners: bit j of the corner number selects lo / hi of the j-th INTERPOLATED component (in mesh order)
      int idim[TQ_PROD_MAX_DIMS], m = 0;
      long base = 0;
      for (int d = 0; d < nd; ++d) {
        if (s_w[d][pl] < 0.0)
          base += s_lo[d][pl];
        else
          idim[m++] = d;
      }
      cd v[1 << TQ_PROD_MAX_DIMS];
      for (int c = 0; c < (1 << m); ++c) {
        long off = base;
        for (int j = 0; j < m; ++j) off += ((c >> j) & 1) ? s_hi[idim[j]][pl] : s_lo[idim[j]][pl];
        v[c] = __ldg(&table[(size_t)off * n_elem + e]);
      }
      // the first component is the innermost sum of the recursion (evaluate.hpp:104-110)
      for (int j = 0; j < m; ++j) {
        const double w = s_w[idim[j]][pl], u = 1 - w;
        for (int c = 0; c < (1 << (m - 1 - j)); ++c) {
          const cd A = v[2 * c], B = v[2 * c + 1];
          v[c] = cmk(__dadd_rn(__dmul_rn(u, A.x), __dmul_rn(w, B.x)), __dadd_rn(__dmul_rn(u, A.y), __dmul_rn(w, B.y)));
        }
      }
      __stcs(&o[t], v[0]);
    }
    __syncthreads();
  }
}

extern "C" tq_status tq_evaluate_prod(tq_ctx *ctx, int n_dims, const tq_axis *axes, long n_elem, const tq_complex *table, long n_pts,
                                      const double *coords, tq_complex *out) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!axes || !table || !coords || !out || n_dims < 1 || n_dims > TQ_PROD_MAX_DIMS || n_elem < 1 || n_pts < 0)
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_evaluate_prod: invalid argument");
  ProdArgs P{};
  P.n_dims = n_dims;
  size_t total = 1;
  for (int d = n_dims - 1; d >= 0; --d) {
    const tq_axis &a = axes[d];
    if (a.size < 1 || (a.kind != TQ_AXIS_LINEAR && a.kind != TQ_AXIS_PERIODIC && a.kind != TQ_AXIS_INDEX))
      return tq_fail(ctx, TQ_ERR_INVALID, "tq_evaluate_prod: invalid axis");
    if (a.kind == TQ_AXIS_LINEAR && (a.size < 2 || !(a.xmax > a.xmin))) // linear.hpp:222 EXPECTS(size() > 1)
      return tq_fail(ctx, TQ_ERR_INVALID, "tq_evaluate_prod: a linear axis needs at least two points and xmax > xmin");
    P.kind[d] = a.kind;
    P.dim[d] = a.size;
    P.stride[d] = (long)total;
    P.xmin[d] = a.xmin;
    if (a.kind == TQ_AXIS_LINEAR) {
      const double delta = (a.xmax - a.xmin) / (double)(a.size - 1); // linear.hpp:51
      P.delta_inv[d] = 1.0 / delta;                                   // linear.hpp:52
    }
    total *= (size_t)a.size;
  }
  if (n_pts == 0) return TQ_OK;
  if (n_elem > (1 << 22)) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tq_evaluate_prod: target too large");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const void *d_tab = nullptr, *d_x = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  const size_t out_bytes = sizeof(cd) * (size_t)n_pts * n_elem;
  TQ_TRY(tq_stage_in(ctx, table, sizeof(cd) * total * (size_t)n_elem, ctx->stage_in, &d_tab));
  TQ_TRY(tq_stage_in(ctx, coords, sizeof(double) * (size_t)n_pts * n_dims, ctx->stage_in2, &d_x));
  TQ_TRY(tq_stage_out_begin(ctx, out, out_bytes, ctx->stage_out, &d_out, &out_host));
  long nb = (n_pts + PROD_PB - 1) / PROD_PB;
  nb = std::min<long>(nb, (long)ctx->sm_count * 8 * 16);
  {
    KernelScope ks(ctx, "evaluate_prod");
    k_evaluate_prod<<<(unsigned)nb, 256, 0, ctx->stream>>>(P, (const cd *)d_tab, (int)n_elem, (const double *)d_x, n_pts, (cd *)d_out);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(tq_stage_out_finish(ctx, out, out_bytes, d_out, out_host));
  return TQ_OK;
}
