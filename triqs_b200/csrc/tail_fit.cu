// Least-squares high-frequency tail fit (fit_tail / fit_hermitian_tail).
//
// Restates c++/triqs/mesh/tail_fitter.hpp:67-298 and the external nda::lapack::gelss_worker{,_hermitian}
// (TRIQS/nda 1.3.x, nda/lapack/gelss_worker.hpp -- not vendored in the reference tree):
//   * mesh-only setup (fit indices :105-128, Vandermonde :34-44/:149-155, adaptive order :165-177, SVD
//     pseudo-inverse) happens once per (mesh, n_fixed, hermitian) on the host and is cached in the
//     context, exactly like tail_fitter::_lss.  The SVD is a one-sided complex Jacobi (high relative
//     accuracy for the small, graded Vandermonde matrices; LAPACK is not linked).
//   * the per-gf work -- gather the fit rows, subtract known moments (:233-254), x = pinv * B, residual
//     error, rescale (:260-268), hermitian symmetrisation -- is one batched kernel: a thread owns a
//     column (and, for the hermitian fit, its transposed partner) and keeps the <= 10 complex unknowns
//     in registers.
#include "tq_common.cuh"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>

typedef std::complex<double> zc;

// ---------------------------------------------------------------------------------------------
// host: mesh-only setup
// ---------------------------------------------------------------------------------------------
struct TailSetup {
  std::vector<long> idx;     // fit indices (Matsubara n)
  std::vector<zc> vander;    // [n_idx][order+1]
  int order = 9;             // columns - 1 of vander
  int n_var = 0;
  int rows_total = 0;        // n_idx (plain) or 2 n_idx (hermitian)
  std::vector<zc> pinv;      // [n_var][rows_total]
  std::vector<double> sv;    // singular values of the plain A, descending
  double w_max = 0;
  long first = 0;
};

// one-sided Jacobi SVD of A [M][N] (row major), M >= N.  Returns un-normalised U columns in Uw (A V = Uw),
// V [N][N] and sigma_j = ||Uw_j||.
static void jacobi_svd(int M, int N, const std::vector<zc> &A, std::vector<zc> &Uw, std::vector<zc> &V, std::vector<double> &sig) {
  Uw = A;
  V.assign((size_t)N * N, zc(0, 0));
  for (int i = 0; i < N; ++i) V[(size_t)i * N + i] = 1.0;
  const double tol = 1e-16;
  for (int sweep = 0; sweep < 100; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        double alpha = 0, beta = 0;
        zc gamma = 0;
        for (int i = 0; i < M; ++i) {
          zc up = Uw[(size_t)i * N + p], uq = Uw[(size_t)i * N + q];
          alpha += std::norm(up);
          beta += std::norm(uq);
          gamma += std::conj(up) * uq;
        }
        double g = std::abs(gamma);
        if (g <= tol * std::sqrt(alpha * beta) || g == 0.0) continue;
        rotated = true;
        zc ph = gamma / g; // e^{i phi}
        double zeta = (beta - alpha) / (2.0 * g);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        zc phc = std::conj(ph);
        for (int i = 0; i < M; ++i) {
          zc up = Uw[(size_t)i * N + p], uq = Uw[(size_t)i * N + q] * phc;
          Uw[(size_t)i * N + p] = c * up - s * uq;
          Uw[(size_t)i * N + q] = s * up + c * uq;
        }
        for (int i = 0; i < N; ++i) {
          zc vp = V[(size_t)i * N + p], vq = V[(size_t)i * N + q] * phc;
          V[(size_t)i * N + p] = c * vp - s * vq;
          V[(size_t)i * N + q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  sig.resize(N);
  for (int j = 0; j < N; ++j) {
    double a = 0;
    for (int i = 0; i < M; ++i) a += std::norm(Uw[(size_t)i * N + j]);
    sig[j] = std::sqrt(a);
  }
}

// pinv = V diag(1/sigma^2) Uw^H   [N][M]
static void pinv_from_svd(int M, int N, const std::vector<zc> &Uw, const std::vector<zc> &V, const std::vector<double> &sig, std::vector<zc> &P) {
  P.assign((size_t)N * M, zc(0, 0));
  for (int v = 0; v < N; ++v)
    for (int i = 0; i < M; ++i) {
      zc acc = 0;
      for (int j = 0; j < N; ++j) acc += V[(size_t)v * N + j] * std::conj(Uw[(size_t)i * N + j]) / (sig[j] * sig[j]);
      P[(size_t)v * M + i] = acc;
    }
}

// The fitter is generic in the mesh (tail_fitter.hpp:105-181 only uses size / first_index / last_index / w_max / to_value):
//   imfreq (stat = TQ_FERMION / TQ_BOSON):  beta, n_iw                                   values i w_n
//   refreq (stat = TAIL_MESH_REFREQ):       `beta` carries w_max, `w_min` the lower end, n_iw = number of points;  values w_i (real)
constexpr int TAIL_MESH_REFREQ = 2;
static std::string tail_setup_host(double beta, int stat, long n_iw, double tail_fraction, int n_tail_max, int expansion_order, int n_fixed,
                                   int hermitian, TailSetup &S, double w_min = 0.0) {
  const bool refreq = stat == TAIL_MESH_REFREQ;
  auto mesh_value = [&](long n) -> zc { // mesh.to_value(n)
    if (!refreq) return zc(0.0, M_PI * (double)(2 * n + stat) / beta);
    if (n_iw == 1) return zc(w_min, 0.0);
    const double wr = double(n) / (n_iw - 1); // linear.hpp:154-160
    return zc(w_min * (1 - wr) + beta * wr, 0.0);
  };
  const int max_order = 9;      // tail_fitter.hpp:69
  const double rcond = 1e-8;    // :74
  const bool adjust = expansion_order < 0;
  S.order = adjust ? max_order : expansion_order;
  const long last = n_iw - 1;
  const long first = refreq ? 0 : -(last + (stat == TQ_FERMION ? 1 : 0));
  S.first = first;
  const long size = last - first + 1;
  // get_tail_fit_indices, :105-128
  int n_rng = (int)std::round(tail_fraction * size / 2);
  int n_tail = std::min(n_rng, n_tail_max);
  S.idx.clear();
  if (n_tail > 0) {
    double step = double(n_rng) / n_tail;
    double idx1 = (double)first;
    double idx2 = (double)(last - n_rng);
    for (int n = 0; n < n_tail; ++n) {
      S.idx.push_back(long(idx1));
      S.idx.push_back(long(idx2));
      idx1 += step;
      idx2 += step;
    }
  }
  const int M = (int)S.idx.size();
  S.w_max = refreq ? std::abs(beta) : std::abs(M_PI * (double)(2 * last + stat) / beta); // |w_max()|, imfreq.hpp:128, refreq.hpp:52
  // vander, :34-44, :149-155
  S.vander.assign((size_t)M * (S.order + 1), zc(0, 0));
  for (int i = 0; i < M; ++i) {
    zc c = S.w_max / mesh_value(S.idx[i]);
    zc z = 1;
    for (int n = 0; n <= S.order; ++n) {
      S.vander[(size_t)i * (S.order + 1) + n] = z;
      z *= c;
    }
  }
  if (n_fixed + 1 > M / 2) return "Insufficient data points for least square procedure"; // :157

  auto build = [&](int n, std::vector<double> &sv_plain, std::vector<zc> &P, int &rows_total) {
    const int N = n + 1 - n_fixed;
    std::vector<zc> A((size_t)M * N);
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) A[(size_t)i * N + j] = S.vander[(size_t)i * (S.order + 1) + n_fixed + j];
    std::vector<zc> Uw, V;
    jacobi_svd(M, N, A, Uw, V, sv_plain);
    if (!hermitian) {
      pinv_from_svd(M, N, Uw, V, sv_plain, P);
      rows_total = M;
    } else {
      // gelss_worker_hermitian: second worker on vstack(A, conj(A))
      std::vector<zc> As((size_t)2 * M * N);
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
          As[(size_t)i * N + j] = A[(size_t)i * N + j];
          As[(size_t)(M + i) * N + j] = std::conj(A[(size_t)i * N + j]);
        }
      std::vector<zc> Uw2, V2;
      std::vector<double> sv2;
      jacobi_svd(2 * M, N, As, Uw2, V2, sv2);
      pinv_from_svd(2 * M, N, Uw2, V2, sv2, P);
      rows_total = 2 * M;
    }
    std::sort(sv_plain.begin(), sv_plain.end(), std::greater<double>());
    return N;
  };

  if (!adjust) {
    if (S.order < n_fixed) return "expansion_order smaller than the number of known moments";
    S.n_var = build(S.order, S.sv, S.pinv, S.rows_total);
    return "";
  }
  // :165-177
  long n_max = std::min<long>(max_order, (long)(1. + 16. / std::log10(1 + S.w_max)));
  n_max = std::min<long>(n_max, M / 2);
  for (int n = (int)n_max; n >= n_fixed; --n) {
    std::vector<double> sv;
    std::vector<zc> P;
    int rt = 0;
    int N = build(n, sv, P, rt);
    if (sv[N - 1] > rcond) {
      S.n_var = N;
      S.sv = sv;
      S.pinv = P;
      S.rows_total = rt;
      return "";
    }
  }
  return "Conditioning of tail-fit violates boundary"; // :180
}

extern "C" tq_status tq_tail_fitter_setup_host(double beta, int statistic, long n_iw, double tail_fraction, int n_tail_max, int expansion_order,
                                               int n_fixed, int hermitian, int *n_idx_out, long *idx_out, int *n_var_out, tq_complex *pinv_out,
                                               double *sv_out, char *errbuf, size_t errbuf_len) {
  if (expansion_order > TQ_MAX_MOMENTS - 1 || n_fixed < 0 || n_iw < 1 || !(beta > 0)) return TQ_ERR_INVALID;
  TailSetup S;
  std::string e = tail_setup_host(beta, statistic, n_iw, tail_fraction, n_tail_max, expansion_order, n_fixed, hermitian, S);
  if (!e.empty()) {
    if (errbuf && errbuf_len) {
      strncpy(errbuf, e.c_str(), errbuf_len - 1);
      errbuf[errbuf_len - 1] = 0;
    }
    return TQ_ERR_RUNTIME;
  }
  if (n_idx_out) *n_idx_out = (int)S.idx.size();
  if (idx_out) std::copy(S.idx.begin(), S.idx.end(), idx_out);
  if (n_var_out) *n_var_out = S.n_var;
  if (pinv_out) memcpy(pinv_out, S.pinv.data(), sizeof(zc) * S.pinv.size());
  if (sv_out) std::copy(S.sv.begin(), S.sv.end(), sv_out);
  return TQ_OK;
}

// ---------------------------------------------------------------------------------------------
// device: batched application
// ---------------------------------------------------------------------------------------------
struct TailFitterPlan {
  TailSetup S;
  int *d_rows = nullptr; // data row of each fit index
  cd *d_vander = nullptr;
  cd *d_pinv = nullptr;
  ~TailFitterPlan() {
    if (d_rows) cudaFree(d_rows);
    if (d_vander) cudaFree(d_vander);
    if (d_pinv) cudaFree(d_pinv);
  }
};

struct FitArgs {
  int M, order1 /* order + 1 */, n_var, n_fixed, rows_total, herm, d, n_cols, n_w;
  double w_max;
  const int *rows;
  const cd *vander, *pinv;
  const cd *g;     // [batch][n_w][n_cols]
  const cd *known; // [batch][n_fixed][n_cols]
  cd *out;         // [batch][n_fixed + n_var][n_cols]
  double *err;     // [batch]
  long total;      // batch * n_cols
};

// fit-row value of column c with the known moments removed   (tail_fitter.hpp:216-254)
TQ_D cd fit_rhs(const FitArgs &A, const cd *g, const cd *known, int i, int c) {
  cd b = g[(size_t)A.rows[i] * A.n_cols + c];
  double z = 1.0;
  for (int p = 0; p < A.n_fixed; ++p) {
    cd km = cscale(z, known[(size_t)p * A.n_cols + c]);
    cd v = __ldg(&A.vander[(size_t)i * A.order1 + p]);
    b = csub(b, cmul(v, km));
    z /= A.w_max;
  }
  return b;
}

// one column of one gf; `rhs(i, c)` returns the fit-row value with the known moments removed
// The fit rows are split over `NL` cooperating lanes (1 = this thread alone, 32 = its warp; partial sums are combined by
// xor-shuffles, so every lane ends up with the same totals and lane 0 writes the results).
template <int NL, class Rhs> TQ_D void fit_column(const FitArgs &A, long gf, int c, const cd *known, Rhs rhs, int lane = 0) {
  int cp = c; // partner column (inner adjoint), hermitian fit only
  if (A.herm) {
    int dd = A.d * A.d;
    int blk = c / dd, r = c - blk * dd;
    int a = r / A.d, b = r - a * A.d;
    cp = blk * dd + b * A.d + a;
  }
  cd x[TQ_MAX_MOMENTS], xp[TQ_MAX_MOMENTS];
#pragma unroll
  for (int v = 0; v < TQ_MAX_MOMENTS; ++v) x[v] = xp[v] = cmk(0, 0);
  const int M = A.M;
  for (int i = lane; i < M; i += NL) {
    cd b = rhs(i, c);
    if (!A.herm) {
#pragma unroll
      for (int v = 0; v < TQ_MAX_MOMENTS; ++v)
        if (v < A.n_var) x[v] = cfma(__ldg(&A.pinv[(size_t)v * A.rows_total + i]), b, x[v]);
    } else {
      cd bp = rhs(i, cp);
      cd bc = cconj(b), bpc = cconj(bp);
#pragma unroll
      for (int v = 0; v < TQ_MAX_MOMENTS; ++v)
        if (v < A.n_var) {
          cd p1 = __ldg(&A.pinv[(size_t)v * A.rows_total + i]), p2 = __ldg(&A.pinv[(size_t)v * A.rows_total + M + i]);
          x[v] = cfma(p2, bpc, cfma(p1, b, x[v]));   // solution of the stacked system for column c
          xp[v] = cfma(p2, bc, cfma(p1, bp, xp[v])); // ... and for the partner column
        }
    }
  }
  if (NL > 1) {
#pragma unroll
    for (int v = 0; v < TQ_MAX_MOMENTS; ++v)
      if (v < A.n_var) {
#pragma unroll
        for (int o = NL / 2; o > 0; o >>= 1) {
          x[v].x += __shfl_xor_sync(0xffffffffu, x[v].x, o);
          x[v].y += __shfl_xor_sync(0xffffffffu, x[v].y, o);
          xp[v].x += __shfl_xor_sync(0xffffffffu, xp[v].x, o);
          xp[v].y += __shfl_xor_sync(0xffffffffu, xp[v].y, o);
        }
      }
  }
  // residual of the (stacked) least-squares problem == || U_null^H B || of gelss_worker
  double r2 = 0.0;
  for (int i = lane; i < M; i += NL) {
    cd b = rhs(i, c);
    cd ax = cmk(0, 0), axc = cmk(0, 0);
#pragma unroll
    for (int v = 0; v < TQ_MAX_MOMENTS; ++v)
      if (v < A.n_var) {
        cd a = __ldg(&A.vander[(size_t)i * A.order1 + A.n_fixed + v]);
        ax = cfma(a, x[v], ax);
        if (A.herm) axc = cfma(cconj(a), x[v], axc);
      }
    r2 += cabs2(csub(b, ax));
    if (A.herm) {
      cd bp = rhs(i, cp);
      r2 += cabs2(csub(cconj(bp), axc));
    }
  }
  if (NL > 1) {
#pragma unroll
    for (int o = NL / 2; o > 0; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
    if (lane != 0) return;
  }
  double e = (A.rows_total == A.n_var) ? 0.0 : sqrt(r2) / sqrt((double)A.rows_total);
  if (e == e) {
    atomicMax((unsigned long long *)&A.err[gf], (unsigned long long)__double_as_longlong(e));
  } else {
    atomicMax((unsigned long long *)&A.err[gf], (unsigned long long)__double_as_longlong(INFINITY));
  }
  // assemble: known rows, then fitted rows rescaled by w_max^n   (:260-287)
  cd *o = A.out + (size_t)gf * (A.n_fixed + A.n_var) * A.n_cols;
  double z = 1.0;
  for (int p = 0; p < A.n_fixed; ++p) {
    o[(size_t)p * A.n_cols + c] = known[(size_t)p * A.n_cols + c];
    z *= A.w_max;
  }
#pragma unroll
  for (int v = 0; v < TQ_MAX_MOMENTS; ++v)
    if (v < A.n_var) {
      cd xv = A.herm ? cscale(0.5, cadd(x[v], cconj(xp[v]))) : x[v];
      o[(size_t)(A.n_fixed + v) * A.n_cols + c] = cscale(z, xv);
      z *= A.w_max;
    }
}

// direct version: every thread walks its column's fit rows in global memory (any n_cols)
__global__ void __launch_bounds__(128) k_fit_tail(const FitArgs A) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.total) return;
  const long gf = t / A.n_cols;
  const int c = (int)(t - gf * A.n_cols);
  const cd *g = A.g + (size_t)gf * A.n_w * A.n_cols;
  const cd *known = A.known + (size_t)gf * A.n_fixed * A.n_cols;
  fit_column<1>(A, gf, c, known, [&](int i, int cc) { return fit_rhs(A, g, known, i, cc); });
}
// warp version (few columns, many gfs): one warp per (gf, column), the fit rows split over its lanes
__global__ void __launch_bounds__(128) k_fit_tail_warp(const FitArgs A) {
  const long t = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= A.total) return; // (whole warps)
  const long gf = t / A.n_cols;
  const int c = (int)(t - gf * A.n_cols);
  const cd *g = A.g + (size_t)gf * A.n_w * A.n_cols;
  const cd *known = A.known + (size_t)gf * A.n_fixed * A.n_cols;
  fit_column<32>(A, gf, c, known, [&](int i, int cc) { return fit_rhs(A, g, known, i, cc); }, (int)(threadIdx.x & 31));
}
// staged version (one CTA per gf): the M x n_cols right-hand sides are gathered into shared memory by the whole CTA first
// (independent loads in flight instead of a chain of ~2 M dependent row reads per thread); same arithmetic, same results
__global__ void __launch_bounds__(128) k_fit_tail_staged(const FitArgs A) {
  extern __shared__ double2 sB[]; // [M][n_cols]
  const long gf = blockIdx.x;
  const cd *g = A.g + (size_t)gf * A.n_w * A.n_cols;
  const cd *known = A.known + (size_t)gf * A.n_fixed * A.n_cols;
  const int n = A.M * A.n_cols;
  for (int idx = threadIdx.x; idx < n; idx += 128) {
    const int i = idx / A.n_cols, c = idx - i * A.n_cols;
    sB[idx] = fit_rhs(A, g, known, i, c);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < A.n_cols; c += 128) fit_column<1>(A, gf, c, known, [&](int i, int cc) { return sB[i * A.n_cols + cc]; });
}

static tq_status get_tail_plan(tq_ctx *ctx, double beta, int stat, long n_iw, double tail_fraction, int n_tail_max, int expansion_order,
                               int n_fixed, int hermitian, std::shared_ptr<TailFitterPlan> *out, double w_min = 0.0) {
  auto key = std::make_tuple(beta, stat, n_iw, tail_fraction, n_tail_max, expansion_order, n_fixed, hermitian, w_min);
  auto it = ctx->tail_plans.find(key);
  if (it != ctx->tail_plans.end()) {
    *out = it->second;
    return TQ_OK;
  }
  auto pl = std::make_shared<TailFitterPlan>();
  std::string e = tail_setup_host(beta, stat, n_iw, tail_fraction, n_tail_max, expansion_order, n_fixed, hermitian, pl->S, w_min);
  if (!e.empty()) return tq_fail(ctx, TQ_ERR_RUNTIME, e);
  TailSetup &S = pl->S;
  std::vector<int> rows(S.idx.size());
  for (size_t i = 0; i < rows.size(); ++i) rows[i] = (int)(S.idx[i] - S.first); // imfreq.hpp:194 to_data_index
  TQ_CUDA(ctx, cudaMalloc(&pl->d_rows, sizeof(int) * rows.size()));
  TQ_CUDA(ctx, cudaMalloc(&pl->d_vander, sizeof(cd) * S.vander.size()));
  TQ_CUDA(ctx, cudaMalloc(&pl->d_pinv, sizeof(cd) * S.pinv.size()));
  TQ_CUDA(ctx, cudaMemcpyAsync(pl->d_rows, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice, ctx->stream));
  TQ_CUDA(ctx, cudaMemcpyAsync(pl->d_vander, S.vander.data(), sizeof(cd) * S.vander.size(), cudaMemcpyHostToDevice, ctx->stream));
  TQ_CUDA(ctx, cudaMemcpyAsync(pl->d_pinv, S.pinv.data(), sizeof(cd) * S.pinv.size(), cudaMemcpyHostToDevice, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->tail_plans[key] = pl;
  *out = pl;
  return TQ_OK;
}

// device-pointer entry used by the inverse transform and by tq_fit_tail.  out is dense [batch][n_moments][n_cols].
static tq_status fit_tail_device_mesh(tq_ctx *ctx, double beta, int statistic, long n_iw, double w_min, double tail_fraction, int n_tail_max,
                                      int expansion_order, int hermitian, int inner_dim, long n_cols, long batch, const cd *d_g,
                                      const cd *d_known, int n_known, cd *d_out_moments, int *n_moments, double *d_err);
tq_status tq_fit_tail_device(tq_ctx *ctx, double beta, int statistic, long n_iw, double tail_fraction, int n_tail_max, int expansion_order,
                             int hermitian, int inner_dim, long n_cols, long batch, const cd *d_g, const cd *d_known, int n_known,
                             cd *d_out_moments, int *n_moments, double *d_err) {
  return fit_tail_device_mesh(ctx, beta, statistic, n_iw, 0.0, tail_fraction, n_tail_max, expansion_order, hermitian, inner_dim, n_cols, batch, d_g,
                              d_known, n_known, d_out_moments, n_moments, d_err);
}
static tq_status fit_tail_device_mesh(tq_ctx *ctx, double beta, int statistic, long n_iw, double w_min, double tail_fraction, int n_tail_max,
                                      int expansion_order, int hermitian, int inner_dim, long n_cols, long batch, const cd *d_g,
                                      const cd *d_known, int n_known, cd *d_out_moments, int *n_moments, double *d_err) {
  if (expansion_order > TQ_MAX_MOMENTS - 1)
    return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tail fit: expansion_order > " + std::to_string(TQ_MAX_MOMENTS - 1) + " is not supported");
  if (hermitian && (inner_dim < 1 || n_cols % ((long)inner_dim * inner_dim) != 0))
    return tq_fail(ctx, TQ_ERR_RUNTIME, "ERROR in hermitian least square fitting: Data shape incompatible with given dimension");
  std::shared_ptr<TailFitterPlan> pl;
  TQ_TRY(get_tail_plan(ctx, beta, statistic, n_iw, tail_fraction, n_tail_max, expansion_order, n_known, hermitian, &pl, w_min));
  const TailSetup &S = pl->S;
  FitArgs A{};
  A.M = (int)S.idx.size();
  A.order1 = S.order + 1;
  A.n_var = S.n_var;
  A.n_fixed = n_known;
  A.rows_total = S.rows_total;
  A.herm = hermitian;
  A.d = inner_dim;
  A.n_cols = (int)n_cols;
  const long last = n_iw - 1;
  const long first = statistic == TAIL_MESH_REFREQ ? 0 : -(last + (statistic == TQ_FERMION ? 1 : 0));
  A.n_w = (int)(last - first + 1);
  A.w_max = S.w_max;
  A.rows = pl->d_rows;
  A.vander = pl->d_vander;
  A.pinv = pl->d_pinv;
  A.g = d_g;
  A.known = d_known;
  A.out = d_out_moments;
  A.err = d_err;
  A.total = batch * n_cols;
  *n_moments = S.n_var + n_known;
  KernelScope ks(ctx, "fit_tail");
  const size_t stage_bytes = sizeof(cd) * (size_t)A.M * (size_t)n_cols;
  if (n_cols < 32 && (A.total + 3) / 4 <= 0x7fffffffL)
    k_fit_tail_warp<<<(unsigned)((A.total + 3) / 4), 128, 0, ctx->stream>>>(A);
  else if (stage_bytes <= 48 * 1024 && batch <= 0x7fffffffL)
    k_fit_tail_staged<<<(unsigned)batch, 128, stage_bytes, ctx->stream>>>(A);
  else
    k_fit_tail<<<(unsigned)((A.total + 127) / 128), 128, 0, ctx->stream>>>(A);
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  return TQ_OK;
}

static tq_status fit_tail_impl(tq_ctx *ctx, double beta, int statistic, long n_iw, double w_min, double tail_fraction, int n_tail_max,
                               int expansion_order, int hermitian, int inner_dim, long n_cols, long batch, const tq_complex *g,
                               const tq_complex *known, int n_known, tq_complex *out_moments, int *out_n_moments, double *out_err) {
  if (!ctx) return TQ_ERR_INVALID;
  const bool refreq = statistic == TAIL_MESH_REFREQ;
  if (!g || !out_moments || !out_n_moments || n_iw < 1 || n_cols < 1 || batch < 1 || (!refreq && !(beta > 0)) || n_known < 0 ||
      (statistic != TQ_FERMION && statistic != TQ_BOSON && !refreq))
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_fit_tail: invalid argument");
  if (hermitian && inner_dim < 1) // tail_fitter.hpp:196-197
    return tq_fail(ctx, TQ_ERR_RUNTIME, "Enforcing the hermiticity in tail_fit requires inner matrix dimension");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!known) n_known = 0;
  const long last = n_iw - 1;
  const long first = refreq ? 0 : -(last + (statistic == TQ_FERMION ? 1 : 0));
  const long n_w = last - first + 1;
  const int order_eff = expansion_order < 0 ? 9 : expansion_order;
  const void *d_g = nullptr, *d_known = nullptr;
  TQ_TRY(tq_stage_in(ctx, g, sizeof(cd) * (size_t)batch * n_w * n_cols, ctx->stage_in, &d_g));
  if (n_known > 0) TQ_TRY(tq_stage_in(ctx, known, sizeof(cd) * (size_t)batch * n_known * n_cols, ctx->stage_in2, &d_known));
  if (n_known > order_eff) { // :204  return {known_moments, 0.0}
    *out_n_moments = n_known;
    size_t bytes = sizeof(cd) * (size_t)batch * n_known * n_cols;
    if (tq_is_device_ptr(out_moments))
      TQ_CUDA(ctx, cudaMemcpyAsync(out_moments, d_known, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    else
      TQ_CUDA(ctx, cudaMemcpyAsync(out_moments, d_known, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_err)
      for (long b = 0; b < batch; ++b) out_err[b] = 0.0;
    return TQ_OK;
  }
  // output staging: dense [batch][n_moments][n_cols]; n_moments is known only after the plan is built, reserve the maximum
  const size_t max_out = sizeof(cd) * (size_t)batch * TQ_MAX_MOMENTS * n_cols;
  const size_t err_off = (max_out + 255) & ~size_t(255);
  TQ_TRY(tq_reserve(ctx, ctx->stage_out, err_off + sizeof(double) * batch));
  double *d_err = (double *)((char *)ctx->stage_out.p + err_off);
  TQ_CUDA(ctx, cudaMemsetAsync(d_err, 0, sizeof(double) * batch, ctx->stream));
  const bool out_dev = tq_is_device_ptr(out_moments);
  cd *d_out = out_dev ? (cd *)out_moments : (cd *)ctx->stage_out.p;
  int nm = 0;
  TQ_TRY(fit_tail_device_mesh(ctx, beta, statistic, n_iw, w_min, tail_fraction, n_tail_max, expansion_order, hermitian, inner_dim, n_cols, batch,
                              (const cd *)d_g, (const cd *)d_known, n_known, d_out, &nm, d_err));
  *out_n_moments = nm;
  TQ_TRY(tq_reserve_pinned(ctx, sizeof(double) * batch));
  TQ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, d_err, sizeof(double) * batch, cudaMemcpyDeviceToHost, ctx->stream));
  if (!out_dev)
    TQ_CUDA(ctx, cudaMemcpyAsync(out_moments, d_out, sizeof(cd) * (size_t)batch * nm * n_cols, cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (out_err) memcpy(out_err, ctx->pinned, sizeof(double) * batch);
  return TQ_OK;
}

extern "C" tq_status tq_fit_tail(tq_ctx *ctx, double beta, int statistic, long n_iw, double tail_fraction, int n_tail_max, int expansion_order,
                                 int hermitian, int inner_dim, long n_cols, long batch, const tq_complex *g, const tq_complex *known,
                                 int n_known, tq_complex *out_moments, int *out_n_moments, double *out_err) {
  if (ctx && statistic != TQ_FERMION && statistic != TQ_BOSON) return tq_fail(ctx, TQ_ERR_INVALID, "tq_fit_tail: invalid argument");
  return fit_tail_impl(ctx, beta, statistic, n_iw, 0.0, tail_fraction, n_tail_max, expansion_order, hermitian, inner_dim, n_cols, batch, g, known,
                       n_known, out_moments, out_n_moments, out_err);
}

// fit_tail of a real-frequency gf: the same fitter on a refreq mesh (tail_fitter.hpp is generic in the mesh; test/c++/gfs/fit_tail_real.cpp)
extern "C" tq_status tq_fit_tail_refreq(tq_ctx *ctx, double w_min, double w_max, long n_w, double tail_fraction, int n_tail_max,
                                        int expansion_order, long n_cols, long batch, const tq_complex *g, const tq_complex *known, int n_known,
                                        tq_complex *out_moments, int *out_n_moments, double *out_err) {
  return fit_tail_impl(ctx, w_max, TAIL_MESH_REFREQ, n_w, w_min, tail_fraction, n_tail_max, expansion_order, 0, 1, n_cols, batch, g, known, n_known,
                       out_moments, out_n_moments, out_err);
}
