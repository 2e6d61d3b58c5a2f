/* tqgpu.h -- C ABI of the B200-native TRIQS Green's-function hot path (libtqgpu.so).
 *
 * Every entry point replaces one seam of the reference (TRIQS 3.3.1, paths relative
 * to its tree); the seam is cited next to the declaration.  All arrays are row-major
 * ("C order") interleaved complex<double> (re, im) exactly as the reference's
 * flattened 2-D views `[rows, n_others]` (c++/triqs/gfs/gf/flatten.hpp:33-47).
 *
 * Pointers may be host or device pointers; the library detects which
 * (cudaPointerGetAttributes).  Host buffers are staged through pinned/device memory
 * inside the call; device buffers are used in place on the context's stream.
 * Every function returns TQ_OK (0) or a negative tq_status; tq_last_error(ctx) holds
 * the message the reference would have thrown (triqs::runtime_error text,
 * c++/triqs/utility/exceptions.hpp:31-37).  There is no CPU fallback: without a CUDA
 * device tq_create fails.
 */
#ifndef TQGPU_H
#define TQGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tq_ctx tq_ctx;

/* interleaved complex<double>; layout-compatible with std::complex<double>, fftw_complex, numpy complex128 */
typedef struct { double re, im; } tq_complex;

typedef enum {
  TQ_OK = 0,
  TQ_ERR_CUDA = -1,            /* a CUDA runtime call failed (message has the cudaError string)           */
  TQ_ERR_INVALID = -2,         /* bad argument                                                            */
  TQ_ERR_RUNTIME = -3,         /* a condition on which the reference throws triqs::runtime_error          */
  TQ_ERR_UNSUPPORTED = -4,     /* shape outside what the kernels implement (e.g. FFT length with a prime
                                  factor > TQ_MAX_GENERIC_RADIX that does not fit the two-pass scheme)     */
  TQ_ERR_NO_DEVICE = -5,       /* no CUDA device: the library never computes on the CPU                   */
  TQ_ERR_COMM = -6             /* NCCL failure                                                            */
} tq_status;

/* c++/triqs/mesh/domains/matsubara.hpp: enum statistic_enum { Boson = 0, Fermion = 1 } */
enum { TQ_BOSON = 0, TQ_FERMION = 1 };

/* warning bits: the reference prints these to std::cerr; the C++ shim re-prints the same text */
enum {
  TQ_WARN_NOISY_TAU = 1,      /* fourier_matsubara.cpp:109-110 (moments cannot be deduced, zero tail used)  */
  TQ_WARN_SHORT_TAU_MESH = 2, /* fourier_matsubara.cpp:131-133 (L < 6 n_iw)                                 */
  TQ_WARN_TAIL_ERR = 4        /* fourier_matsubara.cpp:208-210 (tail error > 1e-4)                          */
};

/* ---- context ------------------------------------------------------------------------------ */

/* One context per host thread / GPU (the reference is single threaded; its FFTW planner and the
 * mutable tail_fitter cache are not thread safe: fourier_common.cpp:30, tail_fitter.hpp:310-324).
 * The context owns a stream, the plan/twiddle/tail-fitter caches keyed by mesh parameters, staging
 * buffers and (optionally) an NCCL communicator. */
tq_status tq_create(int device, tq_ctx **out);
void tq_destroy(tq_ctx *ctx);
const char *tq_last_error(const tq_ctx *ctx);
/* the CUDA stream (cudaStream_t) all work of this context is enqueued on */
void *tq_stream(tq_ctx *ctx);
/* block until all work enqueued by this context has finished */
tq_status tq_synchronize(tq_ctx *ctx);
/* Stream-ordering contract.  The context's stream is created cudaStreamNonBlocking: it does NOT order against the
 * legacy default stream or any other stream.  Calls with HOST pointers are complete when they return.  Calls with
 * DEVICE pointers are asynchronous on the context's stream: a caller that produces inputs or consumes outputs on
 * another stream (cudaStream_t `other`; NULL = the legacy default stream) brackets the call with
 *   tq_stream_wait_external(ctx, other)    -- work enqueued on the context after this waits for everything
 *                                             already enqueued on `other`
 *   tq_stream_signal_external(ctx, other)  -- work enqueued on `other` after this waits for everything already
 *                                             enqueued on the context
 * (event record + cudaStreamWaitEvent, no host synchronisation), or calls tq_synchronize.  The Python binding does
 * this automatically around every call that takes torch CUDA tensors. */
tq_status tq_stream_wait_external(tq_ctx *ctx, void *other);
tq_status tq_stream_signal_external(tq_ctx *ctx, void *other);
/* Engine options that never change results (tests run the same inputs through both engines):
 *   "no_pipe"          1: skip the TMA-pipelined / single-column specialisations, use the generic tile kernels only
 *   "force_planned"    1: the L = 10^5 mesh also runs through the planned (run-time radix) kernels instead of its compile-time
 *                      instantiation (A/B measurements; same results)
 *   "no_col"           1: scalar gfs on the 10^4 mesh do not use the single-column kernel
 *   "no_tr"            1: batches of narrow gfs (1-3 target columns) are not run as one wide virtual gf ("gfs as columns")
 *   "no_two_for_one"   1: tq_fourier_tau_to_iw_real promotes every real column to a complex one instead of packing two real columns into one
 *                      complex column (A/B measurements; results agree to 1e-13)
 *   "no_bluestein"     1: mesh lengths with a prime factor > 127 use the direct O(N^2) window DFT instead of the Bluestein (chirp-z)
 *                      convolution (A/B measurements; results agree to 1e-12)
 *   "force_bluestein"  1: every imtime <-> imfreq transform with L >= 512 uses that convolution (A/B measurements, tests)
 *   "scratch_chunk_mb" > 0: four-step scratch budget per launch pair in MiB (default: 1536 pipelined, 64 generic)
 *   "host_lanes"       1..8: internal lanes over which tq_fourier_tau_to_iw / tq_fourier_iw_to_tau cut a batch passed in HOST
 *                      buffers, so that H2D copies, kernels and D2H copies of different chunks overlap (default 3; 1 = off)
 *   "host_chunk_kb"    input KiB per chunk of such a call (default 49152 = 48 MiB)
 *   "stage_threads"    helper threads that copy a large PAGEABLE host buffer (a NumPy / nda array) through pinned slots, several pieces in flight
 *                      (0 = auto: min(8, cores / 2)); calls with pageable buffers are not cut into lanes, their copies are spread over these threads
 *   "bounce"           1 (default): inside a multi-lane call, pageable host buffers are staged through the lanes' own pinned slots,
 *                      piecewise and overlapped with the DMA (the driver's pageable path is serialised per process at ~7 GB/s);
 *                      0: always plain cudaMemcpyAsync */
tq_status tq_set_option(tq_ctx *ctx, const char *name, long value);
/* Which engine an imtime <-> imfreq transform of this shape takes, as a JSON object (host only: no context, no device needed) -- the
 * analogue of printing an FFTW plan (c++/triqs/gfs/transform/fourier_common.cpp:30 plans per call and shows nothing):
 *   {"L", "n_w", "plan_cost", "engine_fwd", "engine_inv", "plan_fwd", "plan_inv", "bluestein_M"}
 * engines: "planned" (two-radix pipelined kernels, plan = "Nb=250(10x25) Na=400(20x20) ..."), "bluestein" (chirp-z convolution of length M1 x M2),
 * "single-column", "generic" (tile kernels), "direct" (O(N^2) window sum). */
tq_status tq_plan_describe(long n_tau, long n_iw, int statistic, long n_cols, long batch, char *buf, size_t len);
/* number of kernel launches issued by this context since creation (bench.py reports the delta) */
uint64_t tq_launch_count(const tq_ctx *ctx);
/* library build id string ("tqgpu <version> sm_100a") */
const char *tq_version(void);
/* Per-kernel timing for benchmarks: while enabled, every kernel launch of this context is bracketed by CUDA
 * events on the context's stream.  tq_profile_report synchronises, writes a JSON object
 * {"kernel tag": {"launches": n, "ms": total}, ...} into buf (truncated to len) and clears the records. */
tq_status tq_profile_enable(tq_ctx *ctx, int on);
tq_status tq_profile_report(tq_ctx *ctx, char *buf, size_t len);

/* ---- Matsubara transforms ------------------------------------------------------------------ */

/* Replaces  gf_vec_t<imfreq> _fourier_impl(imfreq const&, gf_vec_cvt<imtime>, array_const_view<dcomplex,2>)
 *           c++/triqs/gfs/transform/fourier.hpp:39-40, fourier_matsubara.cpp:96-186
 * for `batch` independent gfs of identical mesh/shape (the reference loops blocks sequentially,
 * c++/triqs/gfs/block/map.hpp:35-40).
 *   gt       [batch][n_tau][n_others]
 *   moments  [batch][n_moments][n_others] known high-frequency moments, or NULL / n_moments = 0:
 *            then the tau-derivative estimate fit_tail(gt) (fourier_matsubara.cpp:39-92) is used, with
 *            the noise heuristic of :100-114.
 *   gw       [batch][n_w][n_others], n_w = 2 n_iw (Fermion) or 2 n_iw - 1 (Boson), index n - first_index
 *   warn     host int[batch] receiving TQ_WARN_* bits, or NULL
 * Errors (TQ_ERR_RUNTIME): |0th moment| >= 1e-8 (:116-118); n_tau - 1 < 2 n_iw (:127-129). */
tq_status tq_fourier_tau_to_iw(tq_ctx *ctx, double beta, int statistic, long n_tau, long n_iw, long n_others, long batch,
                               const tq_complex *gt, const tq_complex *moments, int n_moments, tq_complex *gw, int *warn);

/* The same transform along mesh N > 0 of a product mesh, i.e. what  _fourier<N>  does after flatten_2d
 * (c++/triqs/gfs/transform/fourier.hpp:78-84, gfs/gf/flatten.hpp:33-47), without the two copies: gt is [outer][n_tau][inner],
 * gw [outer][n_w][inner], and the outer x inner columns form ONE gf -- the noise check of fourier_matsubara.cpp:100-114 is a
 * maximum over all of them and *warn is a single word.  moments: [outer][n_moments][inner] or NULL. */
tq_status tq_fourier_tau_to_iw_axis(tq_ctx *ctx, double beta, int statistic, long n_tau, long n_iw, long outer, long inner,
                                    const tq_complex *gt, const tq_complex *moments, int n_moments, tq_complex *gw, int *warn);

/* Replaces  gf_vec_t<imtime> _fourier_impl(imtime const&, gf_vec_cvt<imfreq>, array_const_view<dcomplex,2>)
 *           c++/triqs/gfs/transform/fourier.hpp:47-48, fourier_matsubara.cpp:190-271.
 *   gw       [batch][n_w][n_others] on the FULL mesh (a positive-only mesh throws, :192)
 *   moments  as above; with fewer than 4 moments the least-squares fit_tail(gw, known) of the mesh's
 *            tail fitter (default parameters, tail_fitter.hpp:92,95) completes them (:203-212)
 *   gt       [batch][n_tau][n_others]
 *   tail_err host double[batch] receiving the fit error of each gf (0 if no fit), or NULL
 * Errors: 0th moment (:199-201); fit error >= 1e-2 (:205-207); < 4 moments after fit (:211); mesh too short (:218-220). */
tq_status tq_fourier_iw_to_tau(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_tau, long n_others, long batch,
                               const tq_complex *gw, const tq_complex *moments, int n_moments, tq_complex *gt, int *warn,
                               double *tail_err);

/* Real-valued G(tau)  (gf<imtime, T::real_t>).  The reference promotes a real-valued gf to complex before it transforms it
 *   c++/triqs/gfs/transform/fourier.hpp:78-81  (`gfl_t{flatten_gf_2d<N>(gin)}`, gfl_t = ...::complex_t)
 * and users take `real(g)` after `is_gf_real(g, tol)` on the way back (c++/triqs/gfs/functions/functions2.hpp:223-246).
 * These two entry points move only the 8-byte reals between the caller's buffer and the GPU (half the PCIe / HBM bytes of
 * the large side); the promotion / real-part extraction is a streaming kernel on the device.
 *   tq_fourier_tau_to_iw_real : gt double[batch][n_tau][n_others]; everything else as tq_fourier_tau_to_iw.  With 8 or more columns and moments that
 *                               are absent or real (host pointer) two real columns are transformed as ONE complex column and separated through
 *                               G(i w_-n) = conj G(i w_n): half the transform work; agrees with promoting on the host to rounding (1e-13).
 *                               Otherwise the columns are promoted on the device (same bits as promoting on the host)
 *   tq_fourier_iw_to_tau_real : gt double[batch][n_tau][n_others] receives Re G(tau); max_imag host double[batch] (required)
 *                               receives max |Im G(tau)| per gf, the quantity is_gf_real compares with its tolerance */
tq_status tq_fourier_tau_to_iw_real(tq_ctx *ctx, double beta, int statistic, long n_tau, long n_iw, long n_others, long batch,
                                    const double *gt, const tq_complex *moments, int n_moments, tq_complex *gw, int *warn);
tq_status tq_fourier_iw_to_tau_real(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_tau, long n_others, long batch,
                                    const tq_complex *gw, const tq_complex *moments, int n_moments, double *gt, int *warn,
                                    double *tail_err, double *max_imag);

/* density(G(i w)) of matrix-valued gfs on the full Matsubara mesh -- replaces
 *   nda::matrix<dcomplex> density(gf_const_view<imfreq> g, array_const_view<dcomplex, 3> known_moments)
 *     c++/triqs/gfs/functions/density.cpp:32-120   (scalar overload :123-128)
 * gw [batch][n_w][n][n]; moments [batch][n_moments][n][n] or NULL (then / with fewer than 4 rows the tail is fitted like
 * fit_tail, :40,:47); out [batch][n][n] (upper triangle summed, lower = conjugate, :115).  tail_err[batch] / warn[batch] may
 * be NULL.  Errors as the reference throws them: non-vanishing 0th moment, tail error > 1e-2, fewer than 3 fitted moments. */
tq_status tq_density(tq_ctx *ctx, double beta, int statistic, long n_iw, int n, long batch, const tq_complex *gw,
                     const tq_complex *moments, int n_moments, tq_complex *out, int *warn, double *tail_err);

/* Elementwise companions on the same data; data [batch][n_pts][d][d], d = 1 for scalar targets, d = n1*n2 for a
 * rank-4 target viewed as an (ij),(kl) matrix.  mesh_kind: 0 = full imfreq mesh (-i w_n is the mirrored point), 1 = imtime.
 *   make_hermitian     c++/triqs/gfs/functions/imfreq.hpp:175-215   out = 0.5 (g[w] + dagger(g[-w]))     (out != g)
 *   is_gf_hermitian    imfreq.hpp:81-124                            *result = 1 iff max |dagger(g[-w]) - g[w]| <= tolerance
 *   replace_by_tail    imfreq.hpp:258-261 (+ tail_eval, mesh/tail_fitter.hpp:49-64): g[iw] = sum_n tail[n] / (iw)^n for
 *                      iw.n >= n_min or iw.n < -n_min, in place; g [batch][n_w][n_cols], tail [batch][n_moments][n_cols] */
tq_status tq_make_hermitian(tq_ctx *ctx, int mesh_kind, long n_pts, int d, long batch, const tq_complex *g, tq_complex *out);
tq_status tq_is_gf_hermitian(tq_ctx *ctx, int mesh_kind, long n_pts, int d, long batch, const tq_complex *g, double tolerance, int *result);
tq_status tq_replace_by_tail(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_cols, long batch, tq_complex *g,
                             const tq_complex *tail, int n_moments, int n_min);

/* eps_k = sum_j hop[j] exp(2 pi i k . displ[j]) on the brzone mesh dims[3] -- replaces
 *   tight_binding::fourier(mesh::brzone const &)   c++/triqs/lattice/tight_binding.hpp:88-123
 * displ [n_r][3] integer lattice displacements, hop [n_r][n_orb][n_orb], out [d0 d1 d2][n_orb][n_orb] (the eps_k argument of
 * tq_dyson_ksum). */
tq_status tq_tight_binding_fourier(tq_ctx *ctx, int n_orb, int n_r, const int *displ, const tq_complex *hop, const long dims[3],
                                   tq_complex *out);

/* ---- least-squares tail fit ------------------------------------------------------------------ */

/* Replaces  tail_fitter::fit<0, hermitian>(mesh, g_data, normalize=true, known_moments, inner_matrix_dim)
 *           c++/triqs/mesh/tail_fitter.hpp:192-297 and the free functions fit_tail / fit_hermitian_tail
 *           c++/triqs/gfs/functions/functions2.hpp:50-56, 94-110.
 *   g            [batch][n_w][n_cols] on the full imfreq mesh (beta, statistic, n_iw)
 *   known        [batch][n_known][n_cols] or NULL
 *   tail_fraction, n_tail_max, expansion_order (-1 = adaptive)  tail_fitter.hpp:302-321
 *   hermitian    0 / 1;  inner_dim = d of the d x d target (n_cols must be a multiple of d*d), 1 for scalar
 *   out_moments  [batch][*out_n_moments][n_cols]; the caller provides room for TQ_MAX_MOMENTS rows per gf;
 *                rows are written densely with *out_n_moments rows
 *   out_err      host double[batch]
 * Errors: "Insufficient data points for least square procedure" (:157), "Conditioning of tail-fit violates
 * boundary" (:180), known_moments shape mismatch. */
#define TQ_MAX_MOMENTS 10
tq_status tq_fit_tail(tq_ctx *ctx, double beta, int statistic, long n_iw, double tail_fraction, int n_tail_max,
                      int expansion_order, int hermitian, int inner_dim, long n_cols, long batch, const tq_complex *g,
                      const tq_complex *known, int n_known, tq_complex *out_moments, int *out_n_moments, double *out_err);

/* fit_tail of a real-frequency gf: the same fitter on mesh::refreq{w_min, w_max, n_w} (tail_fitter.hpp:105-181 is generic in the
 * mesh: first_index 0, w_max() = xmax, real mesh values; known answers test/c++/gfs/fit_tail_real.cpp:24-91).
 *   g [batch][n_w][n_cols], known [batch][n_known][n_cols] or NULL; outputs as tq_fit_tail. */
tq_status tq_fit_tail_refreq(tq_ctx *ctx, double w_min, double w_max, long n_w, double tail_fraction, int n_tail_max, int expansion_order,
                             long n_cols, long batch, const tq_complex *g, const tq_complex *known, int n_known, tq_complex *out_moments,
                             int *out_n_moments, double *out_err);

/* The mesh-only part of the fit (fit indices, Vandermonde, SVD pseudo-inverse, chosen order), i.e. what the
 * reference caches in tail_fitter::_lss (tail_fitter.hpp:140-181).  Pure host code (plan creation, no GPU
 * needed, no ctx): exposed so that tests can compare it with LAPACK.
 *   idx_out   long[2 n_tail]                   fit indices (Matsubara n)
 *   pinv_out  [n_var][rows] complex, rows = n_idx (plain) or 2 n_idx (hermitian: stacked [A; conj A])
 *   sv_out    double[n_var] singular values of A (plain, un-stacked; used for the order choice)
 * Buffers must hold 2*n_tail_max indices, TQ_MAX_MOMENTS * 4 * n_tail_max complex, TQ_MAX_MOMENTS doubles. */
tq_status tq_tail_fitter_setup_host(double beta, int statistic, long n_iw, double tail_fraction, int n_tail_max,
                                    int expansion_order, int n_fixed, int hermitian, int *n_idx_out, long *idx_out,
                                    int *n_var_out, tq_complex *pinv_out, double *sv_out, char *errbuf, size_t errbuf_len);

/* ---- retime <-> refreq ----------------------------------------------------------------------------- */

/* Replace  _fourier_impl(refreq const&, gf_vec_cvt<retime>, moments)   c++/triqs/gfs/transform/fourier_real.cpp:35-78
 *     and  _fourier_impl(retime const&, gf_vec_cvt<refreq>, moments)   c++/triqs/gfs/transform/fourier_real.cpp:82-131.
 *   meshes    retime{t_min, t_max, n_pts}, refreq{w_min, w_max, n_pts} (mesh/retime.hpp:39, refreq.hpp:38); they must be
 *             adjoint (delta_t delta_w n_pts = 2 pi within 1e-10, else TQ_ERR_RUNTIME "Meshes are not compatible")
 *   gt, gw    [batch][n_pts][n_others]
 *   moments   [batch][n_moments][n_others] known moments (n_moments >= 3, the 0th must vanish) or NULL: zero tail. */
tq_status tq_fourier_t_to_w(tq_ctx *ctx, double t_min, double t_max, double w_min, double w_max, long n_pts, long n_others, long batch,
                            const tq_complex *gt, const tq_complex *moments, int n_moments, tq_complex *gw);
tq_status tq_fourier_w_to_t(tq_ctx *ctx, double t_min, double t_max, double w_min, double w_max, long n_pts, long n_others, long batch,
                            const tq_complex *gw, const tq_complex *moments, int n_moments, tq_complex *gt);

/* ---- Legendre <-> Matsubara ------------------------------------------------------------------------ */

/* Replace  legendre_matsubara_direct(gw | gt, gl)  and  legendre_matsubara_inverse(gl, gw | gt)
 *          c++/triqs/gfs/transform/legendre_matsubara.hpp:37-50, :56-70, :76-93, :97-118   (T_{nl}: utility/legendre.cpp:24-42)
 *   gl [batch][n_l][n_others] (n_l <= 128),  gw [batch][n_w][n_others] on the full imfreq mesh,  gt [batch][n_tau][n_others].
 * tq_imfreq_to_legendre goes through G(tau) on 50000 points like the reference (:85), the tail is fitted by that transform. */
tq_status tq_legendre_to_imfreq(tq_ctx *ctx, int statistic, long n_l, long n_iw, long n_others, long batch, const tq_complex *gl, tq_complex *gw);
tq_status tq_legendre_to_imtime(tq_ctx *ctx, double beta, long n_l, long n_tau, long n_others, long batch, const tq_complex *gl, tq_complex *gt);
tq_status tq_imtime_to_legendre(tq_ctx *ctx, double beta, long n_tau, long n_l, long n_others, long batch, const tq_complex *gt, tq_complex *gl);
tq_status tq_imfreq_to_legendre(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_l, long n_others, long batch, const tq_complex *gw,
                                tq_complex *gl);

/* ---- plain batched DFT and lattice transforms --------------------------------------------------- */

/* Replaces  _fourier_base(in, out, rank, dims, howmany, sign)   c++/triqs/gfs/transform/fourier_common.cpp:25-45
 * (fftw_plan_many_dft with stride = howmany, dist = 1): unnormalised DFT of rank 1..3 over the leading
 * (C-order flattened) axis of [prod(dims)][howmany]; sign = +1 is FFTW_BACKWARD (e^{+2 pi i jk/N}), -1 FORWARD.
 * in == out (in place) is allowed. */
tq_status tq_fft_many(tq_ctx *ctx, int rank, const int *dims, long howmany, const tq_complex *in, tq_complex *out, int sign);

/* Replaces  _fourier_impl(brzone const&, gf_vec_cvt<cyclat>) / _fourier_impl(cyclat const&, gf_vec_cvt<brzone>)
 *           c++/triqs/gfs/transform/fourier_lattice.cpp:30-59.
 *   to_k = 1: cyclat -> brzone (BACKWARD, no scaling);  to_k = 0: brzone -> cyclat (FORWARD, then / N_k).
 *   in, out  [d0*d1*d2][n_others], linear index i0*d1*d2 + i1*d2 + i2 (mesh/brzone.hpp:187-190). */
tq_status tq_fourier_lattice(tq_ctx *ctx, const long dims[3], int to_k, long n_others, const tq_complex *in, tq_complex *out);

/* ---- matrix inversion and the lattice Dyson k-sum ------------------------------------------------ */

/* Replaces  invert_in_place(gf_view<M, matrix_valued>)  c++/triqs/gfs/functions/functions2.hpp:197-201
 *           _gf_invert_data_in_place(array_view<dcomplex,3>)  c++/triqs/gfs/gf/gf_expr.hpp:219-222
 * a [n_pts][n][n], inverted in place, 1 <= n <= TQ_MAX_MATRIX_DIM (Gauss-Jordan with partial pivoting; n <= 8: one thread per
 * matrix entirely in registers, 9 <= n <= 64: one warp per matrix in shared memory). */
#define TQ_MAX_MATRIX_DIM 64
tq_status tq_invert_batched(tq_ctx *ctx, int n, long n_pts, tq_complex *a);

/* Replaces the k-sum of  SumkDiscrete.__call__   python/triqs/sumk/sumk_discrete.py:140-166
 * (C++ spelling: clef fill + inverse + sum over k, test/c++/gfs/multivar/g_k_om.cpp:64, bubble.cpp:65):
 *   out[w] = sum_k weights[k] * inverse( (i w_n + mu) 1 - eps_k[k] - sigma[w] )
 * fused: G(k, w) is never materialised (n <= 8; larger targets go through slabs of k-points in a <= 256 MB scratch).
 *   eps_k    [n_k][n][n]   this rank's k-points (the caller slices like mpi.slice_array, mpi_mpi4py.py:96-109)
 *   weights  double[n_k] or NULL (then every weight is `uniform_weight`, e.g. 1/N_k_total)
 *   sigma    [n_w][n][n] on the full imfreq mesh (beta, statistic, n_iw)
 *   out      [n_w][n][n]
 *   all_reduce != 0: sum `out` over the ranks of the context's communicator (tq_comm_init) in place,
 *                    the equivalent of  G << mpi.all_reduce(G)  (sumk_discrete.py:163). */
tq_status tq_dyson_ksum(tq_ctx *ctx, int n, long n_k, long n_iw, double beta, int statistic, double mu, const tq_complex *eps_k,
                        const double *weights, double uniform_weight, const tq_complex *sigma, tq_complex *out, int all_reduce);

/* ---- off-mesh evaluation ----------------------------------------------------------------------- */

/* Replaces  gf_evaluator<imtime>::operator()(g, double tau)  c++/triqs/gfs/evaluator.hpp:35-43 with
 *           linear::evaluate  c++/triqs/mesh/bases/linear.hpp:221-228, batched over n_pts points.
 *   table [n_tau][n_elem] (one gf block), taus double[n_pts] in [0, beta], out [n_pts][n_elem]. */
tq_status tq_interp_tau(tq_ctx *ctx, double beta, long n_tau, long n_elem, const tq_complex *table, long n_pts, const double *taus,
                        tq_complex *out);

/* Replaces  gf_evaluator<imfreq>::operator()(g, matsubara_freq)  c++/triqs/gfs/evaluator.hpp:47-75, batched: g(i w_n) for arbitrary
 * Matsubara indices ns[n_pts] -- the mesh value inside the window, the fitted tail  sum_p A_p / (i w_n)^p  outside (one default
 * fit_tail per gf, tail_fitter.hpp:92,95; the reference refits on every call).
 *   g [batch][n_w][n_cols] on the full mesh, ns long[n_pts] (host or device), out [batch][n_pts][n_cols],
 *   tail_err host double[batch] or NULL (0 when every index is inside the mesh and no fit was made). */
tq_status tq_evaluate_imfreq(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_cols, long batch, const tq_complex *g, long n_pts,
                             const long *ns, tq_complex *out, double *tail_err);

/* Replaces the product-mesh evaluation  evaluate(mesh::prod<M...>, f, x...)  c++/triqs/mesh/evaluate.hpp:97-114  (what
 * g(k, w), g(t, t'), g(tau1, tau2) ... call for a gf on a product mesh), batched over n_pts points: multi-linear interpolation,
 * one factor per component mesh, combined in the reference's recursion order (first component innermost).  Per component:
 *   TQ_AXIS_LINEAR    linear::evaluate  c++/triqs/mesh/bases/linear.hpp:221-228  (imtime, retime, refreq): coordinate = the value x in
 *                     [xmin, xmax], `size` points;
 *   TQ_AXIS_PERIODIC  evaluate(brzone1d, f, vi)  c++/triqs/mesh/brzone.hpp:355-359: coordinate = vi in units of the mesh index (the
 *                     caller applies  transpose(units_inv) * k  of brzone.hpp:368; a brzone is three such components), periodic;
 *   TQ_AXIS_INDEX     an index passes through (evaluate.hpp:73): coordinate = the data index along this component (imfreq: n - first).
 *   table [size_0]...[size_{n_dims-1}][n_elem], coords double[n_pts][n_dims], out [n_pts][n_elem]. */
#define TQ_PROD_MAX_DIMS 4
enum { TQ_AXIS_LINEAR = 0, TQ_AXIS_PERIODIC = 1, TQ_AXIS_INDEX = 2 };
typedef struct {
  int kind;
  long size;
  double xmin, xmax; /* TQ_AXIS_LINEAR only */
} tq_axis;
tq_status tq_evaluate_prod(tq_ctx *ctx, int n_dims, const tq_axis *axes, long n_elem, const tq_complex *table, long n_pts,
                           const double *coords, tq_complex *out);

/* ---- multi-GPU (one process per GPU; replaces mpi::all_reduce of gf data, c++/triqs/gfs/gf/gf.hpp:347-351) --- */

#define TQ_COMM_ID_BYTES 128
/* rank 0 creates the id and broadcasts the bytes by whatever means the host program has (MPI_Bcast in a
 * TRIQS application, torch.distributed in this repo's harness); then every rank calls tq_comm_init. */
tq_status tq_comm_unique_id(unsigned char id[TQ_COMM_ID_BYTES]);
tq_status tq_comm_init(tq_ctx *ctx, int n_ranks, int rank, const unsigned char id[TQ_COMM_ID_BYTES]);
/* in-place sum over ranks of `count` doubles at device pointer `data` (NCCL all-reduce on the ctx stream) */
tq_status tq_all_reduce_sum(tq_ctx *ctx, double *data, size_t count);

/* ---- device memory helpers for hosts that want to keep gfs resident in HBM ------------------------ */
tq_status tq_device_alloc(tq_ctx *ctx, size_t bytes, void **dptr);
tq_status tq_device_free(tq_ctx *ctx, void *dptr);
tq_status tq_memcpy_h2d(tq_ctx *ctx, void *dst_device, const void *src_host, size_t bytes);
tq_status tq_memcpy_d2h(tq_ctx *ctx, void *dst_host, const void *src_device, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* TQGPU_H */
