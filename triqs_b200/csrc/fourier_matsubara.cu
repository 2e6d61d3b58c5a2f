// imtime <-> imfreq transforms, fused:  tail model subtraction + phase twist -> tile FFT -> correction + window gather.
//
// Restates (does not copy) c++/triqs/gfs/transform/fourier_matsubara.cpp:
//   :39-92   fit_tail(gt)          -> k_moments_from_tau   (8-point one-sided differences x literal V_inv)
//   :96-186  _fourier_impl(imfreq) -> matsubara_tile_kernel<IN_TAU, OUT_FREQ | OUT_SCRATCH>, <IN_SCRATCH, OUT_FREQ>
//   :190-271 _fourier_impl(imtime) -> matsubara_tile_kernel<IN_FREQ, OUT_TAU | OUT_SCRATCH>, <IN_SCRATCH, OUT_TAU>
// A transform whose columns fit one SM's shared memory runs as ONE kernel (HBM is read once and written
// once); longer ones (L = 10^5) run as a two-kernel four-step FFT L = Na * Nb with an L2-resident scratch.
#include "tq_fft.cuh"
#include "tq_pipe.cuh" // tile_shape (Bluestein picks axis lengths the plain engine has a plan for)

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

// -------------------------------------------------------------------------------------------------
// device-side description of the meshes and the 3-pole tail model
// -------------------------------------------------------------------------------------------------
struct MatsubaraDev {
  int L, n_w, first, last, stat, nhi, lo_shift;
  double beta, fact_fwd, inv_beta;
  double b[3], C[3];
  int rev[3];
  const cd *ph_hi_fwd; // (beta/L) exp(+i pi (h << lo_shift) / L)
  const cd *ph_hi_inv; // exp(-i pi (h << lo_shift) / L)
  const cd *ph_lo;     // exp(+i pi l / L)
  const double *e_hi;  // [3][nhi]   exp(-|b_i| beta (h << lo_shift) / L)
  const double *e_lo;  // [3][1 << lo_shift]
  const double4 *wtab; // [n_w] (omega_n, 1/(b1^2+w^2), 1/(b2^2+w^2), 1/(b3^2+w^2))
};

struct BigTw {
  const cd *hi; // exp(+2 pi i (h << shift) / L)
  const cd *lo; // exp(+2 pi i l / L)
  int shift;
};

struct TileArgs {
  FftPlan P;
  int TC, n_cols, skewmask, direct_tw, smem_tw, rowtab;
  int in_stride, out_stride; // global row = item + stride * local
  int Na;                    // scratch layout [Nb][Na][n_cols]
  const cd *in;
  cd *out;
  cd *scratch;
  const cd *gt_ref;  // G(tau) input (for the trapezoid correction, forward only)
  const cd *moments; // [batch][mom_rows][n_cols]; rows 1..3 are used
  long in_gf_stride, out_gf_stride, scratch_gf_stride, mom_gf_stride, gt_gf_stride;
  MatsubaraDev D;
  BigTw btw;
  long long *dbg; // developer phase timing (TQ_PHASE_DEBUG): 6 clock64 stamps per CTA, or nullptr
};

enum { IN_TAU = 0, IN_FREQ = 1, IN_SCRATCH = 2 };
enum { OUT_FREQ = 0, OUT_TAU = 1, OUT_SCRATCH = 2 };

// h_i(tau_j): one{Fermion,Boson}(a, b_i, tau_j, beta) = a * h_i(tau_j)     fourier_matsubara.cpp:28-34
TQ_D void row_model(const MatsubaraDev &D, int j, double h[3]) {
  const int mask = (1 << D.lo_shift) - 1;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    int jj = D.rev[i] ? D.L - j : j;
    h[i] = D.C[i] * __ldg(&D.e_hi[i * D.nhi + (jj >> D.lo_shift)]) * __ldg(&D.e_lo[(i << D.lo_shift) + (jj & mask)]);
  }
}
TQ_D cd phase_fwd(const MatsubaraDev &D, int j) {
  const int mask = (1 << D.lo_shift) - 1;
  return cmul(__ldg(&D.ph_hi_fwd[j >> D.lo_shift]), __ldg(&D.ph_lo[j & mask]));
}
TQ_D cd phase_inv(const MatsubaraDev &D, int j) {
  const int mask = (1 << D.lo_shift) - 1;
  return cmul(__ldg(&D.ph_hi_inv[j >> D.lo_shift]), cconj(__ldg(&D.ph_lo[j & mask])));
}
TQ_D double4 load_w(const MatsubaraDev &D, int di) {
  const double2 *p = reinterpret_cast<const double2 *>(D.wtab + di);
  double2 a = __ldg(p), b = __ldg(p + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
// sum_i a_i / (i w - b_i) = -S2 - i w S1,  S1 = sum a_i r_i,  S2 = sum a_i b_i r_i,  r_i = 1/(b_i^2 + w^2)
TQ_D cd tail_iw(const MatsubaraDev &D, double4 w, cd a1, cd a2, cd a3) {
  double r1 = w.y, r2 = w.z, r3 = w.w;
  cd S1 = caxpy(r3, a3, caxpy(r2, a2, cscale(r1, a1)));
  cd S2 = caxpy(r3 * D.b[2], a3, caxpy(r2 * D.b[1], a2, cscale(r1 * D.b[0], a1)));
  return cmk(fma(w.x, S1.y, -S2.x), -fma(w.x, S1.x, S2.y));
}
TQ_D cd bigtw(const BigTw &t, int m) { return cmul(__ldg(&t.hi[m >> t.shift]), __ldg(&t.lo[m & ((1 << t.shift) - 1)])); }

// freq-space index k in [0, L) -> mesh data index, or -1 when k is outside the window  (:182, :254  "(iw.n + L) % L")
TQ_D int freq_data_index(const MatsubaraDev &D, int k) {
  int n = (k <= D.last) ? k : k - D.L;
  return (n >= D.first && n <= D.last) ? n - D.first : -1;
}

// Per-row factors.  Everything that depends only on the mesh row (not on the column) is evaluated once per
// row into three double2 -- either into a shared-memory row table (tiles with several columns: the
// table look-ups and the index math are amortised over the row) or inline (single-column tiles).
//   input rows  : IN_TAU   {h0,h1} {h2,-} {phase}          IN_FREQ  {w,r1} {r2,r3} {data index,-}
//   output rows : OUT_FREQ {w,r1} {r2,r3} {data index,pos}  OUT_TAU  {h0,h1} {h2,pos} {phase}   OUT_SCRATCH {twiddle} {pos,-} -
template <int IN> TQ_D void in_row_entry(const TileArgs &A, int item, int r, cd e[3]) {
  const MatsubaraDev &D = A.D;
  if (IN == IN_TAU) {
    const int j = item + A.in_stride * r;
    double h[3];
    row_model(D, j, h);
    e[0] = cmk(h[0], h[1]);
    e[1] = cmk(h[2], 0.0);
    e[2] = D.stat == TQ_FERMION ? phase_fwd(D, j) : cmk(D.fact_fwd, 0.0);
  } else if (IN == IN_FREQ) {
    const int di = freq_data_index(D, item + A.in_stride * r);
    double4 w = di >= 0 ? load_w(D, di) : make_double4(0, 0, 0, 0);
    e[0] = cmk(w.x, w.y);
    e[1] = cmk(w.z, w.w);
    e[2] = cmk((double)di, 0.0);
  }
}
template <int OUT, int SGN> TQ_D void out_row_entry(const TileArgs &A, int item, int kk, cd e[3]) {
  const MatsubaraDev &D = A.D;
  const double pos = (double)__ldg(&A.P.pos[kk]);
  if (OUT == OUT_FREQ) {
    const int di = freq_data_index(D, item + A.out_stride * kk);
    double4 w = di >= 0 ? load_w(D, di) : make_double4(0, 0, 0, 0);
    e[0] = cmk(w.x, w.y);
    e[1] = cmk(w.z, w.w);
    e[2] = cmk((double)di, pos);
  } else if (OUT == OUT_TAU) {
    const int j = item + A.out_stride * kk;
    double h[3];
    row_model(D, j, h);
    e[0] = cmk(h[0], h[1]);
    e[1] = cmk(h[2], pos);
    e[2] = D.stat == TQ_FERMION ? phase_inv(D, j) : cmk(1.0, 0.0);
  } else {
    cd w = bigtw(A.btw, item * kk);
    e[0] = SGN < 0 ? cconj(w) : w;
    e[1] = cmk(pos, 0.0);
    e[2] = cmk(0.0, 0.0);
  }
}
// sum_i a_i / (i w - b_i) from a row entry {w,r1} {r2,r3}
TQ_D cd tail_iw_e(const MatsubaraDev &D, cd e0, cd e1, cd a1, cd a2, cd a3) {
  return tail_iw(D, make_double4(e0.x, e0.y, e1.x, e1.y), a1, a2, a3);
}

template <int IN, int OUT, int SGN, bool SKEW, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) matsubara_tile_kernel(const TileArgs A) {
  extern __shared__ double2 smem[];
  const MatsubaraDev &D = A.D;
  const int item = blockIdx.x, gf = blockIdx.z;
  const int c0 = blockIdx.y * A.TC;
  const int tc = min(A.TC, A.n_cols - c0);
  const int N = A.P.N;
  const int pitch = tc;
  constexpr int sk = SKEW ? ~0 : 0;
  cd *tile = smem;
  cd *coef = tile + fft_rows_alloc(N, sk) * A.TC; // a1[tc], a2[tc], a3[tc], extra[tc]
  cd *twS = coef + 4 * A.TC;                      // [N] copy of the twiddle table (if A.smem_tw)
  cd *rowtab = twS + (A.smem_tw ? N : 0);         // [N][3] per-row factors (if A.rowtab)
  const bool fermion = D.stat == TQ_FERMION;
  const bool use_rt = A.rowtab != 0;
  const TileThread T = tile_thread(threadIdx.x, NT, tc);
  const int n_cols = A.n_cols;
  long long *dbg = A.dbg ? A.dbg + 6 * ((size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z)) : nullptr;
#define TQ_STAMP(i)                                                                                                              \
  if (dbg && threadIdx.x == 0) dbg[i] = clock64();
  TQ_STAMP(0)

  // ---- per-column pole weights from the moments  (:151-169, :238-252)
  {
    const cd *mom = A.moments + (size_t)gf * A.mom_gf_stride;
    for (int c = threadIdx.x; c < tc; c += NT) {
      cd m1 = mom[(size_t)1 * n_cols + c0 + c], m2 = mom[(size_t)2 * n_cols + c0 + c], m3 = mom[(size_t)3 * n_cols + c0 + c];
      cd a1, a2, a3;
      if (fermion) {
        a1 = csub(m1, m3);
        a2 = cscale(0.5, cadd(m2, m3));
        a3 = cscale(0.5, csub(m3, m2));
      } else {
        a1 = cscale(4.0 / 3.0, csub(m1, m3));
        a2 = csub(m3, cscale(0.5, cadd(m1, m2)));
        a3 = cadd(cadd(cscale(1.0 / 6.0, m1), cscale(0.5, m2)), cscale(1.0 / 3.0, m3));
      }
      coef[c] = a1;
      coef[tc + c] = a2;
      coef[2 * tc + c] = a3;
      if (OUT == OUT_FREQ) {
        // trapezoid end-point correction  -0.5 fact (g_0 + m1 -/+ ... g_L)   (:181)
        const cd *g = A.gt_ref + (size_t)gf * A.gt_gf_stride;
        cd g0 = g[c0 + c], gL = g[(size_t)D.L * n_cols + c0 + c];
        cd t = cadd(g0, m1);
        t = fermion ? cadd(t, gL) : csub(t, gL);
        coef[3 * tc + c] = cscale(-0.5 * D.fact_fwd, t);
      } else {
        coef[3 * tc + c] = m1;
      }
    }
    if (A.smem_tw)
      for (int i = threadIdx.x; i < N; i += NT) twS[i] = __ldg(&A.P.tw[i]);
    if (use_rt && IN != IN_SCRATCH)
      for (int r = threadIdx.x; r < N; r += NT) in_row_entry<IN>(A, item, r, rowtab + 3 * r);
    __syncthreads();
  }
  TQ_STAMP(1)

  // ---- load + pre-transform.  LU independent 16-byte loads per thread are issued before any is consumed:
  // with 1-2 CTAs per SM that is what keeps enough bytes in flight to cover the HBM latency.
  constexpr int LU = 8;
  if (T.nbt > 0) {
    for (int c = T.ct; c < tc; c += T.nct) {
      const cd a1 = coef[c], a2 = coef[tc + c], a3 = coef[2 * tc + c];
      // global row pointer of local row r:  base + r * gstep   (IN_TAU: row j = item + in_stride r; IN_SCRATCH: row r)
      const cd *gbase;
      size_t gstep;
      if (IN == IN_TAU) {
        gbase = A.in + (size_t)gf * A.in_gf_stride + (size_t)item * n_cols + c0 + c;
        gstep = (size_t)A.in_stride * n_cols;
      } else if (IN == IN_FREQ) {
        gbase = A.in + (size_t)gf * A.in_gf_stride + c0 + c;
        gstep = (size_t)n_cols;
      } else {
        gbase = A.scratch + (size_t)gf * A.scratch_gf_stride + (size_t)item * A.Na * n_cols + c0 + c;
        gstep = (size_t)n_cols;
      }
      for (int r0 = T.bt; r0 < N; r0 += T.nbt * LU) {
        cd g[LU];
#pragma unroll
        for (int u = 0; u < LU; ++u) {
          const int r = r0 + u * T.nbt;
          g[u] = cmk(0.0, 0.0);
          if (r < N) {
            if (IN == IN_FREQ) {
              const int di = freq_data_index(D, item + A.in_stride * r);
              if (di >= 0) g[u] = gbase[(size_t)di * gstep];
            } else {
              g[u] = gbase[(size_t)r * gstep];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < LU; ++u) {
          const int r = r0 + u * T.nbt;
          if (r >= N) continue;
          cd x = g[u];
          if (IN != IN_SCRATCH) {
            cd e[3];
            if (use_rt) {
              const cd *rt = rowtab + 3 * r;
              e[0] = rt[0];
              e[1] = rt[1];
              e[2] = rt[2];
            } else {
              in_row_entry<IN>(A, item, r, e);
            }
            if (IN == IN_TAU) {
              x = caxpy(-e[1].x, a3, caxpy(-e[0].y, a2, caxpy(-e[0].x, a1, x)));
              x = fermion ? cmul(e[2], x) : cscale(e[2].x, x);
            } else if (e[2].x >= 0.0) {
              x = cscale(D.inv_beta, csub(x, tail_iw_e(D, e[0], e[1], a1, a2, a3)));
            }
          }
          tile[fft_row(r, sk) * pitch + c] = x;
        }
      }
    }
  }
  __syncthreads();
  TQ_STAMP(2)

  // output-row factors overwrite the input-row table (no longer needed); the passes below do not touch it
  if (use_rt)
    for (int kk = threadIdx.x; kk < N; kk += NT) out_row_entry<OUT, SGN>(A, item, kk, rowtab + 3 * kk);

  fft_tile<SGN, SKEW>(A.P, tile, pitch, tc, A.smem_tw ? twS : A.P.tw, A.direct_tw != 0); // ends with __syncthreads()
  if (A.P.npass == 0) __syncthreads();
  TQ_STAMP(3)

  // ---- post-transform + store (applies the digit-reversal permutation)
  if (T.nbt > 0) {
    for (int c = T.ct; c < tc; c += T.nct) {
      const cd a1 = coef[c], a2 = coef[tc + c], a3 = coef[2 * tc + c], ex = coef[3 * tc + c];
      cd *obase;
      if (OUT == OUT_SCRATCH)
        obase = A.scratch + (size_t)gf * A.scratch_gf_stride + (size_t)item * n_cols + c0 + c; // + kk * Na * n_cols
      else
        obase = A.out + (size_t)gf * A.out_gf_stride + c0 + c;
      const size_t sstep = (size_t)A.Na * n_cols;
      for (int kk = T.bt; kk < N; kk += T.nbt) {
        cd e[3];
        if (use_rt) {
          const cd *rt = rowtab + 3 * kk;
          e[0] = rt[0];
          e[1] = rt[1];
          e[2] = rt[2];
        } else {
          out_row_entry<OUT, SGN>(A, item, kk, e);
        }
        if (OUT == OUT_FREQ) {
          const int di = (int)e[2].x;
          if (di < 0) continue;
          cd v = tile[fft_row((int)e[2].y, sk) * pitch + c];
          cd t = tail_iw_e(D, e[0], e[1], a1, a2, a3);
          obase[(size_t)di * n_cols] = cadd(cadd(v, ex), t);
        } else if (OUT == OUT_TAU) {
          const int j = item + A.out_stride * kk;
          cd v = tile[fft_row((int)e[1].y, sk) * pitch + c];
          if (fermion) v = cmul(v, e[2]);
          v = caxpy(e[1].x, a3, caxpy(e[0].y, a2, caxpy(e[0].x, a1, v)));
          obase[(size_t)j * n_cols] = v;
          if (j == 0) { // gt[L] = -/+ (gt[0] + m1)    (:267-268)
            cd t = cadd(v, ex);
            obase[(size_t)D.L * n_cols] = fermion ? cneg(t) : t;
          }
        } else {
          cd v = tile[fft_row((int)e[1].x, sk) * pitch + c];
          obase[(size_t)kk * sstep] = cmul(v, e[0]);
        }
      }
    }
  }
  __syncthreads();
  TQ_STAMP(4)
#undef TQ_STAMP
}

// -------------------------------------------------------------------------------------------------
// Direct evaluation for mesh lengths the tile FFT cannot factor (a prime factor > 128, e.g. L = 49999 of the 50000-point
// mesh legendre_matsubara.hpp:85 goes through).  Only n_w of the L frequencies are non-zero / wanted (:182, :254), so the
// window DFT is a skinny matrix product with n_w L n_cols terms, the matrix e^{+-2 pi i jk/L} generated from the two-level
// twiddle table.  Same pre / post arithmetic as the tile kernel.  The long sum is split into chunks whose partial sums
// are added in fixed order (deterministic).
// -------------------------------------------------------------------------------------------------
TQ_D void pole_weights(bool fermion, cd m1, cd m2, cd m3, cd &a1, cd &a2, cd &a3) { // :151-169, :238-252
  if (fermion) {
    a1 = csub(m1, m3);
    a2 = cscale(0.5, cadd(m2, m3));
    a3 = cscale(0.5, csub(m3, m2));
  } else {
    a1 = cscale(4.0 / 3.0, csub(m1, m3));
    a2 = csub(m3, cscale(0.5, cadd(m1, m2)));
    a3 = cadd(cadd(cscale(1.0 / 6.0, m1), cscale(0.5, m2)), cscale(1.0 / 3.0, m3));
  }
}

struct DirectArgs {
  MatsubaraDev D;
  BigTw btw;
  int n_cols, n_in, n_out, chunk, n_chunks; // rows summed over / produced; inner chunking
  const cd *in;
  cd *out;
  cd *x;       // [gf][n_in][n_cols] pre-transformed input
  cd *partial; // [gf][n_chunks][n_out][n_cols]
  const cd *moments;
  long in_gf_stride, out_gf_stride, mom_gf_stride;
  // Bluestein: the pre kernel writes x chirp into rows [0, n_in) of z[gf][M][n_cols] and zeros into the rest; the post kernel reads
  // row o of z times post_mul[o] instead of the partial sums
  const cd *pre_mul, *post_mul;
  long z_rows;
  int z_wide; // 1: the convolution buffer is [row][gf][column] (all gfs of the launch side by side: narrow targets), 0: [gf][row][column]
};

// x_j = phase_j (g_j - sum_i a_i h_i(tau_j))  (FWD, :159-173)   or   y_di = (g_di - tail(i w)) / beta  (:254)
template <bool FWD> __global__ void k_direct_pre(const DirectArgs A) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long per_gf = (long)A.n_in * A.n_cols;
  if (A.pre_mul && i >= per_gf) { // zero padding of the convolution buffer
    if (i < A.z_rows * A.n_cols) {
      const long rr = i / A.n_cols;
      const size_t zi = A.z_wide ? ((size_t)rr * gridDim.y + blockIdx.y) * A.n_cols + (size_t)(i - rr * A.n_cols) : (size_t)blockIdx.y * A.z_rows * A.n_cols + i;
      A.x[zi] = cmk(0.0, 0.0);
    }
    return;
  }
  if (i >= per_gf) return;
  const int gf = blockIdx.y, r = (int)(i / A.n_cols), c = (int)(i - (long)r * A.n_cols);
  const MatsubaraDev &D = A.D;
  const bool fermion = D.stat == TQ_FERMION;
  const cd *mom = A.moments + (size_t)gf * A.mom_gf_stride;
  cd a1, a2, a3;
  pole_weights(fermion, mom[(size_t)1 * A.n_cols + c], mom[(size_t)2 * A.n_cols + c], mom[(size_t)3 * A.n_cols + c], a1, a2, a3);
  cd x = A.in[(size_t)gf * A.in_gf_stride + i];
  if (FWD) {
    double h[3];
    row_model(D, r, h);
    x = caxpy(-h[2], a3, caxpy(-h[1], a2, caxpy(-h[0], a1, x)));
    x = fermion ? cmul(phase_fwd(D, r), x) : cscale(D.fact_fwd, x);
  } else {
    x = cscale(D.inv_beta, csub(x, tail_iw(D, load_w(D, r), a1, a2, a3)));
  }
  if (A.pre_mul)
    A.x[A.z_wide ? ((size_t)r * gridDim.y + gf) * A.n_cols + c : (size_t)gf * A.z_rows * A.n_cols + i] = cmul(x, A.pre_mul[r]);
  else
    A.x[(size_t)gf * per_gf + i] = x;
}

// The same for SCALAR gfs in the wide layout z[row][gf] (Bluestein, z_wide, n_cols = 1): a 32 x 32 tile goes through shared memory so that both the
// reads (along the rows of a gf) and the writes (along the gfs of a row) are coalesced -- written directly, every 16-byte element lands in a
// sector of its own (1.4 ms instead of 0.4 ms for the C1-sized batch).
template <bool FWD> __global__ void __launch_bounds__(256) k_direct_pre_wide1(const DirectArgs A, int n_gf) {
  __shared__ cd tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
  const long r0 = (long)blockIdx.x * 32;
  const int g0 = blockIdx.y * 32;
  const MatsubaraDev &D = A.D;
  const bool fermion = D.stat == TQ_FERMION;
  const long r = r0 + tx;
  double h[3] = {0, 0, 0};
  cd ph = cmk(0.0, 0.0), pm = cmk(0.0, 0.0);
  double4 wrow = make_double4(0, 0, 0, 0);
  const bool live = r < A.n_in;
  if (live) { // per-row factors once per thread (the four gfs it touches share them)
    if (FWD) {
      row_model(D, (int)r, h);
      ph = fermion ? phase_fwd(D, (int)r) : cmk(D.fact_fwd, 0.0);
    } else {
      wrow = load_w(D, (int)r);
    }
    pm = A.pre_mul[r];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int g = g0 + ty + 8 * k;
    cd v = cmk(0.0, 0.0);
    if (g < n_gf && live) {
      const cd *mom = A.moments + (size_t)g * A.mom_gf_stride;
      cd a1, a2, a3;
      pole_weights(fermion, mom[1], mom[2], mom[3], a1, a2, a3);
      cd x = A.in[(size_t)g * A.in_gf_stride + r];
      if (FWD) {
        x = caxpy(-h[2], a3, caxpy(-h[1], a2, caxpy(-h[0], a1, x)));
        x = cmul(ph, x);
      } else {
        x = cscale(D.inv_beta, csub(x, tail_iw(D, wrow, a1, a2, a3)));
      }
      v = cmul(x, pm);
    }
    tile[ty + 8 * k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long rr = r0 + ty + 8 * k;
    const int g = g0 + tx;
    if (rr < A.z_rows && g < n_gf) A.x[(size_t)rr * n_gf + g] = tile[tx][ty + 8 * k];
  }
}

// partial[chunk][o][c] = sum_{i in chunk} x[i][c] w^{+-(k j)},  FWD: o = frequency data index, i = j;  inverse: o = j, i = data index
constexpr int DIRECT_CW = 8;
template <bool FWD> __global__ void __launch_bounds__(128) k_direct_dft(const DirectArgs A) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;
  const int gf = blockIdx.z / ((A.n_cols + DIRECT_CW - 1) / DIRECT_CW), ct = blockIdx.z - gf * ((A.n_cols + DIRECT_CW - 1) / DIRECT_CW);
  const int c0 = ct * DIRECT_CW, w = min(DIRECT_CW, A.n_cols - c0);
  const MatsubaraDev &D = A.D;
  const long L = D.L;
  const bool active = o < A.n_out;
  const int oo = active ? o : 0;
  // multiplier of the running index and the first product (mod L)
  long s, first_i;
  if (FWD) {
    const long n = (long)D.first + oo;
    s = ((n % L) + L) % L;
    first_i = 0;
  } else {
    s = oo;
    first_i = (((long)D.first % L) + L) % L;
  }
  const int i0 = chunk * A.chunk, i1 = min(i0 + A.chunk, A.n_in);
  long m = (long)(((unsigned long long)((first_i + i0) % L) * (unsigned long long)s) % (unsigned long long)L);
  cd acc[DIRECT_CW];
#pragma unroll
  for (int c = 0; c < DIRECT_CW; ++c) acc[c] = cmk(0.0, 0.0);
  const cd *x = A.x + ((size_t)gf * A.n_in + i0) * A.n_cols + c0;
  for (int i = i0; i < i1; ++i) {
    cd tw = bigtw(A.btw, (int)m);
    if (!FWD) tw = cconj(tw);
    m += s;
    if (m >= L) m -= L;
#pragma unroll
    for (int c = 0; c < DIRECT_CW; ++c)
      if (c < w) acc[c] = cfma(tw, __ldg(x + c), acc[c]);
    x += A.n_cols;
  }
  if (!active) return;
  cd *p = A.partial + (((size_t)gf * A.n_chunks + chunk) * A.n_out + o) * A.n_cols + c0;
  for (int c = 0; c < w; ++c) p[c] = acc[c];
}

// G(i w_n) = X + corr + tail (:181-183)   or   G(tau_j) = Y e^{-i pi j/L} + sum_i a_i h_i  and  G(tau_L) = -/+ (G(0) + m1)  (:261-268)
template <bool FWD> __global__ void k_direct_post(const DirectArgs A) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long per_gf = (long)A.n_out * A.n_cols;
  if (i >= per_gf) return;
  const int gf = blockIdx.y, r = (int)(i / A.n_cols), c = (int)(i - (long)r * A.n_cols);
  const MatsubaraDev &D = A.D;
  const bool fermion = D.stat == TQ_FERMION;
  const cd *mom = A.moments + (size_t)gf * A.mom_gf_stride;
  const cd m1 = mom[(size_t)1 * A.n_cols + c];
  cd a1, a2, a3;
  pole_weights(fermion, m1, mom[(size_t)2 * A.n_cols + c], mom[(size_t)3 * A.n_cols + c], a1, a2, a3);
  cd v = cmk(0.0, 0.0);
  if (A.post_mul)
    v = cmul(A.partial[A.z_wide ? ((size_t)r * gridDim.y + gf) * A.n_cols + c : (size_t)gf * A.z_rows * A.n_cols + i], A.post_mul[r]);
  else
    for (int k = 0; k < A.n_chunks; ++k) v = cadd(v, A.partial[((size_t)gf * A.n_chunks + k) * per_gf + i]);
  cd *out = A.out + (size_t)gf * A.out_gf_stride;
  if (FWD) {
    const cd *g = A.in + (size_t)gf * A.in_gf_stride;
    cd t = cadd(g[c], m1);
    const cd gL = g[(size_t)D.L * A.n_cols + c];
    t = fermion ? cadd(t, gL) : csub(t, gL);
    out[i] = cadd(cadd(v, cscale(-0.5 * D.fact_fwd, t)), tail_iw(D, load_w(D, r), a1, a2, a3));
  } else {
    if (fermion) v = cmul(v, phase_inv(D, r));
    double h[3];
    row_model(D, r, h);
    v = caxpy(h[2], a3, caxpy(h[1], a2, caxpy(h[0], a1, v)));
    out[i] = v;
    if (r == 0) {
      const cd t = cadd(v, m1);
      out[(size_t)D.L * A.n_cols + c] = fermion ? cneg(t) : t;
    }
  }
}

// ... and the way out of the wide layout for scalar gfs: z[row][gf] is read along the gfs, the result written along the rows of each gf
template <bool FWD> __global__ void __launch_bounds__(256) k_direct_post_wide1(const DirectArgs A, int n_gf) {
  __shared__ cd tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
  const long r0 = (long)blockIdx.x * 32;
  const int g0 = blockIdx.y * 32;
  const MatsubaraDev &D = A.D;
  const bool fermion = D.stat == TQ_FERMION;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long rr = r0 + ty + 8 * k;
    const int g = g0 + tx;
    tile[ty + 8 * k][tx] = (rr < A.n_out && g < n_gf) ? A.partial[(size_t)rr * n_gf + g] : cmk(0.0, 0.0);
  }
  __syncthreads();
  const long r = r0 + tx;
  if (r >= A.n_out) return;
  const cd pm = A.post_mul[r];
  double h[3] = {0, 0, 0};
  cd ph = cmk(1.0, 0.0);
  double4 wrow = make_double4(0, 0, 0, 0);
  if (FWD) {
    wrow = load_w(D, (int)r);
  } else {
    row_model(D, (int)r, h);
    if (fermion) ph = phase_inv(D, (int)r);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int g = g0 + ty + 8 * k;
    if (g >= n_gf) continue;
    const cd *mom = A.moments + (size_t)g * A.mom_gf_stride;
    const cd m1 = mom[1];
    cd a1, a2, a3;
    pole_weights(fermion, m1, mom[2], mom[3], a1, a2, a3);
    cd v = cmul(tile[tx][ty + 8 * k], pm);
    cd *out = A.out + (size_t)g * A.out_gf_stride;
    if (FWD) {
      const cd *gin = A.in + (size_t)g * A.in_gf_stride;
      cd t = cadd(gin[0], m1);
      const cd gL = gin[(size_t)D.L];
      t = fermion ? cadd(t, gL) : csub(t, gL);
      out[r] = cadd(cadd(v, cscale(-0.5 * D.fact_fwd, t)), tail_iw(D, wrow, a1, a2, a3));
    } else {
      if (fermion) v = cmul(v, ph);
      v = caxpy(h[2], a3, caxpy(h[1], a2, caxpy(h[0], a1, v)));
      out[r] = v;
      if (r == 0) {
        const cd t = cadd(v, m1);
        out[(size_t)D.L] = fermion ? cneg(t) : t;
      }
    }
  }
}

template <bool FWD>
static tq_status run_direct(tq_ctx *ctx, const MatsubaraDev &D, const BigTw &btw, long n_cols, long batch, const cd *d_in, cd *d_out, const cd *d_mom,
                            int mom_rows) {
  DirectArgs A{};
  A.D = D;
  A.btw = btw;
  A.n_cols = (int)n_cols;
  A.n_in = FWD ? D.L : D.n_w;
  A.n_out = FWD ? D.n_w : D.L;
  A.chunk = 2048;
  A.n_chunks = (A.n_in + A.chunk - 1) / A.chunk;
  A.in_gf_stride = (long)(FWD ? D.L + 1 : D.n_w) * n_cols;
  A.out_gf_stride = (long)(FWD ? D.n_w : D.L + 1) * n_cols;
  A.mom_gf_stride = (long)mom_rows * n_cols;
  const size_t x_gf = sizeof(cd) * (size_t)A.n_in * n_cols, p_gf = sizeof(cd) * (size_t)A.n_chunks * A.n_out * n_cols;
  const int n_ct = (int)((n_cols + DIRECT_CW - 1) / DIRECT_CW);
  // gfs per pass: bounded scratch (256 MB) and grid.z
  long per = std::max<long>(1, std::min<long>((long)((256ull << 20) / (x_gf + p_gf)), 65535 / n_ct));
  per = std::min(per, batch);
  const size_t x_off = (p_gf * per + 255) & ~size_t(255);
  TQ_TRY(tq_reserve(ctx, ctx->scratch, x_off + x_gf * per));
  A.partial = (cd *)ctx->scratch.p;
  A.x = (cd *)((char *)ctx->scratch.p + x_off);
  for (long b0 = 0; b0 < batch; b0 += per) {
    const long nb = std::min(per, batch - b0);
    A.in = d_in + b0 * A.in_gf_stride;
    A.out = d_out + b0 * A.out_gf_stride;
    A.moments = d_mom + b0 * A.mom_gf_stride;
    {
      KernelScope ks(ctx, FWD ? "tau2iw_direct" : "iw2tau_direct");
      k_direct_pre<FWD><<<dim3((unsigned)(((long)A.n_in * n_cols + 255) / 256), (unsigned)nb), 256, 0, ctx->stream>>>(A);
      k_direct_dft<FWD><<<dim3((unsigned)((A.n_out + 127) / 128), (unsigned)A.n_chunks, (unsigned)(nb * n_ct)), 128, 0, ctx->stream>>>(A);
      k_direct_post<FWD><<<dim3((unsigned)(((long)A.n_out * n_cols + 255) / 256), (unsigned)nb), 256, 0, ctx->stream>>>(A);
    }
    tq_count_launch(ctx, 3);
    TQ_CUDA(ctx, cudaGetLastError());
  }
  return TQ_OK;
}

// -------------------------------------------------------------------------------------------------
// Bluestein (chirp-z) evaluation for mesh lengths WITHOUT a plan (a prime factor > 127: L = 1999, 4999, 8191, 19999 ... -- round
// n_tau values are one off a smooth length; FFTW takes them through Rader / Bluestein as well, fourier_common.cpp:30-44).
//   j n = (j^2 + n^2 - (n - j)^2) / 2   =>   X_n = C(n) sum_j [x_j C(j)] B(n - j),   C(m) = e^{i pi m^2 / L},  B = conj C
// i.e. a linear convolution of the chirped input with a chirp, done as a circular one of a SMOOTH length M = M1 M2 >= n_in + n_out - 1
// through the plain axis engine (lattice_pipe.cu): FFT_M as a four-step  [M1][M2]: axis M1, twiddle w_M^{j2 k1}, axis M2  (output
// order [k1][k2], the order the kernel spectrum is stored in), pointwise product, and the same steps backwards.  Only the window is
// wanted / non-zero (:182, :254): n_in = L, n_out = n_w forward and the other way round on the return leg.  The pre / post arithmetic
// of the Matsubara transform is k_direct_pre / k_direct_post.  Cost ~ 9 streaming passes over M n_cols elements instead of
// n_w L n_cols multiply-adds.
// -------------------------------------------------------------------------------------------------
tq_status tq_fft_axis_device(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign); // lattice.cu
double tq_pipe_plan_cost(long L);                                                                                // fourier_pipe.cu
tq_status tq_fft_axis_mul_device(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, const cd *tab, int omod, int os, int rs,
                                 int cdiv, int conj, bool *fused);

struct BlueDir {
  const cd *pre = nullptr;   // [n_in]  chirp on the input
  const cd *post = nullptr;  // [n_out] chirp on the output, times 1 / M
  const cd *khat = nullptr;  // [M1][M2] spectrum of the circular kernel sequence
};
struct BluePlan {
  int M1 = 0, M2 = 0;
  long M = 0;
  const cd *tw = nullptr; // [M1][M2] exp(-2 pi i j2 k1 / M)
  BlueDir dir[2];         // 0: forward (imtime -> imfreq), 1: inverse
  std::vector<void *> allocs;
  ~BluePlan() {
    for (void *p : allocs) cudaFree(p);
  }
};
struct BluePlanCache {
  std::map<std::tuple<long, long, long>, std::shared_ptr<BluePlan>> plans; // (L, n_w, first)
};

// z[gf][m][c] *= tab[m]  (or its conjugate)
static __global__ void k_blue_mul(cd *__restrict__ z, const cd *__restrict__ tab, long M, int n_cols, int conj) {
  const long per = M * n_cols;
  const long gf = blockIdx.y;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long)gridDim.x * blockDim.x) {
    cd t = tab[i / n_cols];
    if (conj) t = cconj(t);
    z[gf * per + i] = cmul(z[gf * per + i], t);
  }
}
static cd blue_chirp(long m, long L, int sign) { // exp(sign i pi m^2 / L), m^2 reduced mod 2 L exactly
  const long r = (long)(((unsigned long long)(m < 0 ? -m : m) * (unsigned long long)(m < 0 ? -m : m)) % (unsigned long long)(2 * L));
  const long double a = 3.141592653589793238462643383279502884L * (long double)r / (long double)L;
  return cmk((double)cosl(a), (double)(sign * sinl(a)));
}

// smooth axis lengths the plain engine runs through its pipelined kernel (two register radices), 16 ... 640
static bool blue_axis_ok(int n) {
  if (n < 16 || n > 640) return false;
  int m = n;
  for (int p : {2, 3, 5})
    while (m % p == 0) m /= p;
  if (m != 1) return false;
  TileShape t;
  return tile_shape(n, &t, false) && !t.generic && !t.heavy;
}
static bool blue_pick_lengths(long need, int *M1, int *M2) {
  long best = 0;
  for (int a = 16; a <= 640; ++a) {
    if (!blue_axis_ok(a)) continue;
    for (int b = a; b <= 640; ++b) {
      if ((long)a * b < need || !blue_axis_ok(b)) continue;
      if (!best || (long)a * b <= best) { // (<=: a ascends, so among equal products the most balanced pair wins)
        best = (long)a * b;
        *M1 = a;
        *M2 = b;
      }
      break; // larger b only grow the product
    }
  }
  return best != 0;
}

static tq_status blue_convolve(tq_ctx *ctx, const BluePlan &bp, const BlueDir *d /* nullptr: transform only (plan creation) */, cd *z, long n_gf,
                               int n_cols) {
  const long M = bp.M;
  const unsigned gx = (unsigned)std::min<long>((M * n_cols + 255) / 256, 148 * 8);
  const auto mul = [&](const cd *tab, int conj) {
    k_blue_mul<<<dim3(gx, (unsigned)n_gf), 256, 0, ctx->stream>>>(z, tab, M, n_cols, conj);
    tq_count_launch(ctx);
  };
  bool fused = false;
  // forward: axis M1 of [gf][M1][M2 cols], x w_M^{j2 k1}, axis M2 of [gf M1][M2][cols]  ->  [k1][k2] order.  The pointwise products ride
  // on the stores of the axis kernels where the pipelined kernel takes the length (always, for the lengths blue_pick_lengths returns,
  // unless "no_pipe" is set).
  TQ_TRY(tq_fft_axis_mul_device(ctx, n_gf, bp.M1, (long)bp.M2 * n_cols, z, z, -1, bp.tw, 0, 0, bp.M2, n_cols, 0, &fused));
  if (!fused) mul(bp.tw, 0);
  if (!d) {
    TQ_TRY(tq_fft_axis_device(ctx, n_gf * bp.M1, bp.M2, n_cols, z, z, -1));
    return TQ_OK;
  }
  TQ_TRY(tq_fft_axis_mul_device(ctx, n_gf * bp.M1, bp.M2, n_cols, z, z, -1, d->khat, bp.M1, bp.M2, 1, 0, 0, &fused));
  if (!fused) mul(d->khat, 0);
  // ... and back
  TQ_TRY(tq_fft_axis_mul_device(ctx, n_gf * bp.M1, bp.M2, n_cols, z, z, +1, bp.tw, bp.M1, bp.M2, 1, 0, 1, &fused));
  if (!fused) mul(bp.tw, 1);
  TQ_TRY(tq_fft_axis_device(ctx, n_gf, bp.M1, (long)bp.M2 * n_cols, z, z, +1));
  TQ_CUDA(ctx, cudaGetLastError());
  return TQ_OK;
}

static tq_status get_blue_plan(tq_ctx *ctx, const MatsubaraDev &D, std::shared_ptr<BluePlan> *out) {
  if (!ctx->blue_cache) ctx->blue_cache = std::make_shared<BluePlanCache>();
  auto key = std::make_tuple((long)D.L, (long)D.n_w, (long)D.first);
  auto it = ctx->blue_cache->plans.find(key);
  if (it != ctx->blue_cache->plans.end()) {
    *out = it->second;
    return TQ_OK;
  }
  auto bp = std::make_shared<BluePlan>();
  const long L = D.L, n_w = D.n_w, first = D.first;
  if (!blue_pick_lengths(L + n_w - 1, &bp->M1, &bp->M2)) { // (meshes beyond 640 x 640 - n_w: the direct evaluation stays)
    *out = nullptr;
    return TQ_OK;
  }
  const long M = bp->M = (long)bp->M1 * bp->M2;
  auto up = [&](const std::vector<cd> &h, const cd **d) -> tq_status {
    void *p = nullptr;
    TQ_CUDA(ctx, cudaMalloc(&p, sizeof(cd) * h.size()));
    bp->allocs.push_back(p);
    TQ_CUDA(ctx, cudaMemcpyAsync(p, h.data(), sizeof(cd) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *d = (const cd *)p;
    return TQ_OK;
  };
  {
    std::vector<cd> tw((size_t)M);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k1 = 0; k1 < bp->M1; ++k1)
      for (int j2 = 0; j2 < bp->M2; ++j2) {
        const long r = ((long)k1 * j2) % M;
        const long double a = two_pi * (long double)r / (long double)M;
        tw[(size_t)k1 * bp->M2 + j2] = cmk((double)cosl(a), (double)-sinl(a));
      }
    TQ_TRY(up(tw, &bp->tw));
  }
  const double inv_M = 1.0 / (double)M;
  for (int dir = 0; dir < 2; ++dir) {
    const long n_in = dir == 0 ? L : n_w, n_out = dir == 0 ? n_w : L;
    std::vector<cd> pre((size_t)n_in), post((size_t)n_out), kc((size_t)M, cmk(0.0, 0.0));
    if (dir == 0) { // X_o = C(first + o) sum_j [x_j C(j)] B(first + o - j)
      for (long j = 0; j < n_in; ++j) pre[j] = blue_chirp(j, L, +1);
      for (long o = 0; o < n_out; ++o) post[o] = cscale(inv_M, blue_chirp(first + o, L, +1));
      for (long t = -(n_in - 1); t <= n_out - 1; ++t) kc[(size_t)(((t % M) + M) % M)] = blue_chirp(first + t, L, -1);
    } else { // Y_j = B(j) sum_i [y_i B(first + i)] C(first + i - j)
      for (long i = 0; i < n_in; ++i) pre[i] = blue_chirp(first + i, L, -1);
      for (long j = 0; j < n_out; ++j) post[j] = cscale(inv_M, blue_chirp(j, L, -1));
      for (long t = -(n_in - 1); t <= n_out - 1; ++t) kc[(size_t)(((t % M) + M) % M)] = blue_chirp(first - t, L, +1);
    }
    TQ_TRY(up(pre, &bp->dir[dir].pre));
    TQ_TRY(up(post, &bp->dir[dir].post));
    const cd *kd = nullptr;
    TQ_TRY(up(kc, &kd));
    TQ_TRY(blue_convolve(ctx, *bp, nullptr, const_cast<cd *>(kd), 1, 1)); // its spectrum, in the engine's own [k1][k2] order
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bp->dir[dir].khat = kd;
  }
  ctx->blue_cache->plans[key] = bp;
  *out = bp;
  return TQ_OK;
}

// ---- the same convolution for a PLAIN axis transform (tq_fft_many / lattice / real-time axes whose length has a prime factor > 127:
// _fourier_base takes any length, fourier_common.cpp:30-44): X_k = sum_j x_j e^{s 2 pi i jk / N} for all k, so M >= 2 N - 1.
struct BlueAxisCache {
  std::map<std::tuple<long, int>, std::shared_ptr<BluePlan>> plans; // (N, sign); dir[0] only
};
// z[g][m][c] = in[o0 + g][m][c0 + c] pre[m]  (m < N), 0 above;   out[o0 + g][k][c0 + c] = z[g][k][c] post[k] scale
static __global__ void k_blue_axis_load(const cd *__restrict__ in, long N, long inner, long c0, int cw, const cd *__restrict__ pre, long M, cd *__restrict__ z,
                                        const cd *__restrict__ rin, const cd *__restrict__ coef, long o0) {
  const long per = M * cw;
  const long g = blockIdx.y;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long)gridDim.x * blockDim.x) {
    const long m = i / cw;
    const int c = (int)(i - m * cw);
    cd v = cmk(0.0, 0.0);
    if (m < N) {
      v = in[(g * N + m) * inner + c0 + c];
      if (rin) { // the affine wrap of the real-time transforms (fourier_real.cu, k_direct_axis<true>): (x - a1 h1 - a2 h2) p_in
        const cd a1 = coef[2 * ((o0 + g) * inner + c0 + c)], a2 = coef[2 * ((o0 + g) * inner + c0 + c) + 1];
        v = cmul(csub(csub(v, cmul(a1, rin[3 * m])), cmul(a2, rin[3 * m + 1])), rin[3 * m + 2]);
      }
      v = cmul(v, pre[m]);
    }
    z[g * per + i] = v;
  }
}
static __global__ void k_blue_axis_store(const cd *__restrict__ z, long N, long inner, long c0, int cw, const cd *__restrict__ post, long M, double scale,
                                         cd *__restrict__ out, const cd *__restrict__ rout, const cd *__restrict__ coef, long o0) {
  const long per = N * cw;
  const long g = blockIdx.y;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long)gridDim.x * blockDim.x) {
    const long k = i / cw;
    const int c = (int)(i - k * cw);
    cd v = cscale(scale, cmul(z[g * M * cw + i], post[k]));
    if (rout) { // p_out v + a1 e1 + a2 e2
      const cd a1 = coef[2 * ((o0 + g) * inner + c0 + c)], a2 = coef[2 * ((o0 + g) * inner + c0 + c) + 1];
      v = cadd(cadd(cmul(rout[3 * k + 2], v), cmul(a1, rout[3 * k])), cmul(a2, rout[3 * k + 1]));
    }
    out[(g * N + k) * inner + c0 + c] = v;
  }
}
bool tq_bluestein_axis_possible(tq_ctx *ctx, long N) {
  int a, b;
  return !ctx->opt_no_bluestein && N >= 512 && blue_pick_lengths(2 * N - 1, &a, &b);
}
tq_status tq_bluestein_plain_axis(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, double scale, const cd *rin,
                                  const cd *rout, const cd *coef) {
  if (!ctx->blue_axis_cache) ctx->blue_axis_cache = std::make_shared<BlueAxisCache>();
  auto &plans = ctx->blue_axis_cache->plans;
  auto key = std::make_tuple(N, sign > 0 ? 1 : -1);
  std::shared_ptr<BluePlan> bp;
  auto it = plans.find(key);
  if (it != plans.end()) {
    bp = it->second;
  } else {
    bp = std::make_shared<BluePlan>();
    if (!blue_pick_lengths(2 * N - 1, &bp->M1, &bp->M2)) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tq_fft_many: axis length too long for the Bluestein path");
    const long M = bp->M = (long)bp->M1 * bp->M2;
    const int sg = sign > 0 ? 1 : -1;
    auto up = [&](const std::vector<cd> &h, const cd **d) -> tq_status {
      void *p = nullptr;
      TQ_CUDA(ctx, cudaMalloc(&p, sizeof(cd) * h.size()));
      bp->allocs.push_back(p);
      TQ_CUDA(ctx, cudaMemcpyAsync(p, h.data(), sizeof(cd) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
      TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      *d = (const cd *)p;
      return TQ_OK;
    };
    std::vector<cd> tw((size_t)M), pre((size_t)N), post((size_t)N), kc((size_t)M, cmk(0.0, 0.0));
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k1 = 0; k1 < bp->M1; ++k1)
      for (int j2 = 0; j2 < bp->M2; ++j2) {
        const long double a = two_pi * (long double)(((long)k1 * j2) % M) / (long double)M;
        tw[(size_t)k1 * bp->M2 + j2] = cmk((double)cosl(a), (double)-sinl(a));
      }
    const double inv_M = 1.0 / (double)M;
    for (long j = 0; j < N; ++j) {
      pre[j] = blue_chirp(j, N, sg);
      post[j] = cscale(inv_M, pre[j]);
    }
    for (long t = -(N - 1); t <= N - 1; ++t) kc[(size_t)(((t % M) + M) % M)] = blue_chirp(t, N, -sg);
    TQ_TRY(up(tw, &bp->tw));
    TQ_TRY(up(pre, &bp->dir[0].pre));
    TQ_TRY(up(post, &bp->dir[0].post));
    const cd *kd = nullptr;
    TQ_TRY(up(kc, &kd));
    TQ_TRY(blue_convolve(ctx, *bp, nullptr, const_cast<cd *>(kd), 1, 1));
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bp->dir[0].khat = kd;
    plans[key] = bp;
  }
  const long M = bp->M;
  const size_t budget = (size_t)ctx->opt_blue_chunk_mb << 20;
  const long cw = std::max<long>(1, std::min<long>(inner, (long)(budget / (sizeof(cd) * (size_t)M))));
  const long per = std::max<long>(1, std::min<long>(std::min<long>(outer, 65535), (long)(budget / (sizeof(cd) * (size_t)M * cw))));
  TQ_TRY(tq_reserve(ctx, ctx->scratch, sizeof(cd) * (size_t)M * cw * per));
  cd *z = (cd *)ctx->scratch.p;
  for (long o0 = 0; o0 < outer; o0 += per) {
    const long g = std::min(per, outer - o0);
    for (long c0 = 0; c0 < inner; c0 += cw) {
      const int w = (int)std::min(cw, inner - c0);
      const unsigned gx = (unsigned)std::min<long>((M * w + 255) / 256, 148 * 8);
      {
        KernelScope ks(ctx, "fft_axis_bluestein");
        k_blue_axis_load<<<dim3(gx, (unsigned)g), 256, 0, ctx->stream>>>(in + o0 * N * inner, N, inner, c0, w, bp->dir[0].pre, M, z, rin, coef, o0);
      }
      tq_count_launch(ctx);
      TQ_TRY(blue_convolve(ctx, *bp, &bp->dir[0], z, g, w));
      k_blue_axis_store<<<dim3(gx, (unsigned)g), 256, 0, ctx->stream>>>(z, N, inner, c0, w, bp->dir[0].post, M, scale, out + o0 * N * inner, rout, coef, o0);
      tq_count_launch(ctx);
      TQ_CUDA(ctx, cudaGetLastError());
    }
  }
  return TQ_OK;
}

template <bool FWD>
static tq_status run_bluestein(tq_ctx *ctx, const MatsubaraDev &D, const BigTw &btw, const BluePlan &bp, long n_cols, long batch, const cd *d_in,
                               cd *d_out, const cd *d_mom, int mom_rows) {
  DirectArgs A{};
  A.D = D;
  A.btw = btw;
  A.n_cols = (int)n_cols;
  A.n_in = FWD ? D.L : D.n_w;
  A.n_out = FWD ? D.n_w : D.L;
  A.chunk = A.n_in;
  A.n_chunks = 1;
  A.in_gf_stride = (long)(FWD ? D.L + 1 : D.n_w) * n_cols;
  A.out_gf_stride = (long)(FWD ? D.n_w : D.L + 1) * n_cols;
  A.mom_gf_stride = (long)mom_rows * n_cols;
  const BlueDir &dir = bp.dir[FWD ? 0 : 1];
  A.pre_mul = dir.pre;
  A.post_mul = dir.post;
  A.z_rows = bp.M;
  const size_t z_gf = sizeof(cd) * (size_t)bp.M * n_cols;
  long per = std::max<long>(1, std::min<long>((long)(((size_t)ctx->opt_blue_chunk_mb << 20) / z_gf), 65535));
  per = std::min(per, batch);
  TQ_TRY(tq_reserve(ctx, ctx->scratch, z_gf * per));
  cd *z = (cd *)ctx->scratch.p;
  A.x = z;       // k_direct_pre writes the chirped, zero-padded input straight into the convolution buffer ...
  A.partial = z; // ... and k_direct_post reads the window rows back out of it
  // narrow targets (scalar gfs ...): all gfs of a group side by side as the columns of ONE wide buffer, so that the axis kernels see rows of
  // hundreds of bytes instead of 16-112 (L = 4999 x 1 column: 115 -> see profiles/r02_bluestein_narrow.log)
  A.z_wide = (n_cols < 16 && per > 1) ? 1 : 0;
  for (long b0 = 0; b0 < batch; b0 += per) {
    const long nb = std::min(per, batch - b0);
    A.in = d_in + b0 * A.in_gf_stride;
    A.out = d_out + b0 * A.out_gf_stride;
    A.moments = d_mom + b0 * A.mom_gf_stride;
    {
      KernelScope ks(ctx, FWD ? "tau2iw_bluestein_pre" : "iw2tau_bluestein_pre");
      if (A.z_wide && n_cols == 1)
        k_direct_pre_wide1<FWD><<<dim3((unsigned)((bp.M + 31) / 32), (unsigned)((nb + 31) / 32)), 256, 0, ctx->stream>>>(A, (int)nb);
      else
        k_direct_pre<FWD><<<dim3((unsigned)((bp.M * n_cols + 255) / 256), (unsigned)nb), 256, 0, ctx->stream>>>(A);
    }
    tq_count_launch(ctx);
    if (A.z_wide)
      TQ_TRY(blue_convolve(ctx, bp, &dir, z, 1, (int)(nb * n_cols)));
    else
      TQ_TRY(blue_convolve(ctx, bp, &dir, z, nb, (int)n_cols));
    {
      KernelScope ks(ctx, FWD ? "tau2iw_bluestein_post" : "iw2tau_bluestein_post");
      if (A.z_wide && n_cols == 1)
        k_direct_post_wide1<FWD><<<dim3((unsigned)((A.n_out + 31) / 32), (unsigned)((nb + 31) / 32)), 256, 0, ctx->stream>>>(A, (int)nb);
      else
        k_direct_post<FWD><<<dim3((unsigned)(((long)A.n_out * n_cols + 255) / 256), (unsigned)nb), 256, 0, ctx->stream>>>(A);
    }
    tq_count_launch(ctx);
    TQ_CUDA(ctx, cudaGetLastError());
  }
  return TQ_OK;
}

// -------------------------------------------------------------------------------------------------
// moments of G(tau) from one-sided finite differences  (fourier_matsubara.cpp:39-92, noise check :100-114)
// -------------------------------------------------------------------------------------------------
__constant__ double c_vinv[2][8] = {
   // rows 0 and 1 of the literal inverse Vandermonde matrix of fourier_matsubara.cpp:52-55 (only these are used, :89-90)
   {8.000000000008853, -28.000000000055913, 56.000000000151815, -70.00000000021936, 56.00000000020795, -28.000000000115904,
    8.00000000003683, -1.0000000000049232},
   {-13.742857142879107, 62.10000000013776, -133.5333333337073, 172.75000000055635, -141.00000000052023, 71.43333333362274,
    -20.600000000091182, 2.592857142869304}};

__device__ double block_max(double v, double *sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (l == 0) sh[0] = r;
  }
  __syncthreads();
  r = sh[0];
  return r;
}

// `joint` (or nullptr): the batch is ONE gf whose columns are spread over [outer][.][inner] (a transform along mesh N > 0 of a
// product mesh, fourier.hpp:78-84): the reference's noise check is then a maximum over all of them.  First pass
// (joint_pass == 1) accumulates the two maxima, second pass (joint_pass == 2) decides with them.
// packed: every complex column holds TWO real-valued columns (Re, Im) -- the noise check looks at them separately, exactly as it would
// at the two promoted columns (hypot(x, 0) = |x|)
__global__ void __launch_bounds__(256) k_moments_from_tau(const cd *gt, long gf_stride, int n_tau, int n_cols, double beta, int stat, cd *mom4,
                                                          int *warn, unsigned long long *joint = nullptr, int joint_pass = 0, int packed = 0) {
  __shared__ double sh[32];
  const int gf = blockIdx.x;
  const cd *g = gt + (size_t)gf * gf_stride;
  const int L = n_tau - 1;
  const double delta = (beta - 0.0) / (n_tau - 1);
  double m1 = 0.0, m2 = 0.0;
  for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
    cd g0 = g[c], g1 = g[(size_t)1 * n_cols + c], g2 = g[(size_t)2 * n_cols + c];
    cd gL = g[(size_t)L * n_cols + c], gL1 = g[(size_t)(L - 1) * n_cols + c], gL2 = g[(size_t)(L - 2) * n_cols + c];
    if (packed) {
      const cd d1 = csub(g1, g0), d1L = csub(gL1, gL), d2 = csub(g2, g0), d2L = csub(gL2, gL);
      m1 = fmax(m1, fmax(fabs(d1.x) + fabs(d1L.x), fabs(d1.y) + fabs(d1L.y)));
      m2 = fmax(m2, fmax(fabs(d2.x) + fabs(d2L.x), fabs(d2.y) + fabs(d2L.y)));
      continue;
    }
    cd d;
    d = csub(g1, g0);
    double a = hypot(d.x, d.y);
    d = csub(gL1, gL);
    a += hypot(d.x, d.y);
    d = csub(g2, g0);
    double b = hypot(d.x, d.y);
    d = csub(gL2, gL);
    b += hypot(d.x, d.y);
    m1 = fmax(m1, a);
    m2 = fmax(m2, b);
  }
  m1 = block_max(m1, sh);
  m2 = block_max(m2, sh);
  if (joint_pass == 1) { // non-negative doubles order like their bit patterns
    if (threadIdx.x == 0) {
      atomicMax(&joint[0], (unsigned long long)__double_as_longlong(m1));
      atomicMax(&joint[1], (unsigned long long)__double_as_longlong(m2));
    }
    return;
  }
  if (joint_pass == 2) {
    m1 = __longlong_as_double((long long)joint[0]);
    m2 = __longlong_as_double((long long)joint[1]);
  }
  double der_1 = m1 / delta;
  double der_2 = 0.5 * m2 / delta;
  const bool noisy = (der_1 < 0.95 * der_2) || (der_1 > 1.05 * der_2);
  if (threadIdx.x == 0 && warn) warn[gf] = noisy ? TQ_WARN_NOISY_TAU : 0;
  cd *mo = mom4 + (size_t)gf * 4 * n_cols;
  const double sign = (stat == TQ_FERMION) ? -1.0 : 1.0;
  for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
    cd t1 = cmk(0, 0), t2 = cmk(0, 0), t3 = cmk(0, 0);
    if (!noisy) {
      cd g0 = g[c], gL = g[(size_t)L * n_cols + c];
      cd gl0 = cmk(0, 0), gl1 = cmk(0, 0), gr0 = cmk(0, 0), gr1 = cmk(0, 0);
#pragma unroll
      for (int m = 1; m <= 8; ++m) {
        double wr = double(m) / double(n_tau - 1);
        double tau_m = 0.0 * (1 - wr) + beta * wr; // linear.hpp:154-160
        cd dl = csub(g[(size_t)m * n_cols + c], g0);
        cd dr = csub(gL, g[(size_t)(L - m) * n_cols + c]);
        dl = cmk(dl.x / tau_m, dl.y / tau_m);
        dr = cmk(dr.x / tau_m, dr.y / tau_m);
        gl0 = caxpy(c_vinv[0][m - 1], dl, gl0);
        gl1 = caxpy(c_vinv[1][m - 1], dl, gl1);
        gr0 = caxpy(c_vinv[0][m - 1], dr, gr0);
        gr1 = caxpy(c_vinv[1][m - 1], dr, gr1);
      }
      t1 = cneg(csub(g0, cscale(sign, gL)));
      t2 = csub(gl0, cscale(sign, gr0));
      t3 = cscale(-2.0 / delta, cadd(gl1, cscale(sign, gr1)));
    }
    mo[c] = cmk(0, 0);
    mo[(size_t)n_cols + c] = t1;
    mo[(size_t)2 * n_cols + c] = t2;
    mo[(size_t)3 * n_cols + c] = t3;
  }
}

// known moments -> dense [batch][4][n_cols] (zero padded) and max |0th moment|   (:116-122, :199-201)
__global__ void k_prepare_known_moments(const cd *known, int n_known, int n_cols, long total_cols /* batch * n_cols */, cd *mom4,
                                        unsigned long long *max_m0_bits) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  double a0 = 0.0;
  if (i < total_cols) {
    long gf = i / n_cols;
    int c = (int)(i - gf * n_cols);
    const cd *k = known + (size_t)gf * n_known * n_cols;
    cd *mo = mom4 + (size_t)gf * 4 * n_cols;
    for (int r = 0; r < 4; ++r) mo[(size_t)r * n_cols + c] = (r < n_known) ? k[(size_t)r * n_cols + c] : cmk(0, 0);
    cd m0 = k[c];
    a0 = hypot(m0.x, m0.y);
    if (!(a0 == a0)) a0 = INFINITY; // NaN fails the "< 1e-8" assertion
  }
  for (int o = 16; o > 0; o >>= 1) a0 = fmax(a0, __shfl_xor_sync(0xffffffffu, a0, o));
  if ((threadIdx.x & 31) == 0 && a0 > 0.0) atomicMax(max_m0_bits, (unsigned long long)__double_as_longlong(a0));
}

// -------------------------------------------------------------------------------------------------
// host: plans
// -------------------------------------------------------------------------------------------------
struct MatsubaraPlan {
  MatsubaraDev dev{};
  BigTw btw{};
  std::vector<void *> allocs;
  ~MatsubaraPlan() {
    for (void *p : allocs) cudaFree(p);
  }
};

template <class T> static tq_status upload(tq_ctx *ctx, MatsubaraPlan &pl, const std::vector<T> &h, const T **d) {
  void *p = nullptr;
  TQ_CUDA(ctx, cudaMalloc(&p, sizeof(T) * std::max<size_t>(h.size(), 1)));
  pl.allocs.push_back(p);
  TQ_CUDA(ctx, cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
  *d = (const T *)p;
  return TQ_OK;
}

static tq_status get_matsubara_plan(tq_ctx *ctx, double beta, int stat, long n_tau, long n_iw, std::shared_ptr<MatsubaraPlan> *out) {
  auto key = std::make_tuple(beta, stat, n_tau, n_iw);
  auto it = ctx->matsubara_plans.find(key);
  if (it != ctx->matsubara_plans.end()) {
    *out = it->second;
    return TQ_OK;
  }
  auto pl = std::make_shared<MatsubaraPlan>();
  MatsubaraDev &D = pl->dev;
  const long L = n_tau - 1;
  D.L = (int)L;
  D.stat = stat;
  D.beta = beta;
  D.fact_fwd = beta / L;  // fourier_matsubara.cpp:141
  D.inv_beta = 1.0 / beta; // :229
  D.last = (int)(n_iw - 1);                              // imfreq.hpp:58-60
  D.first = -(D.last + (stat == TQ_FERMION ? 1 : 0));
  D.n_w = D.last - D.first + 1;
  D.lo_shift = 7;
  const int LO = 1 << D.lo_shift;
  D.nhi = (int)(L >> D.lo_shift) + 1;
  if (stat == TQ_FERMION) {
    D.b[0] = 0; D.b[1] = 1; D.b[2] = -1; // :152-154
  } else {
    D.b[0] = -0.5; D.b[1] = -1; D.b[2] = 1; // :163-165
  }
  const long double pi = 3.141592653589793238462643383279502884L;
  std::vector<cd> ph_hi_f(D.nhi), ph_hi_i(D.nhi), ph_lo(LO);
  for (int h = 0; h < D.nhi; ++h) {
    long double ang = pi * (long double)((long)h << D.lo_shift) / (long double)L;
    ph_hi_f[h] = cmk((double)((long double)D.fact_fwd * cosl(ang)), (double)((long double)D.fact_fwd * sinl(ang)));
    ph_hi_i[h] = cmk((double)cosl(ang), (double)(-sinl(ang)));
  }
  for (int l = 0; l < LO; ++l) {
    long double ang = pi * (long double)l / (long double)L;
    ph_lo[l] = cmk((double)cosl(ang), (double)sinl(ang));
  }
  std::vector<double> e_hi(3 * (size_t)D.nhi), e_lo(3 * (size_t)LO);
  for (int i = 0; i < 3; ++i) {
    long double ab = fabsl((long double)D.b[i]);
    D.rev[i] = D.b[i] < 0 ? 1 : 0;
    long double eb = expl(-(long double)beta * ab);
    if (stat == TQ_FERMION)
      D.C[i] = (double)(-1.0L / (1.0L + eb)); // oneFermion, :28-30
    else
      D.C[i] = D.b[i] >= 0 ? (double)(1.0L / (eb - 1.0L)) : (double)(1.0L / (1.0L - eb)); // oneBoson, :32-34
    for (int h = 0; h < D.nhi; ++h)
      e_hi[(size_t)i * D.nhi + h] = (double)expl(-ab * (long double)beta * (long double)((long)h << D.lo_shift) / (long double)L);
    for (int l = 0; l < LO; ++l) e_lo[(size_t)i * LO + l] = (double)expl(-ab * (long double)beta * (long double)l / (long double)L);
  }
  std::vector<double4> wtab(D.n_w);
  for (int i = 0; i < D.n_w; ++i) {
    long n = D.first + i;
    double w = M_PI * (double)(2 * n + stat) / beta; // matsubara.hpp:55
    double4 t;
    t.x = w;
    t.y = 1.0 / (D.b[0] * D.b[0] + w * w);
    t.z = 1.0 / (D.b[1] * D.b[1] + w * w);
    t.w = 1.0 / (D.b[2] * D.b[2] + w * w);
    wtab[i] = t;
  }
  TQ_TRY(upload(ctx, *pl, ph_hi_f, &D.ph_hi_fwd));
  TQ_TRY(upload(ctx, *pl, ph_hi_i, &D.ph_hi_inv));
  TQ_TRY(upload(ctx, *pl, ph_lo, &D.ph_lo));
  TQ_TRY(upload(ctx, *pl, e_hi, &D.e_hi));
  TQ_TRY(upload(ctx, *pl, e_lo, &D.e_lo));
  TQ_TRY(upload(ctx, *pl, wtab, &D.wtab));
  // two-level table of exp(2 pi i m / L), m in [0, L), for the four-step twiddles
  {
    int shift = 9;
    int nlo = 1 << shift, nhi = (int)(L >> shift) + 1;
    std::vector<cd> hi(nhi), lo(nlo);
    const long double two_pi = 2 * pi;
    for (int h = 0; h < nhi; ++h) {
      long double ang = two_pi * (long double)((long)h << shift) / (long double)L;
      hi[h] = cmk((double)cosl(ang), (double)sinl(ang));
    }
    for (int l = 0; l < nlo; ++l) {
      long double ang = two_pi * (long double)l / (long double)L;
      lo[l] = cmk((double)cosl(ang), (double)sinl(ang));
    }
    pl->btw.shift = shift;
    TQ_TRY(upload(ctx, *pl, hi, &pl->btw.hi));
    TQ_TRY(upload(ctx, *pl, lo, &pl->btw.lo));
  }
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->matsubara_plans[key] = pl;
  *out = pl;
  return TQ_OK;
}

// -------------------------------------------------------------------------------------------------
// host: tiling decisions and launches
// -------------------------------------------------------------------------------------------------
struct Tiling {
  bool two_pass = false;
  int Na = 1, Nb = 1; // L = Na * Nb (two_pass)
  int tc = 1;
  int skew = 0;
};

// tiles with at least this many columns keep a shared-memory copy of the twiddles and a per-row factor table
static bool tile_uses_tables(int tc) { return tc >= 4; }
static size_t tile_bytes(int N, int tc, int skewmask) {
  size_t n = (size_t)fft_rows_alloc(N, skewmask) * tc + 4 * (size_t)tc;
  if (tile_uses_tables(tc)) n += 4 * (size_t)N; // twiddles [N] + row table [N][3]
  return n * sizeof(cd);
}

static bool radix_ok(long N) {
  FftPlan P;
  return tq_fft_factorize((int)N, P);
}

// widest column tile (<= n_cols) whose working set fits the opt-in shared memory
// TQ_TILE_BYTES (tuning knob): preferred shared-memory footprint of one tile; smaller tiles -> more CTAs per SM
static size_t tile_target_bytes(tq_ctx *ctx) {
  static long env = -1;
  if (env < 0) {
    const char *e = TQ_DEV_ENV("TQ_TILE_BYTES");
    env = e ? atol(e) : 0;
  }
  return env > 0 ? std::min<size_t>((size_t)env, (size_t)ctx->max_smem_optin) : (size_t)ctx->max_smem_optin / 2 - 1024;
}

static int pick_tc(tq_ctx *ctx, int N, long n_cols, size_t budget = 0) {
  if (!budget) {
    // prefer the target footprint if that still leaves >= 8 columns (128 contiguous bytes) per row
    budget = (size_t)ctx->max_smem_optin;
    const size_t tgt = tile_target_bytes(ctx);
    if (tile_bytes(N, (int)std::min<long>(8, n_cols), 0) <= tgt) budget = tgt;
  }
  long tc = std::min<long>(n_cols, (long)(budget / (sizeof(cd) * (N + 4))));
  while (tc > 1 && tile_bytes(N, (int)tc, 0) > budget) --tc;
  return (int)std::max<long>(tc, 1);
}

static tq_status choose_tiling(tq_ctx *ctx, long L, long n_cols, long batch, Tiling *t) {
  const size_t budget = (size_t)ctx->max_smem_optin;
  const size_t half = tile_target_bytes(ctx);
  // single pass
  if (tile_bytes((int)L, 1, 0) <= budget && radix_ok(L)) {
    t->two_pass = false;
    int tc_fit_half = tile_bytes((int)L, 1, 0) <= half ? pick_tc(ctx, (int)L, n_cols, half) : 0;
    int tc_fit_full = pick_tc(ctx, (int)L, n_cols);
    // prefer tiles that let two CTAs share an SM unless that would shrink the contiguous run below 128 B
    int tc = (tc_fit_half >= 8 || tc_fit_half >= n_cols) ? tc_fit_half : tc_fit_full;
    // enough CTAs to fill the machine: do not make tiles wider than needed when the batch is small
    while (tc > 8 && batch * ((n_cols + tc - 1) / tc) < 2L * ctx->sm_count && tc > 1) tc = (tc + 1) / 2;
    if (tc < 1) tc = 1;
    t->tc = tc;
    t->skew = (tc == 1 && tile_bytes((int)L, 1, ~0) <= budget) ? ~0 : 0;
    return TQ_OK;
  }
  // two pass: most balanced factor pair with both lengths supported
  long best = -1;
  for (long a = 1; a * a <= L; ++a) {
    if (L % a) continue;
    long b = L / a; // b >= a
    if (!radix_ok(a) || !radix_ok(b)) continue;
    if (tile_bytes((int)b, 1, 0) > budget) continue;
    best = a; // keeps the largest a <= sqrt(L), i.e. the most balanced pair
  }
  if (best < 0)
    return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "Fourier: time mesh length " + std::to_string(L) + " cannot be tiled into shared memory (no factor pair)");
  t->two_pass = true;
  t->Nb = (int)best;      // first pass length (smaller tile)
  t->Na = (int)(L / best); // second pass length
  t->tc = pick_tc(ctx, t->Na, n_cols);
  t->skew = 0;
  return TQ_OK;
}

template <int IN, int OUT, int SGN, bool SKEW> static tq_status launch_tile_impl(tq_ctx *ctx, const TileArgs &A, int n_items, long batch) {
  const int tiles = (A.n_cols + A.TC - 1) / A.TC;
  const size_t smem = tile_bytes(A.P.N, A.TC, A.skewmask);
  dim3 grid(n_items, tiles, (unsigned)batch);
  if (batch > 65535) return tq_fail(ctx, TQ_ERR_INVALID, "batch > 65535 per launch");
  const bool big = smem > (size_t)ctx->max_smem_optin / 2 - 1024;
  static const char *tags[3][3] = {{"tau2iw_single", "", "tau2iw_pass1"}, {"", "iw2tau_single", "iw2tau_pass1"}, {"tau2iw_pass2", "iw2tau_pass2", ""}};
  KernelScope ks(ctx, tags[IN][OUT]);
  TileArgs Ad = A;
  static const bool phase_debug = TQ_DEV_ENV("TQ_PHASE_DEBUG") != nullptr;
  const size_t nblk = (size_t)n_items * tiles * batch;
  if (phase_debug) {
    TQ_CUDA(ctx, cudaMallocManaged(&Ad.dbg, sizeof(long long) * 6 * nblk));
    TQ_CUDA(ctx, cudaMemset(Ad.dbg, 0, sizeof(long long) * 6 * nblk));
  }
  const TileArgs &A2 = Ad;
  if (big) {
    auto k = matsubara_tile_kernel<IN, OUT, SGN, SKEW, 512, 1>;
    TQ_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 512, smem, ctx->stream>>>(A2);
  } else {
    auto k = matsubara_tile_kernel<IN, OUT, SGN, SKEW, 256, 2>;
    TQ_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, ctx->stream>>>(A2);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  if (phase_debug) {
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double ph[4] = {0, 0, 0, 0};
    long long tmin = -1, tmax = 0;
    for (size_t i = 0; i < nblk; ++i) {
      long long *d = Ad.dbg + 6 * i;
      for (int j = 0; j < 4; ++j) ph[j] += double(d[j + 1] - d[j]);
      if (tmin < 0 || d[0] < tmin) tmin = d[0];
      if (d[4] > tmax) tmax = d[4];
    }
    fprintf(stderr, "[phase] %-14s N=%d tc=%d ctas=%zu smem=%zu  avg cycles: prologue %.0f load %.0f fft %.0f store %.0f  (sum %.0f)  span(first..last, one SM clock) %lld\n",
            tags[IN][OUT], A.P.N, A.TC, nblk, smem, ph[0] / nblk, ph[1] / nblk, ph[2] / nblk, ph[3] / nblk,
            (ph[0] + ph[1] + ph[2] + ph[3]) / nblk, tmax - tmin);
    cudaFree(Ad.dbg);
  }
  return TQ_OK;
}
// the skewed tile layout exists only for the single-pass, single-column kernels
template <int IN, int OUT, int SGN> static tq_status launch_tile(tq_ctx *ctx, const TileArgs &A, int n_items, long batch) {
  if (A.skewmask) {
    if constexpr ((IN == IN_TAU && OUT == OUT_FREQ) || (IN == IN_FREQ && OUT == OUT_TAU))
      return launch_tile_impl<IN, OUT, SGN, true>(ctx, A, n_items, batch);
    else
      return tq_fail(ctx, TQ_ERR_INVALID, "internal: skewed tile in a two-pass kernel");
  }
  return launch_tile_impl<IN, OUT, SGN, false>(ctx, A, n_items, batch);
}

// fourier_pipe.cu: TMA double-buffered specialisation for long meshes (runs instead of the generic kernels when it applies)
tq_status tq_pipe_matsubara(tq_ctx *ctx, bool fwd, double beta, int stat, long n_tau, long n_iw, long n_cols, long batch, const cd *d_in,
                            cd *d_out, const cd *d_mom, int mom_rows, bool *handled);

// fourier_col.cu: compile-time single-column kernel for scalar gfs on the 10^4 mesh (direct transform)
tq_status tq_col_matsubara_fwd(tq_ctx *ctx, double beta, int stat, long n_tau, long n_iw, long n_cols, long batch, const cd *d_in, cd *d_out,
                               const cd *d_mom, int mom_rows, bool *handled);

// runs the transform for gfs [0, batch) whose device arrays start at d_in / d_out; moments are [batch][mom_rows][n_cols]
template <int IN_KIND /* IN_TAU: forward, IN_FREQ: inverse */>
static tq_status run_transform(tq_ctx *ctx, MatsubaraPlan &pl, long n_cols, long batch, const cd *d_in, cd *d_out, const cd *d_mom, int mom_rows) {
  const MatsubaraDev &D = pl.dev;
  const long L = D.L;
  constexpr bool fwd = IN_KIND == IN_TAU;
  // Two-radix plans whose second radix is a large prime handled by direct summation (101 | 9999, 89 x 97, ...) cost 4 R FP64 instructions per
  // element and pass: measured 3400 / cost GB/s against 360-410 for the Bluestein convolution (six streaming passes), so beyond a plan cost
  // of ~9.5 the convolution is taken -- 9999 = 99 x 101: 230 -> 395 GB/s, 8633 = 89 x 97: 152 -> 390, 7979 = 79 x 101: 160 -> 375,
  // 5353 = 53 x 101: 184 -> 364 (profiles/r02_bluestein_vs_planned.log).
  const bool costly_plan = (n_cols >= 4 || batch * n_cols >= 8) && !ctx->opt_no_bluestein && tq_pipe_plan_cost(L) > 9.5; // (narrow targets: wide buffer)
  if ((ctx->opt_force_bluestein || costly_plan) && L >= 512) {
    std::shared_ptr<BluePlan> bp;
    TQ_TRY(get_blue_plan(ctx, D, &bp));
    if (bp && sizeof(cd) * (size_t)bp->M * n_cols <= (8ull << 30)) return run_bluestein<fwd>(ctx, D, pl.btw, *bp, n_cols, batch, d_in, d_out, d_mom, mom_rows);
  }
  {
    bool handled = false;
    TQ_TRY(tq_pipe_matsubara(ctx, fwd, D.beta, D.stat, L + 1, (long)D.last + 1, n_cols, batch, d_in, d_out, d_mom, mom_rows, &handled));
    if (handled) return TQ_OK;
    if (fwd) {
      TQ_TRY(tq_col_matsubara_fwd(ctx, D.beta, D.stat, L + 1, (long)D.last + 1, n_cols, batch, d_in, d_out, d_mom, mom_rows, &handled));
      if (handled) return TQ_OK;
    }
  }
  constexpr int SGN = fwd ? 1 : -1;
  Tiling t;
  if (choose_tiling(ctx, L, n_cols, batch, &t) != TQ_OK) { // no factorisation the tile FFT supports: Bluestein, or the direct window DFT
    ctx->err.clear();
    if (!ctx->opt_no_bluestein && L >= 512) { // (short meshes: the direct sum is a handful of launches and as fast)
      std::shared_ptr<BluePlan> bp;
      TQ_TRY(get_blue_plan(ctx, D, &bp));
      // (the convolution buffer of ONE gf has to fit: 8 GB is M = 10^5 with 5000 columns)
      if (bp && sizeof(cd) * (size_t)bp->M * n_cols <= (8ull << 30)) return run_bluestein<fwd>(ctx, D, pl.btw, *bp, n_cols, batch, d_in, d_out, d_mom, mom_rows);
    }
    return run_direct<fwd>(ctx, D, pl.btw, n_cols, batch, d_in, d_out, d_mom, mom_rows);
  }
  TileArgs A{};
  A.n_cols = (int)n_cols;
  A.D = D;
  A.btw = pl.btw;
  A.moments = d_mom;
  A.mom_gf_stride = (long)mom_rows * n_cols;
  A.gt_ref = fwd ? d_in : nullptr;
  A.gt_gf_stride = (L + 1) * n_cols;
  const long in_rows = fwd ? L + 1 : D.n_w, out_rows = fwd ? D.n_w : L + 1;
  A.in = d_in;
  A.out = d_out;
  A.in_gf_stride = in_rows * n_cols;
  A.out_gf_stride = out_rows * n_cols;
  const long maxb = 65535;
  if (!t.two_pass) {
    std::shared_ptr<FftPlanHost> fp;
    TQ_TRY(tq_get_fft_plan(ctx, L, &fp));
    A.P = fp->plan;
    A.TC = t.tc;
    A.skewmask = t.skew;
    A.direct_tw = A.smem_tw = A.rowtab = tile_uses_tables(t.tc);
    A.in_stride = 1;
    A.out_stride = 1;
    for (long b0 = 0; b0 < batch; b0 += maxb) {
      long nb = std::min(maxb, batch - b0);
      TileArgs B = A;
      B.in = A.in + b0 * A.in_gf_stride;
      B.out = A.out + b0 * A.out_gf_stride;
      B.moments = A.moments + b0 * A.mom_gf_stride;
      if (fwd) B.gt_ref = A.gt_ref + b0 * A.gt_gf_stride;
      if (fwd)
        TQ_TRY((launch_tile<IN_TAU, OUT_FREQ, SGN>(ctx, B, 1, nb)));
      else
        TQ_TRY((launch_tile<IN_FREQ, OUT_TAU, SGN>(ctx, B, 1, nb)));
    }
    return TQ_OK;
  }
  // four-step:  index_in = a + Na*b,  index_out = beta + Nb*alpha;  pass 1: length Nb over b for every a, times w_L^{a beta};
  // pass 2: length Na over a for every beta.   scratch[gf][beta][a][col]
  std::shared_ptr<FftPlanHost> fa, fb;
  TQ_TRY(tq_get_fft_plan(ctx, t.Na, &fa));
  TQ_TRY(tq_get_fft_plan(ctx, t.Nb, &fb));
  // chunk the batch so that the scratch of one chunk stays L2 resident between the two kernels
  size_t per_gf = (size_t)L * n_cols * sizeof(cd);
  size_t l2_budget = 64ull << 20;
  if (ctx->opt_scratch_chunk_mb > 0) l2_budget = (size_t)ctx->opt_scratch_chunk_mb << 20;
  long chunk = std::max<long>(1, (long)(l2_budget / per_gf));
  chunk = std::min<long>(chunk, std::min<long>(batch, maxb));
  TQ_TRY(tq_reserve(ctx, ctx->scratch, per_gf * chunk));
  A.scratch = (cd *)ctx->scratch.p;
  A.scratch_gf_stride = L * n_cols;
  A.Na = t.Na;
  A.skewmask = 0;
  for (long b0 = 0; b0 < batch; b0 += chunk) {
    long nb = std::min(chunk, batch - b0);
    TileArgs B = A;
    B.in = A.in + b0 * A.in_gf_stride;
    B.out = A.out + b0 * A.out_gf_stride;
    B.moments = A.moments + b0 * A.mom_gf_stride;
    if (fwd) B.gt_ref = A.gt_ref + b0 * A.gt_gf_stride;
    // pass 1
    B.P = fb->plan;
    B.TC = pick_tc(ctx, t.Nb, n_cols);
    B.direct_tw = B.smem_tw = B.rowtab = tile_uses_tables(B.TC);
    B.in_stride = t.Na;
    B.out_stride = 0;
    if (fwd)
      TQ_TRY((launch_tile<IN_TAU, OUT_SCRATCH, SGN>(ctx, B, t.Na, nb)));
    else
      TQ_TRY((launch_tile<IN_FREQ, OUT_SCRATCH, SGN>(ctx, B, t.Na, nb)));
    // pass 2
    B.P = fa->plan;
    B.TC = t.tc;
    B.direct_tw = B.smem_tw = B.rowtab = tile_uses_tables(B.TC);
    B.in_stride = 0;
    B.out_stride = t.Nb;
    if (fwd)
      TQ_TRY((launch_tile<IN_SCRATCH, OUT_FREQ, SGN>(ctx, B, t.Nb, nb)));
    else
      TQ_TRY((launch_tile<IN_SCRATCH, OUT_TAU, SGN>(ctx, B, t.Nb, nb)));
  }
  return TQ_OK;
}

// tail_fit.cu
tq_status tq_fit_tail_device(tq_ctx *ctx, double beta, int statistic, long n_iw, double tail_fraction, int n_tail_max, int expansion_order,
                             int hermitian, int inner_dim, long n_cols, long batch, const cd *d_g, const cd *d_known, int n_known,
                             cd *d_out_moments /* [batch][TQ_MAX_MOMENTS][n_cols] dense with n_moments rows */, int *n_moments,
                             double *d_err /* device double[batch] */);

struct StatusWords {
  unsigned long long max_m0_bits;
  unsigned long long joint_der[2];
};

// real <-> complex at the boundary (fourier.hpp:78-81 promotes a real-valued gf; functions2.hpp:223-246 is_gf_real / real(g))
static __global__ void k_promote_real(const double *__restrict__ re, cd *__restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n2 = n >> 1; // two reals in, two complex numbers out per thread and step (16-byte loads, 2 x 16-byte stores)
  if ((((uintptr_t)re) & 15) == 0) {
    const double2 *re2 = reinterpret_cast<const double2 *>(re);
    for (size_t j = i; j < n2; j += stride) {
      const double2 v = re2[j];
      out[2 * j] = cmk(v.x, 0.0);
      out[2 * j + 1] = cmk(v.y, 0.0);
    }
    if (i == 0 && (n & 1)) out[n - 1] = cmk(re[n - 1], 0.0);
  } else {
    for (size_t j = i; j < n; j += stride) out[j] = cmk(re[j], 0.0);
  }
}
// real part of every element, and per gf the largest |imaginary part| (is_gf_real's criterion)
static __global__ void k_real_part(const cd *__restrict__ in, double *__restrict__ out, size_t per_gf, unsigned long long *__restrict__ max_im_bits) {
  const size_t gf = blockIdx.y;
  const cd *p = in + gf * per_gf;
  double *q = out + gf * per_gf;
  double m = 0.0;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < per_gf; j += (size_t)gridDim.x * blockDim.x) {
    const cd v = p[j];
    q[j] = v.x;
    m = fmax(m, fabs(v.y));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(max_im_bits + gf, (unsigned long long)__double_as_longlong(m)); // (non-negative doubles order like their bits)
}

// Two real-valued columns per complex column ("two for one"): z[., p] = x[., 2p] + i x[., 2p + 1] (a last odd column: + i 0).  Every step of the
// Matsubara transform is linear over C, so Z = F(x_2p) + i F(x_2p+1), and for real x  F(x)(i w_{-n}) = conj F(x)(i w_n)  (data index i <-> n_w - 1 - i
// for both statistics) separates the two:  F(x_2p)_i = (Z_i + conj Z_i') / 2,  F(x_2p+1)_i = (Z_i - conj Z_i') / (2 i).
static __global__ void k_pack_real_pairs(const double *__restrict__ re, cd *__restrict__ z, long rows, int nc, int nc2) {
  const long total = rows * nc2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / nc2;
    const int p = (int)(i - r * nc2);
    const double *src = re + r * nc + 2 * p;
    z[i] = cmk(src[0], 2 * p + 1 < nc ? src[1] : 0.0);
  }
}
static __global__ void k_unpack_pairs(const cd *__restrict__ Z, cd *__restrict__ out, int n_w, int nc, int nc2) {
  const long gf = blockIdx.y;
  const long per = (long)n_w * nc2;
  const cd *z = Z + gf * per;
  cd *o = out + gf * (long)n_w * nc;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / nc2), p = (int)(i - (long)r * nc2);
    const cd a = z[i], b = z[(long)(n_w - 1 - r) * nc2 + p];
    o[(long)r * nc + 2 * p] = cmk(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
    if (2 * p + 1 < nc) o[(long)r * nc + 2 * p + 1] = cmk(0.5 * (a.y + b.y), 0.5 * (b.x - a.x));
  }
}

// The forward transform of nct = ceil(n_others / 2) packed columns (see k_pack_real_pairs) and the separation of the result.
static tq_status fourier_tau_to_iw_packed(tq_ctx *ctx, MatsubaraPlan &pl, double beta, int statistic, long n_tau, long n_iw, long n_others, long nct,
                                          long batch, const cd *d_in, const tq_complex *packed_moments /* host or nullptr */, int n_moments,
                                          tq_complex *gw, int *warn) {
  const long L = n_tau - 1, last = n_iw - 1, n_w = pl.dev.n_w;
  const bool have_moments = packed_moments != nullptr && n_moments > 0;
  const size_t out_bytes = sizeof(cd) * (size_t)batch * n_w * n_others;
  const void *d_known = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  if (have_moments) TQ_TRY(tq_stage_in(ctx, packed_moments, sizeof(cd) * (size_t)batch * n_moments * nct, ctx->stage_in2, &d_known));
  TQ_TRY(tq_stage_out_begin(ctx, gw, out_bytes, ctx->stage_out, &d_out, &out_host));
  TQ_TRY(tq_reserve(ctx, ctx->stage_in3, sizeof(cd) * (size_t)batch * n_w * nct));
  cd *d_Z = (cd *)ctx->stage_in3.p;
  const size_t mom_bytes = sizeof(cd) * (size_t)batch * 4 * nct;
  const size_t warn_off = (mom_bytes + 255) & ~size_t(255);
  const size_t st_off = (warn_off + sizeof(int) * batch + 255) & ~size_t(255);
  TQ_TRY(tq_reserve(ctx, ctx->small, st_off + sizeof(StatusWords)));
  cd *d_mom = (cd *)ctx->small.p;
  int *d_warn = (int *)((char *)ctx->small.p + warn_off);
  StatusWords *d_st = (StatusWords *)((char *)ctx->small.p + st_off);
  TQ_CUDA(ctx, cudaMemsetAsync(d_warn, 0, st_off - warn_off + sizeof(StatusWords), ctx->stream));
  if (have_moments) {
    const long total = batch * nct;
    KernelScope ks(ctx, "prepare_known_moments");
    k_prepare_known_moments<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_known, n_moments, (int)nct, total, d_mom,
                                                                                        &d_st->max_m0_bits);
  } else {
    if (batch > 0x7fffffffL) return tq_fail(ctx, TQ_ERR_INVALID, "batch too large");
    KernelScope ks(ctx, "moments_from_tau");
    k_moments_from_tau<<<(unsigned)batch, 256, 0, ctx->stream>>>(d_in, n_tau * nct, (int)n_tau, (int)nct, beta, statistic, d_mom, d_warn, nullptr, 0, 1);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(run_transform<IN_TAU>(ctx, pl, nct, batch, d_in, d_Z, d_mom, 4));
  {
    KernelScope ks(ctx, "unpack_pairs");
    const unsigned gx = (unsigned)std::max<long>(1, std::min<long>(((long)n_w * nct + 255) / 256, (148 * 16 + batch - 1) / batch));
    k_unpack_pairs<<<dim3(gx, (unsigned)batch), 256, 0, ctx->stream>>>(d_Z, (cd *)d_out, (int)n_w, (int)n_others, (int)nct);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(tq_reserve_pinned(ctx, sizeof(int) * batch + sizeof(StatusWords)));
  int *h_warn = (int *)ctx->pinned;
  TQ_CUDA(ctx, cudaMemcpyAsync(h_warn, d_warn, sizeof(int) * batch, cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int mesh_warn = (L < 6 * (last + 1)) ? TQ_WARN_SHORT_TAU_MESH : 0; // :131-133
  if (warn)
    for (long b = 0; b < batch; ++b) warn[b] = h_warn[b] | mesh_warn;
  TQ_TRY(tq_stage_out_finish(ctx, gw, out_bytes, d_out, out_host));
  return TQ_OK;
}

static tq_status fourier_tau_to_iw_impl(tq_ctx *ctx, double beta, int statistic, long n_tau, long n_iw, long n_others, long batch,
                                        const tq_complex *gt, const tq_complex *moments, int n_moments, tq_complex *gw, int *warn, bool joint,
                                        bool real_in = false) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!gt || !gw || n_tau < 2 || n_iw < 1 || n_others < 1 || batch < 1 || !(beta > 0) || (statistic != TQ_FERMION && statistic != TQ_BOSON))
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_fourier_tau_to_iw: invalid argument");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long L = n_tau - 1;
  const long last = n_iw - 1;
  // fourier_matsubara.cpp:127-129
  if (L < 2 * (last + 1))
    return tq_fail(ctx, TQ_ERR_RUNTIME,
                   "Fourier: The time mesh mush be at least twice as long as the number of positive frequencies :\n gt.mesh().size() =  " +
                      std::to_string(n_tau) + " gw.mesh().last_index()" + std::to_string(last));
  const bool have_moments = moments != nullptr && n_moments > 0;
  if (!have_moments && n_tau < 17) return tq_fail(ctx, TQ_ERR_INVALID, "tq_fourier_tau_to_iw: n_tau < 17 needs known moments");
  std::shared_ptr<MatsubaraPlan> pl;
  TQ_TRY(get_matsubara_plan(ctx, beta, statistic, n_tau, n_iw, &pl));
  const long n_w = pl->dev.n_w;
  const size_t in_bytes = sizeof(cd) * (size_t)batch * n_tau * n_others, out_bytes = sizeof(cd) * (size_t)batch * n_w * n_others;
  const void *d_in = nullptr, *d_known = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  // columns of the transform itself: n_others, or half of them when two real-valued columns share a complex one
  long nct = n_others;
  std::vector<cd> packed_moments;
  if (real_in) {
    // real-valued G(tau) (gf<imtime, T::real_t>): the reference promotes it to complex before the transform (fourier.hpp:78-81);
    // here only the 8-byte reals cross PCIe / are read from the caller's buffer
    const size_t n_el = (size_t)batch * n_tau * n_others;
    const void *d_re = nullptr;
    TQ_TRY(tq_stage_in(ctx, gt, sizeof(double) * n_el, ctx->stage_in, &d_re));
    // "two for one": needs real moments (those of a real G(tau) are) that the host can look at, and enough columns to stay wide
    bool pack = !joint && n_others >= 8 && !ctx->opt_no_two_for_one && !(have_moments && tq_is_device_ptr(moments));
    const long nc2 = (n_others + 1) / 2;
    if (pack && have_moments) {
      const cd *m = (const cd *)moments;
      packed_moments.assign((size_t)batch * n_moments * nc2, cmk(0.0, 0.0));
      for (long b = 0; b < batch && pack; ++b)
        for (int r = 0; r < n_moments && pack; ++r)
          for (long c = 0; c < n_others; ++c) {
            const cd v = m[((size_t)b * n_moments + r) * n_others + c];
            if (v.y != 0.0) {
              pack = false;
              break;
            }
            if (r == 0 && !(std::fabs(v.x) < 1e-8)) // :116-118 (checked here per column: the packed 0th moment would mix two of them)
              return tq_fail(ctx, TQ_ERR_RUNTIME,
                             "ERROR: Direct Fourier implementation requires vanishing 0th moment\n  error is :" + std::to_string(std::fabs(v.x)));
            cd &z = packed_moments[((size_t)b * n_moments + r) * nc2 + c / 2];
            if (r > 0) (c & 1 ? z.y : z.x) = v.x;
          }
    }
    if (pack) {
      nct = nc2;
      if (n_others % 2 == 0 && ((uintptr_t)d_re & 15) == 0) {
        d_in = d_re; // [rows][2 m] doubles ARE [rows][m] complex numbers
      } else {
        TQ_TRY(tq_reserve(ctx, ctx->promote, sizeof(cd) * (size_t)batch * n_tau * nc2));
        KernelScope ks(ctx, "pack_real_pairs");
        k_pack_real_pairs<<<(unsigned)std::min<size_t>(((size_t)batch * n_tau * nc2 + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(
           (const double *)d_re, (cd *)ctx->promote.p, batch * n_tau, (int)n_others, (int)nc2);
        tq_count_launch(ctx);
        d_in = ctx->promote.p;
      }
    } else {
      TQ_TRY(tq_reserve(ctx, ctx->promote, in_bytes));
      {
        KernelScope ks(ctx, "promote_real");
        k_promote_real<<<(unsigned)std::min<size_t>((n_el + 511) / 512, 148 * 16), 256, 0, ctx->stream>>>((const double *)d_re, (cd *)ctx->promote.p, n_el);
      }
      tq_count_launch(ctx);
      d_in = ctx->promote.p;
    }
    TQ_CUDA(ctx, cudaGetLastError());
  } else {
    TQ_TRY(tq_stage_in(ctx, gt, in_bytes, ctx->stage_in, &d_in));
  }
  const bool packed = nct != n_others;
  if (packed) {
    moments = packed_moments.empty() ? nullptr : (const tq_complex *)packed_moments.data();
    return fourier_tau_to_iw_packed(ctx, *pl, beta, statistic, n_tau, n_iw, n_others, nct, batch, (const cd *)d_in, moments, n_moments, gw, warn);
  }
  if (have_moments) TQ_TRY(tq_stage_in(ctx, moments, sizeof(cd) * (size_t)batch * n_moments * n_others, ctx->stage_in2, &d_known));
  TQ_TRY(tq_stage_out_begin(ctx, gw, out_bytes, ctx->stage_out, &d_out, &out_host));
  // small buffer: [batch][4][n_others] moments, int warn[batch], status words
  const size_t mom_bytes = sizeof(cd) * (size_t)batch * 4 * n_others;
  const size_t warn_off = (mom_bytes + 255) & ~size_t(255);
  const size_t st_off = (warn_off + sizeof(int) * batch + 255) & ~size_t(255);
  TQ_TRY(tq_reserve(ctx, ctx->small, st_off + sizeof(StatusWords)));
  cd *d_mom = (cd *)ctx->small.p;
  int *d_warn = (int *)((char *)ctx->small.p + warn_off);
  StatusWords *d_st = (StatusWords *)((char *)ctx->small.p + st_off);
  TQ_CUDA(ctx, cudaMemsetAsync(d_warn, 0, st_off - warn_off + sizeof(StatusWords), ctx->stream));
  if (have_moments) {
    long total = batch * n_others;
    KernelScope ks(ctx, "prepare_known_moments");
    k_prepare_known_moments<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_known, n_moments, (int)n_others, total,
                                                                                        d_mom, &d_st->max_m0_bits);
  } else {
    if (batch > 0x7fffffffL) return tq_fail(ctx, TQ_ERR_INVALID, "batch too large");
    KernelScope ks(ctx, "moments_from_tau");
    if (joint) {
      k_moments_from_tau<<<(unsigned)batch, 256, 0, ctx->stream>>>((const cd *)d_in, n_tau * n_others, (int)n_tau, (int)n_others, beta,
                                                                    statistic, d_mom, d_warn, d_st->joint_der, 1);
      tq_count_launch(ctx);
    }
    k_moments_from_tau<<<(unsigned)batch, 256, 0, ctx->stream>>>((const cd *)d_in, n_tau * n_others, (int)n_tau, (int)n_others, beta, statistic,
                                                                  d_mom, d_warn, joint ? d_st->joint_der : nullptr, joint ? 2 : 0);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(run_transform<IN_TAU>(ctx, *pl, n_others, batch, (const cd *)d_in, (cd *)d_out, d_mom, 4));
  // status read-back (one small copy; the reference's checks need the values on the host)
  TQ_TRY(tq_reserve_pinned(ctx, sizeof(int) * batch + sizeof(StatusWords)));
  int *h_warn = (int *)ctx->pinned;
  StatusWords *h_st = (StatusWords *)((char *)ctx->pinned + sizeof(int) * batch);
  TQ_CUDA(ctx, cudaMemcpyAsync(h_warn, d_warn, sizeof(int) * batch, cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaMemcpyAsync(h_st, d_st, sizeof(StatusWords), cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (have_moments) {
    double a0;
    memcpy(&a0, &h_st->max_m0_bits, sizeof(double));
    if (!(a0 < 1e-8)) // :116-118
      return tq_fail(ctx, TQ_ERR_RUNTIME, "ERROR: Direct Fourier implementation requires vanishing 0th moment\n  error is :" + std::to_string(a0));
  }
  const int mesh_warn = (L < 6 * (last + 1)) ? TQ_WARN_SHORT_TAU_MESH : 0; // :131-133
  if (warn) {
    if (joint)
      warn[0] = h_warn[0] | mesh_warn; // one gf: one decision
    else
      for (long b = 0; b < batch; ++b) warn[b] = h_warn[b] | mesh_warn;
  }
  TQ_TRY(tq_stage_out_finish(ctx, gw, out_bytes, d_out, out_host));
  return TQ_OK;
}

// -------------------------------------------------------------------------------------------------
// Host-buffer calls: overlap inside the call.  The reference's call model is one host thread and one call for the whole
// BlockGf / batch (fourier.hpp:60-85, block/map.hpp:35-40).  With HOST pointers a single stream would run H2D -> kernels ->
// D2H strictly one after the other; instead the batch is cut into chunks that are dealt round-robin to a few internal
// lanes (child contexts: own stream, staging buffers and plan caches, driven by short-lived host threads), so that the
// H2D copy of one chunk, the kernels of another and the D2H copy of a third overlap (PCIe is full duplex) -- also for
// pageable memory, where every cudaMemcpyAsync is a staged, host-blocking copy.  Results do not depend on the chunking:
// every gf is transformed independently.
// -------------------------------------------------------------------------------------------------
static bool is_pageable(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}
template <class Fn>
static tq_status run_host_lanes(tq_ctx *ctx, long batch, size_t bytes_per_gf, bool pageable, const Fn &fn /* (lane ctx, b0, nb) */) {
  // a copy from / to pageable memory is staged by the driver at the speed of ONE host core per call: more lanes, more cores
  const int want = (int)std::max<long>(1, pageable ? std::max<long>(ctx->opt_host_lanes, std::min<long>(8, std::thread::hardware_concurrency() / 2))
                                                   : ctx->opt_host_lanes);
  const long chunk_bytes = std::max<long>(1, ctx->opt_host_chunk_kb) << 10;
  const long per_chunk = std::max<long>(1, chunk_bytes / (long)std::max<size_t>(bytes_per_gf, 1));
  const long n_chunks = (batch + per_chunk - 1) / per_chunk;
  int lanes = (int)std::min<long>(want, n_chunks);
  // Pageable buffers: a lane stages its chunk through its pinned slots at the speed of ONE core, so the overlap the lanes buy is worth less than
  // spreading every copy over all the helper threads (tq_stage_in / tq_stage_out_finish, tq_ctx.cu: stage_parallel) in ONE un-chunked call.
  // Measured, blocks of configs[1] in NumPy arrays, forward call: 2 blocks 5.0 -> 3.1 ms, 4: 8.0 -> 5.3, 8: 11.0 -> 9.8, 16: 22.9 -> 18.9, 32: 38.3 -> 36.9.
  if (pageable && ctx->opt_bounce == 1 && (size_t)batch * bytes_per_gf >= (12u << 20)) lanes = 1;
  if (lanes <= 1) return fn(ctx, 0, batch);
  while ((int)ctx->lanes.size() < lanes) {
    tq_ctx *c = nullptr;
    tq_status st = tq_create(ctx->device, &c);
    if (st != TQ_OK) return tq_fail(ctx, st, "cannot create an internal lane context");
    c->opt_host_lanes = 1;
    ctx->lanes.push_back(c);
  }
  for (int l = 0; l < lanes; ++l) {
    ctx->lanes[l]->opt_no_pipe = ctx->opt_no_pipe;
    ctx->lanes[l]->opt_no_col = ctx->opt_no_col;
    ctx->lanes[l]->opt_no_tr = ctx->opt_no_tr;
    ctx->lanes[l]->opt_no_two_for_one = ctx->opt_no_two_for_one;
    ctx->lanes[l]->opt_no_bluestein = ctx->opt_no_bluestein;
    ctx->lanes[l]->opt_force_bluestein = ctx->opt_force_bluestein;
    ctx->lanes[l]->opt_blue_chunk_mb = ctx->opt_blue_chunk_mb;
    ctx->lanes[l]->call_batch = batch;
    ctx->lanes[l]->opt_scratch_chunk_mb = ctx->opt_scratch_chunk_mb;
    ctx->lanes[l]->opt_bounce = ctx->opt_bounce ? 2 : 0; // (2: this context is a lane of a multi-lane call)
    ctx->lanes[l]->opt_pipe_lag0 = ctx->opt_pipe_lag0;
    ctx->lanes[l]->opt_force_planned = ctx->opt_force_planned;
    ctx->lanes[l]->profiling = false;
  }
  std::vector<tq_status> sts(lanes, TQ_OK);
  std::vector<std::thread> th;
  auto work = [&](int l) {
    cudaSetDevice(ctx->device);
    for (long c = l; c < n_chunks; c += lanes) {
      const long b0 = c * per_chunk, nb = std::min(per_chunk, batch - b0);
      tq_status st = fn(ctx->lanes[l], b0, nb);
      if (st != TQ_OK) {
        sts[l] = st;
        return;
      }
    }
  };
  for (int l = 1; l < lanes; ++l) th.emplace_back(work, l);
  work(0);
  for (auto &t : th) t.join();
  for (int l = 0; l < lanes; ++l) {
    ctx->launches += ctx->lanes[l]->launches;
    ctx->lanes[l]->launches = 0;
  }
  for (int l = 0; l < lanes; ++l)
    if (sts[l] != TQ_OK) return tq_fail(ctx, sts[l], ctx->lanes[l]->err);
  return TQ_OK;
}

extern "C" tq_status tq_fourier_tau_to_iw(tq_ctx *ctx, double beta, int statistic, long n_tau, long n_iw, long n_others, long batch,
                                          const tq_complex *gt, const tq_complex *moments, int n_moments, tq_complex *gw, int *warn) {
  if (ctx && gt && gw && batch > 1 && n_tau >= 2 && n_iw >= 1 && n_others >= 1 && ctx->opt_host_lanes > 1 && !tq_is_device_ptr(gt) &&
      !tq_is_device_ptr(gw) && !(moments && tq_is_device_ptr(moments))) {
    const long n_w = statistic == TQ_FERMION ? 2 * n_iw : 2 * n_iw - 1;
    return run_host_lanes(ctx, batch, sizeof(cd) * (size_t)n_tau * n_others, is_pageable(gt) || is_pageable(gw), [&](tq_ctx *c, long b0, long nb) {
      return fourier_tau_to_iw_impl(c, beta, statistic, n_tau, n_iw, n_others, nb, gt + (size_t)b0 * n_tau * n_others,
                                    (moments && n_moments > 0) ? moments + (size_t)b0 * n_moments * n_others : nullptr, n_moments,
                                    gw + (size_t)b0 * n_w * n_others, warn ? warn + b0 : nullptr, false);
    });
  }
  return fourier_tau_to_iw_impl(ctx, beta, statistic, n_tau, n_iw, n_others, batch, gt, moments, n_moments, gw, warn, false);
}

extern "C" tq_status tq_fourier_tau_to_iw_real(tq_ctx *ctx, double beta, int statistic, long n_tau, long n_iw, long n_others, long batch,
                                               const double *gt, const tq_complex *moments, int n_moments, tq_complex *gw, int *warn) {
  if (ctx && gt && gw && batch > 1 && n_tau >= 2 && n_iw >= 1 && n_others >= 1 && ctx->opt_host_lanes > 1 && !tq_is_device_ptr(gt) &&
      !tq_is_device_ptr(gw) && !(moments && tq_is_device_ptr(moments))) {
    const long n_w = statistic == TQ_FERMION ? 2 * n_iw : 2 * n_iw - 1;
    // (chunks are sized by the bytes that cross PCIe: twice as many gfs per chunk as in the complex call)
    return run_host_lanes(ctx, batch, sizeof(double) * (size_t)n_tau * n_others, is_pageable(gt) || is_pageable(gw), [&](tq_ctx *c, long b0, long nb) {
      return fourier_tau_to_iw_impl(c, beta, statistic, n_tau, n_iw, n_others, nb, (const tq_complex *)(gt + (size_t)b0 * n_tau * n_others),
                                    (moments && n_moments > 0) ? moments + (size_t)b0 * n_moments * n_others : nullptr, n_moments,
                                    gw + (size_t)b0 * n_w * n_others, warn ? warn + b0 : nullptr, false, true);
    });
  }
  return fourier_tau_to_iw_impl(ctx, beta, statistic, n_tau, n_iw, n_others, batch, (const tq_complex *)gt, moments, n_moments, gw, warn, false, true);
}

extern "C" tq_status tq_fourier_tau_to_iw_axis(tq_ctx *ctx, double beta, int statistic, long n_tau, long n_iw, long outer, long inner,
                                               const tq_complex *gt, const tq_complex *moments, int n_moments, tq_complex *gw, int *warn) {
  return fourier_tau_to_iw_impl(ctx, beta, statistic, n_tau, n_iw, inner, outer, gt, moments, n_moments, gw, warn, true);
}

static tq_status fourier_iw_to_tau_impl(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_tau, long n_others, long batch,
                                        const tq_complex *gw, const tq_complex *moments, int n_moments, tq_complex *gt, int *warn,
                                        double *tail_err, double *max_imag = nullptr /* != nullptr: gt is double[...] (real part) */) {
  const bool real_out = max_imag != nullptr;
  if (!ctx) return TQ_ERR_INVALID;
  if (!gt || !gw || n_tau < 2 || n_iw < 1 || n_others < 1 || batch < 1 || !(beta > 0) || (statistic != TQ_FERMION && statistic != TQ_BOSON))
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_fourier_iw_to_tau: invalid argument");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long L = n_tau - 1;
  const long last = n_iw - 1;
  std::shared_ptr<MatsubaraPlan> pl;
  TQ_TRY(get_matsubara_plan(ctx, beta, statistic, n_tau, n_iw, &pl));
  const long n_w = pl->dev.n_w;
  const size_t in_bytes = sizeof(cd) * (size_t)batch * n_w * n_others, out_bytes = sizeof(cd) * (size_t)batch * n_tau * n_others;
  const void *d_in = nullptr, *d_known = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, gw, in_bytes, ctx->stage_in, &d_in));
  const bool have_moments = moments != nullptr && n_moments > 0;
  if (have_moments) TQ_TRY(tq_stage_in(ctx, moments, sizeof(cd) * (size_t)batch * n_moments * n_others, ctx->stage_in2, &d_known));
  // small buffer: moments [batch][TQ_MAX_MOMENTS][n_others], err double[batch], status
  const size_t mom_bytes = sizeof(cd) * (size_t)batch * TQ_MAX_MOMENTS * n_others;
  const size_t err_off = (mom_bytes + 255) & ~size_t(255);
  const size_t st_off = (err_off + sizeof(double) * batch + 255) & ~size_t(255);
  const size_t zk_off = (st_off + sizeof(StatusWords) + 255) & ~size_t(255);
  TQ_TRY(tq_reserve(ctx, ctx->small, zk_off + sizeof(cd) * (size_t)batch * n_others));
  cd *d_mom = (cd *)ctx->small.p;
  double *d_err = (double *)((char *)ctx->small.p + err_off);
  StatusWords *d_st = (StatusWords *)((char *)ctx->small.p + st_off);
  cd *d_zero_known = (cd *)((char *)ctx->small.p + zk_off);
  TQ_CUDA(ctx, cudaMemsetAsync(d_err, 0, zk_off - err_off + sizeof(cd) * (size_t)batch * n_others, ctx->stream));
  int mom_rows = 0;
  bool fitted = false;
  if (have_moments && n_moments >= 4) {
    long total = batch * n_others;
    k_prepare_known_moments<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_known, n_moments, (int)n_others, total,
                                                                                        d_mom, &d_st->max_m0_bits);
    tq_count_launch(ctx);
    TQ_CUDA(ctx, cudaGetLastError());
    mom_rows = 4;
  } else {
    // :197  no moments -> make_zero_tail(gw, 1);  :203  fewer than 4 -> fit_tail(gw, known_moments)
    const cd *kn = have_moments ? (const cd *)d_known : d_zero_known;
    int nk = have_moments ? n_moments : 1;
    if (have_moments) {
      // the 0th-moment assertion (:199-201) is made on the user's moments before the fit
      long total = batch * n_others;
      TQ_TRY(tq_reserve(ctx, ctx->stage_in3, sizeof(cd) * (size_t)batch * 4 * n_others));
      k_prepare_known_moments<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_known, n_moments, (int)n_others, total,
                                                                                          (cd *)ctx->stage_in3.p, &d_st->max_m0_bits);
      tq_count_launch(ctx);
    }
    tq_status st = tq_fit_tail_device(ctx, beta, statistic, n_iw, 0.2, 30, -1, 0, 1, n_others, batch, (const cd *)d_in, kn, nk, d_mom, &mom_rows,
                                      d_err);
    if (st != TQ_OK) return st;
    fitted = true;
    if (!(mom_rows > 3)) // :211
      return tq_fail(ctx, TQ_ERR_RUNTIME, "ERROR: Inverse Fourier implementation requires at least a proper 3rd high-frequency moment\n");
  }
  if (L < 2 * (last + 1)) // :218-220
    return tq_fail(ctx, TQ_ERR_RUNTIME,
                   "Inverse Fourier: The time mesh mush be at least twice as long as the freq mesh :\n gt.mesh().size() =  " +
                      std::to_string(n_tau) + " gw.mesh().last_index()" + std::to_string(last) + "\n");
  unsigned long long *d_maxim = nullptr;
  if (real_out) {
    // the transform writes complex G(tau) into a device temporary; the caller's buffer receives the real part only
    // (real(g), functions2.hpp:242-246) and max |Im| per gf (the quantity is_gf_real compares with its tolerance, :223-226)
    void *d_re = nullptr;
    TQ_TRY(tq_stage_out_begin(ctx, gt, out_bytes / 2, ctx->stage_out, &d_re, &out_host));
    TQ_TRY(tq_reserve(ctx, ctx->promote, out_bytes + sizeof(unsigned long long) * (size_t)batch));
    d_maxim = (unsigned long long *)((char *)ctx->promote.p + out_bytes);
    TQ_CUDA(ctx, cudaMemsetAsync(d_maxim, 0, sizeof(unsigned long long) * (size_t)batch, ctx->stream));
    TQ_TRY(run_transform<IN_FREQ>(ctx, *pl, n_others, batch, (const cd *)d_in, (cd *)ctx->promote.p, d_mom, mom_rows));
    const size_t per_gf = (size_t)n_tau * n_others;
    {
      KernelScope ks(ctx, "real_part");
      const unsigned gx = (unsigned)std::max<size_t>(1, std::min<size_t>((per_gf + 1023) / 1024, (148 * 16 + batch - 1) / batch));
      k_real_part<<<dim3(gx, (unsigned)batch), 256, 0, ctx->stream>>>((const cd *)ctx->promote.p, (double *)d_re, per_gf, d_maxim);
    }
    tq_count_launch(ctx);
    TQ_CUDA(ctx, cudaGetLastError());
    d_out = d_re;
  } else {
    TQ_TRY(tq_stage_out_begin(ctx, gt, out_bytes, ctx->stage_out, &d_out, &out_host));
    TQ_TRY(run_transform<IN_FREQ>(ctx, *pl, n_others, batch, (const cd *)d_in, (cd *)d_out, d_mom, mom_rows));
  }
  TQ_TRY(tq_reserve_pinned(ctx, sizeof(double) * batch * 2 + sizeof(StatusWords)));
  double *h_err = (double *)ctx->pinned;
  double *h_maxim = h_err + batch;
  StatusWords *h_st = (StatusWords *)((char *)ctx->pinned + sizeof(double) * batch * 2);
  if (real_out) TQ_CUDA(ctx, cudaMemcpyAsync(h_maxim, d_maxim, sizeof(double) * batch, cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaMemcpyAsync(h_err, d_err, sizeof(double) * batch, cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaMemcpyAsync(h_st, d_st, sizeof(StatusWords), cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (have_moments) {
    double a0;
    memcpy(&a0, &h_st->max_m0_bits, sizeof(double));
    if (!(a0 < 1e-8))
      return tq_fail(ctx, TQ_ERR_RUNTIME,
                     "ERROR: Inverse Fourier implementation requires vanishing 0th moment\n  error is :" + std::to_string(a0) + "\n");
  }
  int w_any = 0;
  if (fitted) {
    for (long b = 0; b < batch; ++b) {
      double e = h_err[b];
      if (!(e < 1e-2)) // :205-207
        return tq_fail(ctx, TQ_ERR_RUNTIME,
                       "ERROR: High frequency moments have an error greater than 1e-2.\n  Error = " + std::to_string(e) +
                          "\n Please make sure you treat the constant offset analytically!\n");
      if (warn) warn[b] = (e > 1e-4) ? TQ_WARN_TAIL_ERR : 0; // :208-210
      w_any |= (e > 1e-4);
    }
  } else if (warn) {
    for (long b = 0; b < batch; ++b) warn[b] = 0;
  }
  (void)w_any;
  if (tail_err)
    for (long b = 0; b < batch; ++b) tail_err[b] = fitted ? h_err[b] : 0.0;
  if (real_out)
    for (long b = 0; b < batch; ++b) max_imag[b] = h_maxim[b];
  TQ_TRY(tq_stage_out_finish(ctx, gt, real_out ? out_bytes / 2 : out_bytes, d_out, out_host));
  return TQ_OK;
}

extern "C" tq_status tq_fourier_iw_to_tau_real(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_tau, long n_others, long batch,
                                               const tq_complex *gw, const tq_complex *moments, int n_moments, double *gt, int *warn,
                                               double *tail_err, double *max_imag) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!max_imag) return tq_fail(ctx, TQ_ERR_INVALID, "tq_fourier_iw_to_tau_real: max_imag must not be NULL");
  if (gt && gw && batch > 1 && n_tau >= 2 && n_iw >= 1 && n_others >= 1 && ctx->opt_host_lanes > 1 && !tq_is_device_ptr(gt) &&
      !tq_is_device_ptr(gw) && !(moments && tq_is_device_ptr(moments))) {
    const long n_w = statistic == TQ_FERMION ? 2 * n_iw : 2 * n_iw - 1;
    return run_host_lanes(ctx, batch, sizeof(double) * (size_t)n_tau * n_others, is_pageable(gt) || is_pageable(gw), [&](tq_ctx *c, long b0, long nb) {
      return fourier_iw_to_tau_impl(c, beta, statistic, n_iw, n_tau, n_others, nb, gw + (size_t)b0 * n_w * n_others,
                                    (moments && n_moments > 0) ? moments + (size_t)b0 * n_moments * n_others : nullptr, n_moments,
                                    (tq_complex *)(gt + (size_t)b0 * n_tau * n_others), warn ? warn + b0 : nullptr, tail_err ? tail_err + b0 : nullptr,
                                    max_imag + b0);
    });
  }
  return fourier_iw_to_tau_impl(ctx, beta, statistic, n_iw, n_tau, n_others, batch, gw, moments, n_moments, (tq_complex *)gt, warn, tail_err, max_imag);
}

extern "C" tq_status tq_fourier_iw_to_tau(tq_ctx *ctx, double beta, int statistic, long n_iw, long n_tau, long n_others, long batch,
                                          const tq_complex *gw, const tq_complex *moments, int n_moments, tq_complex *gt, int *warn,
                                          double *tail_err) {
  if (ctx && gt && gw && batch > 1 && n_tau >= 2 && n_iw >= 1 && n_others >= 1 && ctx->opt_host_lanes > 1 && !tq_is_device_ptr(gt) &&
      !tq_is_device_ptr(gw) && !(moments && tq_is_device_ptr(moments))) {
    const long n_w = statistic == TQ_FERMION ? 2 * n_iw : 2 * n_iw - 1;
    return run_host_lanes(ctx, batch, sizeof(cd) * (size_t)n_tau * n_others, is_pageable(gt) || is_pageable(gw), [&](tq_ctx *c, long b0, long nb) {
      return fourier_iw_to_tau_impl(c, beta, statistic, n_iw, n_tau, n_others, nb, gw + (size_t)b0 * n_w * n_others,
                                    (moments && n_moments > 0) ? moments + (size_t)b0 * n_moments * n_others : nullptr, n_moments,
                                    gt + (size_t)b0 * n_tau * n_others, warn ? warn + b0 : nullptr, tail_err ? tail_err + b0 : nullptr);
    });
  }
  return fourier_iw_to_tau_impl(ctx, beta, statistic, n_iw, n_tau, n_others, batch, gw, moments, n_moments, gt, warn, tail_err);
}

// -------------------------------------------------------------------------------------------------
// tq_plan_describe: which engine an imtime <-> imfreq transform of this shape takes (host only, no device needed) -- the analogue of
// printing an FFTW plan.  JSON: {"L":..., "engine_fwd": "...", "engine_inv": "...", "plan_fwd": "...", "plan_inv": "...", "bluestein_M": [M1, M2] | null}
// -------------------------------------------------------------------------------------------------
std::string tq_pipe_plan_text(long L, int dir); // fourier_pipe.cu
extern "C" tq_status tq_plan_describe(long n_tau, long n_iw, int statistic, long n_cols, long batch, char *buf, size_t len) {
  if (!buf || len < 2 || n_tau < 2 || n_iw < 1 || n_cols < 1 || batch < 1 || (statistic != TQ_FERMION && statistic != TQ_BOSON)) return TQ_ERR_INVALID;
  const long L = n_tau - 1;
  const long n_w = statistic == TQ_FERMION ? 2 * n_iw : 2 * n_iw - 1;
  const double cost = tq_pipe_plan_cost(L);
  int M1 = 0, M2 = 0;
  const bool blue_ok = L >= 512 && blue_pick_lengths(L + n_w - 1, &M1, &M2);
  const bool costly = (n_cols >= 4 || batch * n_cols >= 8) && cost > 9.5;
  FftPlan probe;
  const bool generic_ok = tq_fft_factorize((int)std::min<long>(L, 1L << 30), probe);
  auto engine = [&](bool fwd) -> std::string {
    if (costly && blue_ok) return "bluestein";
    if (cost >= 0 && L >= 256) {
      if (n_cols >= 4) return "planned";
      if (fwd && L == 10000 && n_cols == 1) return "single-column";
      if (batch * n_cols >= 8) return "planned (gfs as columns) or single-kernel";
      return "planned or single-kernel";
    }
    if (generic_ok) return "generic";
    return blue_ok ? "bluestein" : "direct";
  };
  std::string s = "{\"L\": " + std::to_string(L) + ", \"n_w\": " + std::to_string(n_w) + ", \"plan_cost\": " + std::to_string(cost) +
                  ", \"engine_fwd\": \"" + engine(true) + "\", \"engine_inv\": \"" + engine(false) + "\", \"plan_fwd\": \"" + tq_pipe_plan_text(L, +1) +
                  "\", \"plan_inv\": \"" + tq_pipe_plan_text(L, -1) + "\", \"bluestein_M\": " +
                  (blue_ok ? "[" + std::to_string(M1) + ", " + std::to_string(M2) + "]" : std::string("null")) + "}";
  snprintf(buf, len, "%s", s.c_str());
  return TQ_OK;
}

// -------------------------------------------------------------------------------------------------
// density(G(i w))   c++/triqs/gfs/functions/density.cpp:32-120        (the consumer of G_loc in a DMFT loop)
//   n = -xi [ (1/beta) sum_n (G(i w_n) - sum_i a_i / (i w_n - b_i)) + m1 + sum_i F(a_i, b_i) ],  F = xi a / (-xi + e^{-beta b})
// One streaming pass over G(i w): k_density_partial sums row chunks (the per-row factors 1/(i w - b_i) are computed once per
// row into shared memory and shared by all columns), k_density_final adds the chunks in fixed order (deterministic),
// applies the closed-form tail terms and mirrors the upper triangle (:115).  Bound: HBM read, 16 B per element.
// -------------------------------------------------------------------------------------------------
struct DensityPoles {
  double b[3], xi;
};
TQ_HD void density_amplitudes(int stat, cd m1, cd m2, cd m3, cd a[3]) {
  if (stat == TQ_FERMION) { // :98-101
    a[0] = csub(m1, m3);
    a[1] = cscale(0.5, cadd(m2, m3));
    a[2] = cscale(0.5, csub(m3, m2));
  } else { // :102-105
    a[0] = cadd(csub(cscale(1.0 / 6.0, m1), cscale(0.5, m2)), cscale(1.0 / 3.0, m3));
    a[1] = cadd(cadd(cscale(-0.5, m1), cscale(0.5, m2)), m3);
    a[2] = csub(cscale(4.0 / 3.0, m1), cscale(4.0 / 3.0, m3));
  }
}
constexpr int DENS_ROWS = 512; // rows per chunk = threads per CTA

__global__ void __launch_bounds__(DENS_ROWS) k_density_partial(const cd *__restrict__ g, long g_gf_stride, int n_w, int n_cols,
                                                               const cd *__restrict__ mom, long mom_gf_stride, double beta, int stat, int first,
                                                               DensityPoles P, cd *__restrict__ partial /* [batch][chunks][n_cols] */) {
  extern __shared__ double2 dens_smem[];
  cd(*kap)[3] = reinterpret_cast<cd(*)[3]>(dens_smem); // [DENS_ROWS][3]
  double2 *red = dens_smem + 3 * DENS_ROWS;            // [row lanes][n_cols]
  const int gf = blockIdx.y, chunk = blockIdx.x;
  const int r0 = chunk * DENS_ROWS;
  const int nr = min(DENS_ROWS, n_w - r0);
  {
    const int r = threadIdx.x;
    if (r < nr) {
      const double w = M_PI * (double)(2 * (long)(first + r0 + r) + stat) / beta; // matsubara.hpp:55
#pragma unroll
      for (int i = 0; i < 3; ++i) { // 1 / (i w - b) = (-b - i w) / (b^2 + w^2)
        const double inv = 1.0 / (P.b[i] * P.b[i] + w * w);
        kap[r][i] = cmk(-P.b[i] * inv, -w * inv);
      }
    }
  }
  __syncthreads();
  const int cl = min(n_cols, DENS_ROWS);        // column lanes
  const int rl = DENS_ROWS / cl;                // row lanes
  const int c_lane = threadIdx.x % cl, r_lane = threadIdx.x / cl;
  const cd *gg = g + (size_t)gf * g_gf_stride + (size_t)r0 * n_cols;
  const cd *mm = mom + (size_t)gf * mom_gf_stride;
  for (int c = c_lane; c < n_cols; c += cl) {
    cd acc = cmk(0.0, 0.0);
    if (r_lane < rl) {
      cd a[3];
      density_amplitudes(stat, mm[(size_t)1 * n_cols + c], mm[(size_t)2 * n_cols + c], mm[(size_t)3 * n_cols + c], a);
      // four independent loads / partial sums in flight per thread (the loop is a pure stream over G(i w))
      cd acc4[4] = {cmk(0.0, 0.0), cmk(0.0, 0.0), cmk(0.0, 0.0), cmk(0.0, 0.0)};
      int r = r_lane;
      for (; r + 3 * rl < nr; r += 4 * rl) {
        cd v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = gg[(size_t)(r + u * rl) * n_cols + c];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = r + u * rl;
          cd t = csub(v[u], cmul(a[0], kap[rr][0]));
          t = csub(t, cmul(a[1], kap[rr][1]));
          t = csub(t, cmul(a[2], kap[rr][2]));
          acc4[u] = cadd(acc4[u], t);
        }
      }
      for (; r < nr; r += rl) {
        cd t = csub(gg[(size_t)r * n_cols + c], cmul(a[0], kap[r][0]));
        t = csub(t, cmul(a[1], kap[r][1]));
        t = csub(t, cmul(a[2], kap[r][2]));
        acc4[0] = cadd(acc4[0], t);
      }
      acc = cadd(cadd(acc4[0], acc4[1]), cadd(acc4[2], acc4[3]));
      red[r_lane * n_cols + c] = acc;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n_cols; c += DENS_ROWS) {
    cd s = cmk(0.0, 0.0);
    for (int l = 0; l < rl; ++l) s = cadd(s, red[l * n_cols + c]);
    partial[((size_t)gf * gridDim.x + chunk) * n_cols + c] = s;
  }
}

__global__ void k_density_final(const cd *__restrict__ partial, int n_chunks, int n, const cd *__restrict__ mom, long mom_gf_stride, double beta,
                                int stat, DensityPoles P, long total /* batch * n * n */, cd *__restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int nn = n * n;
  const long gf = i / nn;
  const int c = (int)(i - gf * nn);
  const int n1 = c / n, n2 = c - n1 * n;
  if (n2 < n1) return; // :89 only the upper triangle is summed
  cd r = cmk(0.0, 0.0);
  for (int k = 0; k < n_chunks; ++k) r = cadd(r, partial[((size_t)gf * n_chunks + k) * nn + c]);
  const cd *mm = mom + (size_t)gf * mom_gf_stride;
  const cd m1 = mm[(size_t)1 * nn + c];
  cd a[3];
  density_amplitudes(stat, m1, mm[(size_t)2 * nn + c], mm[(size_t)3 * nn + c], a);
  cd res = cadd(cscale(1.0 / beta, r), m1);
#pragma unroll
  for (int k = 0; k < 3; ++k) res = cadd(res, cscale(P.xi / (-P.xi + exp(-beta * P.b[k])), a[k])); // F(a, b)  :84
  res = cscale(-P.xi, res); // :112
  out[(size_t)gf * nn + n1 * n + n2] = res;
  if (n2 > n1) out[(size_t)gf * nn + n2 * n + n1] = cconj(res); // :115
}

extern "C" tq_status tq_density(tq_ctx *ctx, double beta, int statistic, long n_iw, int n, long batch, const tq_complex *gw,
                                const tq_complex *moments, int n_moments, tq_complex *out, int *warn, double *tail_err) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!gw || !out || n_iw < 1 || n < 1 || batch < 1 || !(beta > 0) || (statistic != TQ_FERMION && statistic != TQ_BOSON))
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_density: invalid argument");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long n_cols = (long)n * n;
  const long last = n_iw - 1, first = -(last + (statistic == TQ_FERMION ? 1 : 0));
  const long n_w = last - first + 1;
  const size_t in_bytes = sizeof(cd) * (size_t)batch * n_w * n_cols, out_bytes = sizeof(cd) * (size_t)batch * n_cols;
  const void *d_in = nullptr, *d_known = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, gw, in_bytes, ctx->stage_in, &d_in));
  const bool have_moments = moments != nullptr && n_moments > 0;
  if (have_moments) TQ_TRY(tq_stage_in(ctx, moments, sizeof(cd) * (size_t)batch * n_moments * n_cols, ctx->stage_in2, &d_known));
  const int n_chunks = (int)((n_w + DENS_ROWS - 1) / DENS_ROWS);
  // small buffer: moments [batch][TQ_MAX_MOMENTS][n_cols], err double[batch], status, zero known moment, partial sums
  const size_t mom_bytes = sizeof(cd) * (size_t)batch * TQ_MAX_MOMENTS * n_cols;
  const size_t err_off = (mom_bytes + 255) & ~size_t(255);
  const size_t st_off = (err_off + sizeof(double) * batch + 255) & ~size_t(255);
  const size_t zk_off = (st_off + sizeof(StatusWords) + 255) & ~size_t(255);
  const size_t part_off = (zk_off + sizeof(cd) * (size_t)batch * n_cols + 255) & ~size_t(255);
  TQ_TRY(tq_reserve(ctx, ctx->small, part_off + sizeof(cd) * (size_t)batch * n_chunks * n_cols));
  cd *d_mom = (cd *)ctx->small.p;
  double *d_err = (double *)((char *)ctx->small.p + err_off);
  StatusWords *d_st = (StatusWords *)((char *)ctx->small.p + st_off);
  cd *d_zero_known = (cd *)((char *)ctx->small.p + zk_off);
  cd *d_part = (cd *)((char *)ctx->small.p + part_off);
  TQ_CUDA(ctx, cudaMemsetAsync(d_err, 0, part_off - err_off, ctx->stream));
  int mom_rows = 0;
  bool fitted = false;
  const long total = batch * n_cols;
  if (have_moments && n_moments >= 4) {
    k_prepare_known_moments<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_known, n_moments, (int)n_cols, total, d_mom,
                                                                                        &d_st->max_m0_bits);
    tq_count_launch(ctx);
    mom_rows = 4;
  } else {
    // :40 no moments -> make_zero_tail(g, 1);  :47 fewer than 4 -> fit_tail(g, known_moments)
    const cd *kn = have_moments ? (const cd *)d_known : d_zero_known;
    const int nk = have_moments ? n_moments : 1;
    if (have_moments) { // the 0th-moment assertion (:42-44) is made on the user's moments before the fit
      TQ_TRY(tq_reserve(ctx, ctx->stage_in3, sizeof(cd) * (size_t)batch * 4 * n_cols));
      k_prepare_known_moments<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const cd *)d_known, n_moments, (int)n_cols, total,
                                                                                          (cd *)ctx->stage_in3.p, &d_st->max_m0_bits);
      tq_count_launch(ctx);
    }
    tq_status st = tq_fit_tail_device(ctx, beta, statistic, n_iw, 0.2, 30, -1, 0, 1, n_cols, batch, (const cd *)d_in, kn, nk, d_mom, &mom_rows, d_err);
    if (st != TQ_OK) return st;
    fitted = true;
    if (!(mom_rows > 3)) // :57
      return tq_fail(ctx, TQ_ERR_RUNTIME, "ERROR: Density implementation requires at least a proper 3rd high-frequency moment\n");
  }
  TQ_TRY(tq_stage_out_begin(ctx, out, out_bytes, ctx->stage_out, &d_out, &out_host));
  DensityPoles P;
  if (statistic == TQ_FERMION) { // :70-74
    P.xi = -1.0; P.b[0] = 0.0; P.b[1] = 1.0; P.b[2] = -1.0;
  } else { // :75-79
    P.xi = 1.0; P.b[0] = -1.0; P.b[1] = 1.0; P.b[2] = 0.5;
  }
  {
    const int cl = (int)std::min<long>(n_cols, DENS_ROWS), rl = DENS_ROWS / cl;
    const size_t smem = sizeof(cd) * ((size_t)rl * n_cols + 3 * DENS_ROWS);
    if (smem > (size_t)ctx->max_smem_optin) return tq_fail(ctx, TQ_ERR_UNSUPPORTED, "tq_density: target too large for one SM's shared memory");
    TQ_CUDA(ctx, cudaFuncSetAttribute(k_density_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (batch > 65535) return tq_fail(ctx, TQ_ERR_INVALID, "tq_density: batch > 65535");
    KernelScope ks(ctx, "density_partial");
    k_density_partial<<<dim3(n_chunks, (unsigned)batch), DENS_ROWS, smem, ctx->stream>>>((const cd *)d_in, n_w * n_cols, (int)n_w, (int)n_cols, d_mom,
                                                                                         (long)mom_rows * n_cols, beta, statistic, (int)first, P, d_part);
    tq_count_launch(ctx);
    TQ_CUDA(ctx, cudaGetLastError());
  }
  {
    KernelScope ks(ctx, "density_final");
    k_density_final<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(d_part, n_chunks, n, d_mom, (long)mom_rows * n_cols, beta, statistic, P,
                                                                             total, (cd *)d_out);
    tq_count_launch(ctx);
    TQ_CUDA(ctx, cudaGetLastError());
  }
  TQ_TRY(tq_reserve_pinned(ctx, sizeof(double) * batch + sizeof(StatusWords)));
  double *h_err = (double *)ctx->pinned;
  StatusWords *h_st = (StatusWords *)((char *)ctx->pinned + sizeof(double) * batch);
  TQ_CUDA(ctx, cudaMemcpyAsync(h_err, d_err, sizeof(double) * batch, cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaMemcpyAsync(h_st, d_st, sizeof(StatusWords), cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (have_moments) {
    double a0;
    memcpy(&a0, &h_st->max_m0_bits, sizeof(double));
    if (!(a0 < 1e-8)) // :42-44
      return tq_fail(ctx, TQ_ERR_RUNTIME, "ERROR: Density implementation requires vanishing 0th moment\n  error is :" + std::to_string(a0) + "\n");
  }
  for (long b = 0; b < batch; ++b) {
    const double e = fitted ? h_err[b] : 0.0;
    if (fitted && !(e < 1e-2)) // :48-50
      return tq_fail(ctx, TQ_ERR_RUNTIME,
                     "ERROR: High frequency moments have an error greater than 1e-2.\n  Error = " + std::to_string(e) +
                        "\n Please make sure you treat the constant offset analytically!\n");
    if (warn) warn[b] = (fitted && e > 1e-4) ? TQ_WARN_TAIL_ERR : 0; // :51-53
    if (tail_err) tail_err[b] = e;
  }
  TQ_TRY(tq_stage_out_finish(ctx, out, out_bytes, d_out, out_host));
  return TQ_OK;
}
