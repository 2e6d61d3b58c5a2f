// Pipelined four-step imtime <-> imfreq transforms for long time meshes (L = Na * Nb, e.g. 10^5 = 400 * 250).
//
// Same mathematics as the generic tile kernels of fourier_matsubara.cu (which restate
// c++/triqs/gfs/transform/fourier_matsubara.cpp:96-186 and :190-271), different machine mapping:
//   * one persistent, warp-specialised CTA per SM runs a THREE-STAGE pipeline over a ring of three shared-memory tile
//     buffers:   producer warp (TMA: cp.async.bulk[.tensor] + mbarrier complete_tx)  ->  A-group warps (pass A of the
//     tile FFT, in place, pre-transform fused)  ->  B-group warps (pass B, epilogue fused, results stored straight to
//     global memory).  The three stages work on three different tiles at the same time, hand-offs are mbarriers
//     (full -> adone -> free), so the FP64 pipe sees a mix of load-, math- and store-phases instead of lock-step phases;
//   * the tile FFT is the compile-time two-pass engine of tq_fft2.cuh (250 = 10 x 25, 400 = 20 x 20);
//   * every per-row factor is FACTORISED so that no per-row table is read in the inner loops:
//       row = q + RB r (pass A legs)  /  alpha = s + RA t (pass B legs),  j = item + S row:
//         e^{-b delta j} = [per item, folded into the pole weights] x [per q or s: tiny static table] x [per leg: kernel
//         parameter = constant-bank operand];  the e^{+-i pi j / L} twist likewise, its per-q / per-s part folded into
//         the pass-A twiddle rows;
//   * everything that depends only on the tile (four-step twiddles x per-item phase, window-row factors) is precomputed
//     once per plan in global memory (L2 resident) and pulled in by the same TMA transaction as the tile;
//   * the scratch between the two kernels is laid out [beta][column tile][a][column] so that a pass-2 tile is
//     ONE contiguous bulk copy, and the batch is chunked so that it stays in the 126 MB L2.
//
//   kernel F1: G(tau) rows a + Na b  -> FFT_Nb over b -> x w_L^{a beta} (beta/L) e^{i pi a/L}  -> scratch
//   kernel F2: scratch (fixed beta)  -> FFT_Na over a -> + trapezoid correction + tail model   -> G(i w_n), window rows only
//   kernel I1: G(i w) rows a + Na b (window rows only, others are zero) - tail, / beta -> FFT_Nb -> x w_L^{-a beta} e^{-i pi beta/L} -> scratch
//   kernel I2: scratch -> FFT_Na -> x e^{-i pi Nb alpha/L} + tail model in tau -> G(tau)
#pragma once
#include "tq_pipe.cuh"

#include <type_traits>

#include <algorithm>
#include <cmath>
#include <cstring>

// -------------------------------------------------------------------------------------------------
enum { M_F1 = 0, M_F2 = 1, M_I1 = 2, M_I2 = 3 };
// Tile buffers of the ring.  Measured (32 blocks, one B200): the pass-1 kernels gain from a 4-deep ring of 9-column tiles
// (F1 0.559 -> 0.530 ms, I1 0.555 -> 0.482 ms: the TMA row gather of the next-but-one tile has a full tile period more to land),
// the pass-2 kernels lose with the narrower tiles a fourth buffer would force (F2 0.463 -> 0.486 ms, I2 0.480 -> 0.839 ms).
#ifndef TQ_PIPE_NBUF1
#define TQ_PIPE_NBUF1 4
#endif
#ifndef TQ_PIPE_NBUF2
#define TQ_PIPE_NBUF2 3
#endif
__host__ __device__ constexpr int pipe_nbuf(int mode) { return (mode == 0 || mode == 2) ? TQ_PIPE_NBUF1 : TQ_PIPE_NBUF2; } // M_F1, M_I1 : M_F2, M_I2
__host__ __device__ constexpr int pipe_ntab(int mode) { return pipe_nbuf(mode) + 1; } // table sets live until the end of pass B, one stage longer than the tile
constexpr int PIPE_MAXR = 41; // largest register radix of a pass
// worker warps per CTA (+ 1 producer warp).  The register file is 16 K registers per SM sub-partition: 12 warps (3 per
// sub-partition) get 168 registers per thread, 16 warps only 128 (radix 20 / 25 then spill); measured: 7 workers reach 86 %
// of the throughput of 11, 15 workers at 128 registers are slower.
constexpr int PIPE_NW = 11;
// worker warps of the "heavy" instantiations (Rader radices 11, 13, 17, 41 and radix 32): see fourier_pipe_heavy*.cu
#ifndef TQ_NW_HEAVY
#define TQ_NW_HEAVY PIPE_NW
#endif
// per-kernel worker counts (developer A/B switches; > 11 workers means 128 registers per thread)
#ifndef TQ_NW_F1
#define TQ_NW_F1 PIPE_NW
#endif
#ifndef TQ_NW_F2
#define TQ_NW_F2 PIPE_NW
#endif
#ifndef TQ_NW_I1
#define TQ_NW_I1 PIPE_NW
#endif
#ifndef TQ_NW_I2
#define TQ_NW_I2 PIPE_NW
#endif

struct PipeArgs {
  int Na, Nb, n_cols;
  int tcw, n_ct;   // column tiling of this pass
  int tcw2;        // column tiling of pass 2 (defines the scratch layout)
  int n_items;     // Na (pass 1) or Nb (pass 2)
  int wmax;        // row pitch of the compact window-row tables (F2, I1)
  int pitch;       // shared-memory row pitch of a pass-1 tile in columns (>= tcw; pass-2 tiles are dense)
  int i1_rect;     // I1: the window rows are two rectangles of the [b][a] view (loaded as two tensor tiles)
  int i1_edge;     // I1: ... and they are exactly the first and the last leg of every pass-A butterfly (pruned pass A)
  int keep_guard;  // developer A/B switch (TQ_PIPE_GUARD=1): always run the ring-phase guard of the worker loop
  long n_tiles;    // batch * n_items * n_ct
  const cd *in;
  cd *out;
  cd *scratch;
  const cd *polew;   // [batch][n_cols][4] pole weights of this pass (k_pole_weights)
  long in_gf_stride, out_gf_stride, scratch_gf_stride;
  int L, first, last, stat;
  double half_fact_fwd; // 0.5 beta / L
  double inv_beta;
  const cd *twq;          // [RB][RA] pass-A twiddle rows w_N^{SGN q s} x per-q (F1) / per-s (I2) phase
  const double4 *qtab;    // F1: [RB] {E1q, E2q, E3q, -};  I2: [RA] {E1s, E2s, E3s, -}
  const double4 *itemtab; // [n_items] {C1 ea1, C2 ea2, C3 ea3, -}   (F1: item = a, I2: item = beta)
  const cd *Tfull;        // [Na][Nb] four-step twiddle x per-item phase (pass 1)
  const int4 *wintab;     // [n_items] {r1, r2, -, -} window rows of every item (F2, I1)
  const double4 *wperm;   // [n_items][wmax] {b1 r1, w r1, b2 r2, w r2} of the window rows of every item (F2, I1)
  // per-leg constants (constant-bank operands): F1 legs r of pass A, I2 legs t of pass B
  double er[3][PIPE_MAXR];
  cd phr[PIPE_MAXR];
  cd wedge[PIPE_MAXR]; // I1 pruned pass A: exp(-SGN 2 pi i s / RA)
  long long *dbg; // developer phase timing (TQ_PHASE_DEBUG) or nullptr
  // planned (run-time radix) kernels: tile length and radix pair of THIS pass; rb_generic: pass B by direct summation
  int tN, tRA, tRB, rb_generic;
  // "gfs as columns" (planned kernels, narrow targets, tr = target columns of a gf = 1..3, 0 = off): the batch is ONE virtual gf whose
  // columns are all (gf, column) pairs, column cv = gf * tr + c.  Tiles are gathered by tensor maps whose innermost box is the tr
  // contiguous columns of a gf (16 tr bytes), second box dimension = tcw / tr gfs; results go back with the gf stride
  // col_stride_out between groups of tr columns.  Tiles are numbered item-fastest so that the CTAs of a wave touch neighbouring rows.
  int tr;
  long col_stride_out;
  int lag0; // packet order of the planned kernels: 1 = A(k) B(k) | A(k+1) ..., 0 = pass B one tile behind pass A
  int f1_boxes, f1_box_rows; // F1: the tile is fetched as f1_boxes tensor boxes of f1_box_rows rows (a TMA box has <= 256 rows)
  const cd *gtw;             // [tRB] exp(+2 pi i j / tRB)   (generic pass B)
  const double4 *legtab;     // I2, generic pass B: [tRB] per-leg {E1, E2, E3, -}
  const cd *legph;           // I2, generic pass B: [tRB] per-leg phase
};

// Rows r of a tile (global frequency index k = item + S r) that fall inside the Matsubara window
// [first, last] (fourier_matsubara.cpp:182, :254 "(iw.n + L) % L"):  r in [0, r1)  (n = k >= 0)  and  r in [r2, N)  (n = k - L < 0).
struct Window {
  int r1, r2;
};
TQ_HD Window tile_window(int item, int S, int N, int L, int first, int last) {
  Window w;
  w.r1 = last >= item ? ((last - item) / S + 1 < N ? (last - item) / S + 1 : N) : 0;
  const int lo = L + first - item; // k >= L + first
  w.r2 = lo <= 0 ? 0 : ((lo + S - 1) / S < N ? (lo + S - 1) / S : N);
  if (w.r2 < w.r1) w.r2 = w.r1;
  return w;
}

// "gfs as columns": element offset of tile column c = g * nc + cc (nc = 1..3 target columns per gf, gf stride cs); cs = 0: plain columns
TQ_D long tr_col_offset(int c, int nc, long cs) {
  if (cs == 0) return c;
  const int g = nc == 1 ? c : nc == 2 ? (c >> 1) : (c * 43) >> 7; // c / 3 for c < 128
  return (long)g * cs + (c - g * nc);
}

// ---- functors plugged into the two-pass engine -----------------------------------------------
struct PreF1 { // x = phase (g - sum_i at_i E_i[row]),  row = q + RB r        fourier_matsubara.cpp:159-173
  const PipeArgs &A;
  const cd *coef;      // [4][tcw] = a1, a2, a3, m1
  const double4 *qtab; // per-q factors
  double4 ea;          // per-item factors (C_i e_i(item))
  cd a1, a2, a3;
  TQ_D void begin(int q, int c) {
    const double4 e = qtab[q];
    a1 = cscale(e.x * ea.x, coef[c]);
    a2 = cscale(e.y * ea.y, coef[c + A.tcw]);
    a3 = cscale(e.z * ea.z, coef[c + 2 * A.tcw]);
  }
  template <int R> TQ_D cd load(const cd *p) const {
    if (TQ_KNOCK(256)) return *p;
    cd v = caxpy(-A.er[0][R], a1, *p);
    v = caxpy(-A.er[1][R], a2, v);
    v = caxpy(-A.er[2][R], a3, v);
    return R == 0 ? v : cmul(v, A.phr[R]); // the per-q part of the twist sits in the twiddle row
  }
};
template <int RB> struct PreI1 { // x = (g - tail(i w)) / beta for window rows, 0 elsewhere      :254
  const cd *wrec, *coef; // coef [4][tcw] = a1, a2 - a3, a2 + a3, -
  int tcw;
  Window win;
  double inv_beta;
  int rb_rt = 0; // RB == 0: run-time radix
  cd a1, D23, A23;
  int q;
  int nlo, nhi; // legs R < nlo are window rows [0, r1), legs R >= nhi window rows [r2, N); the legs between are zero
  TQ_D int rb() const { return RB ? RB : rb_rt; }
  TQ_D void begin(int qq, int c) {
    q = qq;
    a1 = coef[c];
    D23 = coef[c + tcw];
    A23 = coef[c + 2 * tcw];
    // row = q + rb R < r1  <=>  R < ceil((r1 - q) / rb);   row >= r2  <=>  R >= ceil((r2 - q) / rb)      (one division pair per butterfly
    // instead of two compares and an index computation per leg: most legs are zero rows)
    const int d = rb();
    nlo = win.r1 > q ? (win.r1 - q + d - 1) / d : 0;
    nhi = win.r2 > q ? (win.r2 - q + d - 1) / d : 0;
  }
  // pruned pass A: leg 0 is window row q (compact index q), the last leg is window row RB + q
  TQ_D cd edge(int which, const cd *p) const {
    const int idx = which ? rb() + q : q;
    const cd t = tail_iw4(wrec[2 * idx], wrec[2 * idx + 1], a1, D23, A23);
    return cscale(inv_beta, csub(*p, t));
  }
  template <int R> TQ_D cd load(const cd *p) const {
    int idx;
    if (R < nlo)
      idx = q + rb() * R;
    else if (R >= nhi)
      idx = win.r1 + q + rb() * R - win.r2;
    else
      return cmk(0.0, 0.0);
    const cd t = tail_iw4(wrec[2 * idx], wrec[2 * idx + 1], a1, D23, A23);
    return cscale(inv_beta, csub(*p, t));
  }
};
template <int RA> struct PostScratch { // Y = z T[beta]  -> scratch[beta][column tile][a][column]
  const cd *T;
  cd *base; // scratch of this gf
  long beta_stride; // Na * n_cols
  int n_cols, Na, tcw2, item, c0;
  int ra_rt = 0; // RA == 0: run-time radix
  cd *p;
  const cd *Ts;
  TQ_D int ra() const { return RA ? RA : ra_rt; }
  TQ_D void begin(int sidx, int c) {
    const int cg = c0 + c;
    const int h = cg / tcw2;
    const int cc = cg - h * tcw2;
    const int w = min(tcw2, n_cols - h * tcw2);
    p = base + ((long)h * tcw2 * Na + (long)item * w + cc) + (long)sidx * beta_stride;
    Ts = T + sidx;
  }
  template <int TT> TQ_D void store(cd v) const { p[(long)(ra() * TT) * beta_stride] = TQ_KNOCK(128) ? v : cmul(v, Ts[ra() * TT]); }
  TQ_D void store_rt(int t, cd v) const { p[(long)(ra() * t) * beta_stride] = cmul(v, Ts[ra() * t]); }
};
template <int RA> struct PostF2 { // G(i w_n) = z + corr + tail, window rows only      :181-183
  const cd *wrec, *coef; // coef [4][tcw] = a1, a2 - a3, a2 + a3, trapezoid correction
  int tcw;
  int n_cols, item, Nb, L, first;
  Window win;
  cd *base; // out of this gf + c0
  int ra_rt = 0; // RA == 0: run-time radix
  long cs = 0;   // "gfs as columns": stride between gfs of the output (0: the columns of the tile are contiguous)
  int nc = 1;    // ... and target columns per gf (`n_cols` is the row stride)
  cd a1, D23, A23, ex;
  cd *p;
  int sidx;
  TQ_D int ra() const { return RA ? RA : ra_rt; }
  TQ_D void begin(int s, int c) {
    sidx = s;
    a1 = coef[c];
    D23 = coef[c + tcw];
    A23 = coef[c + 2 * tcw];
    ex = coef[c + 3 * tcw];
    p = base + tr_col_offset(c, nc, cs);
  }
  template <int TT> TQ_D void store(cd v) const { store_rt(TT, v); }
  TQ_D void store_rt(int TT, cd v) const {
    const int k = sidx + ra() * TT;
    int idx, di;
    if (k < win.r1) {
      idx = k;
      di = item + Nb * k - first;
    } else if (k >= win.r2) {
      idx = win.r1 + k - win.r2;
      di = item + Nb * k - L - first;
    } else {
      return;
    }
    const cd t = tail_iw4(wrec[2 * idx], wrec[2 * idx + 1], a1, D23, A23);
    p[(long)di * n_cols] = cadd(cadd(v, ex), t);
  }
};
template <int RA> struct PostI2 { // G(tau_j) = z e^{-i pi j/L} + sum_i at_i E_i,  j = item + Nb (s + RA t)        :261-268
  const PipeArgs &A;
  const cd *coef;      // [4][tcw] = a1, a2, a3, m1
  const double4 *qtab; // per-s factors
  double4 ea;          // per-item factors
  int n_cols, item, Nb, L;
  bool fermion;
  cd *base; // out of this gf + c0
  int ra_rt = 0; // RA == 0: run-time radix
  long cs = 0;   // "gfs as columns": stride between gfs of the output (0: the columns of the tile are contiguous)
  int nc = 1;    // ... and target columns per gf (`n_cols` is the row stride)
  cd a1, a2, a3, m1;
  cd *p;
  int j0;
  TQ_D int ra() const { return RA ? RA : ra_rt; }
  TQ_D void begin(int s, int c) {
    const double4 e = qtab[s];
    a1 = cscale(e.x * ea.x, coef[c]);
    a2 = cscale(e.y * ea.y, coef[c + A.tcw]);
    a3 = cscale(e.z * ea.z, coef[c + 2 * A.tcw]);
    m1 = coef[c + 3 * A.tcw];
    j0 = item + Nb * s;
    p = base + tr_col_offset(c, nc, cs) + (long)j0 * n_cols;
  }
  template <int TT> TQ_D void store(cd z) const {
    cd v = TT == 0 ? z : cmul(z, A.phr[TT]); // the per-s part of the twist was folded into the pass-A twiddle row
    v = caxpy(A.er[0][TT], a1, v);
    v = caxpy(A.er[1][TT], a2, v);
    v = caxpy(A.er[2][TT], a3, v);
    p[(long)(Nb * ra() * TT) * n_cols] = v;
    if (TT == 0 && j0 == 0) { // gt[L] = -/+ (gt[0] + m1)     :267-268
      const cd t = cadd(v, m1);
      p[(long)L * n_cols] = fermion ? cneg(t) : t;
    }
  }
  // generic pass B: the per-leg factors come from a table instead of the constant bank
  TQ_D void store_rt(int t, cd z) const {
    const double2 *lt = reinterpret_cast<const double2 *>(A.legtab + t);
    const double2 e01 = __ldg(lt), e23 = __ldg(lt + 1);
    const double4 e = make_double4(e01.x, e01.y, e23.x, e23.y);
    cd v = t == 0 ? z : cmul(z, __ldg(A.legph + t));
    v = caxpy(e.x, a1, v);
    v = caxpy(e.y, a2, v);
    v = caxpy(e.z, a3, v);
    p[(long)(Nb * ra() * t) * n_cols] = v;
    if (t == 0 && j0 == 0) {
      const cd tt = cadd(v, m1);
      p[(long)L * n_cols] = fermion ? cneg(tt) : tt;
    }
  }
};

// Pole weights of the 3-pole tail model per (gf, column), in the two forms the kernels use   (:151-169, :238-252)
//   w_tau  [gf][4][c] = a1, a2, a3, m1                                   (tau side: F1 pre-transform, I2 post-transform)
//   w_freq [gf][4][c] = a1, a2 - a3, a2 + a3, trapezoid correction        (frequency side: F2 epilogue, I1 pre-transform)
// The trapezoid end-point correction -0.5 (beta/L) (g_0 + m1 +/- g_L) (:181) exists on the forward leg only (gt != nullptr).
static __global__ void k_pole_weights(const cd *mom, long mom_gf_stride, int n_cols, long total, int stat, const cd *gt, long gt_gf_stride, int L,
                               double half_fact_fwd, cd *w_tau, cd *w_freq, int tr) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long gf = i / n_cols;
  const int c = (int)(i - gf * n_cols);
  const cd *m = mom + gf * mom_gf_stride + c;
  const cd m1 = m[(size_t)1 * n_cols], m2 = m[(size_t)2 * n_cols], m3 = m[(size_t)3 * n_cols];
  const bool fermion = stat == TQ_FERMION;
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
  cd ex = cmk(0.0, 0.0);
  if (gt) {
    const cd *g = gt + gf * gt_gf_stride + c;
    const cd g0 = g[0], gL = g[(size_t)L * n_cols];
    cd s = cadd(g0, m1);
    s = fermion ? cadd(s, gL) : csub(s, gL);
    ex = cscale(-half_fact_fwd, s);
  }
  // [gf][4][n_cols]: the four weights of a tile's columns are four contiguous runs, so that lanes with consecutive columns read
  // consecutive 16-byte words of shared memory (the [c][4] layout cost a 4-way bank conflict on every read: 10-18 % of the
  // shared-memory wavefronts of the pipelined kernels, ncu source view)
  // ("gfs as columns": ONE virtual gf whose columns are all (gf, column) pairs -> [4][total])
  const size_t ws = tr ? (size_t)total : (size_t)n_cols;
  cd *wt = tr ? w_tau + i : w_tau + (size_t)gf * 4 * n_cols + c, *wf = tr ? w_freq + i : w_freq + (size_t)gf * 4 * n_cols + c;
  wt[0] = a1;
  wt[ws] = a2;
  wt[2 * ws] = a3;
  wt[3 * ws] = m1;
  wf[0] = a1;
  wf[ws] = csub(a2, a3);
  wf[2 * ws] = cadd(a2, a3);
  wf[3 * ws] = ex;
}

// shared-memory carve-up (bytes); tile buffers first, each padded to 128 B (TMA tensor destination alignment)
template <int MODE> struct PipeSmem {
  static constexpr bool PASS1 = (MODE == M_F1 || MODE == M_I1);
  static constexpr bool QTAB = (MODE == M_F1 || MODE == M_I2);
  static constexpr bool WREC = (MODE == M_F2 || MODE == M_I1);
  int nq;
  size_t buf_stride, stage_off, tw, qtab, gtw, tab0, tab_stride, coef_off, item_off, t_off, win_off, wrec_off, mbar, total;
  __host__ __device__ PipeSmem(int N, int RA, int RB, int tcw, int pitch, int wmax, int rb_generic = 0) {
    nq = (MODE == M_F1) ? RB : RA;
    stage_off = (sizeof(cd) * (size_t)N * (PASS1 ? pitch : tcw) + 127) & ~size_t(127);
    // I1: staging rows for the last butterfly leg (pruned pass A), 128-byte aligned like the tile itself
    buf_stride = stage_off + (MODE == M_I1 ? ((sizeof(cd) * (size_t)RB * pitch + 127) & ~size_t(127)) : 0);
    tw = pipe_nbuf(MODE) * buf_stride;
    qtab = tw + sizeof(cd) * N;
    gtw = qtab + (QTAB ? sizeof(double4) * nq : 0);
    tab0 = gtw + (rb_generic ? sizeof(cd) * RB : 0);
    coef_off = 0;
    item_off = coef_off + sizeof(cd) * 4 * (size_t)tcw;
    t_off = item_off + (QTAB ? sizeof(double4) : 0);
    win_off = t_off + (PASS1 ? sizeof(cd) * N : 0);
    wrec_off = win_off + (WREC ? sizeof(int4) : 0);
    tab_stride = wrec_off + (WREC ? sizeof(cd) * 2 * (size_t)wmax : 0);
    mbar = tab0 + pipe_ntab(MODE) * tab_stride;
    total = mbar + 8 * (3 * pipe_nbuf(MODE) + pipe_ntab(MODE)) + 16;
  }
};

struct TileId {
  long gf;
  int item, c0, w;
};
// tile index -> (gf, item, column tile) without integer division: the divisors are fixed for a launch
struct TileDecoder {
  unsigned long long m_per_gf; // ceil(2^64 / (n_items n_ct)): exact quotient for t < 2^32
  unsigned m_nct;
  int per_gf, n_ct, tcw, n_cols;
  // item_fast ("gfs as columns": one virtual gf): t = item + n_items * column tile
  TQ_D TileDecoder(int n_items, int n_ct_, int tcw_, int n_cols_, bool item_fast) : n_ct(n_ct_), tcw(tcw_), n_cols(n_cols_) {
    per_gf = item_fast ? -n_items : n_items * n_ct;
    const int d = item_fast ? n_items : per_gf;
    m_per_gf = d > 1 ? ~0ull / (unsigned long long)d + 1ull : 0ull;
    m_nct = fastdiv_magic((unsigned)n_ct);
  }
  TQ_D TileId operator()(long t) const {
    TileId id;
    if (per_gf < 0) {
      const int n_items = -per_gf;
      const int ct = n_items > 1 ? (int)__umul64hi((unsigned long long)t, m_per_gf) : (int)t;
      id.gf = 0;
      id.item = (int)(t - (long)ct * n_items);
      id.c0 = ct * tcw;
      id.w = min(tcw, n_cols - id.c0);
      return id;
    }
    id.gf = per_gf > 1 ? (long)__umul64hi((unsigned long long)t, m_per_gf) : t;
    const int rem = (int)(t - id.gf * per_gf);
    id.item = fastdiv(rem, m_nct);
    id.c0 = (rem - id.item * n_ct) * tcw;
    id.w = min(tcw, n_cols - id.c0);
    return id;
  }
};

// registers per thread: the 16 K registers of a sub-partition over its warps (12 warps: 168; 8 warps: 255)
__host__ __device__ constexpr int pipe_maxnreg(int nw) {
  const int r = (16384 / (((nw + 1) + 3) / 4 * 32)) & ~7;
  return r > 255 ? 255 : r;
}
// "dual" packets (compile-time kernels only, bit per mode): 64 butterflies per packet, two per lane (tq_fft2.cuh fft2_pass_*_dual) -- meant
// for 7 worker warps with 255 registers (-DTQ_NW_F1=7 ...)
#ifndef TQ_PIPE_DUAL
#define TQ_PIPE_DUAL 0
#endif

template <int MODE, int N, int RA, int RB, int NW, int SET = 0>
__global__ void __maxnreg__(pipe_maxnreg(NW)) matsubara_pipe_kernel(const __grid_constant__ PipeArgs A,
                                                                                             const __grid_constant__ CUtensorMap tmap,
                                                                                             const __grid_constant__ CUtensorMap tmap2) {
  constexpr bool PASS1 = (MODE == M_F1 || MODE == M_I1);
  constexpr bool FWD = (MODE == M_F1 || MODE == M_F2);
  constexpr int SGN = FWD ? 1 : -1;
  constexpr int NT = (NW + 1) * 32;
  constexpr bool PK_AHEAD = false; // fetching the next packet id one packet ahead was measured slower (register pressure)
  // measured (32 blocks): a static round robin over the warps (TQ_PK_STATIC) 0.57 -> 0.77 ms and fetching 2-3 packets per
  // atomic (TQ_PK_CHUNK) 0.57 -> 0.69-0.75 ms for F1 -- a stalled warp then holds back packets others could have taken
#ifndef TQ_PK_STATIC
#define TQ_PK_STATIC 0
#endif
  constexpr bool PK_STATIC = TQ_PK_STATIC != 0;
#ifndef TQ_PK_CHUNK
#define TQ_PK_CHUNK 1
#endif
  constexpr int PK_CHUNK = TQ_PK_CHUNK;
#ifndef TQ_PK_LATE
#define TQ_PK_LATE 6 // F2 and I1; F1 and I2 have no register to spare (they spill 160-180 B and lose 40-70 %)
#endif
#ifndef TQ_EARLY_FREE
#define TQ_EARLY_FREE 4 // bit per mode; only the pass-1 kernels (PostScratch) have the code path
#endif
  constexpr bool DUAL = N != 0 && ((TQ_PIPE_DUAL >> MODE) & 1) != 0;
  constexpr int PKL = DUAL ? 64 : 32; // butterflies per packet
  constexpr bool EARLY_FREE = PASS1 && ((TQ_EARLY_FREE >> MODE) & 1) != 0 && !DUAL;
#ifndef TQ_PIPE_L2PF
#define TQ_PIPE_L2PF 10 // F2, I2 (contiguous scratch tiles): -1 %; no gain on the tensor-map gathers of F1 / I1
#endif
  constexpr bool L2PF = ((TQ_PIPE_L2PF >> MODE) & 1) != 0; // bit per mode
#ifndef TQ_PK_LAG0
#define TQ_PK_LAG0 15
#endif
  constexpr bool PK_LAG0 = ((TQ_PK_LAG0 >> MODE) & 1) != 0; // bit per mode
  constexpr bool PK_LATE = !PK_STATIC && PK_CHUNK == 1 && !PK_AHEAD && ((TQ_PK_LATE >> MODE) & 1); // bit per mode
  using Smem = PipeSmem<MODE>;
  // tile shape: template constants in the specialised instantiations, run-time values (A.tN = A.tRA * A.tRB) in the planned ones
  constexpr bool PLANNED = (N == 0);
  const int NN = PLANNED ? A.tN : N, RAv = PLANNED ? A.tRA : RA, RBv = PLANNED ? A.tRB : RB;
  const bool rb_generic = PLANNED && A.rb_generic != 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int TCW = A.tcw;
  const Smem S(NN, RAv, RBv, TCW, A.pitch, A.wmax, rb_generic);
  cd *twS = reinterpret_cast<cd *>(smem_raw + S.tw);
  double4 *qtab = reinterpret_cast<double4 *>(smem_raw + S.qtab);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + S.mbar); // TMA landed + tables written   (producer -> pass A)
  uint64_t *adone = full + pipe_nbuf(MODE);                                // pass A complete               (pass A -> pass B)
  uint64_t *freeb = adone + pipe_nbuf(MODE);                               // pass B has read the tile      (pass B -> producer)
  uint64_t *tabfree = freeb + pipe_nbuf(MODE);                             // pass B is done with the tables (pass B -> producer)
  int *next_pk = reinterpret_cast<int *>(tabfree + pipe_ntab(MODE));       // work-packet counter of the worker warps
  volatile int *ready_tile = next_pk + 1;                            // newest tile the producer has started (its buffer is free)
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n_cols = A.n_cols;
  const bool fermion = A.stat == TQ_FERMION;
  // work packets = 32 butterflies (one per lane): pa packets of pass A and pb packets of pass B per tile
  // (a direct-sum pass B splits every butterfly over fft2_generic_groups(RB) work items)
  const int bgroups = rb_generic ? fft2_generic_groups(RBv) : 1;
  const int pa = (RBv * TCW + PKL - 1) / PKL, pb = (RAv * TCW * bgroups + PKL - 1) / PKL;
  cd *gtwS = reinterpret_cast<cd *>(smem_raw + S.gtw); // generic pass B: exp(+2 pi i j / RB)

  for (int i = tid; i < NN; i += NT) twS[i] = A.twq[i];
  if (Smem::QTAB)
    for (int i = tid; i < S.nq; i += NT) qtab[i] = A.qtab[i];
  if (rb_generic)
    for (int i = tid; i < RBv; i += NT) gtwS[i] = A.gtw[i];
  if (tid == 0) {
    for (int b = 0; b < pipe_nbuf(MODE); ++b) {
      mbar_init(&full[b], 1);
      mbar_init(&adone[b], pa);
      mbar_init(&freeb[b], pb);
    }
    for (int b = 0; b < pipe_ntab(MODE); ++b) mbar_init(&tabfree[b], pb);
    *next_pk = 0;
    *ready_tile = -1;
    fence_mbar_init();
  }
  __syncthreads();

  const int nloc = (int)((A.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x); // tiles of this CTA
  const TileDecoder decode(A.n_items, A.n_ct, TCW, n_cols, PLANNED && A.tr != 0);
  const unsigned m_pa = fastdiv_magic((unsigned)pa), m_pab = fastdiv_magic((unsigned)(pa + pb));
  // With at least NW packets per 5 tiles no worker can be two ring cycles ahead of an unfinished packet, and the phase of
  // buffer b can be established by waiting on its `free` barrier; tiny tiles fall back to polling the producer's counter.
  const bool poll_ready = 5 * (pa + pb) < NW;
  // The guard below can be skipped only when the workers cannot run a ring cycle ahead of a tile whose TMA has not landed: the
  // packets of that tile (pass A waits on `full`, pass B on `adone`) block every warp that claims one, so if even the narrowest
  // column tile has at least NW non-empty packets no warp gets past them.  (Measured gain where it applies: 1 %.)
  const int w_min = n_cols - (A.n_ct - 1) * TCW;
  const bool skip_guard = ((RBv * w_min + PKL - 1) / PKL) + ((RAv * w_min * bgroups + PKL - 1) / PKL) >= NW + 2 && !A.keep_guard;
  // item -> (butterfly, column) division magics of the two possible tile widths (full / last column tile)
  const unsigned m_w_full = fastdiv_magic((unsigned)TCW), m_w_last = fastdiv_magic((unsigned)(n_cols - (A.n_ct - 1) * TCW));
  // developer phase timing (TQ_PHASE_DEBUG): compiled in only with -DTQ_DEVELOPER -- the five 64-bit accumulators and the
  // predicated clock reads cost registers and issue slots in kernels that sit at the register limit (-DTQ_PHASE_TIMING on top)
#if defined(TQ_DEVELOPER) && defined(TQ_PHASE_TIMING)
  long long t0 = clock64(), t_wait = 0, t_work = 0, t_sched = 0, t_ready = 0;
#define TQ_LAP(acc)                                                                                                              \
  if (A.dbg && lane == 0) {                                                                                                      \
    const long long now = clock64();                                                                                             \
    acc += now - t0;                                                                                                             \
    t0 = now;                                                                                                                    \
  }
#else
#define TQ_LAP(acc)
#endif

  if (warp == NW) {
    // ===================== producer warp: TMA + per-tile pole weights =====================
    for (int i = 0; i < nloc; ++i) {
      const long t = blockIdx.x + (long)i * gridDim.x;
      const int b = i % pipe_nbuf(MODE), u = i / pipe_nbuf(MODE), ts = i % pipe_ntab(MODE), ut = i / pipe_ntab(MODE);
      if (ut > 0) mbar_wait(&tabfree[ts], (uint32_t)((ut - 1) & 1)); // tile i - NTAB is completely done
      if (u > 0) mbar_wait(&freeb[b], (uint32_t)((u - 1) & 1));       // pass B of tile i - NBUF has read its inputs
      if (lane == 0) *ready_tile = i; // every barrier of buffer b is now in the phase of tile i
      TQ_LAP(t_wait)
      const TileId id = decode(t);
      cd *dst = reinterpret_cast<cd *>(smem_raw + b * S.buf_stride);
      unsigned char *tab = smem_raw + S.tab0 + ts * S.tab_stride;
      Window win{0, 0};
      int cnt = 0;
      if (Smem::WREC) {
        const int4 wv = A.wintab[id.item];
        win.r1 = wv.x;
        win.r2 = wv.y;
        cnt = win.r1 + (NN - win.r2);
      }
      const bool i1_boxes = MODE == M_I1 && A.i1_rect; // window rows form two rectangles of the [b][a] view: two tensor tiles
      if (lane == 0) {
        uint32_t bytes = 0;
        if (MODE == M_F1)
          bytes = (uint32_t)(sizeof(cd) * NN * A.pitch); // the whole box, zero-filled past n_cols
        else if (MODE == M_I1)
          bytes = (i1_boxes || (PLANNED && A.tr)) ? (uint32_t)(sizeof(cd) * cnt * A.pitch) : (uint32_t)(sizeof(cd) * cnt * id.w);
        else
          bytes = (uint32_t)(sizeof(cd) * NN * id.w);
        if (PASS1) bytes += (uint32_t)(sizeof(cd) * NN);
        if (Smem::WREC) bytes += (uint32_t)(sizeof(double4) * cnt + sizeof(int4));
        bytes += (uint32_t)(sizeof(cd) * 4 * id.w);
        if (Smem::QTAB) bytes += (uint32_t)sizeof(double4);
        mbar_arrive_expect_tx(&full[b], bytes);
        // the tile's table set: pole weights of its columns (:151-169, :238-252), per-item factors, four-step twiddles, window rows
        for (int k = 0; k < 4; ++k)
          bulk_g2s(tab + S.coef_off + sizeof(cd) * (size_t)k * TCW, A.polew + ((size_t)id.gf * 4 + k) * n_cols + id.c0, (uint32_t)(sizeof(cd) * id.w),
                   &full[b]);
        if (Smem::QTAB) bulk_g2s(tab + S.item_off, A.itemtab + id.item, (uint32_t)sizeof(double4), &full[b]);
        if (MODE == M_F1) {
          if constexpr (!PLANNED) {
            tma_load_4d(dst, &tmap, 2 * id.c0, id.item, 0, (int)id.gf, &full[b]); // [gf][b = 0..Nb)[a = item][columns]
          } else {
            for (int bx = 0; bx < A.f1_boxes; ++bx) { // a TMA box has at most 256 rows
              if (A.tr) // tensor [b][a][gf][2 tr doubles]: the columns of the tile are (gf, column) pairs
                tma_load_4d(dst + (size_t)bx * A.f1_box_rows * A.pitch, &tmap, 0, id.c0 / A.tr, id.item, bx * A.f1_box_rows, &full[b]);
              else
                tma_load_4d(dst + (size_t)bx * A.f1_box_rows * A.pitch, &tmap, 2 * id.c0, id.item, bx * A.f1_box_rows, (int)id.gf, &full[b]);
            }
          }
        } else if (MODE == M_I1) {
          if (i1_boxes) {
            tma_load_4d(dst, &tmap, 2 * id.c0, id.item, 0, (int)id.gf, &full[b]);                                  // rows [0, r1): n >= 0
            cd *dst2 = A.i1_edge ? reinterpret_cast<cd *>(smem_raw + b * S.buf_stride + S.stage_off) : dst + (size_t)win.r2 * A.pitch;
            tma_load_4d(dst2, &tmap2, 2 * id.c0, id.item, 0, (int)id.gf, &full[b]); // rows [r2, N): n < 0
          }
        } else {
          bulk_g2s(dst, A.scratch + id.gf * A.scratch_gf_stride + (long)id.item * A.Na * n_cols + (long)id.c0 * A.Na,
                   (uint32_t)(sizeof(cd) * NN * id.w), &full[b]);
        }
        if (PASS1) bulk_g2s(tab + S.t_off, A.Tfull + (size_t)id.item * NN, (uint32_t)(sizeof(cd) * NN), &full[b]);
        if (L2PF && i + 1 < nloc) { // the next tile of this CTA: in L2 by the time its buffer is free
          const TileId nx = decode(t + gridDim.x);
          if (MODE == M_F1)
            tma_prefetch_4d(&tmap, 2 * nx.c0, nx.item, 0, (int)nx.gf);
          else if (MODE == M_F2 || MODE == M_I2)
            bulk_prefetch_l2(A.scratch + nx.gf * A.scratch_gf_stride + (long)nx.item * A.Na * n_cols + (long)nx.c0 * A.Na,
                             (uint32_t)(sizeof(cd) * NN * nx.w));
          else if (i1_boxes) {
            tma_prefetch_4d(&tmap, 2 * nx.c0, nx.item, 0, (int)nx.gf);
            tma_prefetch_4d(&tmap2, 2 * nx.c0, nx.item, 0, (int)nx.gf);
          }
        }
        if (Smem::WREC) bulk_g2s(tab + S.win_off, A.wintab + id.item, (uint32_t)sizeof(int4), &full[b]);
        if (Smem::WREC && cnt > 0)
          bulk_g2s(tab + S.wrec_off, A.wperm + (size_t)id.item * A.wmax, (uint32_t)(sizeof(double4) * cnt), &full[b]);
      }
      if (MODE == M_I1 && !i1_boxes) {
        // window rows only (the rest of the frequency axis is zero, :254): one bulk row copy each
        for (int r0 = lane; r0 < cnt; r0 += 32) {
          const int r = r0 < win.r1 ? r0 : win.r2 + (r0 - win.r1);
          const int k = id.item + A.Na * r;
          const long di = (k <= A.last ? k : k - A.L) - A.first;
          if (PLANNED && A.tr) // one row of tcw / tr gfs: tensor [row][gf][2 tr doubles], box {2 tr, pitch / tr, 1}
            tma_load_3d(dst + (size_t)r * A.pitch, &tmap, 0, id.c0 / A.tr, (int)di, &full[b]);
          else
            bulk_g2s(dst + (size_t)r * A.pitch, A.in + id.gf * A.in_gf_stride + di * n_cols + id.c0, (uint32_t)(sizeof(cd) * id.w), &full[b]);
        }
      }
      TQ_LAP(t_work)
    }
  } else {
    // ===================== worker warps: pull packets in a fixed order (below) =====================
    // (a packet only ever waits for packets that were handed out before it, hence no deadlock)
    int pk_next = 0, pk_left = 0;
    if (!PK_STATIC && (PK_AHEAD || PK_LATE) && lane == 0) pk_next = atomicAdd(next_pk, 1);
    // PK_LATE: the next packet is claimed between the arithmetic and the stores of the current one, so that the round trip of the
    // shared-memory atomic (it queues behind the data traffic) overlaps the stores instead of sitting in front of the next packet
    const auto fetch_next = [&]() {
      if (PK_LATE && lane == 0) pk_next = atomicAdd(next_pk, 1);
    };
    for (int pk_static = warp;; pk_static += NW) {
      int pk;
      if (PK_LATE) {
        pk = __shfl_sync(0xffffffffu, pk_next, 0);
      } else if (PK_STATIC) {
        pk = pk_static; // packets cost the same: round robin over the worker warps balances as well as a shared counter
      } else if (PK_CHUNK > 1) {
        // one shared-memory atomic per PK_CHUNK packets (its latency behind the data traffic in the LSU queue is paid per fetch)
        if (pk_left == 0) {
          if (lane == 0) pk_next = atomicAdd(next_pk, PK_CHUNK);
          pk_next = __shfl_sync(0xffffffffu, pk_next, 0);
          pk_left = PK_CHUNK;
        }
        pk = pk_next++;
        --pk_left;
      } else {
        if (!PK_AHEAD && lane == 0) pk_next = atomicAdd(next_pk, 1);
        pk = __shfl_sync(0xffffffffu, pk_next, 0);
        if (PK_AHEAD && lane == 0) pk_next = atomicAdd(next_pk, 1); // fetched while this packet is processed (hides the atomic's latency)
      }
      int blk, r;
      bool is_b = false;
      if (PLANNED ? A.lag0 != 0 : PK_LAG0) {
        // order A(0) B(0) | A(1) B(1) | ...: pass B right behind pass A of the same tile.  Its first packets wait for the last
        // pass-A packets, but the tile buffer goes back to the producer one tile period earlier, and the ring -- not the
        // arithmetic -- is what limits these kernels (buffer residency = TMA flight + pass A + pass B).  Measured, 32 blocks:
        // F1 0.512 -> 0.478 ms, F2 0.462 -> 0.443, I1 0.466 -> 0.460, I2 0.483 -> 0.462; placing pass B a quarter or half a
        // tile later gives the same within noise, a full tile later (the old order) is 5 % slower.
        blk = fastdiv(pk, m_pab);
        r = pk - blk * (pa + pb);
        is_b = r >= pa;
        if (is_b) r -= pa;
        if (blk >= nloc) break;
        blk += is_b ? 2 : 0; // (keeps the common "k = is_b ? blk - 2 : blk" below)
      } else {
        // order A(0) A(1) | B(0) A(2) | B(1) A(3) | ...: pass B one tile behind pass A (nobody waits for pass-A stragglers)
        if (pk < 2 * pa) {
          blk = fastdiv(pk, m_pa);
          r = pk - blk * pa;
        } else {
          const int q = pk - 2 * pa;
          blk = fastdiv(q, m_pab);
          r = q - blk * (pa + pb);
          blk += 2;
          is_b = r < pb;
          if (!is_b) r -= pb;
        }
        if (blk >= nloc + 2) break;
      }
      const int k = is_b ? blk - 2 : blk; // tile of this CTA
      if (k >= nloc) {
        fetch_next();
        continue;
      }
      const long t = blockIdx.x + (long)k * gridDim.x;
      const int b = k % pipe_nbuf(MODE), u = k / pipe_nbuf(MODE);
      const TileId id = decode(t);
      cd *tile = reinterpret_cast<cd *>(smem_raw + b * S.buf_stride);
      const int ts = k % pipe_ntab(MODE);
      unsigned char *tab = smem_raw + S.tab0 + ts * S.tab_stride;
      const cd *coef = reinterpret_cast<const cd *>(tab + S.coef_off);
      const int pitch = (MODE == M_F1 || MODE == M_I1) ? A.pitch : id.w;
      const int it = r * PKL + lane; // butterfly of this lane (the engine skips it when past the end; dual packets: it and it + 32)
      // a packet of a narrower (last) column tile can be empty for lane 0: its hook never runs, claim the next packet here
      if (PK_LATE && r * PKL >= (is_b ? RAv * bgroups : RBv) * id.w) fetch_next();
      // (I2 sits at the register limit: one more live value makes it spill, so it recomputes the magic per packet)
      const unsigned m_w = (MODE == M_I2 && !PLANNED) ? ~0u : (id.w == TCW ? m_w_full : m_w_last);
      TQ_LAP(t_sched)
      // an mbarrier parity wait is only meaningful for the current or the previous phase: do not look at the barriers of
      // buffer b before the producer has recycled it for tile k (matters when a tile has fewer packets than there are workers)
      if (skip_guard) {
        // nothing: every packet of tile k - NBUF is finished, so full / adone of buffer b are in phase u - 1 (complete) or u
      } else if (poll_ready) {
        while (*ready_tile < k) __nanosleep(32);
      } else if (u > 0) {
        mbar_wait(&freeb[b], (uint32_t)((u - 1) & 1)); // pass B of tile k - NBUF is complete: buffer b is in (or entering) its phase u
      }
      TQ_LAP(t_ready)
      // pass A / pass B of this lane's butterfly: compile-time tile shape in the specialised instantiations, a switch over the
      // radix table in the planned ones (tq_fft2.cuh)
      const auto run_a = [&](auto &pre, auto tw0) {
        constexpr bool TW0 = decltype(tw0)::value;
        if constexpr (DUAL) {
          auto pre1 = pre;
          fft2_pass_a_dual<N, RA, RB, SGN, TW0>(tile, pitch, id.w, twS, it, pre, pre1, m_w, fetch_next);
        } else if constexpr (!PLANNED) {
          fft2_pass_a<N, RA, RB, SGN, TW0>(tile, pitch, id.w, twS, it, 1 << 30, pre, m_w, fetch_next);
        } else {
#define TQ_CALL_A(R) fft2_item_a<R, SGN, TW0>(tile, pitch, id.w, RBv, twS, it, pre, m_w, fetch_next);
          if constexpr (SET == 1) {
            TQ_RADIX_SWITCH_X(RAv, TQ_CALL_A)
          } else {
            TQ_RADIX_SWITCH(RAv, TQ_CALL_A)
          }
#undef TQ_CALL_A
        }
      };
      const auto run_b = [&](auto &post, const auto &hook) {
        if constexpr (DUAL) {
          auto post1 = post;
          fft2_pass_b_dual<N, RA, RB, SGN>(tile, pitch, id.w, it, post, post1, m_w, hook);
        } else if constexpr (!PLANNED) {
          fft2_pass_b<N, RA, RB, SGN>(tile, pitch, id.w, it, 1 << 30, post, m_w, hook);
        } else if (rb_generic) {
          fft2_item_b_generic<SGN>(tile, pitch, id.w, RAv, RBv, gtwS, it, post, m_w, hook);
        } else {
#define TQ_CALL_B(R) fft2_item_b<R, SGN>(tile, pitch, id.w, RAv, it, post, m_w, hook);
          if constexpr (SET == 1) {
            TQ_RADIX_SWITCH_X(RBv, TQ_CALL_B)
          } else {
            TQ_RADIX_SWITCH(RBv, TQ_CALL_B)
          }
#undef TQ_CALL_B
        }
      };
      if (!is_b) {
        mbar_wait(&full[b], (uint32_t)(u & 1));
        TQ_LAP(t_wait)
        if (MODE == M_F1) {
          PreF1 pre{A, coef, qtab, *reinterpret_cast<const double4 *>(tab + S.item_off)};
          run_a(pre, std::true_type{});
        } else if (MODE == M_I1) {
          const int4 wv = *reinterpret_cast<const int4 *>(tab + S.win_off);
          PreI1<RB> pre{reinterpret_cast<const cd *>(tab + S.wrec_off), coef, TCW, Window{wv.x, wv.y}, A.inv_beta, RBv};
          bool done = false;
          if constexpr (!PLANNED) {
            if (A.i1_edge) {
              auto wedge = [&](int rr) { return A.wedge[rr]; };
              fft2_pass_a_edge<N, RA, RB, false>(tile, reinterpret_cast<const cd *>(smem_raw + b * S.buf_stride + S.stage_off), pitch, id.w, twS, it,
                                                 1 << 30, pre, wedge, m_w);
              if constexpr (DUAL)
                fft2_pass_a_edge<N, RA, RB, false>(tile, reinterpret_cast<const cd *>(smem_raw + b * S.buf_stride + S.stage_off), pitch, id.w, twS,
                                                   it + 32, 1 << 30, pre, wedge, m_w);
              if (r * PKL < RB * id.w) fetch_next(); // (an empty packet has claimed its successor above)
              done = true;
            }
          }
          if (!done) run_a(pre, std::false_type{});
        } else {
          PreIdentity pre;
          run_a(pre, std::integral_constant<bool, MODE == M_I2>{});
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&adone[b]);
        TQ_LAP(t_work)
      } else {
        mbar_wait(&adone[b], (uint32_t)(u & 1));
        TQ_LAP(t_wait)
        if constexpr (!EARLY_FREE) {
          if (PASS1) {
            PostScratch<RA> post{reinterpret_cast<const cd *>(tab + S.t_off), A.scratch + id.gf * A.scratch_gf_stride, (long)A.Na * n_cols, n_cols,
                                 A.Na, A.tcw2, id.item, id.c0, RAv};
            run_b(post, fetch_next);
          } else if (MODE == M_F2) {
            const int4 wv = *reinterpret_cast<const int4 *>(tab + S.win_off);
            const bool tr = PLANNED && A.tr != 0;
            PostF2<RA> post{reinterpret_cast<const cd *>(tab + S.wrec_off), coef, TCW, tr ? A.tr : n_cols, id.item, A.Nb, A.L, A.first, Window{wv.x, wv.y},
                            tr ? A.out + (long)(id.c0 / A.tr) * A.col_stride_out : A.out + id.gf * A.out_gf_stride + id.c0, RAv, tr ? A.col_stride_out : 0L,
                            tr ? A.tr : 1};
            run_b(post, fetch_next);
          } else {
            const bool tr = PLANNED && A.tr != 0;
            PostI2<RA> post{A, coef, qtab, *reinterpret_cast<const double4 *>(tab + S.item_off), tr ? A.tr : n_cols, id.item, A.Nb, A.L, fermion,
                            tr ? A.out + (long)(id.c0 / A.tr) * A.col_stride_out : A.out + id.gf * A.out_gf_stride + id.c0, RAv, tr ? A.col_stride_out : 0L,
                            tr ? A.tr : 1};
            run_b(post, fetch_next);
          }
          __syncwarp(); // every lane's reads of the tile are done ...
          if (lane == 0) {
            fence_proxy_async(); // ... and ordered before the TMA engine's (async proxy) next write to the buffer
            mbar_arrive(&freeb[b]);
            mbar_arrive(&tabfree[ts]);
          }
        } else {
          // EARLY_FREE (I1): the tile buffer goes back to the producer as soon as the packet's loads are in registers (the butterfly
          // has consumed them when the hook runs), not after the epilogue and its scratch stores; the tables stay until the
          // stores are done (`tabfree`).  I1: 0.454 -> 0.436 ms.  No gain on F1 / F2, and I2 has no register to spare.
          const unsigned act = __ballot_sync(0xffffffffu, it < RAv * bgroups * id.w);
          const auto release = [&]() {
            __syncwarp(act);
            if (lane == 0) {
              fence_proxy_async();
              mbar_arrive(&freeb[b]);
            }
            fetch_next();
          };
          PostScratch<RA> post{reinterpret_cast<const cd *>(tab + S.t_off), A.scratch + id.gf * A.scratch_gf_stride, (long)A.Na * n_cols, n_cols,
                               A.Na, A.tcw2, id.item, id.c0, RAv};
          run_b(post, release);
          __syncwarp();
          if (lane == 0) {
            if (act == 0u) { // empty packet: nobody ran the hook
              fence_proxy_async();
              mbar_arrive(&freeb[b]);
            }
            mbar_arrive(&tabfree[ts]);
          }
        }
        TQ_LAP(t_work)
      }
    }
  }
#if defined(TQ_DEVELOPER) && defined(TQ_PHASE_TIMING)
  if (A.dbg && lane == 0) {
    A.dbg[(size_t)blockIdx.x * 64 + 4 * warp] = t_wait;
    A.dbg[(size_t)blockIdx.x * 64 + 4 * warp + 1] = t_work;
    A.dbg[(size_t)blockIdx.x * 64 + 4 * warp + 2] = t_sched;
    A.dbg[(size_t)blockIdx.x * 64 + 4 * warp + 3] = t_ready;
  }
#endif
#undef TQ_LAP
}

template <int MODE, int N, int RA, int RB, int NW, int SET = 0>
static tq_status launch_pipe(tq_ctx *ctx, const PipeArgs &A, const CUtensorMap &tmap, const CUtensorMap &tmap2, const char *tag) {
  constexpr int NT = (NW + 1) * 32;
  const size_t smem = PipeSmem<MODE>(N ? N : A.tN, RA ? RA : A.tRA, RB ? RB : A.tRB, A.tcw, A.pitch, A.wmax, N ? 0 : A.rb_generic).total;
  auto k = matsubara_pipe_kernel<MODE, N, RA, RB, NW, SET>;
  TQ_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<long>(A.n_tiles, (long)ctx->sm_count);
  static const bool phase_debug = TQ_DEV_ENV("TQ_PHASE_DEBUG") != nullptr;
  PipeArgs Ad = A;
  if (phase_debug) {
    TQ_CUDA(ctx, cudaMallocManaged(&Ad.dbg, sizeof(long long) * 64 * grid));
    TQ_CUDA(ctx, cudaMemset(Ad.dbg, 0, sizeof(long long) * 64 * grid));
  }
  {
    KernelScope ks(ctx, tag);
    k<<<grid, NT, smem, ctx->stream>>>(Ad, tmap, tmap2);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  if (phase_debug) {
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // per role: average over CTAs and warps of (cycles waiting, cycles working) per tile
    double w[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int i = 0; i < grid; ++i)
      for (int wp = 0; wp < NW + 1; ++wp) {
        const int role = wp < NW ? 0 : 1;
        const int nw = role == 0 ? NW : 1;
        for (int j = 0; j < 4; ++j) w[role][j] += (double)Ad.dbg[(size_t)i * 64 + 4 * wp + j] / nw;
      }
    const double per = (double)A.n_tiles;
    fprintf(stderr,
            "[pipe] %-13s N=%d tcw=%d tiles=%ld grid=%d smem=%zu thr=%d  cycles/tile  workers: sched %.0f ready-spin %.0f mbar-wait %.0f work %.0f | "
            "producer: wait %.0f work %.0f\n",
            tag, N, A.tcw, A.n_tiles, grid, smem, NT, w[0][2] / per, w[0][3] / per, w[0][0] / per, w[0][1] / per, w[1][0] / per, w[1][1] / per);
    cudaFree(Ad.dbg);
  }
  return TQ_OK;
}

