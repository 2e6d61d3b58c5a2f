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
#include "tq_fft2.cuh"

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstring>

// -------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + bulk async copies (TMA engine; SASS: UBLKCP / UTMALDG) + proxy fence
// -------------------------------------------------------------------------------------------------
TQ_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
TQ_D void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
TQ_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TQ_D void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
TQ_D void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
TQ_D void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
     "{\n"
     ".reg .pred P1;\n"
     "LAB_WAIT:\n"
     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
     "@P1 bra DONE;\n"
     "bra LAB_WAIT;\n"
     "DONE:\n"
     "}\n" ::"r"(smem_u32(bar)),
     "r"(parity)
     : "memory");
}
TQ_D void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
TQ_D void tma_load_4d(void *dst_smem, const CUtensorMap *tmap, int c0, int c1, int c2, int c3, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                  smem_u32(dst_smem)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
               : "memory");
}
TQ_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// -------------------------------------------------------------------------------------------------
enum { M_F1 = 0, M_F2 = 1, M_I1 = 2, M_I2 = 3 };
constexpr int PIPE_NBUF = 3;
constexpr int PIPE_MAXR = 25; // largest radix of a pass

struct PipeArgs {
  int Na, Nb, n_cols;
  int tcw, n_ct;   // column tiling of this pass
  int tcw2;        // column tiling of pass 2 (defines the scratch layout)
  int n_items;     // Na (pass 1) or Nb (pass 2)
  int wmax;        // row pitch of the compact window-row tables (F2, I1)
  long n_tiles;    // batch * n_items * n_ct
  const cd *in;
  cd *out;
  cd *scratch;
  const cd *moments; // [batch][mom_rows][n_cols]; rows 1..3 are used
  long in_gf_stride, out_gf_stride, scratch_gf_stride, mom_gf_stride;
  int L, first, last, stat;
  double half_fact_fwd; // 0.5 beta / L
  double inv_beta;
  const cd *twq;          // [RB][RA] pass-A twiddle rows w_N^{SGN q s} x per-q (F1) / per-s (I2) phase
  const double4 *qtab;    // F1: [RB] {E1q, E2q, E3q, -};  I2: [RA] {E1s, E2s, E3s, -}
  const double4 *itemtab; // [n_items] {C1 ea1, C2 ea2, C3 ea3, -}   (F1: item = a, I2: item = beta)
  const cd *Tfull;        // [Na][Nb] four-step twiddle x per-item phase (pass 1)
  const double4 *wperm;   // [n_items][wmax] {b1 r1, w r1, b2 r2, w r2} of the window rows of every item (F2, I1)
  // per-leg constants (constant-bank operands): F1 legs r of pass A, I2 legs t of pass B
  double er[3][PIPE_MAXR];
  cd phr[PIPE_MAXR];
  long long *dbg; // developer phase timing (TQ_PHASE_DEBUG) or nullptr
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

// sum_i a_i / (i w - b_i) with b3 = -b2:   -(b1 r1) a1 - i (w r1) a1 - (b2 r2)(a2 - a3) - i (w r2)(a2 + a3)
TQ_D cd tail_iw4(cd w01, cd w23, cd a1, cd D23, cd A23) {
  const double p1 = w01.x, u = w01.y, p2 = w23.x, v = w23.y;
  double re = fma(u, a1.y, -p1 * a1.x);
  re = fma(-p2, D23.x, re);
  re = fma(v, A23.y, re);
  double im = fma(-u, a1.x, -p1 * a1.y);
  im = fma(-p2, D23.y, im);
  im = fma(-v, A23.x, im);
  return cmk(re, im);
}

// ---- functors plugged into the two-pass engine -----------------------------------------------
struct PreF1 { // x = phase (g - sum_i at_i E_i[row]),  row = q + RB r        fourier_matsubara.cpp:159-173
  const PipeArgs &A;
  const cd *coef;
  const double4 *qtab;
  int TCW;
  cd a1, a2, a3;
  TQ_D void begin(int q, int c) {
    const double4 e = qtab[q];
    a1 = cscale(e.x, coef[c]);
    a2 = cscale(e.y, coef[TCW + c]);
    a3 = cscale(e.z, coef[2 * TCW + c]);
  }
  template <int R> TQ_D cd load(const cd *p) const {
    cd v = caxpy(-A.er[0][R], a1, *p);
    v = caxpy(-A.er[1][R], a2, v);
    v = caxpy(-A.er[2][R], a3, v);
    return R == 0 ? v : cmul(v, A.phr[R]); // the per-q part of the twist sits in the twiddle row
  }
};
template <int RB> struct PreI1 { // x = (g - tail(i w)) / beta for window rows, 0 elsewhere      :254
  const cd *wrec, *coef;
  int TCW;
  Window win;
  double inv_beta;
  cd a1, D23, A23;
  int q;
  TQ_D void begin(int qq, int c) {
    q = qq;
    a1 = coef[c];
    D23 = coef[TCW + c];
    A23 = coef[2 * TCW + c];
  }
  template <int R> TQ_D cd load(const cd *p) const {
    const int row = q + RB * R;
    int idx;
    if (row < win.r1)
      idx = row;
    else if (row >= win.r2)
      idx = win.r1 + row - win.r2;
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
  cd *p;
  const cd *Ts;
  TQ_D void begin(int sidx, int c) {
    const int cg = c0 + c;
    const int h = cg / tcw2;
    const int cc = cg - h * tcw2;
    const int w = min(tcw2, n_cols - h * tcw2);
    p = base + ((long)h * tcw2 * Na + (long)item * w + cc) + (long)sidx * beta_stride;
    Ts = T + sidx;
  }
  template <int TT> TQ_D void store(cd v) const { p[(long)(RA * TT) * beta_stride] = cmul(v, Ts[RA * TT]); }
};
template <int RA> struct PostF2 { // G(i w_n) = z + corr + tail, window rows only      :181-183
  const cd *wrec, *coef;
  int TCW, n_cols, item, Nb, L, first;
  Window win;
  cd *base; // out of this gf + c0
  cd a1, D23, A23, ex;
  cd *p;
  int sidx;
  TQ_D void begin(int s, int c) {
    sidx = s;
    a1 = coef[c];
    D23 = coef[TCW + c];
    A23 = coef[2 * TCW + c];
    ex = coef[3 * TCW + c];
    p = base + c;
  }
  template <int TT> TQ_D void store(cd v) const {
    const int k = sidx + RA * TT;
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
  const cd *coef;
  const double4 *qtab;
  int TCW, n_cols, item, Nb, L;
  bool fermion;
  cd *base; // out of this gf + c0
  cd a1, a2, a3, m1;
  cd *p;
  int j0;
  TQ_D void begin(int s, int c) {
    const double4 e = qtab[s];
    a1 = cscale(e.x, coef[c]);
    a2 = cscale(e.y, coef[TCW + c]);
    a3 = cscale(e.z, coef[2 * TCW + c]);
    m1 = coef[3 * TCW + c];
    j0 = item + Nb * s;
    p = base + c + (long)j0 * n_cols;
  }
  template <int TT> TQ_D void store(cd z) const {
    cd v = TT == 0 ? z : cmul(z, A.phr[TT]); // the per-s part of the twist was folded into the pass-A twiddle row
    v = caxpy(A.er[0][TT], a1, v);
    v = caxpy(A.er[1][TT], a2, v);
    v = caxpy(A.er[2][TT], a3, v);
    p[(long)(Nb * RA * TT) * n_cols] = v;
    if (TT == 0 && j0 == 0) { // gt[L] = -/+ (gt[0] + m1)     :267-268
      const cd t = cadd(v, m1);
      p[(long)L * n_cols] = fermion ? cneg(t) : t;
    }
  }
};

// shared-memory carve-up (bytes); tile buffers first, each padded to 128 B (TMA tensor destination alignment)
template <int MODE, int N, int RA, int RB> struct PipeSmem {
  static constexpr bool PASS1 = (MODE == M_F1 || MODE == M_I1);
  static constexpr bool QTAB = (MODE == M_F1 || MODE == M_I2);
  static constexpr bool WREC = (MODE == M_F2 || MODE == M_I1);
  static constexpr int NQ = (MODE == M_F1) ? RB : RA;
  size_t buf_stride, tw, qtab, tab0, tab_stride, coef_off, t_off, wrec_off, mbar, total;
  __host__ __device__ PipeSmem(int tcw, int wmax) {
    buf_stride = (sizeof(cd) * (size_t)N * tcw + 127) & ~size_t(127);
    tw = PIPE_NBUF * buf_stride;
    qtab = tw + sizeof(cd) * N;
    tab0 = qtab + (QTAB ? sizeof(double4) * NQ : 0);
    coef_off = 0;
    t_off = coef_off + sizeof(cd) * 4 * (size_t)tcw;
    wrec_off = t_off + (PASS1 ? sizeof(cd) * N : 0);
    tab_stride = wrec_off + (WREC ? sizeof(cd) * 2 * (size_t)wmax : 0);
    mbar = tab0 + PIPE_NBUF * tab_stride;
    total = mbar + 8 * 3 * PIPE_NBUF;
  }
};

struct TileId {
  long gf;
  int item, c0, w;
};
TQ_D TileId decode_tile(long t, int n_items, int n_ct, int tcw, int n_cols) {
  TileId id;
  const long per_gf = (long)n_items * n_ct;
  id.gf = t / per_gf;
  const int rem = (int)(t - id.gf * per_gf);
  id.item = rem / n_ct;
  id.c0 = (rem - id.item * n_ct) * tcw;
  id.w = min(tcw, n_cols - id.c0);
  return id;
}

template <int MODE, int N, int RA, int RB, int WA, int WB>
__global__ void __maxnreg__((16384 / (((WA + WB + 1) + 3) / 4 * 32)) & ~7) matsubara_pipe_kernel(const __grid_constant__ PipeArgs A,
                                                                                     const __grid_constant__ CUtensorMap tmap) {
  constexpr bool PASS1 = (MODE == M_F1 || MODE == M_I1);
  constexpr bool FWD = (MODE == M_F1 || MODE == M_F2);
  constexpr int SGN = FWD ? 1 : -1;
  constexpr int NT = (WA + WB + 1) * 32;
  using Smem = PipeSmem<MODE, N, RA, RB>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int TCW = A.tcw;
  const Smem S(TCW, A.wmax);
  cd *twS = reinterpret_cast<cd *>(smem_raw + S.tw);
  double4 *qtab = reinterpret_cast<double4 *>(smem_raw + S.qtab);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + S.mbar); // TMA landed + tables written   (producer -> A)
  uint64_t *adone = full + PIPE_NBUF;                                // pass A complete               (A -> B)
  uint64_t *freeb = adone + PIPE_NBUF;                               // pass B has consumed the tile  (B -> producer)
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n_cols = A.n_cols;
  const bool fermion = A.stat == TQ_FERMION;

  for (int i = tid; i < N; i += NT) twS[i] = A.twq[i];
  if (Smem::QTAB)
    for (int i = tid; i < Smem::NQ; i += NT) qtab[i] = A.qtab[i];
  if (tid == 0) {
    for (int b = 0; b < PIPE_NBUF; ++b) {
      mbar_init(&full[b], 2);
      mbar_init(&adone[b], WA);
      mbar_init(&freeb[b], WB);
    }
    fence_mbar_init();
  }
  __syncthreads();

  long long t0 = clock64(), t_wait = 0, t_work = 0;
#define TQ_LAP(acc)                                                                                                              \
  if (A.dbg && lane == 0) {                                                                                                      \
    const long long now = clock64();                                                                                             \
    acc += now - t0;                                                                                                             \
    t0 = now;                                                                                                                    \
  }
  // timeline of CTA 0 (first warp of each role, first 16 tiles): trace[role][tile][0 = got the buffer, 1 = handed it on]
  long long *trace = (A.dbg && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == WA || warp == WA + WB)) ? A.dbg + (size_t)gridDim.x * 64 : nullptr;
  const int trole = warp == 0 ? 0 : (warp == WA ? 1 : 2);
#define TQ_TRACE(i, ev)                                                                                                          \
  if (trace && (i) < 16) trace[(trole * 16 + (i)) * 2 + (ev)] = clock64();

  if (warp == WA + WB) {
    // ===================== producer warp: TMA + per-tile pole weights =====================
    int i = 0;
    for (long t = blockIdx.x; t < A.n_tiles; t += gridDim.x, ++i) {
      const int b = i % PIPE_NBUF, u = i / PIPE_NBUF;
      if (u > 0) mbar_wait(&freeb[b], (uint32_t)((u - 1) & 1));
      TQ_LAP(t_wait)
      TQ_TRACE(i, 0)
      const TileId id = decode_tile(t, A.n_items, A.n_ct, TCW, n_cols);
      cd *dst = reinterpret_cast<cd *>(smem_raw + b * S.buf_stride);
      unsigned char *tab = smem_raw + S.tab0 + b * S.tab_stride;
      cd *coef = reinterpret_cast<cd *>(tab + S.coef_off);
      Window win{0, 0};
      int cnt = 0;
      if (Smem::WREC) {
        win = tile_window(id.item, MODE == M_F2 ? A.Nb : A.Na, N, A.L, A.first, A.last);
        cnt = win.r1 + (N - win.r2);
      }
      if (lane == 0) {
        uint32_t bytes = 0;
        if (MODE == M_F1)
          bytes = (uint32_t)(sizeof(cd) * N * TCW); // the whole box, zero-filled past n_cols
        else if (MODE == M_I1)
          bytes = (uint32_t)(sizeof(cd) * cnt * id.w);
        else
          bytes = (uint32_t)(sizeof(cd) * N * id.w);
        if (PASS1) bytes += (uint32_t)(sizeof(cd) * N);
        if (Smem::WREC) bytes += (uint32_t)(sizeof(double4) * cnt);
        mbar_arrive_expect_tx(&full[b], bytes);
        if (MODE == M_F1)
          tma_load_4d(dst, &tmap, 2 * id.c0, id.item, 0, (int)id.gf, &full[b]); // [gf][b = 0..Nb)[a = item][columns]
        else if (MODE != M_I1)
          bulk_g2s(dst, A.scratch + id.gf * A.scratch_gf_stride + (long)id.item * A.Na * n_cols + (long)id.c0 * A.Na,
                   (uint32_t)(sizeof(cd) * N * id.w), &full[b]);
        if (PASS1) bulk_g2s(tab + S.t_off, A.Tfull + (size_t)id.item * N, (uint32_t)(sizeof(cd) * N), &full[b]);
        if (Smem::WREC && cnt > 0)
          bulk_g2s(tab + S.wrec_off, A.wperm + (size_t)id.item * A.wmax, (uint32_t)(sizeof(double4) * cnt), &full[b]);
      }
      if (MODE == M_I1) {
        // window rows only (the rest of the frequency axis is zero, :254): one bulk row copy each, tile pitch = w
        for (int r0 = lane; r0 < cnt; r0 += 32) {
          const int r = r0 < win.r1 ? r0 : win.r2 + (r0 - win.r1);
          const int k = id.item + A.Na * r;
          const long di = (k <= A.last ? k : k - A.L) - A.first;
          bulk_g2s(dst + (size_t)r * id.w, A.in + id.gf * A.in_gf_stride + di * n_cols + id.c0, (uint32_t)(sizeof(cd) * id.w), &full[b]);
        }
      }
      // pole weights of the tile's columns  (:151-169, :238-252)
      for (int c = lane; c < id.w; c += 32) {
        const cd *mom = A.moments + id.gf * A.mom_gf_stride + id.c0 + c;
        const cd m1 = mom[(size_t)1 * n_cols], m2 = mom[(size_t)2 * n_cols], m3 = mom[(size_t)3 * n_cols];
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
        if (MODE == M_F1 || MODE == M_I2) {
          const double4 ea = A.itemtab[id.item];
          coef[c] = cscale(ea.x, a1);
          coef[TCW + c] = cscale(ea.y, a2);
          coef[2 * TCW + c] = cscale(ea.z, a3);
          coef[3 * TCW + c] = m1;
        } else {
          coef[c] = a1;
          coef[TCW + c] = csub(a2, a3);
          coef[2 * TCW + c] = cadd(a2, a3);
          if (MODE == M_F2) {
            // trapezoid end-point correction  -0.5 (beta/L) (g_0 + m1 +/- g_L)   (:181); A.in is G(tau) here
            const cd *g = A.in + id.gf * A.in_gf_stride + id.c0 + c;
            const cd g0 = g[0], gL = g[(size_t)A.L * n_cols];
            cd s = cadd(g0, m1);
            s = fermion ? cadd(s, gL) : csub(s, gL);
            coef[3 * TCW + c] = cscale(-A.half_fact_fwd, s);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[b]); // second arrival (release): the table set is visible to whoever sees the phase flip
      TQ_LAP(t_work)
      TQ_TRACE(i, 1)
    }
  } else if (warp < WA) {
    // ===================== A group: pass A in place =====================
    int i = 0;
    for (long t = blockIdx.x; t < A.n_tiles; t += gridDim.x, ++i) {
      const int b = i % PIPE_NBUF, u = i / PIPE_NBUF;
      const TileId id = decode_tile(t, A.n_items, A.n_ct, TCW, n_cols);
      cd *tile = reinterpret_cast<cd *>(smem_raw + b * S.buf_stride);
      unsigned char *tab = smem_raw + S.tab0 + b * S.tab_stride;
      const cd *coef = reinterpret_cast<const cd *>(tab + S.coef_off);
      const int pitch = (MODE == M_F1) ? TCW : id.w;
      mbar_wait(&full[b], (uint32_t)(u & 1));
      TQ_LAP(t_wait)
      TQ_TRACE(i, 0)
      if (MODE == M_F1) {
        PreF1 pre{A, coef, qtab, TCW};
        fft2_pass_a<N, RA, RB, SGN, true>(tile, pitch, id.w, twS, tid, WA * 32, pre);
      } else if (MODE == M_I1) {
        PreI1<RB> pre{reinterpret_cast<const cd *>(tab + S.wrec_off), coef, TCW, tile_window(id.item, A.Na, N, A.L, A.first, A.last),
                      A.inv_beta};
        fft2_pass_a<N, RA, RB, SGN, false>(tile, pitch, id.w, twS, tid, WA * 32, pre);
      } else {
        PreIdentity pre;
        fft2_pass_a<N, RA, RB, SGN, MODE == M_I2>(tile, pitch, id.w, twS, tid, WA * 32, pre);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&adone[b]);
      TQ_LAP(t_work)
      TQ_TRACE(i, 1)
    }
  } else {
    // ===================== B group: pass B + epilogue -> global =====================
    const int tb = tid - WA * 32;
    int i = 0;
    for (long t = blockIdx.x; t < A.n_tiles; t += gridDim.x, ++i) {
      const int b = i % PIPE_NBUF, u = i / PIPE_NBUF;
      const TileId id = decode_tile(t, A.n_items, A.n_ct, TCW, n_cols);
      const cd *tile = reinterpret_cast<const cd *>(smem_raw + b * S.buf_stride);
      unsigned char *tab = smem_raw + S.tab0 + b * S.tab_stride;
      const cd *coef = reinterpret_cast<const cd *>(tab + S.coef_off);
      const int pitch = (MODE == M_F1) ? TCW : id.w;
      mbar_wait(&adone[b], (uint32_t)(u & 1));
      TQ_LAP(t_wait)
      TQ_TRACE(i, 0)
      if (PASS1) {
        PostScratch<RA> post{reinterpret_cast<const cd *>(tab + S.t_off), A.scratch + id.gf * A.scratch_gf_stride, (long)A.Na * n_cols, n_cols,
                             A.Na, A.tcw2, id.item, id.c0};
        fft2_pass_b<N, RA, RB, SGN>(tile, pitch, id.w, tb, WB * 32, post);
      } else if (MODE == M_F2) {
        PostF2<RA> post{reinterpret_cast<const cd *>(tab + S.wrec_off), coef, TCW, n_cols, id.item, A.Nb, A.L, A.first,
                        tile_window(id.item, A.Nb, N, A.L, A.first, A.last), A.out + id.gf * A.out_gf_stride + id.c0};
        fft2_pass_b<N, RA, RB, SGN>(tile, pitch, id.w, tb, WB * 32, post);
      } else {
        PostI2<RA> post{A, coef, qtab, TCW, n_cols, id.item, A.Nb, A.L, fermion, A.out + id.gf * A.out_gf_stride + id.c0};
        fft2_pass_b<N, RA, RB, SGN>(tile, pitch, id.w, tb, WB * 32, post);
      }
      fence_proxy_async(); // our generic-proxy reads of the tile are ordered before the TMA engine's next write to it
      __syncwarp();
      if (lane == 0) mbar_arrive(&freeb[b]);
      TQ_LAP(t_work)
      TQ_TRACE(i, 1)
    }
  }
  if (A.dbg && lane == 0) {
    A.dbg[(size_t)blockIdx.x * 64 + 2 * warp] = t_wait;
    A.dbg[(size_t)blockIdx.x * 64 + 2 * warp + 1] = t_work;
  }
#undef TQ_LAP
#undef TQ_TRACE
}

// -------------------------------------------------------------------------------------------------
// host
// -------------------------------------------------------------------------------------------------
struct ModeTables {
  const cd *twq = nullptr;
  const double4 *qtab = nullptr, *itemtab = nullptr, *wperm = nullptr;
  const cd *Tfull = nullptr;
  double er[3][PIPE_MAXR];
  cd phr[PIPE_MAXR];
};
struct PipePlan {
  int L = 0, Na = 0, Nb = 0, first = 0, last = 0, n_w = 0, stat = 0, wmax1 = 0, wmax2 = 0;
  double beta = 0;
  ModeTables m[4];
  std::vector<void *> allocs;
  ~PipePlan() {
    for (void *p : allocs) cudaFree(p);
  }
};
struct PipePlanCache {
  std::map<std::tuple<double, int, long, long>, std::shared_ptr<PipePlan>> plans;
};

template <class T> static tq_status pp_upload(tq_ctx *ctx, PipePlan &pl, const std::vector<T> &h, const T **d) {
  void *p = nullptr;
  TQ_CUDA(ctx, cudaMalloc(&p, sizeof(T) * std::max<size_t>(h.size(), 1)));
  pl.allocs.push_back(p);
  TQ_CUDA(ctx, cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // h may be a temporary
  *d = (const T *)p;
  return TQ_OK;
}

static const long double PI_L = 3.141592653589793238462643383279502884L;
static cd expi(long double ang) { return cmk((double)cosl(ang), (double)sinl(ang)); }

// Factor tables of the tau-side pre/post transform for  j = item + S row,  row = x + X y  (x in [0,X): static table,
// y in [0,Y): per-leg kernel parameters),  X * Y = R rows:
//   exp(-|b_i| delta idx_i(j)),  idx_i = j (b_i >= 0)  or  L - j = (S - item) + S (X-1-x) + S X (Y-1-y)  (b_i < 0)
//   phase e^{sgn i pi j / L} = [item] x [x] x [y]
struct TauFactors {
  std::vector<double4> xtab, itemtab; // [X], [S]
  std::vector<cd> xph, itemph;        // [X], [S]
  double er[3][PIPE_MAXR];
  cd phr[PIPE_MAXR];
};
static void tau_factors(double beta, int stat, long L, int S, int X, int Y, int sgn, double scale, const double b[3], const double C[3],
                        TauFactors &f) {
  const long double delta = (long double)beta / (long double)L;
  f.xtab.assign(X, make_double4(0, 0, 0, 0));
  f.itemtab.assign(S, make_double4(0, 0, 0, 0));
  f.xph.assign(X, cmk(1, 0));
  f.itemph.assign(S, cmk(scale, 0));
  auto ex = [&](int i, long idx) { return (double)expl(-fabsl((long double)b[i]) * delta * (long double)idx); };
  for (int x = 0; x < X; ++x) {
    double e[3];
    for (int i = 0; i < 3; ++i) e[i] = ex(i, b[i] >= 0 ? (long)S * x : (long)S * (X - 1 - x));
    f.xtab[x] = make_double4(e[0], e[1], e[2], 0.0);
    if (stat == TQ_FERMION) f.xph[x] = expi(sgn * PI_L * (long double)((long)S * x) / (long double)L);
  }
  for (int y = 0; y < PIPE_MAXR; ++y) {
    for (int i = 0; i < 3; ++i) f.er[i][y] = y < Y ? ex(i, b[i] >= 0 ? (long)S * X * y : (long)S * X * (Y - 1 - y)) : 0.0;
    f.phr[y] = (stat == TQ_FERMION && y < Y) ? expi(sgn * PI_L * (long double)((long)S * X * y) / (long double)L) : cmk(1, 0);
  }
  for (int a = 0; a < S; ++a) {
    double e[3];
    for (int i = 0; i < 3; ++i) e[i] = (double)((long double)C[i] * (long double)ex(i, b[i] >= 0 ? a : (S - a)));
    f.itemtab[a] = make_double4(e[0], e[1], e[2], 0.0);
    if (stat == TQ_FERMION) {
      long double ang = sgn * PI_L * (long double)a / (long double)L;
      f.itemph[a] = cmk((double)((long double)scale * cosl(ang)), (double)((long double)scale * sinl(ang)));
    }
  }
}

// pass-A twiddle rows  twq[q][s] = w_N^{sgn q s} x rowph[q] x colph[s]
static std::vector<cd> twiddle_rows(int N, int RA, int RB, int sgn, const std::vector<cd> *rowph, const std::vector<cd> *colph) {
  std::vector<cd> w;
  tq_fft_twiddles_host(N, w);
  std::vector<cd> out((size_t)N);
  for (int q = 0; q < RB; ++q)
    for (int s = 0; s < RA; ++s) {
      cd t = w[(q * s) % N];
      if (sgn < 0) t = cconj(t);
      if (rowph) t = cmul(t, (*rowph)[q]);
      if (colph) t = cmul(t, (*colph)[s]);
      out[(size_t)q * RA + s] = t;
    }
  return out;
}

static bool pipe_supported_length(long L, int *Na, int *Nb) {
  if (L == 100000) {
    *Na = 400;
    *Nb = 250;
    return true;
  }
  return false;
}
// radices of the two tile lengths (must match the kernel instantiations below)
constexpr int P1_N = 250, P1_RA = 10, P1_RB = 25;
constexpr int P2_N = 400, P2_RA = 20, P2_RB = 20;

static tq_status get_pipe_plan(tq_ctx *ctx, double beta, int stat, long n_tau, long n_iw, std::shared_ptr<PipePlan> *out) {
  if (!ctx->pipe_cache) ctx->pipe_cache = std::make_shared<PipePlanCache>();
  auto &plans = ctx->pipe_cache->plans;
  auto key = std::make_tuple(beta, stat, n_tau, n_iw);
  auto it = plans.find(key);
  if (it != plans.end()) {
    *out = it->second;
    return TQ_OK;
  }
  auto pl = std::make_shared<PipePlan>();
  const long L = n_tau - 1;
  int Na = 0, Nb = 0;
  if (!pipe_supported_length(L, &Na, &Nb)) return tq_fail(ctx, TQ_ERR_INVALID, "internal: pipe plan for an unsupported length");
  pl->L = (int)L;
  pl->Na = Na;
  pl->Nb = Nb;
  pl->stat = stat;
  pl->beta = beta;
  pl->last = (int)(n_iw - 1);                                  // imfreq.hpp:58-60
  pl->first = -(pl->last + (stat == TQ_FERMION ? 1 : 0));
  pl->n_w = pl->last - pl->first + 1;
  pl->wmax1 = (int)std::min<long>(Nb, pl->n_w / Na + 2);
  pl->wmax2 = (int)std::min<long>(Na, pl->n_w / Nb + 2);
  double b[3], C[3];
  if (stat == TQ_FERMION) {
    b[0] = 0; b[1] = 1; b[2] = -1; // fourier_matsubara.cpp:152-154
  } else {
    b[0] = -0.5; b[1] = -1; b[2] = 1; // :163-165
  }
  for (int i = 0; i < 3; ++i) {
    long double eb = expl(-(long double)beta * fabsl((long double)b[i]));
    if (stat == TQ_FERMION)
      C[i] = (double)(-1.0L / (1.0L + eb)); // oneFermion :28-30
    else
      C[i] = b[i] >= 0 ? (double)(1.0L / (eb - 1.0L)) : (double)(1.0L / (1.0L - eb)); // oneBoson :32-34
  }
  TauFactors f1, f2;
  // F1: j = a + Na (q + RB r): static over q in [0,RB), legs r in [0,RA); phase (beta/L) e^{+i pi j/L}
  tau_factors(beta, stat, L, Na, P1_RB, P1_RA, +1, beta / (double)L, b, C, f1);
  // I2: j = beta + Nb (s + RA t): static over s in [0,RA), legs t in [0,RB); phase e^{-i pi j/L} (item part goes into pass 1)
  tau_factors(beta, stat, L, Nb, P2_RA, P2_RB, -1, 1.0, b, C, f2);
  {
    ModeTables &m = pl->m[M_F1];
    TQ_TRY(pp_upload(ctx, *pl, twiddle_rows(P1_N, P1_RA, P1_RB, +1, &f1.xph, nullptr), &m.twq));
    TQ_TRY(pp_upload(ctx, *pl, f1.xtab, &m.qtab));
    TQ_TRY(pp_upload(ctx, *pl, f1.itemtab, &m.itemtab));
    memcpy(m.er, f1.er, sizeof(m.er));
    memcpy(m.phr, f1.phr, sizeof(m.phr));
  }
  {
    ModeTables &m = pl->m[M_I2];
    TQ_TRY(pp_upload(ctx, *pl, twiddle_rows(P2_N, P2_RA, P2_RB, -1, nullptr, &f2.xph), &m.twq));
    TQ_TRY(pp_upload(ctx, *pl, f2.xtab, &m.qtab));
    TQ_TRY(pp_upload(ctx, *pl, f2.itemtab, &m.itemtab));
    memcpy(m.er, f2.er, sizeof(m.er));
    memcpy(m.phr, f2.phr, sizeof(m.phr));
  }
  TQ_TRY(pp_upload(ctx, *pl, twiddle_rows(P2_N, P2_RA, P2_RB, +1, nullptr, nullptr), &pl->m[M_F2].twq));
  TQ_TRY(pp_upload(ctx, *pl, twiddle_rows(P1_N, P1_RA, P1_RB, -1, nullptr, nullptr), &pl->m[M_I1].twq));
  {
    // four-step twiddles x per-item phase:  F1: (beta/L) e^{i pi a/L} w_L^{a beta};  I1: e^{-i pi beta/L} w_L^{-a beta}
    std::vector<cd> tf((size_t)Na * Nb), ti((size_t)Na * Nb);
    for (int a = 0; a < Na; ++a)
      for (int be = 0; be < Nb; ++be) {
        const cd w = expi(2 * PI_L * (long double)((long)a * be) / (long double)L);
        tf[(size_t)a * Nb + be] = cmul(w, f1.itemph[a]);
        ti[(size_t)a * Nb + be] = cmul(cconj(w), f2.itemph[be]);
      }
    TQ_TRY(pp_upload(ctx, *pl, tf, &pl->m[M_F1].Tfull));
    TQ_TRY(pp_upload(ctx, *pl, ti, &pl->m[M_I1].Tfull));
  }
  {
    // window-row factors, permuted per item into the compact order the kernels use
    auto wrow = [&](int di) {
      long n = pl->first + di;
      double w = M_PI * (double)(2 * n + stat) / beta; // matsubara.hpp:55   (stat: Fermion = 1, Boson = 0)
      double r1 = 1.0 / (b[0] * b[0] + w * w), r2 = 1.0 / (b[1] * b[1] + w * w);
      return make_double4(b[0] * r1, w * r1, b[1] * r2, w * r2);
    };
    auto perm = [&](int n_items, int S, int N, int wmax, std::vector<double4> &out) -> bool {
      out.assign((size_t)n_items * wmax, make_double4(0, 0, 0, 0));
      for (int item = 0; item < n_items; ++item) {
        Window win = tile_window(item, S, N, (int)L, pl->first, pl->last);
        int cnt = win.r1 + (N - win.r2);
        if (cnt > wmax) return false;
        for (int i = 0; i < cnt; ++i) {
          int r = i < win.r1 ? i : win.r2 + (i - win.r1);
          int k = item + S * r;
          int di = (k <= pl->last ? k : k - (int)L) - pl->first;
          out[(size_t)item * wmax + i] = wrow(di);
        }
      }
      return true;
    };
    std::vector<double4> w1, w2;
    if (!perm(Na, Na, Nb, pl->wmax1, w1) || !perm(Nb, Nb, Na, pl->wmax2, w2))
      return tq_fail(ctx, TQ_ERR_INVALID, "internal: window-row table capacity");
    TQ_TRY(pp_upload(ctx, *pl, w1, &pl->m[M_I1].wperm));
    TQ_TRY(pp_upload(ctx, *pl, w2, &pl->m[M_F2].wperm));
  }
  plans[key] = pl;
  *out = pl;
  return TQ_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// G(tau) of `nb` gfs as the 4-D tensor [gf][b][a][2 n_cols doubles] (row j = a + Na b); box = one F1 tile [Nb][1][2 tcw]
static tq_status make_f1_tensor_map(tq_ctx *ctx, const cd *in, long nb, int Na, int Nb, long n_cols, long L, int tcw, CUtensorMap *out) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {(cuuint64_t)(2 * n_cols), (cuuint64_t)Na, (cuuint64_t)Nb, (cuuint64_t)nb};
  cuuint64_t strides[3] = {(cuuint64_t)(n_cols * sizeof(cd)), (cuuint64_t)((long)Na * n_cols * sizeof(cd)), (cuuint64_t)((L + 1) * n_cols * sizeof(cd))};
  cuuint32_t box[4] = {(cuuint32_t)(2 * tcw), 1u, (cuuint32_t)Nb, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return TQ_OK;
}

template <int MODE, int N, int RA, int RB, int WA, int WB>
static tq_status launch_pipe(tq_ctx *ctx, const PipeArgs &A, const CUtensorMap &tmap, const char *tag) {
  constexpr int NT = (WA + WB + 1) * 32;
  const size_t smem = PipeSmem<MODE, N, RA, RB>(A.tcw, A.wmax).total;
  auto k = matsubara_pipe_kernel<MODE, N, RA, RB, WA, WB>;
  TQ_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<long>(A.n_tiles, (long)ctx->sm_count);
  static const bool phase_debug = getenv("TQ_PHASE_DEBUG") != nullptr;
  PipeArgs Ad = A;
  if (phase_debug) {
    TQ_CUDA(ctx, cudaMallocManaged(&Ad.dbg, sizeof(long long) * (64 * grid + 128)));
    TQ_CUDA(ctx, cudaMemset(Ad.dbg, 0, sizeof(long long) * (64 * grid + 128)));
  }
  {
    KernelScope ks(ctx, tag);
    k<<<grid, NT, smem, ctx->stream>>>(Ad, tmap);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  if (phase_debug) {
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // per role: average over CTAs and warps of (cycles waiting, cycles working) per tile
    double w[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int i = 0; i < grid; ++i)
      for (int wp = 0; wp < WA + WB + 1; ++wp) {
        const int role = wp < WA ? 0 : (wp < WA + WB ? 1 : 2);
        const int nw = role == 0 ? WA : (role == 1 ? WB : 1);
        w[role][0] += (double)Ad.dbg[(size_t)i * 64 + 2 * wp] / nw;
        w[role][1] += (double)Ad.dbg[(size_t)i * 64 + 2 * wp + 1] / nw;
      }
    const double per = (double)A.n_tiles;
    fprintf(stderr,
            "[pipe] %-13s N=%d tcw=%d tiles=%ld grid=%d smem=%zu thr=%d  cycles/tile  A: wait %.0f work %.0f | B: wait %.0f work %.0f | P: wait %.0f work "
            "%.0f\n",
            tag, N, A.tcw, A.n_tiles, grid, smem, NT, w[0][0] / per, w[0][1] / per, w[1][0] / per, w[1][1] / per, w[2][0] / per, w[2][1] / per);
    if (getenv("TQ_PHASE_TRACE")) {
      const long long *tr = Ad.dbg + (size_t)grid * 64;
      const long long t00 = tr[(2 * 16 + 0) * 2 + 0];
      fprintf(stderr, "   tile:   P got-slot  P issued |  A got-data  A done |  B got-A  B done     (cycles since the producer's first stamp, CTA 0)\n");
      for (int i = 0; i < 16; ++i)
        fprintf(stderr, "   %4d: %10lld %10lld | %10lld %8lld | %8lld %8lld\n", i, tr[(2 * 16 + i) * 2] - t00, tr[(2 * 16 + i) * 2 + 1] - t00,
                tr[(0 * 16 + i) * 2] - t00, tr[(0 * 16 + i) * 2 + 1] - t00, tr[(1 * 16 + i) * 2] - t00, tr[(1 * 16 + i) * 2 + 1] - t00);
    }
    cudaFree(Ad.dbg);
  }
  return TQ_OK;
}

// widest balanced column tile for a tile length N: n_ct tiles of width ceil(n_cols / n_ct)
template <int MODE, int N, int RA, int RB> static int pipe_pick_tcw(tq_ctx *ctx, long n_cols, int wmax, int cap) {
  int fit = 0;
  for (int tc = 1; tc <= n_cols && tc <= cap; ++tc)
    if (PipeSmem<MODE, N, RA, RB>(tc, wmax).total <= (size_t)ctx->max_smem_optin) fit = tc;
  if (fit < 1) return 0;
  long n_ct = (n_cols + fit - 1) / fit;
  return (int)((n_cols + n_ct - 1) / n_ct);
}

// warps of the A / B groups (+ 1 producer warp).  The register file is 16K registers per SM sub-partition, so 12 warps
// (3 per sub-partition) can have 168 registers each, 16 warps only 128.
constexpr int P1_WA = 6, P1_WB = 5;
static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Runs the transform if a pipelined specialisation exists for this mesh.  *handled tells the caller.
tq_status tq_pipe_matsubara(tq_ctx *ctx, bool fwd, double beta, int stat, long n_tau, long n_iw, long n_cols, long batch, const cd *d_in,
                            cd *d_out, const cd *d_mom, int mom_rows, bool *handled) {
  *handled = false;
  const long L = n_tau - 1;
  int Na, Nb;
  if (!pipe_supported_length(L, &Na, &Nb)) return TQ_OK;
  if (getenv("TQ_NO_PIPE")) return TQ_OK;
  if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return TQ_OK; // bulk copies need 16-byte aligned rows
  std::shared_ptr<PipePlan> pl;
  TQ_TRY(get_pipe_plan(ctx, beta, stat, n_tau, n_iw, &pl));
  const int wmax1 = pl->wmax1, wmax2 = pl->wmax2;
  static const int p2_variant = env_int("TQ_PIPE_P2", 0), cap1 = env_int("TQ_PIPE_TCW1", 64), cap2 = env_int("TQ_PIPE_TCW2", 64);
  const int tcw1 = fwd ? pipe_pick_tcw<M_F1, P1_N, P1_RA, P1_RB>(ctx, n_cols, wmax1, cap1) : pipe_pick_tcw<M_I1, P1_N, P1_RA, P1_RB>(ctx, n_cols, wmax1, cap1);
  const int tcw2 = fwd ? pipe_pick_tcw<M_F2, P2_N, P2_RA, P2_RB>(ctx, n_cols, wmax2, cap2) : pipe_pick_tcw<M_I2, P2_N, P2_RA, P2_RB>(ctx, n_cols, wmax2, cap2);
  if (tcw1 < 1 || tcw2 < 1) return TQ_OK;
  PipeArgs A{};
  A.Na = Na;
  A.Nb = Nb;
  A.n_cols = (int)n_cols;
  A.tcw2 = tcw2;
  A.moments = d_mom;
  A.mom_gf_stride = (long)mom_rows * n_cols;
  A.L = (int)L;
  A.first = pl->first;
  A.last = pl->last;
  A.stat = stat;
  A.half_fact_fwd = 0.5 * (beta / (double)L);
  A.inv_beta = 1.0 / beta;
  const long in_rows = fwd ? L + 1 : pl->n_w, out_rows = fwd ? pl->n_w : L + 1;
  A.in_gf_stride = in_rows * n_cols;
  A.out_gf_stride = out_rows * n_cols;
  A.scratch_gf_stride = L * n_cols;
  // chunk the batch so that the scratch of one chunk stays L2 resident between the two kernels
  const size_t per_gf = (size_t)L * n_cols * sizeof(cd);
  size_t l2_budget = 96ull << 20;
  if (const char *e = getenv("TQ_TWOPASS_CHUNK_MB")) l2_budget = (size_t)atol(e) << 20;
  long chunk = std::max<long>(1, (long)(l2_budget / per_gf));
  chunk = std::min<long>(chunk, batch);
  TQ_TRY(tq_reserve(ctx, ctx->scratch, per_gf * chunk));
  A.scratch = (cd *)ctx->scratch.p;
  auto set_mode = [&](PipeArgs &P, int mode) {
    const ModeTables &m = pl->m[mode];
    P.twq = m.twq;
    P.qtab = m.qtab;
    P.itemtab = m.itemtab;
    P.Tfull = m.Tfull;
    P.wperm = m.wperm;
    memcpy(P.er, m.er, sizeof(P.er));
    memcpy(P.phr, m.phr, sizeof(P.phr));
  };
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  for (long b0 = 0; b0 < batch; b0 += chunk) {
    const long nb = std::min(chunk, batch - b0);
    PipeArgs P1 = A;
    P1.in = d_in + b0 * A.in_gf_stride;
    P1.out = d_out + b0 * A.out_gf_stride;
    P1.moments = d_mom + b0 * A.mom_gf_stride;
    P1.tcw = tcw1;
    P1.n_ct = (int)((n_cols + tcw1 - 1) / tcw1);
    P1.n_items = Na;
    P1.n_tiles = nb * Na * P1.n_ct;
    P1.wmax = wmax1;
    PipeArgs P2 = P1;
    P2.tcw = tcw2;
    P2.n_ct = (int)((n_cols + tcw2 - 1) / tcw2);
    P2.n_items = Nb;
    P2.n_tiles = nb * Nb * P2.n_ct;
    P2.wmax = wmax2;
    if (fwd) {
      set_mode(P1, M_F1);
      set_mode(P2, M_F2);
      TQ_TRY(make_f1_tensor_map(ctx, P1.in, nb, Na, Nb, n_cols, L, tcw1, &tmap));
      TQ_TRY((launch_pipe<M_F1, P1_N, P1_RA, P1_RB, P1_WA, P1_WB>(ctx, P1, tmap, "tau2iw_pass1")));
      if (p2_variant == 0)
        TQ_TRY((launch_pipe<M_F2, P2_N, P2_RA, P2_RB, 6, 5>(ctx, P2, tmap, "tau2iw_pass2")));
      else if (p2_variant == 1)
        TQ_TRY((launch_pipe<M_F2, P2_N, P2_RA, P2_RB, 5, 5>(ctx, P2, tmap, "tau2iw_pass2")));
      else
        TQ_TRY((launch_pipe<M_F2, P2_N, P2_RA, P2_RB, 8, 7>(ctx, P2, tmap, "tau2iw_pass2")));
    } else {
      set_mode(P1, M_I1);
      set_mode(P2, M_I2);
      TQ_TRY((launch_pipe<M_I1, P1_N, P1_RA, P1_RB, P1_WA, P1_WB>(ctx, P1, tmap, "iw2tau_pass1")));
      if (p2_variant == 0)
        TQ_TRY((launch_pipe<M_I2, P2_N, P2_RA, P2_RB, 6, 5>(ctx, P2, tmap, "iw2tau_pass2")));
      else if (p2_variant == 1)
        TQ_TRY((launch_pipe<M_I2, P2_N, P2_RA, P2_RB, 5, 5>(ctx, P2, tmap, "iw2tau_pass2")));
      else
        TQ_TRY((launch_pipe<M_I2, P2_N, P2_RA, P2_RB, 8, 7>(ctx, P2, tmap, "iw2tau_pass2")));
    }
  }
  *handled = true;
  return TQ_OK;
}
