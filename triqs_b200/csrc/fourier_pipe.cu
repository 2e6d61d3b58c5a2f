// Pipelined four-step imtime <-> imfreq transforms for long time meshes (L = Na * Nb, e.g. 10^5 = 400 * 250).
//
// Same mathematics as the generic tile kernels of fourier_matsubara.cu (which restate
// c++/triqs/gfs/transform/fourier_matsubara.cpp:96-186 and :190-271), different machine mapping:
//   * one persistent CTA per SM loops over tiles; the NEXT tile is pulled into the second shared-memory buffer
//     by the TMA engine (cp.async.bulk + mbarrier complete_tx) while the CURRENT one is transformed, so HBM/L2
//     latency is off the critical path and no thread ever waits on a global load;
//   * the tile FFT is the compile-time two-pass engine of tq_fft2.cuh (250 = 10 x 25, 400 = 20 x 20): pass A
//     consumes the raw tile with the pre-transform fused (tail-model subtraction, e^{i pi tau/beta} twist),
//     pass B applies the epilogue in registers (four-step twiddle / window gather + tail / tau post-twist)
//     and stores straight to global memory;
//   * every per-row factor is FACTORISED into (static per-row table in shared memory) x (per-tile scalar):
//       j = item + S * row :  e^{i pi j / L} = e^{i pi item / L} e^{i pi S row / L},
//                             e^{-b tau_j}  = e^{-b delta item} e^{-b delta S row}            (b >= 0)
//                             e^{-|b| (beta - tau_j)} = e^{-|b| delta (S - item)} e^{-|b| delta S (rows-1-row)}  (b < 0)
//     so a tile's prologue is O(columns + rows) table look-ups, not O(rows) exponentials;
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
// PTX wrappers: mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP) + proxy fence
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
TQ_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// -------------------------------------------------------------------------------------------------
enum { M_F1 = 0, M_F2 = 1, M_I1 = 2, M_I2 = 3 };

// thread block = consumer warps (butterflies) + 1 producer warp (TMA issue and the next tile's tables)

struct BigTw2 {
  const cd *hi; // exp(+2 pi i (h << shift) / L)
  const cd *lo; // exp(+2 pi i l / L)
  int shift;
};

struct PipeArgs {
  int Na, Nb, n_cols;
  int tcw, n_ct;   // column tiling of this pass
  int tcw2;        // column tiling of pass 2 (defines the scratch layout)
  int n_items;     // Na (pass 1) or Nb (pass 2)
  int wmax;        // capacity of the compact window-row table (F2, I1)
  long n_tiles;    // batch * n_items * n_ct
  const cd *in;
  cd *out;
  cd *scratch;
  const cd *moments; // [batch][mom_rows][n_cols]; rows 1..3 are used
  long in_gf_stride, out_gf_stride, scratch_gf_stride, mom_gf_stride;
  int L, first, last, stat;
  double half_fact_fwd; // 0.5 beta / L
  double inv_beta;
  const cd *twN;          // [N] exp(2 pi i j / N), N = tile length of this pass
  const cd *rowtab;       // [N][3] {phase} {E1,E2} {E3,-}    (F1: row = b, I2: row = alpha)
  const double4 *itemtab; // [n_items] {C1 ea1, C2 ea2, C3 ea3, -}   (F1: item = a, I2: item = beta)
  const cd *itemph;       // [Na]  F1: (beta/L) e^{i pi a / L} (Fermion) or beta/L (Boson)
  const cd *outph;        // [Nb]  I1: e^{-i pi beta_idx / L}  (Fermion) or 1
  BigTw2 btw;
  const double4 *wtab; // [n_w] {b1 r1, w r1, b2 r2, w r2},  r_i = 1/(b_i^2 + w^2)   (b3 = -b2 for both statistics)
  long long *dbg;      // developer phase timing (TQ_PHASE_DEBUG): 8 accumulated clock64 spans per CTA, or nullptr
};

// Rows r of a tile (global frequency index k = item + S r) that fall inside the Matsubara window
// [first, last] (fourier_matsubara.cpp:182, :254 "(iw.n + L) % L"):  r in [0, r1)  (n = k >= 0)  and  r in [r2, N)  (n = k - L < 0).
struct Window {
  int r1, r2;
};
TQ_D Window tile_window(int item, int S, int N, int L, int first, int last) {
  Window w;
  w.r1 = last >= item ? min(N, (last - item) / S + 1) : 0;
  const int lo = L + first - item; // k >= L + first
  w.r2 = lo <= 0 ? 0 : min(N, (lo + S - 1) / S);
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
struct PreF1 { // x = phase[row] (g - sum_i at_i E_i[row])        fourier_matsubara.cpp:159-173
  const cd *rowtab, *coef;
  int TCW;
  cd a1, a2, a3;
  TQ_D void begin(int c) {
    a1 = coef[c];
    a2 = coef[TCW + c];
    a3 = coef[2 * TCW + c];
  }
  TQ_D cd load(int row, const cd *p) const {
    const cd g = *p;
    const cd ph = rowtab[3 * row], e12 = rowtab[3 * row + 1], e3 = rowtab[3 * row + 2];
    cd v = caxpy(-e12.x, a1, g);
    v = caxpy(-e12.y, a2, v);
    v = caxpy(-e3.x, a3, v);
    return cmul(v, ph);
  }
};
struct PreI1 { // x = (g - tail(i w)) / beta for window rows, 0 elsewhere      :254
  const cd *wrec, *coef;
  int TCW;
  Window win;
  double inv_beta;
  cd a1, D23, A23;
  TQ_D void begin(int c) {
    a1 = coef[c];
    D23 = coef[TCW + c];
    A23 = coef[2 * TCW + c];
  }
  TQ_D cd load(int row, const cd *p) const {
    int idx;
    if (row < win.r1)
      idx = row;
    else if (row >= win.r2)
      idx = win.r1 + row - win.r2;
    else
      return cmk(0.0, 0.0);
    const cd g = *p;
    const cd t = tail_iw4(wrec[2 * idx], wrec[2 * idx + 1], a1, D23, A23);
    return cscale(inv_beta, csub(g, t));
  }
};
struct PostScratch { // Y = z T[beta]  -> scratch[beta][column tile][a][column]
  const cd *T;
  cd *base; // scratch of this gf
  long beta_stride; // Na * n_cols
  int n_cols, Na, tcw2, item, c0;
  cd *p;
  TQ_D void begin(int c) {
    const int cg = c0 + c;
    const int h = cg / tcw2;
    const int cc = cg - h * tcw2;
    const int w = min(tcw2, n_cols - h * tcw2);
    p = base + ((long)h * tcw2 * Na + (long)item * w + cc);
  }
  TQ_D void store(int k, cd v) const { p[(long)k * beta_stride] = cmul(v, T[k]); }
};
struct PostF2 { // G(i w_n) = z + corr + tail, window rows only      :181-183
  const cd *wrec, *coef;
  int TCW, n_cols, item, Nb, L, first;
  Window win;
  cd *base; // out of this gf + c0
  cd a1, D23, A23, ex;
  cd *p;
  TQ_D void begin(int c) {
    a1 = coef[c];
    D23 = coef[TCW + c];
    A23 = coef[2 * TCW + c];
    ex = coef[3 * TCW + c];
    p = base + c;
  }
  TQ_D void store(int k, cd v) const {
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
struct PostI2 { // G(tau_j) = z e^{-i pi j/L} + sum_i at_i E_i,  j = item + Nb alpha        :261-268
  const cd *rowtab, *coef;
  int TCW, n_cols, item, Nb, L;
  bool fermion;
  cd *base; // out of this gf + c0
  cd a1, a2, a3, m1;
  cd *p;
  TQ_D void begin(int c) {
    a1 = coef[c];
    a2 = coef[TCW + c];
    a3 = coef[2 * TCW + c];
    m1 = coef[3 * TCW + c];
    p = base + c;
  }
  TQ_D void store(int k, cd z) const {
    const cd ph = rowtab[3 * k], e12 = rowtab[3 * k + 1], e3 = rowtab[3 * k + 2];
    cd v = cmul(z, ph);
    v = caxpy(e12.x, a1, v);
    v = caxpy(e12.y, a2, v);
    v = caxpy(e3.x, a3, v);
    const int j = item + Nb * k;
    p[(long)j * n_cols] = v;
    if (j == 0) { // gt[L] = -/+ (gt[0] + m1)     :267-268
      const cd t = cadd(v, m1);
      p[(long)L * n_cols] = fermion ? cneg(t) : t;
    }
  }
};

// shared-memory carve-up (bytes); tile buffers first, each padded to 128 B (TMA tensor destination alignment)
template <int MODE, int N> struct PipeSmem {
  static constexpr bool PASS1 = (MODE == M_F1 || MODE == M_I1);
  static constexpr bool ROWTAB = (MODE == M_F1 || MODE == M_I2);
  static constexpr bool WREC = (MODE == M_F2 || MODE == M_I1);
  size_t buf_stride, tw, rowtab, tab0, tab_stride, coef_off, t_off, wrec_off, mbar, total;
  __host__ __device__ PipeSmem(int tcw, int wmax) {
    buf_stride = (sizeof(cd) * (size_t)N * tcw + 127) & ~size_t(127);
    tw = 2 * buf_stride;
    rowtab = tw + sizeof(cd) * N;
    tab0 = rowtab + (ROWTAB ? sizeof(cd) * 3 * N : 0);
    coef_off = 0;
    t_off = coef_off + sizeof(cd) * 4 * (size_t)tcw;
    wrec_off = t_off + (PASS1 ? sizeof(cd) * N : 0);
    tab_stride = wrec_off + (WREC ? sizeof(cd) * 2 * (size_t)wmax : 0);
    mbar = tab0 + 2 * tab_stride;
    total = mbar + 16;
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

TQ_D void bar_sync_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
TQ_D void tma_load_4d(void *dst_smem, const CUtensorMap *tmap, int c0, int c1, int c2, int c3, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                  smem_u32(dst_smem)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
               : "memory");
}

template <int MODE, int N, int RA, int RB, int PIPE_NT, int MINB>
__global__ void __launch_bounds__(PIPE_NT, MINB) matsubara_pipe_kernel(const PipeArgs A, const __grid_constant__ CUtensorMap tmap) {
  constexpr int PIPE_NC = PIPE_NT - 32;
  constexpr bool PASS1 = (MODE == M_F1 || MODE == M_I1);
  constexpr bool FWD = (MODE == M_F1 || MODE == M_F2);
  constexpr int SGN = FWD ? 1 : -1;
  constexpr int NC = PIPE_NC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int TCW = A.tcw;
  const PipeSmem<MODE, N> S(TCW, A.wmax);
  cd *twS = reinterpret_cast<cd *>(smem_raw + S.tw);
  cd *rowtab = reinterpret_cast<cd *>(smem_raw + S.rowtab);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + S.mbar);
  const int tid = threadIdx.x;
  const int n_cols = A.n_cols;
  const bool fermion = A.stat == TQ_FERMION;
  const bool producer = tid >= NC;
  const int lane = tid & 31;

  const long long tstart = clock64();
  for (int i = tid; i < N; i += PIPE_NT) twS[i] = A.twN[i];
  if (S.ROWTAB)
    for (int i = tid; i < 3 * N; i += PIPE_NT) rowtab[i] = A.rowtab[i];
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  // ---- producer side: ask the TMA engine for tile t -> buffer b  (whole warp calls; one elected lane issues)
  auto issue = [&](long t, int b) {
    const TileId id = decode_tile(t, A.n_items, A.n_ct, TCW, n_cols);
    cd *dst = reinterpret_cast<cd *>(smem_raw + b * S.buf_stride);
    if (MODE == M_F1) {
      // one 4-D tensor tile: [gf][b = 0..Nb)[a = item][2 c0 .. 2 (c0 + TCW)) doubles; columns past n_cols are zero-filled
      if (lane == 0) {
        mbar_arrive_expect_tx(&mbar[b], (uint32_t)(sizeof(cd) * N * TCW));
        tma_load_4d(dst, &tmap, 2 * id.c0, id.item, 0, (int)id.gf, &mbar[b]);
      }
    } else if (MODE == M_I1) {
      // window rows only (the rest of the frequency axis is zero, :254): one bulk row copy each, tile pitch = w
      const Window win = tile_window(id.item, A.Na, N, A.L, A.first, A.last);
      const int cnt = win.r1 + (N - win.r2);
      if (lane == 0) mbar_arrive_expect_tx(&mbar[b], (uint32_t)(sizeof(cd) * cnt * id.w));
      __syncwarp();
      for (int i = lane; i < cnt; i += 32) {
        const int r = i < win.r1 ? i : win.r2 + (i - win.r1);
        const int k = id.item + A.Na * r;
        const long di = (k <= A.last ? k : k - A.L) - A.first;
        bulk_g2s(dst + (size_t)r * id.w, A.in + id.gf * A.in_gf_stride + di * n_cols + id.c0, (uint32_t)(sizeof(cd) * id.w), &mbar[b]);
      }
    } else if (lane == 0) {
      const uint32_t bytes = (uint32_t)(sizeof(cd) * N * id.w);
      mbar_arrive_expect_tx(&mbar[b], bytes);
      bulk_g2s(dst, A.scratch + id.gf * A.scratch_gf_stride + (long)id.item * A.Na * n_cols + (long)id.c0 * A.Na, bytes, &mbar[b]);
    }
  };
  // ---- producer side: per-tile tables -> table set b  (pole weights :151-169 / :238-252, window rows, four-step twiddles)
  auto build_tables = [&](long t, int b) {
    const TileId id = decode_tile(t, A.n_items, A.n_ct, TCW, n_cols);
    unsigned char *tab = smem_raw + S.tab0 + b * S.tab_stride;
    cd *coef = reinterpret_cast<cd *>(tab + S.coef_off);
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
    if (S.WREC) {
      cd *wrec = reinterpret_cast<cd *>(tab + S.wrec_off);
      const int stride = (MODE == M_F2) ? A.Nb : A.Na;
      const Window win = tile_window(id.item, stride, N, A.L, A.first, A.last);
      const int cnt = win.r1 + (N - win.r2);
#pragma unroll 4
      for (int i = lane; i < cnt; i += 32) {
        const int r = i < win.r1 ? i : win.r2 + (i - win.r1);
        const int k = id.item + stride * r;
        const int di = (k <= A.last ? k : k - A.L) - A.first;
        const double4 wv = A.wtab[di];
        wrec[2 * i] = cmk(wv.x, wv.y);
        wrec[2 * i + 1] = cmk(wv.z, wv.w);
      }
    }
    if (PASS1) {
      cd *Ttab = reinterpret_cast<cd *>(tab + S.t_off);
      const cd iph = (MODE == M_F1) ? A.itemph[id.item] : cmk(1.0, 0.0);
#pragma unroll 4
      for (int r = lane; r < N; r += 32) {
        const int m = id.item * r; // < L
        cd wv = cmul(A.btw.hi[m >> A.btw.shift], A.btw.lo[m & ((1 << A.btw.shift) - 1)]);
        if (MODE == M_F1)
          wv = cmul(wv, iph);
        else
          wv = cmul(cconj(wv), A.outph[r]);
        Ttab[r] = wv;
      }
    }
  };

  long long ph[4] = {0, 0, 0, 0}, tprev = clock64();
#define TQ_STAMP(i)                                                                                                              \
  if (A.dbg && tid == 0) {                                                                                                       \
    long long now = clock64();                                                                                                   \
    ph[i] += now - tprev;                                                                                                        \
    tprev = now;                                                                                                                 \
  }
  long t = blockIdx.x;
  if (A.dbg && tid == 0) A.dbg[8 * blockIdx.x + 4] = clock64() - tstart;
  if (producer && t < A.n_tiles) {
    issue(t, 0);
    build_tables(t, 0);
  }
  __syncthreads();
  if (A.dbg && tid == 0) A.dbg[8 * blockIdx.x + 5] = clock64() - tstart;
  tprev = clock64();
  long long prod_cycles = 0;
  for (int it = 0; t < A.n_tiles; t += gridDim.x, ++it) {
    const int cur = it & 1;
    if (producer) {
      // buffer cur^1 and table set cur^1 were released by the barrier that ended the previous tile
      const long tn = t + gridDim.x;
      const long long p0 = clock64();
      if (tn < A.n_tiles) {
        issue(tn, cur ^ 1);
        build_tables(tn, cur ^ 1);
      }
      prod_cycles += clock64() - p0;
    } else {
      const TileId id = decode_tile(t, A.n_items, A.n_ct, TCW, n_cols);
      unsigned char *tab = smem_raw + S.tab0 + cur * S.tab_stride;
      const cd *coef = reinterpret_cast<const cd *>(tab + S.coef_off);
      const cd *Ttab = reinterpret_cast<const cd *>(tab + S.t_off);
      const cd *wrec = reinterpret_cast<const cd *>(tab + S.wrec_off);
      cd *tile = reinterpret_cast<cd *>(smem_raw + cur * S.buf_stride);
      const int pitch = (MODE == M_F1) ? TCW : id.w;
      TQ_STAMP(0)
      mbar_wait(&mbar[cur], (uint32_t)((it >> 1) & 1));
      TQ_STAMP(1)
      if (MODE == M_F1) {
        PreF1 pre{rowtab, coef, TCW};
        fft2_pass_a<N, RA, RB, SGN>(tile, pitch, id.w, twS, tid, NC, pre);
      } else if (MODE == M_I1) {
        PreI1 pre{wrec, coef, TCW, tile_window(id.item, A.Na, N, A.L, A.first, A.last), A.inv_beta};
        fft2_pass_a<N, RA, RB, SGN>(tile, pitch, id.w, twS, tid, NC, pre);
      } else {
        PreIdentity pre;
        fft2_pass_a<N, RA, RB, SGN>(tile, pitch, id.w, twS, tid, NC, pre);
      }
      bar_sync_named(1, NC); // consumers only: pass A complete
      TQ_STAMP(2)
      if (PASS1) {
        PostScratch post{Ttab, A.scratch + id.gf * A.scratch_gf_stride, (long)A.Na * n_cols, n_cols, A.Na, A.tcw2, id.item, id.c0};
        fft2_pass_b<N, RA, RB, SGN>(tile, pitch, id.w, tid, NC, post);
      } else if (MODE == M_F2) {
        PostF2 post{wrec, coef, TCW, n_cols, id.item, A.Nb, A.L, A.first, tile_window(id.item, A.Nb, N, A.L, A.first, A.last),
                    A.out + id.gf * A.out_gf_stride + id.c0};
        fft2_pass_b<N, RA, RB, SGN>(tile, pitch, id.w, tid, NC, post);
      } else {
        PostI2 post{rowtab, coef, TCW, n_cols, id.item, A.Nb, A.L, fermion, A.out + id.gf * A.out_gf_stride + id.c0};
        fft2_pass_b<N, RA, RB, SGN>(tile, pitch, id.w, tid, NC, post);
      }
      fence_proxy_async();
    }
    __syncthreads(); // releases the tile buffer and the table set
    TQ_STAMP(3)
  }
  if (A.dbg && tid == 0)
    for (int i = 0; i < 4; ++i) A.dbg[8 * blockIdx.x + i] = ph[i];
  if (A.dbg && tid == NC) A.dbg[8 * blockIdx.x + 6] = prod_cycles;
#undef TQ_STAMP
}

// -------------------------------------------------------------------------------------------------
// host
// -------------------------------------------------------------------------------------------------
struct PipePlan {
  int L = 0, Na = 0, Nb = 0, first = 0, last = 0, n_w = 0, stat = 0;
  double beta = 0;
  const cd *twNa = nullptr, *twNb = nullptr;
  const cd *rowtab_f1 = nullptr, *rowtab_i2 = nullptr;
  const double4 *itemtab_f1 = nullptr, *itemtab_i2 = nullptr;
  const cd *itemph = nullptr, *outph = nullptr;
  BigTw2 btw{};
  const double4 *wtab = nullptr;
  std::vector<void *> allocs;
  ~PipePlan() {
    for (void *p : allocs) cudaFree(p);
  }
};

template <class T> static tq_status pp_upload(tq_ctx *ctx, PipePlan &pl, const std::vector<T> &h, const T **d) {
  void *p = nullptr;
  TQ_CUDA(ctx, cudaMalloc(&p, sizeof(T) * std::max<size_t>(h.size(), 1)));
  pl.allocs.push_back(p);
  TQ_CUDA(ctx, cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
  *d = (const T *)p;
  return TQ_OK;
}

// tau-side factor tables for  j = item + S * row,  item in [0, S),  row in [0, R)
//   rowtab[row]  = { e^{sgn i pi S row / L} (Fermion; 1 for Boson) } { E1, E2 } { E3, 0 }
//   itemtab[item] = { C_i e_i(item) },   itemph[item] = scale * e^{sgn i pi item / L}
static void tau_tables(double beta, int stat, long L, int S, int R, int sgn, double scale, const double b[3], const double C[3],
                       std::vector<cd> &rowtab, std::vector<double4> &itemtab, std::vector<cd> &itemph) {
  const long double pi = 3.141592653589793238462643383279502884L;
  const long double delta = (long double)beta / (long double)L;
  rowtab.assign((size_t)3 * R, cmk(0, 0));
  itemtab.assign(S, make_double4(0, 0, 0, 0));
  itemph.assign(S, cmk(0, 0));
  for (int r = 0; r < R; ++r) {
    long double ang = pi * (long double)((long)S * r) / (long double)L;
    rowtab[3 * r] = stat == TQ_FERMION ? cmk((double)cosl(ang), (double)(sgn * sinl(ang))) : cmk(1.0, 0.0);
    double E[3];
    for (int i = 0; i < 3; ++i) {
      long double ab = fabsl((long double)b[i]);
      long idx = b[i] >= 0 ? (long)S * r : (long)S * (R - 1 - r);
      E[i] = (double)expl(-ab * delta * (long double)idx);
    }
    rowtab[3 * r + 1] = cmk(E[0], E[1]);
    rowtab[3 * r + 2] = cmk(E[2], 0.0);
  }
  for (int a = 0; a < S; ++a) {
    long double ang = pi * (long double)a / (long double)L;
    itemph[a] = stat == TQ_FERMION ? cmk((double)((long double)scale * cosl(ang)), (double)((long double)scale * sgn * sinl(ang))) : cmk(scale, 0.0);
    double e[3];
    for (int i = 0; i < 3; ++i) {
      long double ab = fabsl((long double)b[i]);
      long idx = b[i] >= 0 ? a : (S - a);
      e[i] = (double)((long double)C[i] * expl(-ab * delta * (long double)idx));
    }
    itemtab[a] = make_double4(e[0], e[1], e[2], 0.0);
  }
}

struct PipePlanCache {
  std::map<std::tuple<double, int, long, long>, std::shared_ptr<PipePlan>> plans;
};

static bool pipe_supported_length(long L, int *Na, int *Nb) {
  if (L == 100000) {
    *Na = 400;
    *Nb = 250;
    return true;
  }
  return false;
}

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
  std::vector<cd> twa, twb;
  tq_fft_twiddles_host(Na, twa);
  tq_fft_twiddles_host(Nb, twb);
  TQ_TRY(pp_upload(ctx, *pl, twa, &pl->twNa));
  TQ_TRY(pp_upload(ctx, *pl, twb, &pl->twNb));
  {
    std::vector<cd> rt, iph;
    std::vector<double4> itab;
    // F1: j = a + Na b, rows b in [0, Nb), phase (beta/L) e^{+i pi j/L}
    tau_tables(beta, stat, L, Na, Nb, +1, beta / (double)L, b, C, rt, itab, iph);
    TQ_TRY(pp_upload(ctx, *pl, rt, &pl->rowtab_f1));
    TQ_TRY(pp_upload(ctx, *pl, itab, &pl->itemtab_f1));
    TQ_TRY(pp_upload(ctx, *pl, iph, &pl->itemph));
    // I2: j = beta_idx + Nb alpha, rows alpha in [0, Na), phase e^{-i pi j/L}; its per-item phase goes into pass 1 (outph)
    tau_tables(beta, stat, L, Nb, Na, -1, 1.0, b, C, rt, itab, iph);
    TQ_TRY(pp_upload(ctx, *pl, rt, &pl->rowtab_i2));
    TQ_TRY(pp_upload(ctx, *pl, itab, &pl->itemtab_i2));
    TQ_TRY(pp_upload(ctx, *pl, iph, &pl->outph));
  }
  {
    std::vector<double4> wtab(pl->n_w);
    for (int i = 0; i < pl->n_w; ++i) {
      long n = pl->first + i;
      double w = M_PI * (double)(2 * n + stat) / beta; // matsubara.hpp:55   (stat: Fermion = 1, Boson = 0)
      double r1 = 1.0 / (b[0] * b[0] + w * w), r2 = 1.0 / (b[1] * b[1] + w * w);
      wtab[i] = make_double4(b[0] * r1, w * r1, b[1] * r2, w * r2);
    }
    TQ_TRY(pp_upload(ctx, *pl, wtab, &pl->wtab));
  }
  {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    int shift = 9;
    int nlo = 1 << shift, nhi = (int)(L >> shift) + 1;
    std::vector<cd> hi(nhi), lo(nlo);
    for (int h = 0; h < nhi; ++h) {
      long double ang = two_pi * (long double)((long)h << shift) / (long double)L;
      hi[h] = cmk((double)cosl(ang), (double)sinl(ang));
    }
    for (int l = 0; l < nlo; ++l) {
      long double ang = two_pi * (long double)l / (long double)L;
      lo[l] = cmk((double)cosl(ang), (double)sinl(ang));
    }
    pl->btw.shift = shift;
    TQ_TRY(pp_upload(ctx, *pl, hi, &pl->btw.hi));
    TQ_TRY(pp_upload(ctx, *pl, lo, &pl->btw.lo));
  }
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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

static int pipe_variant() {
  static int v = [] {
    const char *e = getenv("TQ_PIPE_VARIANT");
    return e ? atoi(e) : 0;
  }();
  return v;
}
static size_t pipe_smem_budget(tq_ctx *ctx) {
  return pipe_variant() == 1 ? ((size_t)ctx->max_smem_optin + 1024) / 2 - 1024 - 512 : (size_t)ctx->max_smem_optin;
}

template <int MODE, int N, int RA, int RB, int PIPE_NT, int MINB>
static tq_status launch_pipe_v(tq_ctx *ctx, const PipeArgs &A, const CUtensorMap &tmap, const char *tag);
template <int MODE, int N, int RA, int RB> static tq_status launch_pipe(tq_ctx *ctx, const PipeArgs &A, const CUtensorMap &tmap, const char *tag) {
  if (pipe_variant() == 1) return launch_pipe_v<MODE, N, RA, RB, 256, 2>(ctx, A, tmap, tag);
  return launch_pipe_v<MODE, N, RA, RB, 352, 1>(ctx, A, tmap, tag);
}
template <int MODE, int N, int RA, int RB, int PIPE_NT, int MINB>
static tq_status launch_pipe_v(tq_ctx *ctx, const PipeArgs &A, const CUtensorMap &tmap, const char *tag) {
  const size_t smem = PipeSmem<MODE, N>(A.tcw, A.wmax).total;
  auto k = matsubara_pipe_kernel<MODE, N, RA, RB, PIPE_NT, MINB>;
  TQ_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<long>(A.n_tiles, (long)ctx->sm_count * MINB);
  static const bool phase_debug = getenv("TQ_PHASE_DEBUG") != nullptr;
  PipeArgs Ad = A;
  if (phase_debug) {
    TQ_CUDA(ctx, cudaMallocManaged(&Ad.dbg, sizeof(long long) * 8 * grid));
    TQ_CUDA(ctx, cudaMemset(Ad.dbg, 0, sizeof(long long) * 8 * grid));
  }
  {
    KernelScope ks(ctx, tag);
    k<<<grid, PIPE_NT, smem, ctx->stream>>>(Ad, tmap);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  if (phase_debug) {
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double ph[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < 7; ++j) ph[j] += (double)Ad.dbg[8 * i + j];
    const double per_cta_tiles = (double)A.n_tiles / grid;
    fprintf(stderr,
            "[pipe] %-13s N=%d tcw=%d tiles=%ld grid=%d smem=%zu  cycles/tile: sync %.0f wait %.0f passA %.0f passB %.0f  (sum %.0f)  producer/tile %.0f | "
            "prologue/CTA: static tables %.0f, +first tile tables %.0f\n",
            tag, N, A.tcw, A.n_tiles, grid, smem, ph[0] / grid / per_cta_tiles, ph[1] / grid / per_cta_tiles, ph[2] / grid / per_cta_tiles,
            ph[3] / grid / per_cta_tiles, (ph[0] + ph[1] + ph[2] + ph[3]) / grid / per_cta_tiles, ph[6] / grid / per_cta_tiles, ph[4] / grid,
            ph[5] / grid);
    cudaFree(Ad.dbg);
  }
  return TQ_OK;
}

// widest balanced column tile for a tile length N: n_ct tiles of width ceil(n_cols / n_ct)
template <int MODE, int N> static int pipe_pick_tcw(tq_ctx *ctx, long n_cols, int wmax) {
  int fit = 0;
  for (int tc = 1; tc <= n_cols && tc <= 64; ++tc)
    if (PipeSmem<MODE, N>(tc, wmax).total <= pipe_smem_budget(ctx)) fit = tc;
  if (fit < 1) return 0;
  long n_ct = (n_cols + fit - 1) / fit;
  return (int)((n_cols + n_ct - 1) / n_ct);
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
  // capacity of the compact window-row tables: rows of a tile that can fall inside the frequency window
  const long n_w = 2 * n_iw - (stat == TQ_FERMION ? 0 : 1);
  const int wmax1 = (int)std::min<long>(Nb, n_w / Na + 2), wmax2 = (int)std::min<long>(Na, n_w / Nb + 2);
  const int tcw1 = fwd ? pipe_pick_tcw<M_F1, 250>(ctx, n_cols, wmax1) : pipe_pick_tcw<M_I1, 250>(ctx, n_cols, wmax1);
  const int tcw2 = fwd ? pipe_pick_tcw<M_F2, 400>(ctx, n_cols, wmax2) : pipe_pick_tcw<M_I2, 400>(ctx, n_cols, wmax2);
  if (tcw1 < 1 || tcw2 < 1) return TQ_OK;
  std::shared_ptr<PipePlan> pl;
  TQ_TRY(get_pipe_plan(ctx, beta, stat, n_tau, n_iw, &pl));
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
  A.btw = pl->btw;
  A.wtab = pl->wtab;
  A.itemph = pl->itemph;
  A.outph = pl->outph;
  const long in_rows = fwd ? L + 1 : pl->n_w, out_rows = fwd ? pl->n_w : L + 1;
  A.in_gf_stride = in_rows * n_cols;
  A.out_gf_stride = out_rows * n_cols;
  A.scratch_gf_stride = L * n_cols;
  // chunk the batch so that the scratch of one chunk stays L2 resident between the two kernels
  const size_t per_gf = (size_t)L * n_cols * sizeof(cd);
  size_t l2_budget = 48ull << 20;
  if (const char *e = getenv("TQ_TWOPASS_CHUNK_MB")) l2_budget = (size_t)atol(e) << 20;
  long chunk = std::max<long>(1, (long)(l2_budget / per_gf));
  chunk = std::min<long>(chunk, batch);
  TQ_TRY(tq_reserve(ctx, ctx->scratch, per_gf * chunk));
  A.scratch = (cd *)ctx->scratch.p;
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
    P1.twN = pl->twNb;
    P1.wmax = wmax1;
    PipeArgs P2 = P1;
    P2.wmax = wmax2;
    P2.tcw = tcw2;
    P2.n_ct = (int)((n_cols + tcw2 - 1) / tcw2);
    P2.n_items = Nb;
    P2.n_tiles = nb * Nb * P2.n_ct;
    P2.twN = pl->twNa;
    if (fwd) {
      P1.rowtab = pl->rowtab_f1;
      P1.itemtab = pl->itemtab_f1;
      TQ_TRY(make_f1_tensor_map(ctx, P1.in, nb, Na, Nb, n_cols, L, tcw1, &tmap));
      TQ_TRY((launch_pipe<M_F1, 250, 10, 25>(ctx, P1, tmap, "tau2iw_pass1")));
      TQ_TRY((launch_pipe<M_F2, 400, 20, 20>(ctx, P2, tmap, "tau2iw_pass2")));
    } else {
      P2.rowtab = pl->rowtab_i2;
      P2.itemtab = pl->itemtab_i2;
      TQ_TRY((launch_pipe<M_I1, 250, 10, 25>(ctx, P1, tmap, "iw2tau_pass1")));
      TQ_TRY((launch_pipe<M_I2, 400, 20, 20>(ctx, P2, tmap, "iw2tau_pass2")));
    }
  }
  *handled = true;
  return TQ_OK;
}
