// Direct transform imtime -> imfreq of SCALAR gfs on the L = 10^4 mesh (configs[0]: n_tau = 10001, batch of thousands).
//
// Same mathematics as matsubara_tile_kernel<IN_TAU, OUT_FREQ> (c++/triqs/gfs/transform/fourier_matsubara.cpp:96-186); what
// changes is that everything is compile time.  With one column per gf there is no column axis to share twiddles or row
// factors over, so the run-time engine spends most of its issue slots on index arithmetic and table look-ups.  Here:
//   * a persistent CTA per SM pulls a whole G(tau) column (160 016 B, rows 0..L) into shared memory with ONE bulk TMA copy;
//   * four in-place radix-10 DIF passes with compile-time strides 1000 / 100 / 10 / 1 (butterflies in registers);
//   * pass 0 fuses the pre-transform: row j = q + 1000 r, so  e^{i pi j/L} and the three tail exponentials factorise into a
//     per-q table (coalesced, one entry per butterfly) times per-leg kernel parameters (constant-bank operands); its
//     twiddles w_L^{q s} are built from the single table entry w_L^q by a product chain (the full 10^4-entry table would not
//     fit beside the tile); passes 1 and 2 read their 1000- / 100-entry twiddle tables through L1;
//   * the epilogue gathers the 2 n_iw window outputs from their digit-reversed positions and adds the trapezoid
//     correction and the tail model (:181-183).
// Bound: shared-memory bandwidth ~ FP64 (4 passes x 32 B per element vs ~64 FP64 instructions per element).
#include "tq_pipe.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

constexpr int COL_L = 10000;

struct ColArgs {
  const cd *in;
  cd *out;
  const cd *mom; // [batch][mom_rows][1]
  long batch, mom_gf_stride;
  int n_w, first, last, stat;
  double half_fact_fwd;
  const cd *wq;          // [1000] e^{+2 pi i q / L}
  const cd *tw1;         // [1000] e^{+2 pi i j / 1000}
  const cd *tw2;         // [100]  e^{+2 pi i j / 100}
  const double4 *qtab;   // [1000] {E1q, E2q, E3q, -}
  const cd *qph;         // [1000] (beta/L) e^{i pi q / L} (Fermion) or beta/L
  const double4 *wtab;   // [n_w] {b1 r1, w r1, b2 r2, w r2}
  double C[3];           // one{Fermion,Boson} prefactors
  double er[3][10];      // per-leg exponentials
  cd phr[10];            // per-leg twist e^{i pi r / 10} (Fermion) or 1
  int skip;              // developer timing experiment (TQ_COL_SKIP): 1 passes 1-3, 2 pass 0, 4 epilogue, 8 load
};

TQ_D double4 ldg4(const double4 *p) {
  const double2 *q = reinterpret_cast<const double2 *>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// one radix-10 DIF pass over the whole column: sub-blocks of length M = 10 MSUB, twiddles w_M^{q s} = tw[q s]
template <int MSUB, int COL_NT> TQ_D void col_pass(cd *tile, const cd *__restrict__ tw, int tid) {
  for (int i = tid; i < COL_L / 10; i += COL_NT) {
    const int blk = i / MSUB, q = i - blk * MSUB;
    cd *p = tile + blk * (10 * MSUB) + q;
    cd a[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) a[r] = p[r * MSUB];
    Dft<10, 1>::run(a);
    if (MSUB > 1 && q != 0) {
      // w_M^{q r} by a product chain from the one (coalesced) entry w_M^q: a gather of tw[q r] costs r x 4 wavefronts per warp
      const cd w = __ldg(&tw[q]);
      cd pw[10];
      pw[1] = w;
#pragma unroll
      for (int r = 2; r < 10; ++r) pw[r] = (r & 1) ? cmul(pw[r - 1], w) : cmul(pw[r >> 1], pw[r >> 1]);
#pragma unroll
      for (int r = 1; r < 10; ++r) a[r] = cmul(a[r], pw[r]);
    }
#pragma unroll
    for (int r = 0; r < 10; ++r) p[r * MSUB] = a[r];
  }
}

template <int COL_NT> __global__ void __launch_bounds__(COL_NT, 1) matsubara_col_kernel(const __grid_constant__ ColArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd *tile = reinterpret_cast<cd *>(smem_raw); // rows 0..L
  cd *small = reinterpret_cast<cd *>(smem_raw + ((sizeof(cd) * (COL_L + 1) + 127) & ~size_t(127)));
  uint64_t *mbar = reinterpret_cast<uint64_t *>(small + 8);
  const int tid = threadIdx.x;
  const bool fermion = A.stat == TQ_FERMION;
  if (tid == 0) {
    mbar_init(mbar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  int it = 0;
  for (long gf = blockIdx.x; gf < A.batch; gf += gridDim.x, ++it) {
    if (tid == 0 && !(A.skip & 8 && it > 0)) {
      mbar_arrive_expect_tx(mbar, (uint32_t)(sizeof(cd) * (COL_L + 1)));
      bulk_g2s(tile, A.in + gf * (COL_L + 1), (uint32_t)(sizeof(cd) * (COL_L + 1)), mbar);
    }
    // pole weights of this gf (:151-169) while the column is in flight
    cd a1, a2, a3, m1;
    {
      const cd *mom = A.mom + gf * A.mom_gf_stride;
      m1 = __ldg(&mom[1]);
      const cd m2 = __ldg(&mom[2]), m3 = __ldg(&mom[3]);
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
    if (!(A.skip & 8 && it > 0)) mbar_wait(mbar, (uint32_t)(it & 1));
    // trapezoid end-point correction -0.5 (beta/L) (g_0 + m1 +/- g_L)   (:181) -- read before pass 0 overwrites row 0
    cd ex;
    {
      const cd g0 = tile[0], gL = tile[COL_L];
      cd s = cadd(g0, m1);
      s = fermion ? cadd(s, gL) : csub(s, gL);
      ex = cscale(-A.half_fact_fwd, s);
    }
    __syncthreads();
    // ---- pass 0 (stride 1000) with the fused pre-transform:  x_j = phase_j (g_j - sum_i a_i h_i(tau_j)),  j = q + 1000 r
    if (!(A.skip & 2)) {
      const cd c1 = cscale(A.C[0], a1), c2 = cscale(A.C[1], a2), c3 = cscale(A.C[2], a3);
      for (int q = tid; q < 1000; q += COL_NT) {
        cd *p = tile + q;
        const double4 e = ldg4(&A.qtab[q]);
        const cd b1 = cscale(e.x, c1), b2 = cscale(e.y, c2), b3 = cscale(e.z, c3);
        cd a[10];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          cd v = caxpy(-A.er[0][r], b1, p[r * 1000]);
          v = caxpy(-A.er[1][r], b2, v);
          v = caxpy(-A.er[2][r], b3, v);
          a[r] = r == 0 ? v : cmul(v, A.phr[r]);
        }
        Dft<10, 1>::run(a);
        // outputs s: times w_L^{q s} (product chain from w_L^q) times the per-q part of the twist
        const cd w = __ldg(&A.wq[q]), ph = __ldg(&A.qph[q]);
        cd pw[10];
        pw[1] = w;
#pragma unroll
        for (int r = 2; r < 10; ++r) pw[r] = (r & 1) ? cmul(pw[r - 1], w) : cmul(pw[r >> 1], pw[r >> 1]);
        a[0] = cmul(a[0], ph);
#pragma unroll
        for (int r = 1; r < 10; ++r) a[r] = cmul(a[r], cmul(pw[r], ph));
#pragma unroll
        for (int r = 0; r < 10; ++r) p[r * 1000] = a[r];
      }
    }
    __syncthreads();
    if (!(A.skip & 1)) {
      col_pass<100, COL_NT>(tile, A.tw1, tid);
      __syncthreads();
      col_pass<10, COL_NT>(tile, A.tw2, tid);
      __syncthreads();
      col_pass<1, COL_NT>(tile, nullptr, tid);
      __syncthreads();
    }
    // ---- epilogue: gw[n] = X[(n + L) % L] + corr + sum_i a_i / (i w_n - b_i)    (:181-183)
    if (!(A.skip & 4)) {
      const cd D23 = csub(a2, a3), A23 = cadd(a2, a3);
      cd *o = A.out + gf * A.n_w;
      for (int di = tid; di < A.n_w; di += COL_NT) {
        const int n = A.first + di;
        const int k = n < 0 ? n + COL_L : n;
        const int k0 = k % 10, k1 = (k / 10) % 10, k2 = (k / 100) % 10, k3 = k / 1000;
        const cd x = tile[k0 * 1000 + k1 * 100 + k2 * 10 + k3]; // digit-reversed position of output k
        const double4 wv = ldg4(&A.wtab[di]);
        const cd t = tail_iw4(cmk(wv.x, wv.y), cmk(wv.z, wv.w), a1, D23, A23);
        o[di] = cadd(cadd(x, ex), t);
      }
    }
    __syncthreads(); // the tile may be overwritten by the next column
    if (tid == 0) fence_proxy_async();
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
struct ColPlan {
  ColArgs A{};
  std::vector<void *> allocs;
  ~ColPlan() {
    for (void *p : allocs) cudaFree(p);
  }
};
struct ColPlanCache {
  std::map<std::tuple<double, int, long>, std::shared_ptr<ColPlan>> plans;
};

template <class T> static tq_status cp_upload(tq_ctx *ctx, ColPlan &pl, const std::vector<T> &h, const T **d) {
  void *p = nullptr;
  TQ_CUDA(ctx, cudaMalloc(&p, sizeof(T) * std::max<size_t>(h.size(), 1)));
  pl.allocs.push_back(p);
  TQ_CUDA(ctx, cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *d = (const T *)p;
  return TQ_OK;
}

static tq_status get_col_plan(tq_ctx *ctx, double beta, int stat, long n_iw, std::shared_ptr<ColPlan> *out) {
  if (!ctx->col_cache) ctx->col_cache = std::make_shared<ColPlanCache>();
  auto &plans = ctx->col_cache->plans;
  auto key = std::make_tuple(beta, stat, n_iw);
  auto it = plans.find(key);
  if (it != plans.end()) {
    *out = it->second;
    return TQ_OK;
  }
  auto pl = std::make_shared<ColPlan>();
  ColArgs &A = pl->A;
  const long L = COL_L;
  const long double PI = 3.141592653589793238462643383279502884L;
  A.stat = stat;
  A.last = (int)(n_iw - 1);                                 // imfreq.hpp:58-60
  A.first = -(A.last + (stat == TQ_FERMION ? 1 : 0));
  A.n_w = A.last - A.first + 1;
  A.half_fact_fwd = 0.5 * (beta / (double)L);
  double b[3];
  if (stat == TQ_FERMION) {
    b[0] = 0; b[1] = 1; b[2] = -1; // fourier_matsubara.cpp:152-154
  } else {
    b[0] = -0.5; b[1] = -1; b[2] = 1; // :163-165
  }
  for (int i = 0; i < 3; ++i) {
    long double eb = expl(-(long double)beta * fabsl((long double)b[i]));
    if (stat == TQ_FERMION)
      A.C[i] = (double)(-1.0L / (1.0L + eb)); // oneFermion :28-30
    else
      A.C[i] = b[i] >= 0 ? (double)(1.0L / (eb - 1.0L)) : (double)(1.0L / (1.0L - eb)); // oneBoson :32-34
  }
  const long double delta = (long double)beta / (long double)L;
  // j = q + 1000 r:  b >= 0: e^{-b delta q} e^{-b delta 1000 r};  b < 0: L - j = (1000 - q) + 1000 (9 - r)
  std::vector<double4> qtab(1000);
  std::vector<cd> qph(1000), wq(1000), tw1(1000), tw2(100);
  for (int q = 0; q < 1000; ++q) {
    double e[3];
    for (int i = 0; i < 3; ++i) e[i] = (double)expl(-fabsl((long double)b[i]) * delta * (long double)(b[i] >= 0 ? q : 1000 - q));
    qtab[q] = make_double4(e[0], e[1], e[2], 0.0);
    const double scale = beta / (double)L;
    if (stat == TQ_FERMION) {
      long double ang = PI * (long double)q / (long double)L;
      qph[q] = cmk((double)((long double)scale * cosl(ang)), (double)((long double)scale * sinl(ang)));
    } else {
      qph[q] = cmk(scale, 0.0);
    }
    long double a2 = 2 * PI * (long double)q / (long double)L;
    wq[q] = cmk((double)cosl(a2), (double)sinl(a2));
  }
  tq_fft_twiddles_host(1000, tw1);
  tq_fft_twiddles_host(100, tw2);
  for (int r = 0; r < 10; ++r) {
    for (int i = 0; i < 3; ++i) A.er[i][r] = (double)expl(-fabsl((long double)b[i]) * delta * (long double)(b[i] >= 0 ? 1000 * r : 1000 * (9 - r)));
    long double ang = PI * (long double)r / 10.0L;
    A.phr[r] = stat == TQ_FERMION ? cmk((double)cosl(ang), (double)sinl(ang)) : cmk(1.0, 0.0);
  }
  std::vector<double4> wtab(A.n_w);
  for (int i = 0; i < A.n_w; ++i) {
    long n = A.first + i;
    double w = M_PI * (double)(2 * n + stat) / beta; // matsubara.hpp:55
    double r1 = 1.0 / (b[0] * b[0] + w * w), r2 = 1.0 / (b[1] * b[1] + w * w);
    wtab[i] = make_double4(b[0] * r1, w * r1, b[1] * r2, w * r2);
  }
  TQ_TRY(cp_upload(ctx, *pl, qtab, &A.qtab));
  TQ_TRY(cp_upload(ctx, *pl, qph, &A.qph));
  TQ_TRY(cp_upload(ctx, *pl, wq, &A.wq));
  TQ_TRY(cp_upload(ctx, *pl, tw1, &A.tw1));
  TQ_TRY(cp_upload(ctx, *pl, tw2, &A.tw2));
  TQ_TRY(cp_upload(ctx, *pl, wtab, &A.wtab));
  plans[key] = pl;
  *out = pl;
  return TQ_OK;
}

// imtime -> imfreq for scalar gfs on the 10^4 mesh; *handled = false when the specialisation does not apply
tq_status tq_col_matsubara_fwd(tq_ctx *ctx, double beta, int stat, long n_tau, long n_iw, long n_cols, long batch, const cd *d_in, cd *d_out,
                               const cd *d_mom, int mom_rows, bool *handled) {
  *handled = false;
  if (n_tau - 1 != COL_L || n_cols != 1 || ctx->opt_no_pipe || ctx->opt_no_col || TQ_DEV_ENV("TQ_NO_COL")) return TQ_OK;
  if (((uintptr_t)d_in) & 15) return TQ_OK; // bulk copies need 16-byte aligned sources (rows are 16 B, so every gf is)
  std::shared_ptr<ColPlan> pl;
  TQ_TRY(get_col_plan(ctx, beta, stat, n_iw, &pl));
  ColArgs A = pl->A;
  A.in = d_in;
  A.out = d_out;
  A.mom = d_mom;
  A.mom_gf_stride = mom_rows;
  A.batch = batch;
  A.skip = 0;
  if (const char *e = TQ_DEV_ENV("TQ_COL_SKIP")) A.skip = atoi(e);
  const size_t smem = ((sizeof(cd) * (COL_L + 1) + 127) & ~size_t(127)) + 16 * sizeof(cd);
  const int grid = (int)std::min<long>(batch, ctx->sm_count);
  static const int nt = [] {
    const char *e = TQ_DEV_ENV("TQ_COL_NT");
    return e ? atoi(e) : 1024;
  }();
  {
    KernelScope ks(ctx, "tau2iw_col");
    if (nt == 512) {
      TQ_CUDA(ctx, cudaFuncSetAttribute(matsubara_col_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      matsubara_col_kernel<512><<<grid, 512, smem, ctx->stream>>>(A);
    } else {
      TQ_CUDA(ctx, cudaFuncSetAttribute(matsubara_col_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      matsubara_col_kernel<1024><<<grid, 1024, smem, ctx->stream>>>(A);
    }
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  *handled = true;
  return TQ_OK;
}
