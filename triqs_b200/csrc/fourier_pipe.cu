// Host side of the pipelined four-step imtime <-> imfreq transforms: the plan (factorisation L = Na * Nb, radix pairs, tables),
// tensor maps and launches.  The kernels are in fourier_pipe_kernel.cuh.
#include "fourier_pipe_kernel.cuh"

// fourier_pipe_planned{0..3}.cu / fourier_pipe_heavy{0..3}.cu: the run-time-radix instantiations of each mode with the base / the
// extended radix set (one translation unit each)
#define TQ_DECL_PLANNED(name) tq_status name(tq_ctx *, const PipeArgs &, const CUtensorMap &, const CUtensorMap &, const char *);
TQ_DECL_PLANNED(tq_launch_pipe_planned_m0) TQ_DECL_PLANNED(tq_launch_pipe_planned_m1) TQ_DECL_PLANNED(tq_launch_pipe_planned_m2)
TQ_DECL_PLANNED(tq_launch_pipe_planned_m3) TQ_DECL_PLANNED(tq_launch_pipe_heavy_m0) TQ_DECL_PLANNED(tq_launch_pipe_heavy_m1)
TQ_DECL_PLANNED(tq_launch_pipe_heavy_m2) TQ_DECL_PLANNED(tq_launch_pipe_heavy_m3)
#undef TQ_DECL_PLANNED
static tq_status tq_launch_pipe_planned(tq_ctx *ctx, int mode, bool heavy, const PipeArgs &A, const CUtensorMap &t1, const CUtensorMap &t2,
                                        const char *tag) {
  switch (mode) {
    case M_F1: return heavy ? tq_launch_pipe_heavy_m0(ctx, A, t1, t2, tag) : tq_launch_pipe_planned_m0(ctx, A, t1, t2, tag);
    case M_F2: return heavy ? tq_launch_pipe_heavy_m1(ctx, A, t1, t2, tag) : tq_launch_pipe_planned_m1(ctx, A, t1, t2, tag);
    case M_I1: return heavy ? tq_launch_pipe_heavy_m2(ctx, A, t1, t2, tag) : tq_launch_pipe_planned_m2(ctx, A, t1, t2, tag);
    default: return heavy ? tq_launch_pipe_heavy_m3(ctx, A, t1, t2, tag) : tq_launch_pipe_planned_m3(ctx, A, t1, t2, tag);
  }
}

// -------------------------------------------------------------------------------------------------
// host
// -------------------------------------------------------------------------------------------------
struct ModeTables {
  const cd *twq = nullptr;
  const double4 *qtab = nullptr, *itemtab = nullptr, *wperm = nullptr;
  const int4 *wintab = nullptr;
  const cd *Tfull = nullptr;
  const cd *gtw = nullptr;          // generic pass B: exp(+2 pi i j / RB)
  const double4 *legtab = nullptr;  // I2 with a generic pass B: per-leg factors as tables
  const cd *legph = nullptr;
  double er[3][PIPE_MAXR];
  cd phr[PIPE_MAXR];
};
struct PipePlan {
  int L = 0, Na = 0, Nb = 0, first = 0, last = 0, n_w = 0, stat = 0, wmax1 = 0, wmax2 = 0;
  TileShape t1, t2; // pass 1 (length Nb), pass 2 (length Na)
  bool specialised = false; // the compile-time instantiations (400 = 20 x 20, 250 = 10 x 25) apply
  double beta = 0;
  ModeTables m[4];
  std::vector<void *> allocs;
  ~PipePlan() {
    for (void *p : allocs) cudaFree(p);
  }
};
struct PipePlanCache {
  std::map<std::tuple<double, int, long, long, int>, std::shared_ptr<PipePlan>> plans; // (..., orientation: see pipe_plan_lengths)
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
  std::vector<double4> legtab; // [Y] the same per-leg factors as tables (Y may exceed PIPE_MAXR with a generic pass B)
  std::vector<cd> legph;
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
  f.legtab.assign(std::max(Y, 1), make_double4(0, 0, 0, 0));
  f.legph.assign(std::max(Y, 1), cmk(1, 0));
  for (int y = 0; y < std::max(Y, PIPE_MAXR); ++y) {
    double e[3];
    for (int i = 0; i < 3; ++i) e[i] = y < Y ? ex(i, b[i] >= 0 ? (long)S * X * y : (long)S * X * (Y - 1 - y)) : 0.0;
    const cd ph = (stat == TQ_FERMION && y < Y) ? expi(sgn * PI_L * (long double)((long)S * X * y) / (long double)L) : cmk(1, 0);
    if (y < PIPE_MAXR) {
      for (int i = 0; i < 3; ++i) f.er[i][y] = e[i];
      f.phr[y] = ph;
    }
    if (y < Y) {
      f.legtab[y] = make_double4(e[0], e[1], e[2], 0.0);
      f.legph[y] = ph;
    }
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

// ---- the plan: L = Na * Nb, Nb = RA1 * RB1 (pass 1), Na = RA2 * RB2 (pass 2) ---------------------------------------------------
// FFTW takes any length (fourier_common.cpp:30-44); the default adjoint mesh is L = 6 n_iw (mesh/adjoint.hpp:35-38), the reference's
// own test mesh L = 9999 = 3^2 11 101.  Every divisor pair is tried; a tile length is usable when it splits into a register radix
// RA and a second radix RB that is either a register radix too or small enough for the direct-sum pass B.
constexpr int PIPE_MAX_TILE = 640;      // longest tile (rows) the shared-memory ring is sized for

// dir: +1 imtime -> imfreq, -1 imfreq -> imtime, 0 either.  A tile length with a heavy (Rader) radix costs very differently in the four kernels
// (1.2 GB per launch: F1 1.2 ms with 7 workers, I2 1.15, F2 2.0, I1 2.8 -- smooth tiles 0.45-0.65), so the forward leg wants it in pass 1 and the
// return leg in pass 2: the two directions of such a mesh use transposed plans (L = 6150: forward 2.7 -> 1.8 ms; 61500: inverse 3.5 -> 1.9 ms).
static bool pipe_plan_lengths(long L, TileShape *t1, TileShape *t2, int dir = 0) {
  if (L < 64 || L > (long)PIPE_MAX_TILE * PIPE_MAX_TILE) return false;
  if (L == 100000) { // the headline mesh keeps its tuned, fully compile-time kernels
    tile_shape(250, t1);
    tile_shape(400, t2);
    *t1 = TileShape{250, 10, 25, false, false, 2.0};
    *t2 = TileShape{400, 20, 20, false, false, 2.0};
    return true;
  }
  bool found = false;
  double best = 0;
  const char *force_nb_f = TQ_DEV_ENV("TQ_PLAN_NB_FWD"), *force_nb_i = TQ_DEV_ENV("TQ_PLAN_NB_INV"); // (developer: pass-1 tile length per direction)
  const long force_nb = (dir > 0 && force_nb_f) ? atol(force_nb_f) : (dir < 0 && force_nb_i) ? atol(force_nb_i) : 0;
  for (long nb = 1; nb <= PIPE_MAX_TILE; ++nb) {
    if (L % nb) continue;
    if (force_nb && nb != force_nb) continue;
    const long na = L / nb;
    if (na > PIPE_MAX_TILE) continue;
    TileShape a, b;
    if (!tile_shape((int)nb, &a) || !tile_shape((int)na, &b)) continue;
    // both passes stream the whole array once; very short tiles pay per-tile overhead (barriers, tables), very long ones
    // leave room for few columns in the ring
    double cost = a.cost + b.cost + 0.15 * std::fabs(std::log((double)na / (double)nb));
    if (nb < 16) cost += 0.5;
    if (na < 16) cost += 0.5;
    if (nb > na) cost += 0.05;  // pass 1 keeps a deeper ring of (padded) tiles: give it the shorter length
    if (std::max(na, nb) > 512) cost += 0.3; // few columns of such a tile fit the ring
    if (dir > 0 && b.heavy && !a.heavy) cost += 0.8; // (a: pass-1 tile, b: pass-2 tile)
    if (dir < 0 && a.heavy && !b.heavy) cost += 0.8;
    if (!found || cost < best) {
      best = cost;
      *t1 = a;
      *t2 = b;
      found = true;
    }
  }
  return found;
}

static tq_status get_pipe_plan(tq_ctx *ctx, double beta, int stat, long n_tau, long n_iw, bool fwd, std::shared_ptr<PipePlan> *out) {
  if (!ctx->pipe_cache) ctx->pipe_cache = std::make_shared<PipePlanCache>();
  auto &plans = ctx->pipe_cache->plans;
  // orientation: 0 when both directions pick the same plan (every smooth length), +-1 when a heavy tile makes them differ
  int orient = 0;
  {
    TileShape f1, f2, i1, i2;
    if (pipe_plan_lengths(n_tau - 1, &f1, &f2, +1) && pipe_plan_lengths(n_tau - 1, &i1, &i2, -1) && (f1.N != i1.N || f1.RA != i1.RA))
      orient = fwd ? +1 : -1;
    if (TQ_DEV_ENV("TQ_PLAN_NB_FWD") || TQ_DEV_ENV("TQ_PLAN_NB_INV")) orient = fwd ? +1 : -1;
  }
  auto key = std::make_tuple(beta, stat, n_tau, n_iw, orient);
  auto it = plans.find(key);
  if (it != plans.end()) {
    *out = it->second;
    return TQ_OK;
  }
  auto pl = std::make_shared<PipePlan>();
  const long L = n_tau - 1;
  if (!pipe_plan_lengths(L, &pl->t1, &pl->t2, orient)) return tq_fail(ctx, TQ_ERR_INVALID, "internal: pipe plan for an unsupported length");
  const TileShape &T1 = pl->t1, &T2 = pl->t2;
  const int Na = T2.N, Nb = T1.N;
  pl->specialised = (T1.N == 250 && T1.RA == 10 && T1.RB == 25 && T2.N == 400 && T2.RA == 20 && T2.RB == 20);
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
  tau_factors(beta, stat, L, Na, T1.RB, T1.RA, +1, beta / (double)L, b, C, f1);
  // I2: j = beta + Nb (s + RA t): static over s in [0,RA), legs t in [0,RB); phase e^{-i pi j/L} (item part goes into pass 1)
  tau_factors(beta, stat, L, Nb, T2.RA, T2.RB, -1, 1.0, b, C, f2);
  {
    ModeTables &m = pl->m[M_F1];
    TQ_TRY(pp_upload(ctx, *pl, twiddle_rows(T1.N, T1.RA, T1.RB, +1, &f1.xph, nullptr), &m.twq));
    TQ_TRY(pp_upload(ctx, *pl, f1.xtab, &m.qtab));
    TQ_TRY(pp_upload(ctx, *pl, f1.itemtab, &m.itemtab));
    memcpy(m.er, f1.er, sizeof(m.er));
    memcpy(m.phr, f1.phr, sizeof(m.phr));
  }
  {
    ModeTables &m = pl->m[M_I2];
    TQ_TRY(pp_upload(ctx, *pl, twiddle_rows(T2.N, T2.RA, T2.RB, -1, nullptr, &f2.xph), &m.twq));
    TQ_TRY(pp_upload(ctx, *pl, f2.xtab, &m.qtab));
    TQ_TRY(pp_upload(ctx, *pl, f2.itemtab, &m.itemtab));
    memcpy(m.er, f2.er, sizeof(m.er));
    memcpy(m.phr, f2.phr, sizeof(m.phr));
    TQ_TRY(pp_upload(ctx, *pl, f2.legtab, &m.legtab));
    TQ_TRY(pp_upload(ctx, *pl, f2.legph, &m.legph));
  }
  {
    // direct-sum pass B: exp(+2 pi i j / RB) of each pass (the kernels conjugate for the inverse direction)
    std::vector<cd> g1, g2;
    tq_fft_twiddles_host(std::max(T1.RB, 1), g1);
    tq_fft_twiddles_host(std::max(T2.RB, 1), g2);
    TQ_TRY(pp_upload(ctx, *pl, g1, &pl->m[M_F1].gtw));
    TQ_TRY(pp_upload(ctx, *pl, g2, &pl->m[M_F2].gtw));
    pl->m[M_I1].gtw = pl->m[M_F1].gtw;
    pl->m[M_I2].gtw = pl->m[M_F2].gtw;
  }
  TQ_TRY(pp_upload(ctx, *pl, twiddle_rows(T2.N, T2.RA, T2.RB, +1, nullptr, nullptr), &pl->m[M_F2].twq));
  TQ_TRY(pp_upload(ctx, *pl, twiddle_rows(T1.N, T1.RA, T1.RB, -1, nullptr, nullptr), &pl->m[M_I1].twq));
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
    auto perm = [&](int n_items, int S, int N, int wmax, std::vector<double4> &out, std::vector<int4> &wins) -> bool {
      out.assign((size_t)n_items * wmax, make_double4(0, 0, 0, 0));
      wins.assign(n_items, make_int4(0, 0, 0, 0));
      for (int item = 0; item < n_items; ++item) {
        Window win = tile_window(item, S, N, (int)L, pl->first, pl->last);
        wins[item] = make_int4(win.r1, win.r2, 0, 0);
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
    std::vector<int4> i1, i2;
    if (!perm(Na, Na, Nb, pl->wmax1, w1, i1) || !perm(Nb, Nb, Na, pl->wmax2, w2, i2))
      return tq_fail(ctx, TQ_ERR_INVALID, "internal: window-row table capacity");
    TQ_TRY(pp_upload(ctx, *pl, w1, &pl->m[M_I1].wperm));
    TQ_TRY(pp_upload(ctx, *pl, w2, &pl->m[M_F2].wperm));
    TQ_TRY(pp_upload(ctx, *pl, i1, &pl->m[M_I1].wintab));
    TQ_TRY(pp_upload(ctx, *pl, i2, &pl->m[M_F2].wintab));
  }
  plans[key] = pl;
  *out = pl;
  return TQ_OK;
}

// G(tau) of `nb` gfs as the 4-D tensor [gf][b][a][2 n_cols doubles] (row j = a + Na b); box = one F1 tile [Nb][1][2 tcw]
static tq_status make_f1_tensor_map(tq_ctx *ctx, const cd *in, long nb, int Na, int Nb, long n_cols, long L, int tcw, int box_rows,
                                    CUtensorMap *out) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {(cuuint64_t)(2 * n_cols), (cuuint64_t)Na, (cuuint64_t)Nb, (cuuint64_t)nb};
  cuuint64_t strides[3] = {(cuuint64_t)(n_cols * sizeof(cd)), (cuuint64_t)((long)Na * n_cols * sizeof(cd)), (cuuint64_t)((L + 1) * n_cols * sizeof(cd))};
  cuuint32_t box[4] = {(cuuint32_t)(2 * tcw), 1u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return TQ_OK;
}

// G(i w) of `nb` gfs, one half of the window, as the 4-D tensor [gf][b'][a][2 n_cols doubles]: row = row0 + a + Na b'.
// Valid only when the half is a whole number of Na-row groups (n_rows % Na == 0), so that the box never leaves the window.
static tq_status make_i1_tensor_map(tq_ctx *ctx, const cd *in, long row0, long n_rows, long nb, int Na, long n_cols, long n_w, int pitch,
                                    CUtensorMap *out) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const long nbp = n_rows / Na;
  cuuint64_t dims[4] = {(cuuint64_t)(2 * n_cols), (cuuint64_t)Na, (cuuint64_t)nbp, (cuuint64_t)nb};
  cuuint64_t strides[3] = {(cuuint64_t)(n_cols * sizeof(cd)), (cuuint64_t)((long)Na * n_cols * sizeof(cd)), (cuuint64_t)(n_w * n_cols * sizeof(cd))};
  cuuint32_t box[4] = {(cuuint32_t)(2 * pitch), 1u, (cuuint32_t)nbp, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)(in + row0 * n_cols), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled (I1) failed with CUresult " + std::to_string((int)r));
  return TQ_OK;
}

// "gfs as columns" (narrow targets, nc = 1..3 columns per gf): G(tau) of `batch` gfs as the 4-D tensor [b][a][gf][2 nc doubles];
// box = one F1 tile [box_rows][1][tcw / nc gfs][2 nc]  -> shared memory [row][tcw columns], column = gf * nc + c
static tq_status make_f1_tr_tensor_map(tq_ctx *ctx, const cd *in, long batch, int Na, int Nb, int nc, long L, int tcw, int box_rows, CUtensorMap *out) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {(cuuint64_t)(2 * nc), (cuuint64_t)batch, (cuuint64_t)Na, (cuuint64_t)Nb};
  cuuint64_t strides[3] = {(cuuint64_t)((L + 1) * nc * sizeof(cd)), (cuuint64_t)(nc * sizeof(cd)), (cuuint64_t)((long)Na * nc * sizeof(cd))};
  cuuint32_t box[4] = {(cuuint32_t)(2 * nc), (cuuint32_t)(tcw / nc), 1u, (cuuint32_t)box_rows};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled (F1, gfs as columns) failed with CUresult " + std::to_string((int)r));
  return TQ_OK;
}
// ... and G(i w) as the 3-D tensor [row][gf][2 nc doubles]; box = one window row of a tile, [1][tcw / nc gfs][2 nc]
static tq_status make_i1_tr_tensor_map(tq_ctx *ctx, const cd *in, long batch, int nc, long n_w, int tcw, CUtensorMap *out) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {(cuuint64_t)(2 * nc), (cuuint64_t)batch, (cuuint64_t)n_w};
  cuuint64_t strides[2] = {(cuuint64_t)(n_w * nc * sizeof(cd)), (cuuint64_t)(nc * sizeof(cd))};
  cuuint32_t box[3] = {(cuuint32_t)(2 * nc), (cuuint32_t)(tcw / nc), 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled (I1, gfs as columns) failed with CUresult " + std::to_string((int)r));
  return TQ_OK;
}

// widest balanced column tile for a tile length N: n_ct tiles of width ceil(n_cols / n_ct)
static int env_int_early(const char *name, int dflt) {
  const char *e = TQ_DEV_ENV_NAME(name);
  return e ? atoi(e) : dflt;
}
// pad: 0 = dense rows, 1 = rows padded to 8 columns (128 B), 2 = padded only if `rows_per_box` rows of the tile are not a multiple of 128 B
template <int MODE> static int pipe_pitch(int tc, int pad, int rows_per_box) {
  if (pad == 1 || (pad == 2 && (rows_per_box * tc) % 8 != 0)) return (tc + 7) & ~7;
  return tc;
}
// step > 1 ("gfs as columns"): widths are multiples of `step` (whole gfs, and rows of a multiple of 128 bytes), no balancing
template <int MODE> static int pipe_pick_tcw(tq_ctx *ctx, const TileShape &T, long n_cols, int wmax, int cap, int pad = 0, int rows_per_box = 0,
                                             int step = 1) {
  int fit = 0;
  for (int tc = step; (tc <= n_cols || tc == step) && tc <= cap; tc += step)
    if (PipeSmem<MODE>(T.N, T.RA, T.RB, tc, pipe_pitch<MODE>(tc, pad, rows_per_box), wmax, T.generic).total <= (size_t)ctx->max_smem_optin) fit = tc;
  if (fit < 1) return 0;
  if (step > 1) return fit;
  static const int unbalanced = env_int_early("TQ_PIPE_UNBALANCED", 0); // (tuning experiment: widest tiles + a narrow remainder)
  if (unbalanced) return fit;
  long n_ct = (n_cols + fit - 1) / fit;
  return (int)((n_cols + n_ct - 1) / n_ct);
}

static int env_int(const char *name, int dflt) {
  const char *e = TQ_DEV_ENV_NAME(name);
  return e ? atoi(e) : dflt;
}

// Cost of the two-radix plan of a mesh length in units of a register-radix pass (2 = both passes in registers; a direct-sum radix R adds
// ~R / 9); < 0: no plan.  fourier_matsubara.cu prefers the Bluestein convolution above ~9.5 (measured: 3400 / cost GB/s against 360-410).
double tq_pipe_plan_cost(long L) {
  TileShape a, b;
  if (!pipe_plan_lengths(L, &a, &b)) return -1.0;
  return a.cost + b.cost;
}

// Host-only description of the plan of a mesh length and direction (tq_plan_describe): "Nb=250(10x25) Na=400(20x20) cost=4.00 heavy=0 generic=0"
std::string tq_pipe_plan_text(long L, int dir) {
  TileShape a, b;
  if (!pipe_plan_lengths(L, &a, &b, dir)) return "none";
  char buf[192];
  snprintf(buf, sizeof buf, "Nb=%d(%dx%d) Na=%d(%dx%d) cost=%.2f heavy=%d%d generic=%d%d", a.N, a.RA, a.RB, b.N, b.RA, b.RB, a.cost + b.cost, (int)a.heavy,
           (int)b.heavy, (int)a.generic, (int)b.generic);
  return buf;
}

// Runs the transform if a pipelined specialisation exists for this mesh.  *handled tells the caller.
tq_status tq_pipe_matsubara(tq_ctx *ctx, bool fwd, double beta, int stat, long n_tau, long n_iw, long n_cols, long batch, const cd *d_in,
                            cd *d_out, const cd *d_mom, int mom_rows, bool *handled) {
  *handled = false;
  const long L = n_tau - 1;
  if (ctx->opt_no_pipe) return TQ_OK;
  // "gfs as columns": a batch of narrow gfs (1-3 target columns) runs as ONE virtual gf whose columns are all (gf, column) pairs;
  // the tiles are gathered / scattered with the gf stride by tensor maps (fourier_pipe_kernel.cuh, PipeArgs::tr)
  int tr = 0;
  {
    TileShape a, b;
    if (!pipe_plan_lengths(L, &a, &b)) return TQ_OK;
    if (L < 256) return TQ_OK;
    if (n_cols < 4) {
      // the single-column kernel (fourier_col.cu: one HBM read, one write, no scratch) keeps its mesh
      const bool col_kernel = fwd && L == 10000 && n_cols == 1 && !ctx->opt_no_col;
      // (a lane of a chunked host-buffer call decides as the whole call would: the chunking must not change a bit)
      // Single-column gfs, inverse leg: I2 writes one 16-byte element per (gf, time point) at the gf stride -- half a 32-byte sector each, 3x the
      // time of the two-column case whether the lanes or a tensor-map store issue the writes (both measured).  On smooth lengths whose column
      // fits one SM's shared memory the single-kernel path is faster for this one case (L = 6000: 1015 vs 750 GB/s, 10^4: 941 vs 683); lengths
      // with a heavy / direct-sum radix (6150: 281 vs 643) and longer meshes (2 10^4: 175 vs 671) keep the planned kernels.
      const bool single_kernel_fits = sizeof(cd) * (size_t)(L + 64) <= (size_t)ctx->max_smem_optin;
      const bool smooth = !a.heavy && !a.generic && !b.heavy && !b.generic;
      const bool tr_loses = !fwd && n_cols == 1 && single_kernel_fits && smooth;
      if (std::max(batch, ctx->call_batch) * n_cols >= 8 && !col_kernel && !tr_loses && !ctx->opt_no_tr) {
        tr = (int)n_cols;
      } else {
        // Narrow targets on meshes whose whole column fits one SM's shared memory stay with the single-kernel paths (one HBM read
        // and one write, no scratch): the pipelined tiles would have rows of 16-48 bytes.
        if (single_kernel_fits) return TQ_OK;
      }
    }
  }
  if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return TQ_OK; // bulk copies need 16-byte aligned rows
#ifdef TQ_KNOCKOUT
  {
    const int kv = env_int("TQ_KNOCK", 0);
    cudaMemcpyToSymbol(g_tq_knock, &kv, sizeof(int));
  }
#endif
  std::shared_ptr<PipePlan> pl;
  TQ_TRY(get_pipe_plan(ctx, beta, stat, n_tau, n_iw, fwd, &pl));
  const TileShape &T1 = pl->t1, &T2 = pl->t2;
  const int Na = pl->Na, Nb = pl->Nb;
  const bool spec = pl->specialised && !ctx->opt_force_planned && !tr;
  const long n_cols_real = n_cols; // target columns of a gf
  const size_t tr_budget = ctx->opt_scratch_chunk_mb > 0 ? (size_t)ctx->opt_scratch_chunk_mb << 20 : 1536ull << 20;
  // gfs per launch pair ("gfs as columns"): bounded by the scratch and by 2^31 tiles / 2^31 columns
  const long tr_chunk = std::min<long>(batch, std::max<long>(8, std::min<long>((long)(tr_budget / ((size_t)L * n_cols * sizeof(cd))), (1L << 22) / n_cols)));
  if (tr) n_cols = tr_chunk * n_cols; // columns of the virtual gf of one chunk
  const int tr_step = tr == 3 ? 24 : 8; // whole gfs per tile and rows of a multiple of 128 bytes (TMA destinations)
  const int wmax1 = pl->wmax1, wmax2 = pl->wmax2;
  static const int cap1 = env_int("TQ_PIPE_TCW1", 64), cap2 = env_int("TQ_PIPE_TCW2", 64); // (tile-width caps: tuning experiments)
  // I1: window rows with n >= 0 / n < 0; when both are whole groups of Na rows they are two rectangles of the [b][a] view
  const long n_pos = (long)pl->last + 1, n_neg = -(long)pl->first;
  bool i1_rect = n_pos % Na == 0 && n_neg % Na == 0 && n_pos > 0 && n_neg > 0 && n_pos / Na <= 256 && n_neg / Na <= 256 && !tr &&
                 !TQ_DEV_ENV("TQ_PIPE_NO_I1_BOXES");
  // (the pruned pass A exists in the compile-time instantiations only)
  const bool i1_edge = spec && i1_rect && n_pos == (long)Na * T1.RB && n_neg == (long)Na * T1.RB && !TQ_DEV_ENV("TQ_PIPE_NO_I1_EDGE");
  // F1 fetches its tile as tensor boxes of at most 256 rows (each box lands 128-byte aligned: rows padded to 8 columns then)
  int f1_boxes = 1;
  while (Nb % f1_boxes != 0 || Nb / f1_boxes > 256) ++f1_boxes;
  const int f1_pad = (f1_boxes > 1 && !tr) ? 2 : 0;
  // I1: the two tensor tiles of the rectangle path land 128-byte aligned only with rows padded to 8 columns; when that does not fit
  // the ring (long tiles), the row-by-row path with dense rows is used instead
  int tcw1 = tr ? (fwd ? pipe_pick_tcw<M_F1>(ctx, T1, n_cols, wmax1, cap1, 0, 0, tr_step) : pipe_pick_tcw<M_I1>(ctx, T1, n_cols, wmax1, cap1, 0, 0, tr_step))
             : fwd ? pipe_pick_tcw<M_F1>(ctx, T1, n_cols, wmax1, cap1, f1_pad, Nb / f1_boxes)
                   : pipe_pick_tcw<M_I1>(ctx, T1, n_cols, wmax1, cap1, (i1_rect && !i1_edge) ? 1 : 0);
  if (!fwd && i1_rect && !i1_edge) {
    const int dense = pipe_pick_tcw<M_I1>(ctx, T1, n_cols, wmax1, cap1, 0);
    if (tcw1 < 1 || (n_cols + tcw1 - 1) / tcw1 > (n_cols + dense - 1) / std::max(dense, 1)) { // fewer column tiles without the padding
      i1_rect = false;
      tcw1 = dense;
    }
  }
  const int tcw2 = fwd ? pipe_pick_tcw<M_F2>(ctx, T2, n_cols, wmax2, cap2, 0, 0, tr ? tr_step : 1)
                       : pipe_pick_tcw<M_I2>(ctx, T2, n_cols, wmax2, cap2, 0, 0, tr ? tr_step : 1);
  if (tcw1 < 1 || tcw2 < 1) return TQ_OK;
  PipeArgs A{};
  A.Na = Na;
  A.Nb = Nb;
  A.n_cols = (int)n_cols;
  A.tcw2 = tcw2;
  // pole weights of all gfs (tiny): [2][batch][4][n_cols]   ("gfs as columns": [2][4][all columns], written per scratch chunk below)
  const size_t pw_elems = (size_t)batch * n_cols_real * 4;
  TQ_TRY(tq_reserve(ctx, ctx->pipe_w, sizeof(cd) * 2 * pw_elems));
  cd *w_tau = (cd *)ctx->pipe_w.p, *w_freq = w_tau + pw_elems;
  const auto pole_weights = [&](long g0, long n_gf, cd *wt, cd *wf) -> tq_status {
    const long total = n_gf * n_cols_real;
    KernelScope ks(ctx, "pole_weights");
    k_pole_weights<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(d_mom + g0 * mom_rows * n_cols_real, (long)mom_rows * n_cols_real, (int)n_cols_real,
                                                                            total, stat, fwd ? d_in + g0 * (L + 1) * n_cols_real : nullptr,
                                                                            (L + 1) * n_cols_real, (int)L, 0.5 * (beta / (double)L), wt, wf, tr);
    tq_count_launch(ctx);
    TQ_CUDA(ctx, cudaGetLastError());
    return TQ_OK;
  };
  if (!tr) TQ_TRY(pole_weights(0, batch, w_tau, w_freq));
  A.L = (int)L;
  A.first = pl->first;
  A.last = pl->last;
  A.stat = stat;
  A.half_fact_fwd = 0.5 * (beta / (double)L);
  A.inv_beta = 1.0 / beta;
  A.keep_guard = env_int("TQ_PIPE_GUARD", 0);
  const long in_rows = fwd ? L + 1 : pl->n_w, out_rows = fwd ? pl->n_w : L + 1;
  A.in_gf_stride = in_rows * n_cols;
  A.out_gf_stride = out_rows * n_cols;
  A.scratch_gf_stride = env_int("TQ_SCRATCH_ALIAS", 0) ? 0 : L * n_cols; // (developer timing experiment: every gf through the same L2-resident scratch)
  // Chunk the batch by scratch size.  Measured on B200 (tools/gpu_probe.py c2): keeping the scratch L2 resident (chunks of
  // <= 3 blocks = 120 MB) costs more in per-launch prologue / pipeline fill and drain than it saves in HBM traffic while the
  // kernels are FP64-bound, so the default is one launch pair per 1.5 GB of scratch.
  const size_t per_gf = (size_t)L * n_cols * sizeof(cd);
  size_t l2_budget = 1536ull << 20;
  if (ctx->opt_scratch_chunk_mb > 0) l2_budget = (size_t)ctx->opt_scratch_chunk_mb << 20;
  long chunk = std::max<long>(1, (long)(l2_budget / per_gf));
  chunk = std::min<long>(std::min<long>(chunk, batch), 1L << 18); // (keeps the tile count of a launch below 2^31)
  if (!tr) TQ_TRY(tq_reserve(ctx, ctx->scratch, per_gf * chunk));
  A.scratch = (cd *)ctx->scratch.p;
  auto set_mode = [&](PipeArgs &P, int mode) {
    const ModeTables &m = pl->m[mode];
    P.twq = m.twq;
    P.qtab = m.qtab;
    P.itemtab = m.itemtab;
    P.Tfull = m.Tfull;
    P.wperm = m.wperm;
    P.wintab = m.wintab;
    P.gtw = m.gtw;
    P.legtab = m.legtab;
    P.legph = m.legph;
    memcpy(P.er, m.er, sizeof(P.er));
    memcpy(P.phr, m.phr, sizeof(P.phr));
    const TileShape &T = (mode == M_F1 || mode == M_I1) ? T1 : T2;
    P.tN = T.N;
    P.tRA = T.RA;
    P.tRB = T.RB;
    P.rb_generic = T.generic ? 1 : 0;
    P.lag0 = (int)ctx->opt_pipe_lag0;
  };
  if (tr) {
    const long nc = n_cols_real;
    TQ_TRY(tq_reserve(ctx, ctx->scratch, (size_t)L * n_cols * sizeof(cd)));
    A.scratch = (cd *)ctx->scratch.p;
    A.tr = tr;
    A.col_stride_out = out_rows * nc;
    for (long g0 = 0; g0 < batch; g0 += tr_chunk) {
      const long ng = std::min(tr_chunk, batch - g0), ncv = ng * nc;
      TQ_TRY(pole_weights(g0, ng, w_tau, w_freq));
      PipeArgs P1 = A;
      P1.n_cols = (int)ncv;
      P1.in = d_in + g0 * in_rows * nc;
      P1.out = d_out + g0 * out_rows * nc;
      P1.in_gf_stride = P1.out_gf_stride = P1.scratch_gf_stride = 0; // one virtual gf
      P1.polew = fwd ? w_tau : w_freq;
      P1.tcw = P1.pitch = tcw1;
      P1.n_ct = (int)((ncv + tcw1 - 1) / tcw1);
      P1.n_items = Na;
      P1.n_tiles = (long)Na * P1.n_ct;
      P1.wmax = wmax1;
      PipeArgs P2 = P1;
      P2.tcw = P2.pitch = tcw2;
      P2.n_ct = (int)((ncv + tcw2 - 1) / tcw2);
      P2.n_items = Nb;
      P2.n_tiles = (long)Nb * P2.n_ct;
      P2.wmax = wmax2;
      P2.polew = fwd ? w_freq : w_tau;
      CUtensorMap tm1, tm2;
      memset(&tm1, 0, sizeof(tm1));
      memset(&tm2, 0, sizeof(tm2));
      if (fwd) {
        set_mode(P1, M_F1);
        set_mode(P2, M_F2);
        P1.f1_boxes = f1_boxes;
        P1.f1_box_rows = Nb / f1_boxes;
        TQ_TRY(make_f1_tr_tensor_map(ctx, P1.in, ng, Na, Nb, (int)nc, L, tcw1, P1.f1_box_rows, &tm1));
        TQ_TRY(tq_launch_pipe_planned(ctx, M_F1, T1.heavy, P1, tm1, tm2, "tau2iw_pass1"));
        TQ_TRY(tq_launch_pipe_planned(ctx, M_F2, T2.heavy, P2, tm1, tm2, "tau2iw_pass2"));
      } else {
        set_mode(P1, M_I1);
        set_mode(P2, M_I2);
        P1.i1_rect = P1.i1_edge = 0;
        TQ_TRY(make_i1_tr_tensor_map(ctx, P1.in, ng, (int)nc, pl->n_w, tcw1, &tm1));
        TQ_TRY(tq_launch_pipe_planned(ctx, M_I1, T1.heavy, P1, tm1, tm2, "iw2tau_pass1"));
        TQ_TRY(tq_launch_pipe_planned(ctx, M_I2, T2.heavy, P2, tm1, tm2, "iw2tau_pass2"));
      }
    }
    *handled = true;
    return TQ_OK;
  }
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  for (long b0 = 0; b0 < batch; b0 += chunk) {
    const long nb = std::min(chunk, batch - b0);
    PipeArgs P1 = A;
    P1.in = d_in + b0 * A.in_gf_stride;
    P1.out = d_out + b0 * A.out_gf_stride;
    P1.polew = (fwd ? w_tau : w_freq) + (size_t)b0 * n_cols * 4;
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
    P2.polew = (fwd ? w_freq : w_tau) + (size_t)b0 * n_cols * 4;
    CUtensorMap tmap2;
    memset(&tmap2, 0, sizeof(tmap2));
    P1.pitch = tcw1;
    P2.pitch = tcw2;
    if (fwd) {
      set_mode(P1, M_F1);
      set_mode(P2, M_F2);
      P1.f1_boxes = f1_boxes;
      P1.f1_box_rows = Nb / f1_boxes;
      P1.pitch = pipe_pitch<M_F1>(tcw1, f1_pad, Nb / f1_boxes);
      TQ_TRY(make_f1_tensor_map(ctx, P1.in, nb, Na, Nb, n_cols, L, P1.pitch, P1.f1_box_rows, &tmap));
      if (spec) {
        TQ_TRY((launch_pipe<M_F1, 250, 10, 25, TQ_NW_F1>(ctx, P1, tmap, tmap2, "tau2iw_pass1")));
        TQ_TRY((launch_pipe<M_F2, 400, 20, 20, TQ_NW_F2>(ctx, P2, tmap, tmap2, "tau2iw_pass2")));
      } else {
        TQ_TRY(tq_launch_pipe_planned(ctx, M_F1, T1.heavy, P1, tmap, tmap2, "tau2iw_pass1"));
        TQ_TRY(tq_launch_pipe_planned(ctx, M_F2, T2.heavy, P2, tmap, tmap2, "tau2iw_pass2"));
      }
    } else {
      set_mode(P1, M_I1);
      set_mode(P2, M_I2);
      P1.i1_rect = i1_rect;
      P1.i1_edge = i1_edge;
      // tile rows land at a pitch rounded up to 8 columns (128 B) so that the second tensor tile (rows r2..) is 128-byte
      // aligned; the pruned path stages that tile separately and keeps the conflict-free dense pitch
      P1.pitch = (P1.i1_rect && !P1.i1_edge) ? ((tcw1 + 7) & ~7) : tcw1;
      for (int r = 0; r < PIPE_MAXR; ++r) P1.wedge[r] = expi(2 * PI_L * (long double)r / (long double)T1.RA); // SGN = -1: exp(+2 pi i r / RA)
      if (P1.i1_rect) {
        TQ_TRY(make_i1_tensor_map(ctx, P1.in, n_neg, n_pos, nb, Na, n_cols, pl->n_w, P1.pitch, &tmap));  // k in [0, last]: data rows n - first
        TQ_TRY(make_i1_tensor_map(ctx, P1.in, 0, n_neg, nb, Na, n_cols, pl->n_w, P1.pitch, &tmap2));     // k in [L + first, L): data rows 0..
      }
      if (spec) {
        TQ_TRY((launch_pipe<M_I1, 250, 10, 25, TQ_NW_I1>(ctx, P1, tmap, tmap2, "iw2tau_pass1")));
        TQ_TRY((launch_pipe<M_I2, 400, 20, 20, TQ_NW_I2>(ctx, P2, tmap, tmap2, "iw2tau_pass2")));
      } else {
        TQ_TRY(tq_launch_pipe_planned(ctx, M_I1, T1.heavy, P1, tmap, tmap2, "iw2tau_pass1"));
        TQ_TRY(tq_launch_pipe_planned(ctx, M_I2, T2.heavy, P2, tmap, tmap2, "iw2tau_pass2"));
      }
    }
  }
  *handled = true;
  return TQ_OK;
}
