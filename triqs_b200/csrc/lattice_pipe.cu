// TMA-pipelined batched DFT along one mesh axis (cyclat <-> brzone, the _fourier_base seam): the fast path of lattice.cu
// for axis lengths with a compile-time two-pass plan (256 = 16 x 16, 128 = 8 x 16, 64 = 8 x 8).
//
// Restates c++/triqs/gfs/transform/fourier_common.cpp:25-45 / fourier_lattice.cpp:30-59 exactly like lattice.cu; only the
// machine mapping differs.  The array is [outer][N][inner] (inner contiguous).  A tile is [N][TC] adjacent inner columns:
//   producer warp : ONE 3-D tensor-map TMA load per tile into a ring of three shared-memory buffers (mbarrier complete_tx);
//   worker warps  : pull 32-butterfly packets (pass A in place, then pass B whose outputs go straight to global memory,
//                   scaled on the way when this is the last axis of a k -> r transform), same scheduler as fourier_pipe.cu.
// No thread ever waits on a global load, the FFT of a tile overlaps the loads of the next two, and every global access is a
// TC*16-byte run.  Bound: HBM (read + write 16 B per element per axis; ~25 FP64 instructions per element).
#include "tq_pipe.cuh"

#include <algorithm>
#include <cstring>

#ifndef TQ_LP_NBUF
#define TQ_LP_NBUF 3
#endif
constexpr int LP_NBUF = TQ_LP_NBUF;

struct LatPipeArgs {
  cd *out;
  const cd *twq; // [RB][RA] pass-A twiddle rows
  long inner;    // contiguous elements per (outer, n)
  long n_tiles;  // outer * tiles
  int tiles;     // column tiles per outer index
  int TC;
  double scale;
  // planned (run-time radix pair) instantiation
  int tN, tRA, tRB, rb_generic;
  int boxes, box_rows; // the tile is fetched as `boxes` tensor boxes of box_rows rows (a TMA box has <= 256 rows)
  const cd *gtw;       // [tRB] exp(+2 pi i j / tRB) for the direct-sum pass B
  // fused pointwise product on the way out (MUL instantiations; the Bluestein convolution of fourier_matsubara.cu multiplies by the
  // four-step twiddles / the kernel spectrum between its axis transforms):  out(o, k, i) *= mul[(o % mul_omod) * mul_os + k * mul_rs + i / mul_cdiv]
  const cd *mul;
  int mul_omod, mul_os, mul_rs, mul_cdiv, mul_conj; // mul_omod = 0: no outer term; mul_cdiv = 0: no column term
};

template <int N, int RA> struct PostLattice {
  cd *base; // out + o*N*inner + c0
  long inner;
  double scale;
  bool scaled;
  cd *p;
  int ra_rt = 0; // RA == 0: run-time radix
  TQ_D int ra() const { return RA ? RA : ra_rt; }
  TQ_D void begin(int sidx, int c) { p = base + (long)sidx * inner + c; }
  template <int TT> TQ_D void store(cd v) const { store_rt(TT, v); }
  TQ_D void store_rt(int t, cd v) const {
    if (scaled) v = cscale(scale, v);
    p[(long)(ra() * t) * inner] = v;
  }
};
// ... with the fused pointwise product (planned instantiation only)
struct PostLatticeMul {
  cd *base;
  long inner;
  const cd *mul; // mul + outer term
  int mul_rs, mul_cdiv, mul_conj, c0;
  int ra_rt;
  cd *p;
  const cd *m;
  TQ_D int ra() const { return ra_rt; }
  TQ_D void begin(int sidx, int c) {
    p = base + (long)sidx * inner + c;
    m = mul + (long)sidx * mul_rs + (mul_cdiv ? (c0 + c) / mul_cdiv : 0);
  }
  template <int TT> TQ_D void store(cd v) const { store_rt(TT, v); }
  TQ_D void store_rt(int t, cd v) const {
    cd f = __ldg(m + (long)(ra_rt * t) * mul_rs);
    if (mul_conj) f = cconj(f);
    p[(long)(ra_rt * t) * inner] = cmul(v, f);
  }
};

template <int N, int RA, int RB, int SGN, int NW, int SET = 0, bool MUL = false>
__global__ void __maxnreg__((16384 / (((NW + 1) + 3) / 4 * 32)) & ~7) lattice_pipe_kernel(const __grid_constant__ LatPipeArgs A,
                                                                                           const __grid_constant__ CUtensorMap tmap) {
  constexpr int NT = (NW + 1) * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int TC = A.TC;
  constexpr bool PLANNED = (N == 0);
  const int NN = PLANNED ? A.tN : N, RAv = PLANNED ? A.tRA : RA, RBv = PLANNED ? A.tRB : RB;
  const bool rb_generic = PLANNED && A.rb_generic != 0;
  const size_t buf_stride = (sizeof(cd) * (size_t)NN * TC + 127) & ~size_t(127);
  cd *twS = reinterpret_cast<cd *>(smem_raw + LP_NBUF * buf_stride);
  cd *gtwS = twS + NN; // [RB] (direct-sum pass B only)
  uint64_t *full = reinterpret_cast<uint64_t *>(gtwS + (rb_generic ? RBv : 0));
  uint64_t *adone = full + LP_NBUF;
  uint64_t *freeb = adone + LP_NBUF;
  int *next_pk = reinterpret_cast<int *>(freeb + LP_NBUF);
  volatile int *ready_tile = next_pk + 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bgroups = rb_generic ? fft2_generic_groups(RBv) : 1;
  const int pa = (RBv * TC + 31) >> 5, pb = (RAv * TC * bgroups + 31) >> 5;
  for (int i = tid; i < NN; i += NT) twS[i] = A.twq[i];
  if (rb_generic)
    for (int i = tid; i < RBv; i += NT) gtwS[i] = A.gtw[i];
  if (tid == 0) {
    for (int b = 0; b < LP_NBUF; ++b) {
      mbar_init(&full[b], 1);
      mbar_init(&adone[b], pa);
      mbar_init(&freeb[b], pb);
    }
    *next_pk = 0;
    *ready_tile = -1;
    fence_mbar_init();
  }
  __syncthreads();
  const int nloc = (int)((A.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const unsigned long long m_tiles = A.tiles > 1 ? ~0ull / (unsigned long long)A.tiles + 1ull : 0ull;
  auto tile_of = [&](long t, long *o, int *c0) {
    *o = A.tiles > 1 ? (long)__umul64hi((unsigned long long)t, m_tiles) : t;
    *c0 = (int)(t - *o * A.tiles) * TC;
  };

  if (warp == NW) {
    for (int i = 0; i < nloc; ++i) {
      const long t = blockIdx.x + (long)i * gridDim.x;
      const int b = i % LP_NBUF, u = i / LP_NBUF;
      if (u > 0) mbar_wait(&freeb[b], (uint32_t)((u - 1) & 1));
      if (lane == 0) {
        *ready_tile = i;
        long o;
        int c0;
        tile_of(t, &o, &c0);
        mbar_arrive_expect_tx(&full[b], (uint32_t)(sizeof(cd) * NN * TC)); // the whole box; columns past `inner` are zero-filled
        const int nbox = PLANNED ? A.boxes : 1, brows = PLANNED ? A.box_rows : N;
        for (int bx = 0; bx < nbox; ++bx)
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                          smem_u32(smem_raw + b * buf_stride + sizeof(cd) * (size_t)bx * brows * TC)),
                       "l"(&tmap), "r"(2 * c0), "r"(bx * brows), "r"((int)o), "r"(smem_u32(&full[b]))
                       : "memory");
      }
    }
  } else {
    const unsigned m_pa = fastdiv_magic((unsigned)pa), m_pab = fastdiv_magic((unsigned)(pa + pb));
    const bool scaled = A.scale != 1.0;
    for (;;) {
      int pk = 0;
      if (lane == 0) pk = atomicAdd(next_pk, 1);
      pk = __shfl_sync(0xffffffffu, pk, 0);
      int blk, r;
      bool is_b = false;
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
      const int k = is_b ? blk - 2 : blk;
      if (k >= nloc) continue;
      const long t = blockIdx.x + (long)k * gridDim.x;
      const int b = k % LP_NBUF, u = k / LP_NBUF;
      long o;
      int c0;
      tile_of(t, &o, &c0);
      const int w = (int)min((long)TC, A.inner - c0);
      cd *tile = reinterpret_cast<cd *>(smem_raw + b * buf_stride);
      const int it = r * 32 + lane;
      while (*ready_tile < k) __nanosleep(32); // (see fourier_pipe.cu: never look at a barrier more than one phase ahead)
      if (!is_b) {
        mbar_wait(&full[b], (uint32_t)(u & 1));
        PreIdentity pre;
        if constexpr (!PLANNED) {
          fft2_pass_a<N, RA, RB, SGN, false>(tile, TC, w, twS, it, 1 << 30, pre);
        } else {
          const unsigned m_w = fastdiv_magic((unsigned)w);
#define TQ_CALL_A(R) fft2_item_a<R, SGN, false>(tile, TC, w, RBv, twS, it, pre, m_w);
          if constexpr (SET == 1) {
            TQ_RADIX_SWITCH_X(RAv, TQ_CALL_A)
          } else {
            TQ_RADIX_SWITCH32(RAv, TQ_CALL_A)
          }
#undef TQ_CALL_A
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&adone[b]);
      } else {
        mbar_wait(&adone[b], (uint32_t)(u & 1));
        if constexpr (MUL) {
          static_assert(!MUL || (N == 0 && SET == 0), "the fused product exists for the planned instantiation only");
          PostLatticeMul post{A.out + (o * NN) * A.inner + c0, A.inner, A.mul + (A.mul_omod ? (long)(o % A.mul_omod) * A.mul_os : 0L), A.mul_rs, A.mul_cdiv,
                              A.mul_conj, c0, RAv, nullptr, nullptr};
          const unsigned m_w = fastdiv_magic((unsigned)w);
#define TQ_CALL_B(R) fft2_item_b<R, SGN>(tile, TC, w, RAv, it, post, m_w);
          TQ_RADIX_SWITCH32(RBv, TQ_CALL_B)
#undef TQ_CALL_B
        } else {
        PostLattice<N, RA> post{A.out + (o * NN) * A.inner + c0, A.inner, A.scale, scaled, nullptr, RAv};
        if constexpr (!PLANNED) {
          fft2_pass_b<N, RA, RB, SGN>(tile, TC, w, it, 1 << 30, post);
        } else {
          const unsigned m_w = fastdiv_magic((unsigned)w);
          if (rb_generic) {
            fft2_item_b_generic<SGN>(tile, TC, w, RAv, RBv, gtwS, it, post, m_w);
          } else {
#define TQ_CALL_B(R) fft2_item_b<R, SGN>(tile, TC, w, RAv, it, post, m_w);
            if constexpr (SET == 1) {
              TQ_RADIX_SWITCH_X(RBv, TQ_CALL_B)
            } else {
              TQ_RADIX_SWITCH32(RBv, TQ_CALL_B)
            }
#undef TQ_CALL_B
          }
        }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&freeb[b]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
struct LatTwiddles {
  std::map<std::tuple<int, int>, cd *> tw; // (64 N + RA, sign) -> device [RB][RA]; (-RB, 0) -> exp(+2 pi i j / RB)
  ~LatTwiddles() {
    for (auto &kv : tw) cudaFree(kv.second);
  }
};

static size_t lat_smem_bytes(int N, int RB, bool generic, long tc) {
  const size_t buf_stride = (sizeof(cd) * (size_t)N * tc + 127) & ~size_t(127);
  return LP_NBUF * buf_stride + sizeof(cd) * (N + (generic ? RB : 0)) + 8 * 3 * LP_NBUF + 16;
}
template <int N, int RA, int RB, int SGN, int SET = 0, bool MUL = false> static tq_status launch_lat(tq_ctx *ctx, LatPipeArgs A, const CUtensorMap &tmap) {
  constexpr int NW = 11;
  const size_t smem = lat_smem_bytes(N ? N : A.tN, N ? RB : A.tRB, N ? false : A.rb_generic != 0, A.TC);
  auto k = lattice_pipe_kernel<N, RA, RB, SGN, NW, SET, MUL>;
  TQ_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<long>(A.n_tiles, (long)ctx->sm_count);
  {
    KernelScope ks(ctx, "fft_axis_pipe");
    k<<<grid, (NW + 1) * 32, smem, ctx->stream>>>(A, tmap);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  return TQ_OK;
}

// N0 / RA0 / RB0 = 0: the planned instantiation with the run-time shape T
struct LatMul { // fused pointwise product of the planned instantiation (LatPipeArgs::mul ...)
  const cd *tab = nullptr;
  int omod = 0, os = 0, rs = 0, cdiv = 0, conj = 0;
};
template <int N0, int RA0, int RB0, int SET = 0>
static tq_status run_lat(tq_ctx *ctx, const TileShape &T, long outer, long inner, const cd *in, cd *out, int sign, double scale, bool *handled,
                         const LatMul *mul = nullptr) {
  const int N = T.N, RA = T.RA, RB = T.RB;
  const size_t budget = (size_t)ctx->max_smem_optin - sizeof(cd) * (N + RB) - 256;
  long tc = (long)(budget / LP_NBUF / (sizeof(cd) * N));
  tc = std::min<long>(tc, 128);
  if (tc >= inner)
    tc = inner; // all columns in ONE tile (inner = 25: a 24-column tile would leave a second tile with a single column)
  else if (tc >= 8)
    tc &= ~7L; // 128-byte runs
  if (tc < inner) { // balance the column tiles (inner = 200, tc = 128: 2 x 104 instead of 128 + 72)
    const long nt = (inner + tc - 1) / tc;
    long bal = (inner + nt - 1) / nt;
    if (bal >= 8) bal = (bal + 7) & ~7L;
    tc = std::min(tc, bal);
  }
  int boxes = 1;
  while (N % boxes != 0 || N / boxes > 256) ++boxes;
  if (boxes > 1 && ((long)(N / boxes) * tc) % 8 != 0 && tc > 8) tc &= ~7L; // every box must land 128-byte aligned: whole 128-byte rows do
  if (boxes > 1 && ((long)(N / boxes) * tc) % 8 != 0) return TQ_OK;
  if (const char *e = TQ_DEV_ENV("TQ_LP_TC")) tc = std::min<long>(tc, atol(e)); // (tuning experiments)
  if (tc < 1) return TQ_OK;
  // pointers: TMA needs 16-byte alignment (always true for complex<double> arrays from cudaMalloc / numpy)
  if (((uintptr_t)in | (uintptr_t)out) & 15) return TQ_OK;
  if (2 * inner >= (1L << 32) || outer >= (1L << 31)) return TQ_OK;
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return TQ_OK;
  if (!ctx->lat_tw) ctx->lat_tw = std::make_shared<LatTwiddles>();
  cd *&twd = ctx->lat_tw->tw[std::make_tuple(N * 64 + RA, sign)]; // (the rows depend on the radix split, and a length can have two: heavy radices allowed or not)
  if (!twd) {
    std::vector<cd> w, rows((size_t)N);
    tq_fft_twiddles_host(N, w);
    for (int q = 0; q < RB; ++q)
      for (int s = 0; s < RA; ++s) rows[(size_t)q * RA + s] = sign > 0 ? w[(q * s) % N] : cconj(w[(q * s) % N]);
    TQ_CUDA(ctx, cudaMalloc(&twd, sizeof(cd) * N));
    TQ_CUDA(ctx, cudaMemcpyAsync(twd, rows.data(), sizeof(cd) * N, cudaMemcpyHostToDevice, ctx->stream));
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  CUtensorMap tmap;
  cuuint64_t dims[3] = {(cuuint64_t)(2 * inner), (cuuint64_t)N, (cuuint64_t)outer};
  cuuint64_t strides[2] = {(cuuint64_t)(inner * sizeof(cd)), (cuuint64_t)((long)N * inner * sizeof(cd))};
  cuuint32_t box[3] = {(cuuint32_t)(2 * tc), (cuuint32_t)(N / boxes), 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tq_fail(ctx, TQ_ERR_CUDA, "cuTensorMapEncodeTiled (lattice) failed with CUresult " + std::to_string((int)r));
  LatPipeArgs A{};
  A.out = out;
  A.twq = twd;
  A.inner = inner;
  A.TC = (int)tc;
  A.tiles = (int)((inner + tc - 1) / tc);
  A.n_tiles = outer * A.tiles;
  A.scale = scale;
  A.tN = N;
  A.tRA = RA;
  A.tRB = RB;
  A.rb_generic = T.generic ? 1 : 0;
  A.boxes = boxes;
  A.box_rows = N / boxes;
  if (T.generic) {
    cd *&g = ctx->lat_tw->tw[std::make_tuple(-RB, 0)];
    if (!g) {
      std::vector<cd> w;
      tq_fft_twiddles_host(RB, w);
      TQ_CUDA(ctx, cudaMalloc(&g, sizeof(cd) * RB));
      TQ_CUDA(ctx, cudaMemcpyAsync(g, w.data(), sizeof(cd) * RB, cudaMemcpyHostToDevice, ctx->stream));
      TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    A.gtw = g;
  }
  if (A.n_tiles >= (1L << 31)) return TQ_OK;
  if constexpr (N0 == 0 && SET == 0) {
    if (mul) {
      if (T.generic || scale != 1.0) return TQ_OK; // (not handled: the caller multiplies in a pass of its own)
      A.mul = mul->tab;
      A.mul_omod = mul->omod;
      A.mul_os = mul->os;
      A.mul_rs = mul->rs;
      A.mul_cdiv = mul->cdiv;
      A.mul_conj = mul->conj;
      if (sign > 0)
        TQ_TRY((launch_lat<0, 0, 0, 1, 0, true>(ctx, A, tmap)));
      else
        TQ_TRY((launch_lat<0, 0, 0, -1, 0, true>(ctx, A, tmap)));
      *handled = true;
      return TQ_OK;
    }
  }
  if (mul) return TQ_OK;
  if (sign > 0)
    TQ_TRY((launch_lat<N0, RA0, RB0, 1, SET>(ctx, A, tmap)));
  else
    TQ_TRY((launch_lat<N0, RA0, RB0, -1, SET>(ctx, A, tmap)));
  *handled = true;
  return TQ_OK;
}

// Axis transform with a fused pointwise product on the output (Bluestein): only lengths the planned (smooth, non-heavy) instantiation takes
tq_status tq_lattice_pipe_axis_mul(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, const cd *tab, int omod, int os,
                                   int rs, int cdiv, int conj, bool *handled) {
  *handled = false;
  if (ctx->opt_no_pipe) return TQ_OK;
  TileShape T;
  if (N < 4 || N > 1024 || !tile_shape((int)N, &T, false) || T.generic || T.heavy) return TQ_OK;
  LatMul m;
  m.tab = tab;
  m.omod = omod;
  m.os = os;
  m.rs = rs;
  m.cdiv = cdiv;
  m.conj = conj;
  return run_lat<0, 0, 0, 0>(ctx, T, outer, inner, in, out, sign, 1.0, handled, &m);
}

// DFT along the middle axis of [outer][N][inner]; *handled = false when there is no specialisation for N (caller falls back)
tq_status tq_lattice_pipe_axis(tq_ctx *ctx, long outer, long N, long inner, const cd *in, cd *out, int sign, double scale, bool *handled) {
  *handled = false;
  if (ctx->opt_no_pipe) return TQ_OK;
  // compile-time instantiations of the benchmark shapes (256 = 16 x 16, 128 = 8 x 16, 64 = 8 x 8) ...
  switch (N) {
    case 256: return run_lat<256, 16, 16>(ctx, TileShape{256, 16, 16, false, false, 2.0}, outer, inner, in, out, sign, scale, handled);
    case 128: return run_lat<128, 8, 16>(ctx, TileShape{128, 8, 16, false, false, 2.0}, outer, inner, in, out, sign, scale, handled);
    case 64: return run_lat<64, 8, 8>(ctx, TileShape{64, 8, 8, false, false, 2.0}, outer, inner, in, out, sign, scale, handled);
    default: break;
  }
  // ... and the planned one for every other axis length that splits into a register radix and a second radix (register or direct
  // sum): 32, 48, 96, 100, 512, 3 2^k, primes up to 127 ...   (fourier_common.cpp:30-44: FFTW takes any length)
  TileShape T;
  if (N < 4 || N > 1024 || !tile_shape((int)N, &T, true)) return TQ_OK;
  // (the Rader primes 11, 13, 17, 41 live in their own instantiation: their spills stay out of the kernel the smooth axes use)
  const bool primes = T.heavy && (T.RA != 32 && T.RB != 32 ? true : (tq_radix_is_heavy(T.RA) && T.RA != 32) || (tq_radix_is_heavy(T.RB) && T.RB != 32));
  if (primes) return run_lat<0, 0, 0, 1>(ctx, T, outer, inner, in, out, sign, scale, handled);
  return run_lat<0, 0, 0, 0>(ctx, T, outer, inner, in, out, sign, scale, handled);
}
