// Compile-time two-pass tile FFT, N = RA * RB, for the pipelined (TMA double-buffered) kernels.
//
// The tile s[n][c] (n in [0,N) transform axis, c in [0,tcw) columns, row pitch `pitch` >= tcw) sits in shared memory.
//   pass A (in place): for every q in [0,RB): radix-RA butterfly over the rows q + RB r, times w_N^{SGN q s};
//                      result s in [0,RA) goes back to row q + RB s.  The `Pre` functor produces the input
//                      values (it reads the raw tile element and applies the fused pre-transform), so the
//                      raw data the TMA engine dropped into the tile is touched exactly once.
//   pass B (read only): for every s in [0,RA): radix-RB butterfly over the contiguous rows s RB + t; output
//                      t is X[s + RA t] and is handed to the `Post` functor, which applies the fused epilogue
//                      and stores straight to global memory (no write-back to the tile, no digit-reversal pass).
// Both radices are template parameters: every stride is a multiple of a run-time tile width only, the
// butterflies are fully unrolled in registers, and there is one division per work item (item -> row, column).
// Work items are (butterfly, column) pairs with the column fastest, so a warp touches contiguous shared
// memory (conflict free) and reads its twiddles / row factors as broadcasts.
//
// __host__ __device__ so tests/hostsim can run the same code on the CPU box.
#pragma once
#include "tq_fft.cuh"
#include "tq_fft_rader.cuh"


// compile-time loops (the functors take the butterfly leg as a template argument so that per-leg constants
// index kernel parameters with immediate offsets)
template <int R, int I = 0, class Pre> TQ_HD void tq_unroll_load(cd *a, Pre &pre, const cd *p, int stride) {
  if constexpr (I < R) {
    a[I] = pre.template load<I>(p + I * stride);
    tq_unroll_load<R, I + 1>(a, pre, p, stride);
  }
}
template <int R, int I = 0, class Post> TQ_HD void tq_unroll_store(const cd *a, Post &post) {
  if constexpr (I < R) {
    post.template store<I>(a[I]);
    tq_unroll_store<R, I + 1>(a, post);
  }
}

// Developer knock-out switches (compiled in only with -DTQ_KNOCKOUT, never in the product): remove one ingredient of the
// packet loops at a time to see what the tile period is made of.  Results are wrong by construction.
#ifdef TQ_KNOCKOUT
__device__ int g_tq_knock = 0;
#endif
#if defined(TQ_KNOCKOUT) && defined(__CUDA_ARCH__)
#define TQ_KNOCK(bit) (g_tq_knock & (bit))
#else
#define TQ_KNOCK(bit) 0
#endif

struct NoHook {
  TQ_HD void operator()() const {}
};

struct PreIdentity {
  TQ_HD void begin(int, int) {}
  template <int R> TQ_HD cd load(const cd *p) const { return *p; }
};

// twq[q * RA + s] = w_N^{SGN q s} (times whatever per-q / per-s factor the caller folded in).  TW0: entry s = 0 is not 1.
// `before_store` runs once per butterfly between the arithmetic and the stores (the pipelined kernels fetch their next work
// packet there, so that the fetch latency overlaps the stores).
template <int N, int RA, int RB, int SGN, bool TW0, class Pre, class Hook = NoHook>
TQ_HD void fft2_pass_a(cd *s, int pitch, int tcw, const cd *__restrict__ twq, int tid, int nthreads, Pre &pre, unsigned magic = ~0u,
                       const Hook &before_store = Hook()) {
  static_assert(RA * RB == N, "N = RA * RB");
  const int nitems = RB * tcw;
  const int stride = RB * pitch;
  if (magic == ~0u) magic = fastdiv_magic((unsigned)tcw); // (a 64-bit division: hot callers pass the precomputed value)
  for (int i = tid; i < nitems; i += nthreads) {
    const int q = fastdiv(i, magic);
    const int c = i - q * tcw;
    cd *p = s + q * pitch + c; // row q, column c
    pre.begin(q, c);
    cd a[RA];
    if (!TQ_KNOCK(32)) {
      tq_unroll_load<RA>(a, pre, p, stride);
    } else {
#pragma unroll
      for (int r = 0; r < RA; ++r) a[r] = cmk(1.0 + i, (double)r);
    }
    if (!TQ_KNOCK(1)) Dft<RA, SGN>::run(a);
    const cd *tw = twq + q * RA;
    if ((TW0 || q != 0) && !TQ_KNOCK(16)) {
#pragma unroll
      for (int r = TW0 ? 0 : 1; r < RA; ++r) a[r] = cmul(a[r], tw[r]);
    }
    before_store();
    if (!TQ_KNOCK(2)) {
#pragma unroll
      for (int r = 0; r < RA; ++r) p[r * stride] = a[r];
    } else if (a[0].x == 1.2345e300) {
      p[0] = a[1];
    }
  }
}

template <int N, int RA, int RB, int SGN, class Post, class Hook = NoHook>
TQ_HD void fft2_pass_b(const cd *s, int pitch, int tcw, int tid, int nthreads, Post &post, unsigned magic = ~0u,
                       const Hook &before_store = Hook()) {
  static_assert(RA * RB == N, "N = RA * RB");
  const int nitems = RA * tcw;
  if (magic == ~0u) magic = fastdiv_magic((unsigned)tcw);
  for (int i = tid; i < nitems; i += nthreads) {
    const int sidx = fastdiv(i, magic);
    const int c = i - sidx * tcw;
    const cd *p = s + sidx * (RB * pitch) + c;
    post.begin(sidx, c);
    cd a[RB];
    if (!TQ_KNOCK(8)) {
#pragma unroll
      for (int t = 0; t < RB; ++t) a[t] = p[t * pitch];
    } else {
#pragma unroll
      for (int t = 0; t < RB; ++t) a[t] = cmk(1.0 + i, (double)t);
    }
    if (!TQ_KNOCK(1)) Dft<RB, SGN>::run(a);
    before_store();
    if (!TQ_KNOCK(4)) {
      tq_unroll_store<RB>(a, post);
    } else {
      cd acc = a[0];
#pragma unroll
      for (int t = 1; t < RB; ++t) acc = cadd(acc, a[t]);
      if (acc.x == 1.2345e300) post.template store<0>(acc);
    }
  }
}

// ---- two butterflies per lane ("dual" packets of 64) ---------------------------------------------------------------------
// The pipelined kernels with 8 fat warps (255 registers) run TWO independent butterflies per lane: both sets of loads are issued
// first, the stores of the first butterfly overlap the arithmetic of the second, so a warp hides its own shared-memory / global
// latencies instead of relying on the other two warps of its scheduler being in a different phase.  The second item is i0 + 32; when
// it is past the end it recomputes the first one with its stores predicated off (straight-line code, no divergent phases).
template <int N, int RA, int RB, int SGN, bool TW0, class Pre, class Hook = NoHook>
TQ_HD void fft2_pass_a_dual(cd *s, int pitch, int tcw, const cd *__restrict__ twq, int i0, Pre &pre0, Pre &pre1, unsigned magic,
                            const Hook &before_store = Hook()) {
  static_assert(RA * RB == N, "N = RA * RB");
  const int nitems = RB * tcw;
  const int stride = RB * pitch;
  if (i0 >= nitems) return;
  if (magic == ~0u) magic = fastdiv_magic((unsigned)tcw);
  const bool v1 = i0 + 32 < nitems;
  const int i1 = v1 ? i0 + 32 : i0;
  const int q0 = fastdiv(i0, magic), q1 = fastdiv(i1, magic);
  const int c0 = i0 - q0 * tcw, c1 = i1 - q1 * tcw;
  cd *p0 = s + q0 * pitch + c0, *p1 = s + q1 * pitch + c1;
  cd a[RA], b[RA];
  pre0.begin(q0, c0);
  tq_unroll_load<RA>(a, pre0, p0, stride);
  pre1.begin(q1, c1);
  tq_unroll_load<RA>(b, pre1, p1, stride);
  Dft<RA, SGN>::run(a);
  {
    const cd *tw = twq + q0 * RA;
    if (TW0 || q0 != 0) {
#pragma unroll
      for (int r = TW0 ? 0 : 1; r < RA; ++r) a[r] = cmul(a[r], tw[r]);
    }
  }
  before_store();
  // (the second butterfly's inputs are in registers: when it duplicates the first one, overwriting its rows here is harmless)
#pragma unroll
  for (int r = 0; r < RA; ++r) p0[r * stride] = a[r];
  Dft<RA, SGN>::run(b);
  {
    const cd *tw = twq + q1 * RA;
    if (TW0 || q1 != 0) {
#pragma unroll
      for (int r = TW0 ? 0 : 1; r < RA; ++r) b[r] = cmul(b[r], tw[r]);
    }
  }
  if (v1) {
#pragma unroll
    for (int r = 0; r < RA; ++r) p1[r * stride] = b[r];
  }
}

template <int N, int RA, int RB, int SGN, class Post, class Hook = NoHook>
TQ_HD void fft2_pass_b_dual(const cd *s, int pitch, int tcw, int i0, Post &post0, Post &post1, unsigned magic, const Hook &before_store = Hook()) {
  static_assert(RA * RB == N, "N = RA * RB");
  const int nitems = RA * tcw;
  if (i0 >= nitems) return;
  if (magic == ~0u) magic = fastdiv_magic((unsigned)tcw);
  const bool v1 = i0 + 32 < nitems;
  const int i1 = v1 ? i0 + 32 : i0;
  const int s0 = fastdiv(i0, magic), s1 = fastdiv(i1, magic);
  const int c0 = i0 - s0 * tcw, c1 = i1 - s1 * tcw;
  const cd *p0 = s + s0 * (RB * pitch) + c0, *p1 = s + s1 * (RB * pitch) + c1;
  cd a[RB], b[RB];
#pragma unroll
  for (int t = 0; t < RB; ++t) a[t] = p0[t * pitch];
#pragma unroll
  for (int t = 0; t < RB; ++t) b[t] = p1[t * pitch];
  post0.begin(s0, c0);
  Dft<RB, SGN>::run(a);
  before_store();
  tq_unroll_store<RB>(a, post0);
  post1.begin(s1, c1);
  Dft<RB, SGN>::run(b);
  if (v1) tq_unroll_store<RB>(b, post1);
}

// Pass A when only the first and the last leg of every butterfly are non-zero (inverse Matsubara transform: the frequency
// window occupies rows [0, RB) and [N - RB, N) of the tile, everything between is zero-fill):
//   y_s = x_0 + x_{RA-1} w_RA^{SGN (RA-1) s} = x_0 + x_{RA-1} wedge[s],   wedge[s] = exp(-SGN 2 pi i s / RA)
// The last-leg rows come from a separate staging area `hi` (same pitch), so the in-place write of all RA legs is hazard free.
template <int N, int RA, int RB, bool TW0, class Pre, class Wedge>
TQ_HD void fft2_pass_a_edge(cd *s, const cd *hi, int pitch, int tcw, const cd *__restrict__ twq, int tid, int nthreads, Pre &pre, const Wedge &wedge,
                            unsigned magic = ~0u) {
  static_assert(RA * RB == N, "N = RA * RB");
  const int nitems = RB * tcw;
  const int stride = RB * pitch;
  if (magic == ~0u) magic = fastdiv_magic((unsigned)tcw);
  for (int i = tid; i < nitems; i += nthreads) {
    const int q = fastdiv(i, magic);
    const int c = i - q * tcw;
    cd *p = s + q * pitch + c;
    pre.begin(q, c);
    const cd x0 = pre.edge(0, p), x1 = pre.edge(1, hi + q * pitch + c);
    const cd *tw = twq + q * RA;
    cd a[RA];
    a[0] = cadd(x0, x1);
    if (TW0) a[0] = cmul(a[0], tw[0]);
#pragma unroll
    for (int r = 1; r < RA; ++r) {
      a[r] = cfma(x1, wedge(r), x0);
      if (TW0 || q != 0) a[r] = cmul(a[r], tw[r]);
    }
#pragma unroll
    for (int r = 0; r < RA; ++r) p[r * stride] = a[r];
  }
}

// Pass B of ONE butterfly, split at the point where the tile has been read into registers (the pipelined kernels release
// the tile buffer there, one stage before the epilogue is done).
template <int N, int RA, int RB> TQ_HD bool fft2_b_load(const cd *s, int pitch, int tcw, int item, cd *a, int *sidx, int *c) {
  if (item >= RA * tcw) return false;
  *sidx = fastdiv(item, fastdiv_magic((unsigned)tcw));
  *c = item - *sidx * tcw;
  const cd *p = s + *sidx * (RB * pitch) + *c;
#pragma unroll
  for (int t = 0; t < RB; ++t) a[t] = p[t * pitch];
  return true;
}
template <int N, int RA, int RB, int SGN, class Post> TQ_HD void fft2_b_finish(cd *a, int sidx, int c, Post &post) {
  post.begin(sidx, c);
  Dft<RB, SGN>::run(a);
  tq_unroll_store<RB>(a, post);
}

// -------------------------------------------------------------------------------------------------
// Planned engine: the same two passes for ANY tile length N = RA * RB.  One butterfly (`item`) per call -- the pipelined
// kernels hand out packets of 32 butterflies, one per lane.  The radix whose butterfly is unrolled in registers is the
// template parameter (pass A: RA, pass B: RB); the other one only enters strides and is a run-time value, so a kernel can
// `switch` over a radix table (TQ_RADIX_SWITCH) and serve every pair without one instantiation per pair.
// -------------------------------------------------------------------------------------------------
// radices with a register butterfly (Dft<R, SGN>)
#define TQ_RADIX_SWITCH(r, CALL)                                                                                                 \
  switch (r) {                                                                                                                   \
    case 1: CALL(1) break;                                                                                                       \
    case 2: CALL(2) break;                                                                                                       \
    case 3: CALL(3) break;                                                                                                       \
    case 4: CALL(4) break;                                                                                                       \
    case 5: CALL(5) break;                                                                                                       \
    case 6: CALL(6) break;                                                                                                       \
    case 7: CALL(7) break;                                                                                                       \
    case 8: CALL(8) break;                                                                                                       \
    case 9: CALL(9) break;                                                                                                       \
    case 10: CALL(10) break;                                                                                                     \
    case 12: CALL(12) break;                                                                                                     \
    case 15: CALL(15) break;                                                                                                     \
    case 16: CALL(16) break;                                                                                                     \
    case 20: CALL(20) break;                                                                                                     \
    case 25: CALL(25) break;                                                                                                     \
    default: break;                                                                                                              \
  }
// base set + radix 32: the default instantiation of the lattice kernel (power-of-two axes 512, 1024)
#define TQ_RADIX_SWITCH32(r, CALL)                                                                                                 \
  switch (r) {                                                                                                                   \
    case 1: CALL(1) break;                                                                                                       \
    case 2: CALL(2) break;                                                                                                       \
    case 3: CALL(3) break;                                                                                                       \
    case 4: CALL(4) break;                                                                                                       \
    case 5: CALL(5) break;                                                                                                       \
    case 6: CALL(6) break;                                                                                                       \
    case 7: CALL(7) break;                                                                                                       \
    case 8: CALL(8) break;                                                                                                       \
    case 9: CALL(9) break;                                                                                                       \
    case 10: CALL(10) break;                                                                                                     \
    case 12: CALL(12) break;                                                                                                     \
    case 15: CALL(15) break;                                                                                                     \
    case 16: CALL(16) break;                                                                                                     \
    case 20: CALL(20) break;                                                                                                     \
    case 25: CALL(25) break;                                                                                                     \
    case 32: CALL(32) break;                                                                                                     \
    default: break;                                                                                                              \
  }
// ... plus the "heavy" radices: 32 (power-of-two lattice axes 512, 1024) and the primes 11, 13, 17, 41 by Rader's algorithm
// (tq_fft_rader.cuh; 41 | 6150, the default Python mesh).  Their butterflies need more than the 168 registers of the 12-warp
// kernels, so they live in separate kernel instantiations (SET = 1): the spills stay out of the kernels every smooth length uses.
#define TQ_RADIX_SWITCH_X(r, CALL)                                                                                               \
  switch (r) {                                                                                                                   \
    case 1: CALL(1) break;                                                                                                       \
    case 2: CALL(2) break;                                                                                                       \
    case 3: CALL(3) break;                                                                                                       \
    case 4: CALL(4) break;                                                                                                       \
    case 5: CALL(5) break;                                                                                                       \
    case 6: CALL(6) break;                                                                                                       \
    case 7: CALL(7) break;                                                                                                       \
    case 8: CALL(8) break;                                                                                                       \
    case 9: CALL(9) break;                                                                                                       \
    case 10: CALL(10) break;                                                                                                     \
    case 11: CALL(11) break;                                                                                                     \
    case 12: CALL(12) break;                                                                                                     \
    case 13: CALL(13) break;                                                                                                     \
    case 15: CALL(15) break;                                                                                                     \
    case 16: CALL(16) break;                                                                                                     \
    case 17: CALL(17) break;                                                                                                     \
    case 20: CALL(20) break;                                                                                                     \
    case 25: CALL(25) break;                                                                                                     \
    case 32: CALL(32) break;                                                                                                     \
    case 41: CALL(41) break;                                                                                                     \
    default: break;                                                                                                              \
  }
// (-DTQ_NO_RADER=mask, developer A/B: bit 0 drops radix 41, bit 1 radix 17, bit 2 radix 13, bit 3 radix 11 from the register set, so that the
//  planner takes the direct-sum pass B for them)
#ifndef TQ_NO_RADER
#define TQ_NO_RADER 0
#endif
TQ_HD bool tq_radix_is_heavy(int r) {
  return (r == 11 && !(TQ_NO_RADER & 8)) || (r == 13 && !(TQ_NO_RADER & 4)) || (r == 17 && !(TQ_NO_RADER & 2)) || r == 32 || (r == 41 && !(TQ_NO_RADER & 1));
}
TQ_HD bool tq_radix_in_registers(int r, bool allow_heavy = false) {
  return r == 1 || r == 2 || r == 3 || r == 4 || r == 5 || r == 6 || r == 7 || r == 8 || r == 9 || r == 10 || r == 12 || r == 15 || r == 16 ||
         r == 20 || r == 25 || (allow_heavy && tq_radix_is_heavy(r));
}

template <int RA, int SGN, bool TW0, class Pre, class Hook = NoHook>
TQ_HD void fft2_item_a(cd *s, int pitch, int tcw, int rb, const cd *__restrict__ twq, int item, Pre &pre, unsigned magic,
                       const Hook &before_store = Hook()) {
  if (item >= rb * tcw) return;
  const int stride = rb * pitch;
  const int q = fastdiv(item, magic);
  const int c = item - q * tcw;
  cd *p = s + q * pitch + c; // row q, column c
  pre.begin(q, c);
  cd a[RA];
  tq_unroll_load<RA>(a, pre, p, stride);
  Dft<RA, SGN>::run(a);
  const cd *tw = twq + q * RA;
  if (TW0 || q != 0) {
#pragma unroll
    for (int r = TW0 ? 0 : 1; r < RA; ++r) a[r] = cmul(a[r], tw[r]);
  }
  before_store();
#pragma unroll
  for (int r = 0; r < RA; ++r) p[r * stride] = a[r];
}

template <int RB, int SGN, class Post, class Hook = NoHook>
TQ_HD void fft2_item_b(const cd *s, int pitch, int tcw, int ra, int item, Post &post, unsigned magic, const Hook &before_store = Hook()) {
  if (item >= ra * tcw) return;
  const int sidx = fastdiv(item, magic);
  const int c = item - sidx * tcw;
  const cd *p = s + sidx * (RB * pitch) + c;
  post.begin(sidx, c);
  cd a[RB];
#pragma unroll
  for (int t = 0; t < RB; ++t) a[t] = p[t * pitch];
  Dft<RB, SGN>::run(a);
  before_store();
  tq_unroll_store<RB>(a, post);
}

// Pass B for a radix without a register butterfly (a prime 11, 13, 41, 101 ... or any rb <= TQ_MAX_GENERIC_RADIX): direct sum
//   X_k = sum_t x_t w_rb^{SGN t k}.  One work item = (output group kg, butterfly s, column c): TQ_GENERIC_KO outputs out of registers;
//   x_t is read once per item, the twiddle w_rb^{(t k) mod rb} comes from the table wtab[rb] (a broadcast: the lanes of a warp
//   differ in (butterfly, column), rarely in the group).  Splitting a butterfly over ceil(rb / KO) items is what keeps the warps busy when
//   a tile has few columns (a 101-point butterfly is 40 k FP64 instructions).  The tile is read-only in pass B, so no local
//   copy of the butterfly is needed.  4 rb FP64 instructions per element.
#define TQ_GENERIC_KO 8
TQ_HD int fft2_generic_groups(int rb) { return (rb + TQ_GENERIC_KO - 1) / TQ_GENERIC_KO; }
template <int SGN, class Post, class Hook = NoHook>
TQ_HD void fft2_item_b_generic(const cd *s, int pitch, int tcw, int ra, int rb, const cd *__restrict__ wtab, int item, Post &post, unsigned magic,
                               const Hook &before_store = Hook()) {
  constexpr int KO = TQ_GENERIC_KO;
  const int per_group = ra * tcw;
  if (item >= per_group * fft2_generic_groups(rb)) return;
  const int kg = item / per_group; // (one division per 4 rb KO FP64 instructions)
  const int rem = item - kg * per_group;
  const int sidx = fastdiv(rem, magic);
  const int c = rem - sidx * tcw;
  const cd *p = s + sidx * (rb * pitch) + c;
  post.begin(sidx, c);
  const int k0 = kg * KO;
  cd acc[KO];
  int idx[KO], step[KO];
#pragma unroll
  for (int u = 0; u < KO; ++u) {
    acc[u] = cmk(0.0, 0.0);
    idx[u] = 0;
    step[u] = k0 + u < rb ? k0 + u : 0; // outputs past rb are computed (as X_0) and not stored
  }
  for (int t = 0; t < rb; ++t) {
    const cd x = p[t * pitch];
#pragma unroll
    for (int u = 0; u < KO; ++u) {
      cd w = wtab[idx[u]];
      if (SGN < 0) w = cconj(w);
      acc[u] = cfma(x, w, acc[u]);
      idx[u] += step[u]; // (t k) mod rb, incrementally
      if (idx[u] >= rb) idx[u] -= rb;
    }
  }
  before_store(); // every read of the tile is done
#pragma unroll
  for (int u = 0; u < KO; ++u)
    if (k0 + u < rb) post.store_rt(k0 + u, acc[u]);
}
