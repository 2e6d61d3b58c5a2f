// TEST-ONLY host emulation of the device math of libtqgpu (never linked into the product library).
// The tile-FFT engine and the small dense kernels are __host__ __device__; here a CTA is emulated by
// calling each pass with (tid = 0, nthreads = 1), which is equivalent because passes are hazard free.
#include "../../triqs_b200/csrc/tq_fft.cuh"
#include <vector>
#include <cstring>


extern "C" int hostsim_fft(int N, int TC, int sign, int direct_tw, int skew, double *data /* [N][TC] complex */) {
  FftPlan P;
  if (!tq_fft_factorize(N, P)) return -1;
  std::vector<cd> tw;
  tq_fft_twiddles_host(N, tw);
  P.tw = tw.data();
  int sm = skew ? ~0 : 0;
  int rows = fft_rows_alloc(N, sm);
  std::vector<cd> s((size_t)rows * TC);
  cd *in = (cd *)data;
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < TC; ++c) s[(size_t)fft_row(n, sm) * TC + c] = in[(size_t)n * TC + c];
  // emulate a CTA of `nthreads` threads: passes are hazard free, so threads can run one after the other
  const int nthreads = 7;
  for (int p = 0; p < P.npass; ++p)
    for (int tid = 0; tid < nthreads; ++tid) {
      TileThread T = tile_thread(tid, nthreads, TC);
      if (sign > 0) {
        if (skew) fft_pass<1, true>(P, p, s.data(), TC, TC, P.tw, T, direct_tw != 0); else fft_pass<1, false>(P, p, s.data(), TC, TC, P.tw, T, direct_tw != 0);
      } else {
        if (skew) fft_pass<-1, true>(P, p, s.data(), TC, TC, P.tw, T, direct_tw != 0); else fft_pass<-1, false>(P, p, s.data(), TC, TC, P.tw, T, direct_tw != 0);
      }
    }
  for (int k = 0; k < N; ++k) {
    int pos = fft_pos(P, k);
    for (int c = 0; c < TC; ++c) in[(size_t)k * TC + c] = s[(size_t)fft_row(pos, sm) * TC + c];
  }
  return P.npass;
}

#include "../../triqs_b200/csrc/tq_fft2.cuh"
template <int RA> struct PostCollect {
  cd *out;
  int tcw, c, sidx;
  TQ_HD void begin(int s, int cc) { sidx = s; c = cc; }
  template <int T> TQ_HD void store(cd v) { out[(size_t)(sidx + RA * T) * tcw + c] = v; }
};
template <int N, int RA, int RB> static void run_fft2(int tcw, int sign, cd *data) {
  std::vector<cd> tw1, tw((size_t)N);
  tq_fft_twiddles_host(N, tw1);
  for (int q = 0; q < RB; ++q)
    for (int r = 0; r < RA; ++r) tw[(size_t)q * RA + r] = sign > 0 ? tw1[(q * r) % N] : cconj(tw1[(q * r) % N]);
  const int pitch = tcw + 2; // padded rows, like a partially filled TMA box
  std::vector<cd> s((size_t)N * pitch);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < tcw; ++c) s[(size_t)n * pitch + c] = data[(size_t)n * tcw + c];
  const int nthreads = 37;
  PreIdentity pre;
  PostCollect<RA> post{data, tcw, 0, 0};
  for (int tid = 0; tid < nthreads; ++tid) {
    if (sign > 0) fft2_pass_a<N, RA, RB, 1, false>(s.data(), pitch, tcw, tw.data(), tid, nthreads, pre);
    else fft2_pass_a<N, RA, RB, -1, false>(s.data(), pitch, tcw, tw.data(), tid, nthreads, pre);
  }
  for (int tid = 0; tid < nthreads; ++tid) {
    if (sign > 0) fft2_pass_b<N, RA, RB, 1>(s.data(), pitch, tcw, tid, nthreads, post);
    else fft2_pass_b<N, RA, RB, -1>(s.data(), pitch, tcw, tid, nthreads, post);
  }
}
// compile-time two-pass tile FFT of the pipelined kernels (tq_fft2.cuh)
extern "C" int hostsim_fft2(int N, int RA, int tcw, int sign, double *data /* [N][tcw] complex, in place */) {
  cd *d = (cd *)data;
  if (N == 250 && RA == 10) run_fft2<250, 10, 25>(tcw, sign, d);
  else if (N == 400 && RA == 20) run_fft2<400, 20, 20>(tcw, sign, d);
  else if (N == 256 && RA == 16) run_fft2<256, 16, 16>(tcw, sign, d);
  else if (N == 100 && RA == 10) run_fft2<100, 10, 10>(tcw, sign, d);
  else if (N == 500 && RA == 20) run_fft2<500, 20, 25>(tcw, sign, d);
  else if (N == 625 && RA == 25) run_fft2<625, 25, 25>(tcw, sign, d);
  else if (N == 128 && RA == 8) run_fft2<128, 8, 16>(tcw, sign, d);
  else return -1;
  return 0;
}


// planned two-pass engine (run-time radix pair, TQ_RADIX_SWITCH; generic = direct-sum pass B)
struct PostCollectRt {
  cd *out;
  int tcw, ra, c, sidx;
  TQ_HD void begin(int s, int cc) { sidx = s; c = cc; }
  template <int T> TQ_HD void store(cd v) { out[(size_t)(sidx + ra * T) * tcw + c] = v; }
  TQ_HD void store_rt(int t, cd v) { out[(size_t)(sidx + ra * t) * tcw + c] = v; }
};
template <int SGN> static int run_fft2_rt(int RA, int RB, int generic, int tcw, cd *data) {
  const int N = RA * RB;
  std::vector<cd> tw1, tw((size_t)N), gtw;
  tq_fft_twiddles_host(N, tw1);
  tq_fft_twiddles_host(RB, gtw);
  for (int q = 0; q < RB; ++q)
    for (int r = 0; r < RA; ++r) tw[(size_t)q * RA + r] = SGN > 0 ? tw1[(q * r) % N] : cconj(tw1[(q * r) % N]);
  const int pitch = tcw + 3;
  std::vector<cd> s((size_t)N * pitch);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < tcw; ++c) s[(size_t)n * pitch + c] = data[(size_t)n * tcw + c];
  PreIdentity pre;
  PostCollectRt post{data, tcw, RA, 0, 0};
  const unsigned magic = fastdiv_magic((unsigned)tcw);
  if (!tq_radix_in_registers(RA, true)) return -1;
  for (int item = 0; item < RB * tcw + 5; ++item) {
#define CALL_A(R) fft2_item_a<R, SGN, false>(s.data(), pitch, tcw, RB, tw.data(), item, pre, magic);
    TQ_RADIX_SWITCH_X(RA, CALL_A)
#undef CALL_A
  }
  for (int item = 0; item < RA * tcw * (generic ? fft2_generic_groups(RB) : 1) + 5; ++item) {
    if (generic) {
      fft2_item_b_generic<SGN>(s.data(), pitch, tcw, RA, RB, gtw.data(), item, post, magic);
    } else {
      if (!tq_radix_in_registers(RB, true)) return -1;
#define CALL_B(R) fft2_item_b<R, SGN>(s.data(), pitch, tcw, RA, item, post, magic);
      TQ_RADIX_SWITCH_X(RB, CALL_B)
#undef CALL_B
    }
  }
  return 0;
}
extern "C" int hostsim_fft2_rt(int RA, int RB, int generic, int tcw, int sign, double *data /* [RA*RB][tcw] complex, in place */) {
  return sign > 0 ? run_fft2_rt<1>(RA, RB, generic, tcw, (cd *)data) : run_fft2_rt<-1>(RA, RB, generic, tcw, (cd *)data);
}

// three-pass small-radix tile FFT (tools/tq_fft3.cuh: an experiment that lost against the two-pass engine, kept with its test)
#include "../../tools/tq_fft3.cuh"
template <int STEP> struct PostCollect3 {
  cd *out;
  int tcw, c, klow;
  TQ_HD void begin(int k, int cc) { klow = k; c = cc; }
  template <int T> TQ_HD void store(cd v) { out[(size_t)(klow + STEP * T) * tcw + c] = v; }
};
template <int N, int R1, int R2, int R3> static void run_fft3(int tcw, int sign, cd *data) {
  constexpr int M = R2 * R3;
  std::vector<cd> wN, wM, tw1((size_t)N), tw2((size_t)M);
  tq_fft_twiddles_host(N, wN);
  tq_fft_twiddles_host(M, wM);
  for (int m = 0; m < M; ++m)
    for (int r = 0; r < R1; ++r) tw1[(size_t)m * R1 + r] = sign > 0 ? wN[(m * r) % N] : cconj(wN[(m * r) % N]);
  for (int q = 0; q < R3; ++q)
    for (int r = 0; r < R2; ++r) tw2[(size_t)q * R2 + r] = sign > 0 ? wM[(q * r) % M] : cconj(wM[(q * r) % M]);
  const int pitch = tcw + 1;
  std::vector<cd> s((size_t)N * pitch);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < tcw; ++c) s[(size_t)n * pitch + c] = data[(size_t)n * tcw + c];
  const int nthreads = 41;
  PreIdentity pre;
  PostCollect3<R1 * R2> post{data, tcw, 0, 0};
  for (int tid = 0; tid < nthreads; ++tid) {
    if (sign > 0) fft3_pass1<N, R1, R2, R3, 1>(s.data(), pitch, tcw, tw1.data(), tid, nthreads, pre);
    else fft3_pass1<N, R1, R2, R3, -1>(s.data(), pitch, tcw, tw1.data(), tid, nthreads, pre);
  }
  for (int tid = 0; tid < nthreads; ++tid) {
    if (sign > 0) fft3_pass2<N, R1, R2, R3, 1>(s.data(), pitch, tcw, tw2.data(), tid, nthreads);
    else fft3_pass2<N, R1, R2, R3, -1>(s.data(), pitch, tcw, tw2.data(), tid, nthreads);
  }
  for (int tid = 0; tid < nthreads; ++tid) {
    if (sign > 0) fft3_pass3<N, R1, R2, R3, 1>(s.data(), pitch, tcw, tid, nthreads, post);
    else fft3_pass3<N, R1, R2, R3, -1>(s.data(), pitch, tcw, tid, nthreads, post);
  }
}
extern "C" int hostsim_fft3(int N, int R1, int R2, int tcw, int sign, double *data /* [N][tcw] complex, in place */) {
  cd *d = (cd *)data;
  if (N == 250 && R1 == 10 && R2 == 5) run_fft3<250, 10, 5, 5>(tcw, sign, d);
  else if (N == 250 && R1 == 5 && R2 == 10) run_fft3<250, 5, 10, 5>(tcw, sign, d);
  else if (N == 400 && R1 == 10 && R2 == 8) run_fft3<400, 10, 8, 5>(tcw, sign, d);
  else if (N == 400 && R1 == 8 && R2 == 10) run_fft3<400, 8, 10, 5>(tcw, sign, d);
  else if (N == 400 && R1 == 5 && R2 == 8) run_fft3<400, 5, 8, 10>(tcw, sign, d);
  else return -1;
  return 0;
}

extern "C" int hostsim_radices(int N, int *radix) {
  FftPlan P;
  if (!tq_fft_factorize(N, P)) return -1;
  for (int i = 0; i < P.npass; ++i) radix[i] = P.radix[i];
  return P.npass;
}

#include "../../triqs_b200/csrc/tq_linalg.cuh"
template <int N> static void inv_all(long n_pts, cd *a) {
  for (long p = 0; p < n_pts; ++p) {
    cd M[N][N];
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) M[i][j] = a[(size_t)p * N * N + i * N + j];
    tq_invert_inplace<N>(M);
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) a[(size_t)p * N * N + i * N + j] = M[i][j];
  }
}
// guarded no-interchange variant: ok[p] tells whether the natural pivots were accepted; rejected matrices are left to the caller
template <int N> static void inv_all_nopivot(long n_pts, cd *a, int *ok) {
  for (long p = 0; p < n_pts; ++p) {
    cd M[N][N];
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) M[i][j] = a[(size_t)p * N * N + i * N + j];
    ok[p] = tq_invert_inplace_nopivot<N>(M) ? 1 : 0;
    if (ok[p])
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) a[(size_t)p * N * N + i * N + j] = M[i][j];
  }
}
extern "C" int hostsim_invert_nopivot(int n, long n_pts, double *data, int *ok) {
  cd *a = (cd *)data;
  switch (n) {
    case 2: inv_all_nopivot<2>(n_pts, a, ok); break;
    case 3: inv_all_nopivot<3>(n_pts, a, ok); break;
    case 5: inv_all_nopivot<5>(n_pts, a, ok); break;
    case 8: inv_all_nopivot<8>(n_pts, a, ok); break;
    default: return -1;
  }
  return 0;
}
extern "C" int hostsim_invert(int n, long n_pts, double *data) {
  cd *a = (cd *)data;
  switch (n) {
    case 1: inv_all<1>(n_pts, a); break;
    case 2: inv_all<2>(n_pts, a); break;
    case 3: inv_all<3>(n_pts, a); break;
    case 4: inv_all<4>(n_pts, a); break;
    case 5: inv_all<5>(n_pts, a); break;
    case 6: inv_all<6>(n_pts, a); break;
    case 7: inv_all<7>(n_pts, a); break;
    case 8: inv_all<8>(n_pts, a); break;
    default: return -1;
  }
  return 0;
}
