// Three-pass in-place tile FFT with small compile-time radices (N = R1 * R2 * R3), the round-2 candidate for the pipelined
// Matsubara kernels: radices <= 10 need ~45 registers of data instead of the ~100 of a radix-20 / radix-25 butterfly, so three
// times as many warps fit the register file, at the price of one more shared-memory sweep.  Same conventions as tq_fft2.cuh
// (tile s[N][pitch] complex, rows = transform index, `tcw` active columns; work item i = butterfly x column, one per thread
// and loop step; functors for the first-pass loads and the last-pass stores).  An EXPERIMENT, not part of the product:
// validated on the host (tests/hostsim) and timed against the two-pass engine in tools/ubench5.cu -- where it LOSES:
// 2.15 cycles per element at 512 threads (worse with more threads, which spill at 64-80 registers) against 1.53-1.67 for the
// two-pass radix-10x25 / 20x20 engine at 11 warps (one B200 SM, tile in shared memory, lock-step passes).
//
//   M = R2 R3.  n = m + M r1:          pass 1  butterflies (m, c), legs r1 at stride M rows, output s1 in place,
//                                              times tw1[m R1 + s1] = w_N^{SGN m s1}
//   inside block s1:  m = q + R3 r2:   pass 2  butterflies (s1, q, c), legs r2 at stride R3 rows, output s2 in place,
//                                              times tw2[q R2 + s2] = w_M^{SGN q s2}
//   inside block (s1, s2):             pass 3  butterflies (s1, s2, c), legs q contiguous rows, output s3:
//                                              k = s1 + R1 (s2 + R2 s3)   ->  post.begin(s1 + R1 s2, c); post.store<S3>(v)
#pragma once
#include "../triqs_b200/csrc/tq_fft2.cuh"

template <int N, int R1, int R2, int R3, int SGN, class Pre>
TQ_HD void fft3_pass1(cd *s, int pitch, int tcw, const cd *__restrict__ tw1, int tid, int nthreads, Pre &pre) {
  static_assert(R1 * R2 * R3 == N, "N = R1 * R2 * R3");
  constexpr int M = R2 * R3;
  const int nitems = M * tcw, stride = M * pitch;
  const unsigned magic = fastdiv_magic((unsigned)tcw);
  for (int i = tid; i < nitems; i += nthreads) {
    const int m = fastdiv(i, magic), c = i - m * tcw;
    cd *p = s + m * pitch + c;
    pre.begin(m, c);
    cd a[R1];
    tq_unroll_load<R1>(a, pre, p, stride);
    Dft<R1, SGN>::run(a);
    const cd *tw = tw1 + m * R1;
    if (m != 0) {
#pragma unroll
      for (int r = 1; r < R1; ++r) a[r] = cmul(a[r], tw[r]);
    }
#pragma unroll
    for (int r = 0; r < R1; ++r) p[r * stride] = a[r];
  }
}

template <int N, int R1, int R2, int R3, int SGN>
TQ_HD void fft3_pass2(cd *s, int pitch, int tcw, const cd *__restrict__ tw2, int tid, int nthreads) {
  constexpr int M = R2 * R3;
  const int nitems = R1 * R3 * tcw, stride = R3 * pitch;
  const unsigned magic = fastdiv_magic((unsigned)tcw), magic3 = fastdiv_magic((unsigned)R3);
  for (int i = tid; i < nitems; i += nthreads) {
    const int j = fastdiv(i, magic), c = i - j * tcw;
    const int s1 = fastdiv(j, magic3), q = j - s1 * R3;
    cd *p = s + (s1 * M + q) * pitch + c;
    cd a[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) a[r] = p[r * stride];
    Dft<R2, SGN>::run(a);
    const cd *tw = tw2 + q * R2;
    if (q != 0) {
#pragma unroll
      for (int r = 1; r < R2; ++r) a[r] = cmul(a[r], tw[r]);
    }
#pragma unroll
    for (int r = 0; r < R2; ++r) p[r * stride] = a[r];
  }
}

template <int N, int R1, int R2, int R3, int SGN, class Post>
TQ_HD void fft3_pass3(const cd *s, int pitch, int tcw, int tid, int nthreads, Post &post) {
  constexpr int M = R2 * R3;
  const int nitems = R1 * R2 * tcw;
  const unsigned magic = fastdiv_magic((unsigned)tcw), magic2 = fastdiv_magic((unsigned)R2);
  for (int i = tid; i < nitems; i += nthreads) {
    const int j = fastdiv(i, magic), c = i - j * tcw;
    const int s1 = fastdiv(j, magic2), s2 = j - s1 * R2;
    const cd *p = s + (s1 * M + R3 * s2) * pitch + c;
    post.begin(s1 + R1 * s2, c);
    cd a[R3];
#pragma unroll
    for (int r = 0; r < R3; ++r) a[r] = p[r * pitch];
    Dft<R3, SGN>::run(a);
    tq_unroll_store<R3>(a, post);
  }
}
