// In-shared-memory mixed-radix FFT engine (complex<double>).
//
// A CTA owns a tile s[n][c] (n in [0,N) transform axis, c in [0,TC) independent columns) in shared
// memory and runs an IN-PLACE decimation-in-frequency transform over n for all columns:
// pass p works on sub-blocks of length M_p (M_0 = N, M_{p+1} = M_p / R_p) and, for every q in
// [0, M_p/R_p), replaces the R_p elements {base + q + r*M_p/R_p} by their R_p-point DFT times the
// twiddles w_{M_p}^{q s}.  No butterfly reads what another one writes inside a pass, so one
// __syncthreads per pass is enough and no ping-pong buffer is needed (that is what lets a
// 10^4-point complex<double> column = 160 KB live in one SM's shared memory).
// The result is left in mixed-radix digit-reversed order: X[k] sits at position fft_pos(plan, k);
// the fused epilogues (frequency-window gather, tau post-twist, lattice store) apply that
// permutation while they stream the result to HBM, so it costs no extra pass.
//
// Everything below is __host__ __device__ so the index math and butterflies are unit-tested on the
// CPU build box (tests/hostsim), while the product only ever calls them from kernels.
#pragma once
#include "tq_common.cuh"

#define TQ_MAX_PASSES 16
#define TQ_MAX_GENERIC_RADIX 128

struct FftPlan {
  int N;
  int npass;
  int radix[TQ_MAX_PASSES];
  int M[TQ_MAX_PASSES]; // sub-block length entering pass p
  const cd *tw;         // tw[j] = exp(+2 pi i j / N), j in [0, N)
  const int *pos;       // pos[k] = fft_pos(plan, k): where output k sits after the last pass
};

// position of output k after all passes: k = s_1 + R_1 s_2 + R_1 R_2 s_3 + ...,  pos = sum s_i * (M_i / R_i)
TQ_HD int fft_pos(const FftPlan &P, int k) {
  int pos = 0;
#pragma unroll 1
  for (int p = 0; p < P.npass; ++p) {
    int R = P.radix[p];
    int s = k % R;
    k /= R;
    pos += s * (P.M[p] / R);
  }
  return pos;
}

// optional skew of the row index (one pad row every 32) -- breaks the power-of-two/large strides of the
// digit-reversed gather when TC == 1
// (skewmask = ~0 enables it, 0 disables it; a run-time value keeps the number of kernel instantiations down)
TQ_HD int fft_row(int n, int skewmask) { return n + ((n >> 5) & skewmask); }
TQ_HD int fft_rows_alloc(int N, int skewmask) { return skewmask ? N + (N >> 5) + 1 : N; }

// ---------------------------------------------------------------------------------------------
// butterflies: y_s = sum_r a_r exp(SGN 2 pi i r s / R), in registers
// ---------------------------------------------------------------------------------------------
template <int SGN> TQ_HD cd mul_i_sgn(cd a) { return SGN > 0 ? cmuli(a) : cmulmi(a); }

template <int SGN> TQ_HD void dft2(cd &a0, cd &a1) {
  cd t = a0;
  a0 = cadd(t, a1);
  a1 = csub(t, a1);
}

template <int SGN> TQ_HD void dft3(cd &a0, cd &a1, cd &a2) {
  const double S = 0.86602540378443864676; // sin(2 pi / 3)
  cd t = cadd(a1, a2);
  cd u = caxpy(-0.5, t, a0);
  cd d = mul_i_sgn<SGN>(cscale(S, csub(a1, a2)));
  a0 = cadd(a0, t);
  a1 = cadd(u, d);
  a2 = csub(u, d);
}

template <int SGN> TQ_HD void dft4(cd &a0, cd &a1, cd &a2, cd &a3) {
  cd t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_i_sgn<SGN>(csub(a1, a3));
  a0 = cadd(t0, t2);
  a1 = cadd(t1, t3);
  a2 = csub(t0, t2);
  a3 = csub(t1, t3);
}

template <int SGN> TQ_HD void dft5(cd &a0, cd &a1, cd &a2, cd &a3, cd &a4) {
  const double C1 = 0.30901699437494742410;  // cos(2 pi/5)
  const double C2 = -0.80901699437494742410; // cos(4 pi/5)
  const double S1 = 0.95105651629515357212;  // sin(2 pi/5)
  const double S2 = 0.58778525229247312917;  // sin(4 pi/5)
  cd t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
  cd b1 = caxpy(C2, t2, caxpy(C1, t1, a0));
  cd b2 = caxpy(C1, t2, caxpy(C2, t1, a0));
  cd d1 = mul_i_sgn<SGN>(caxpy(S2, t4, cscale(S1, t3)));
  cd d2 = mul_i_sgn<SGN>(caxpy(-S1, t4, cscale(S2, t3)));
  a0 = cadd(a0, cadd(t1, t2));
  a1 = cadd(b1, d1);
  a4 = csub(b1, d1);
  a2 = cadd(b2, d2);
  a3 = csub(b2, d2);
}

template <int R, int SGN> struct Dft;
template <int SGN> struct Dft<2, SGN> {
  static TQ_HD void run(cd *a) { dft2<SGN>(a[0], a[1]); }
};
template <int SGN> struct Dft<3, SGN> {
  static TQ_HD void run(cd *a) { dft3<SGN>(a[0], a[1], a[2]); }
};
template <int SGN> struct Dft<4, SGN> {
  static TQ_HD void run(cd *a) { dft4<SGN>(a[0], a[1], a[2], a[3]); }
};
template <int SGN> struct Dft<5, SGN> {
  static TQ_HD void run(cd *a) { dft5<SGN>(a[0], a[1], a[2], a[3], a[4]); }
};
// 8 = 2 x 4, decimation in frequency inside the registers
template <int SGN> struct Dft<8, SGN> {
  static TQ_HD void run(cd *a) {
    const double H = 0.70710678118654752440;
    cd u0 = cadd(a[0], a[4]), u1 = cadd(a[1], a[5]), u2 = cadd(a[2], a[6]), u3 = cadd(a[3], a[7]);
    cd v0 = csub(a[0], a[4]), v1 = csub(a[1], a[5]), v2 = csub(a[2], a[6]), v3 = csub(a[3], a[7]);
    // v_n *= w8^n, w8 = (1 + SGN i)/sqrt2
    cd r1 = mul_i_sgn<SGN>(v1);
    v1 = cscale(H, cadd(v1, r1));
    v2 = mul_i_sgn<SGN>(v2);
    cd r3 = mul_i_sgn<SGN>(v3);
    v3 = cscale(H, csub(r3, v3));
    dft4<SGN>(u0, u1, u2, u3);
    dft4<SGN>(v0, v1, v2, v3);
    a[0] = u0; a[2] = u1; a[4] = u2; a[6] = u3;
    a[1] = v0; a[3] = v1; a[5] = v2; a[7] = v3;
  }
};
// 10 = 2 x 5 by the prime-factor (Good-Thomas) map: n = (5 n1 + 2 n2) mod 10, k = (5 k1 + 6 k2) mod 10, no twiddles
template <int SGN> struct Dft<10, SGN> {
  static TQ_HD void run(cd *a) {
    cd e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], e4 = a[8]; // n1 = 0
    cd o0 = a[5], o1 = a[7], o2 = a[9], o3 = a[1], o4 = a[3]; // n1 = 1
    dft5<SGN>(e0, e1, e2, e3, e4);
    dft5<SGN>(o0, o1, o2, o3, o4);
    a[0] = cadd(e0, o0); a[5] = csub(e0, o0);
    a[6] = cadd(e1, o1); a[1] = csub(e1, o1);
    a[2] = cadd(e2, o2); a[7] = csub(e2, o2);
    a[8] = cadd(e3, o3); a[3] = csub(e3, o3);
    a[4] = cadd(e4, o4); a[9] = csub(e4, o4);
  }
};
// 16 = 4 x 4, decimation in frequency inside the registers
template <int SGN> struct Dft<16, SGN> {
  static TQ_HD void run(cd *a) {
    const double H = 0.70710678118654752440;
    const double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173; // cos, sin(pi/8)
    cd b[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      cd x0 = a[n], x1 = a[n + 4], x2 = a[n + 8], x3 = a[n + 12];
      dft4<SGN>(x0, x1, x2, x3);
      b[n][0] = x0; b[n][1] = x1; b[n][2] = x2; b[n][3] = x3;
    }
    // b[n][s] *= w16^{n s}
    const cd w1 = cmk(C1, SGN * S1), w2 = cmk(H, SGN * H), w3 = cmk(S1, SGN * C1);
    b[1][1] = cmul(b[1][1], w1);
    b[1][2] = cmul(b[1][2], w2);
    b[1][3] = cmul(b[1][3], w3);
    b[2][1] = cmul(b[2][1], w2);
    b[2][2] = mul_i_sgn<SGN>(b[2][2]);                 // w16^4
    b[2][3] = cmul(b[2][3], cmk(-H, SGN * H));         // w16^6
    b[3][1] = cmul(b[3][1], w3);
    b[3][2] = cmul(b[3][2], cmk(-H, SGN * H));         // w16^6
    b[3][3] = cmul(b[3][3], cmk(-C1, -SGN * S1));      // w16^9
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      cd x0 = b[0][s], x1 = b[1][s], x2 = b[2][s], x3 = b[3][s];
      dft4<SGN>(x0, x1, x2, x3);
      a[s] = x0; a[s + 4] = x1; a[s + 8] = x2; a[s + 12] = x3;
    }
  }
};

// 20 = 4 x 5 by the prime-factor map: n = (5 n1 + 4 n2) mod 20, k = (5 k1 + 16 k2) mod 20, no twiddles
template <int SGN> struct Dft<20, SGN> {
  static TQ_HD void run(cd *a) {
    cd e[4][5];
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) {
      cd x0 = a[(5 * n1) % 20], x1 = a[(5 * n1 + 4) % 20], x2 = a[(5 * n1 + 8) % 20], x3 = a[(5 * n1 + 12) % 20], x4 = a[(5 * n1 + 16) % 20];
      dft5<SGN>(x0, x1, x2, x3, x4);
      e[n1][0] = x0; e[n1][1] = x1; e[n1][2] = x2; e[n1][3] = x3; e[n1][4] = x4;
    }
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) {
      cd x0 = e[0][k2], x1 = e[1][k2], x2 = e[2][k2], x3 = e[3][k2];
      dft4<SGN>(x0, x1, x2, x3);
      a[(16 * k2) % 20] = x0; a[(5 + 16 * k2) % 20] = x1; a[(10 + 16 * k2) % 20] = x2; a[(15 + 16 * k2) % 20] = x3;
    }
  }
};
// 25 = 5 x 5: n = n1 + 5 n2, k = k2 + 5 k1;  u[n1][k2] = dft5_{n2}, times w25^{n1 k2}, then dft5_{n1}
template <int SGN> struct Dft<25, SGN> {
  static TQ_HD cd w25(int m) {
    // exp(2 pi i m / 25), m = n1 * k2 in {1,2,3,4,6,8,9,12,16}
    switch (m) {
      case 1: return cmk(0.9685831611286311194901684, SGN * 0.2486898871648547882422837);
      case 2: return cmk(0.8763066800438635873081159, SGN * 0.4817536741017152749871915);
      case 3: return cmk(0.7289686274214115231467303, SGN * 0.6845471059286886737322834);
      case 4: return cmk(0.5358267949789966182713088, SGN * 0.8443279255020150785485581);
      case 6: return cmk(0.06279051952931337607617822, SGN * 0.9980267284282715619523368);
      case 8: return cmk(-0.4257792915650726488625024, SGN * 0.9048270524660195277136686);
      case 9: return cmk(-0.6374239897486897101767128, SGN * 0.7705132427757892308030096);
      case 12: return cmk(-0.992114701314477831049793, SGN * 0.1253332335643042453731188);
      default: return cmk(-0.6374239897486897101767128, SGN * -0.7705132427757892308030096); // 16
    }
  }
  static TQ_HD void run(cd *a) {
    cd u[5][5];
#pragma unroll
    for (int n1 = 0; n1 < 5; ++n1) {
      cd x0 = a[n1], x1 = a[n1 + 5], x2 = a[n1 + 10], x3 = a[n1 + 15], x4 = a[n1 + 20];
      dft5<SGN>(x0, x1, x2, x3, x4);
      u[n1][0] = x0; u[n1][1] = x1; u[n1][2] = x2; u[n1][3] = x3; u[n1][4] = x4;
    }
#pragma unroll
    for (int n1 = 1; n1 < 5; ++n1)
#pragma unroll
      for (int k2 = 1; k2 < 5; ++k2) u[n1][k2] = cmul(u[n1][k2], w25(n1 * k2));
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) {
      cd x0 = u[0][k2], x1 = u[1][k2], x2 = u[2][k2], x3 = u[3][k2], x4 = u[4][k2];
      dft5<SGN>(x0, x1, x2, x3, x4);
      a[k2] = x0; a[k2 + 5] = x1; a[k2 + 10] = x2; a[k2 + 15] = x3; a[k2 + 20] = x4;
    }
  }
};

// ---- radices of the planned two-pass engine (tq_fft2.cuh): 1, 6, 7, 9, 12, 15 ------------------
template <int SGN> struct Dft<1, SGN> {
  static TQ_HD void run(cd *) {}
};
// 7: symmetric form  y_k, y_{7-k} = (x0 + sum_j c_{jk} t_j) +- i SGN (sum_j s_{jk} d_j),  t_j = x_j + x_{7-j},  d_j = x_j - x_{7-j}
template <int SGN> struct Dft<7, SGN> {
  static TQ_HD void run(cd *a) {
    const double C1 = 0.62348980185873353053, C2 = -0.22252093395631440429, C3 = -0.90096886790241912624; // cos(2 pi j/7)
    const double S1 = 0.78183148246802980871, S2 = 0.97492791218182360702, S3 = 0.43388373911755812048;  // sin(2 pi j/7)
    const cd t1 = cadd(a[1], a[6]), t2 = cadd(a[2], a[5]), t3 = cadd(a[3], a[4]);
    const cd d1 = csub(a[1], a[6]), d2 = csub(a[2], a[5]), d3 = csub(a[3], a[4]);
    const cd x0 = a[0];
    a[0] = cadd(cadd(x0, t1), cadd(t2, t3));
    // k = 1: cos (C1, C2, C3), sin (S1, S2, S3);  k = 2: cos (C2, C3, C1), sin (S2, -S3, -S1);  k = 3: cos (C3, C1, C2), sin (S3, -S1, S2)
    const cd A1 = caxpy(C3, t3, caxpy(C2, t2, caxpy(C1, t1, x0)));
    const cd A2 = caxpy(C1, t3, caxpy(C3, t2, caxpy(C2, t1, x0)));
    const cd A3 = caxpy(C2, t3, caxpy(C1, t2, caxpy(C3, t1, x0)));
    const cd B1 = mul_i_sgn<SGN>(caxpy(S3, d3, caxpy(S2, d2, cscale(S1, d1))));
    const cd B2 = mul_i_sgn<SGN>(caxpy(-S1, d3, caxpy(-S3, d2, cscale(S2, d1))));
    const cd B3 = mul_i_sgn<SGN>(caxpy(S2, d3, caxpy(-S1, d2, cscale(S3, d1))));
    a[1] = cadd(A1, B1); a[6] = csub(A1, B1);
    a[2] = cadd(A2, B2); a[5] = csub(A2, B2);
    a[3] = cadd(A3, B3); a[4] = csub(A3, B3);
  }
};
// 9 = 3 x 3: n = n1 + 3 n2, k = k2 + 3 k1;  u[n1][k2] = dft3_{n2}, times w9^{n1 k2}, then dft3_{n1}
template <int SGN> struct Dft<9, SGN> {
  static TQ_HD void run(cd *a) {
    const cd w1 = cmk(0.76604444311897803520, SGN * 0.64278760968653932632);  // w9^1
    const cd w2 = cmk(0.17364817766693034885, SGN * 0.98480775301220805937);  // w9^2
    const cd w4 = cmk(-0.93969262078590838405, SGN * 0.34202014332566873304); // w9^4
    cd u[3][3];
#pragma unroll
    for (int n1 = 0; n1 < 3; ++n1) {
      cd x0 = a[n1], x1 = a[n1 + 3], x2 = a[n1 + 6];
      dft3<SGN>(x0, x1, x2);
      u[n1][0] = x0; u[n1][1] = x1; u[n1][2] = x2;
    }
    u[1][1] = cmul(u[1][1], w1);
    u[1][2] = cmul(u[1][2], w2);
    u[2][1] = cmul(u[2][1], w2);
    u[2][2] = cmul(u[2][2], w4);
#pragma unroll
    for (int k2 = 0; k2 < 3; ++k2) {
      cd x0 = u[0][k2], x1 = u[1][k2], x2 = u[2][k2];
      dft3<SGN>(x0, x1, x2);
      a[k2] = x0; a[k2 + 3] = x1; a[k2 + 6] = x2;
    }
  }
};
// 32 = 4 x 8 with twiddles: n = n1 + 4 n2, k = k2 + 8 k1;  u[n1][k2] = dft8_{n2}, times w32^{n1 k2}, then dft4_{n1}
template <int SGN> struct Dft<32, SGN> {
  static TQ_HD cd w32(int m) {
    constexpr double W[22][2] = {{1.00000000000000000000, 0.00000000000000000000}, {0.98078528040323043058, 0.19509032201612824808}, {0.92387953251128673848, 0.38268343236508978178}, {0.83146961230254523567, 0.55557023301960217765}, {0.70710678118654757274, 0.70710678118654746172}, {0.55557023301960228867, 0.83146961230254523567}, {0.38268343236508983729, 0.92387953251128673848}, {0.19509032201612833135, 0.98078528040323043058}, {0.00000000000000006123, 1.00000000000000000000}, {-0.19509032201612819257, 0.98078528040323043058}, {-0.38268343236508972627, 0.92387953251128673848}, {-0.55557023301960195560, 0.83146961230254545772}, {-0.70710678118654746172, 0.70710678118654757274}, {-0.83146961230254534669, 0.55557023301960217765}, {-0.92387953251128673848, 0.38268343236508989280}, {-0.98078528040323043058, 0.19509032201612860891}, {-1.00000000000000000000, 0.00000000000000012246}, {-0.98078528040323043058, -0.19509032201612835911}, {-0.92387953251128684951, -0.38268343236508967076}, {-0.83146961230254545772, -0.55557023301960195560}, {-0.70710678118654768376, -0.70710678118654746172}, {-0.55557023301960217765, -0.83146961230254523567}};
    return cmk(W[m][0], SGN * W[m][1]);
  }
  static TQ_HD void run(cd *a) {
    cd u[4][8];
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) {
      cd x[8];
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) x[n2] = a[n1 + 4 * n2];
      Dft<8, SGN>::run(x);
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) u[n1][k2] = (n1 == 0 || k2 == 0) ? x[k2] : cmul(x[k2], w32(n1 * k2));
    }
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      cd x0 = u[0][k2], x1 = u[1][k2], x2 = u[2][k2], x3 = u[3][k2];
      dft4<SGN>(x0, x1, x2, x3);
      a[k2] = x0; a[k2 + 8] = x1; a[k2 + 16] = x2; a[k2 + 24] = x3;
    }
  }
};
// Good-Thomas prime-factor map for coprime R1, R2 (no twiddles):
//   n = (R2 n1 + R1 n2) mod N,   k = (E1 k1 + E2 k2) mod N  with  E1 = 1 mod R1, 0 mod R2  and  E2 = 0 mod R1, 1 mod R2
//   => n k / N = n1 k1 / R1 + n2 k2 / R2 (mod 1):  plain DFT_R1 over n1 for every n2, then plain DFT_R2 over n2 for every k1.
constexpr int tq_modinv(int a, int m) {
  a %= m;
  for (int x = 1; x < m; ++x)
    if ((a * x) % m == 1) return x;
  return 1;
}
template <int R1, int R2, int SGN> struct DftPfa {
  static constexpr int N = R1 * R2;
  static constexpr int E1 = R2 * tq_modinv(R2, R1); // = 1 mod R1, 0 mod R2
  static constexpr int E2 = R1 * tq_modinv(R1, R2); // = 0 mod R1, 1 mod R2
  static TQ_HD void run(cd *a) {
    cd u[R2][R1];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
      cd x[R1];
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) x[n1] = a[(R2 * n1 + R1 * n2) % N];
      Dft<R1, SGN>::run(x);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) u[n2][k1] = x[k1];
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
      cd x[R2];
#pragma unroll
      for (int n2 = 0; n2 < R2; ++n2) x[n2] = u[n2][k1];
      Dft<R2, SGN>::run(x);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) a[(E1 * k1 + E2 * k2) % N] = x[k2];
    }
  }
};
template <int SGN> struct Dft<6, SGN> {
  static TQ_HD void run(cd *a) { DftPfa<3, 2, SGN>::run(a); }
};
template <int SGN> struct Dft<12, SGN> {
  static TQ_HD void run(cd *a) { DftPfa<3, 4, SGN>::run(a); }
};
template <int SGN> struct Dft<15, SGN> {
  static TQ_HD void run(cd *a) { DftPfa<3, 5, SGN>::run(a); }
};

// ---------------------------------------------------------------------------------------------
// Thread layout inside a CTA: a thread owns a fixed column lane `ct` (of `nct` column lanes) and a
// butterfly/row lane `bt` (of `nbt`): no integer division in any inner loop, adjacent threads touch
// adjacent columns (coalesced global rows, conflict-free shared rows).
// ---------------------------------------------------------------------------------------------
struct TileThread {
  int ct, nct, bt, nbt;
};
TQ_HD TileThread tile_thread(int tid, int nthreads, int TC) {
  TileThread t;
  t.nct = TC < nthreads ? TC : nthreads;
  t.nbt = nthreads / t.nct;
  t.bt = tid / t.nct;
  t.ct = tid - t.bt * t.nct;
  if (t.bt >= t.nbt) { // left-over threads (nthreads not a multiple of TC) idle
    t.bt = 0;
    t.nbt = 0;
  }
  return t;
}

// n / d for n * d < 2^32 with m = ceil(2^32 / d) (m = 0 encodes d = 1)
TQ_HD unsigned fastdiv_magic(unsigned d) { return d <= 1 ? 0u : (unsigned)((0x100000000ull + d - 1) / d); }
TQ_HD int fastdiv(int n, unsigned m) {
#ifdef __CUDA_ARCH__
  return m ? (int)__umulhi((unsigned)n, m) : n;
#else
  return m ? (int)(((unsigned long long)(unsigned)n * m) >> 32) : n;
#endif
}

// one pass over the tile.
//   direct_tw = true : every twiddle w_M^{q s} is read from the table `tw` (a shared-memory copy when a
//                      warp shares a butterfly, i.e. many columns: the loads are broadcasts)
//   direct_tw = false: only w_M^{q} is read, the powers are built by a log-depth product chain (TC == 1)
template <int R, int SGN, bool SKEW>
TQ_HD void fft_pass_fixed(cd *s, int pitch, int TC, int N, int M, const cd *__restrict__ tw, TileThread T, bool direct_tw) {
  const int Msub = M / R;
  const int nb = N / R;
  const int twstride = N / M;
  const unsigned magic = fastdiv_magic((unsigned)Msub);
  const int st = Msub * pitch; // distance between the legs of a butterfly (elements), un-skewed tiles
  if (T.nbt == 0) return;
  for (int c = T.ct; c < TC; c += T.nct) {
    for (int b = T.bt; b < nb; b += T.nbt) {
      const int blk = fastdiv(b, magic);
      const int q = b - blk * Msub;
      const int base = blk * M + q;
      cd *p = s + (base * pitch + c);
      cd a[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = SKEW ? s[fft_row(base + r * Msub, ~0) * pitch + c] : p[r * st];
      Dft<R, SGN>::run(a);
      if (Msub > 1 && q != 0) {
        const int qt = q * twstride;
        if (direct_tw) {
#pragma unroll
          for (int r = 1; r < R; ++r) {
            cd t = tw[qt * r];
            if (SGN < 0) t = cconj(t);
            a[r] = cmul(a[r], t);
          }
        } else {
          cd pw[R];
          pw[1] = tw[qt];
          if (SGN < 0) pw[1] = cconj(pw[1]);
#pragma unroll
          for (int r = 2; r < R; ++r) pw[r] = (r & 1) ? cmul(pw[r - 1], pw[1]) : cmul(pw[r >> 1], pw[r >> 1]);
#pragma unroll
          for (int r = 1; r < R; ++r) a[r] = cmul(a[r], pw[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (SKEW)
          s[fft_row(base + r * Msub, ~0) * pitch + c] = a[r];
        else
          p[r * st] = a[r];
      }
    }
  }
}

// any radix (odd primes 7, 11, 13, 41, 101, ...): O(R^2) DFT per butterfly out of a thread-local array.
// The R positions belong to this butterfly only and a[] is a private copy, so in-place is safe.
template <int SGN>
TQ_HD void fft_pass_generic(cd *s, int pitch, int TC, int N, int M, int R, const cd *__restrict__ tw, TileThread T, int skewmask) {
  const int Msub = M / R;
  const int nb = N / R;
  const int twstride = N / M;
  const int rstride = N / R; // w_R^j = tw[j * N / R]
  const unsigned magic = fastdiv_magic((unsigned)Msub);
  if (T.nbt == 0) return;
  for (int c = T.ct; c < TC; c += T.nct) {
    for (int b = T.bt; b < nb; b += T.nbt) {
      int blk = fastdiv(b, magic);
      int q = b - blk * Msub;
      int base = blk * M + q;
      cd a[TQ_MAX_GENERIC_RADIX];
      for (int r = 0; r < R; ++r) a[r] = s[fft_row(base + r * Msub, skewmask) * pitch + c];
      for (int k = 0; k < R; ++k) {
        cd acc = a[0];
        int j = 0;
        for (int r = 1; r < R; ++r) {
          j += k;
          if (j >= R) j -= R;
          cd t = tw[j * rstride];
          if (SGN < 0) t = cconj(t);
          acc = cfma(a[r], t, acc);
        }
        if (Msub > 1 && q != 0 && k != 0) {
          cd t = tw[(q * k) * twstride];
          if (SGN < 0) t = cconj(t);
          acc = cmul(acc, t);
        }
        s[fft_row(base + k * Msub, skewmask) * pitch + c] = acc;
      }
    }
  }
}

template <int SGN, bool SKEW>
TQ_HD void fft_pass(const FftPlan &P, int p, cd *s, int pitch, int TC, const cd *__restrict__ tw, TileThread T, bool direct_tw) {
  const int skewmask = SKEW ? ~0 : 0;
  const int R = P.radix[p], M = P.M[p], N = P.N;
  switch (R) {
    case 2: fft_pass_fixed<2, SGN, SKEW>(s, pitch, TC, N, M, tw, T, direct_tw); break;
    case 3: fft_pass_fixed<3, SGN, SKEW>(s, pitch, TC, N, M, tw, T, direct_tw); break;
    case 4: fft_pass_fixed<4, SGN, SKEW>(s, pitch, TC, N, M, tw, T, direct_tw); break;
    case 5: fft_pass_fixed<5, SGN, SKEW>(s, pitch, TC, N, M, tw, T, direct_tw); break;
    case 8: fft_pass_fixed<8, SGN, SKEW>(s, pitch, TC, N, M, tw, T, direct_tw); break;
    case 10: fft_pass_fixed<10, SGN, SKEW>(s, pitch, TC, N, M, tw, T, direct_tw); break;
    case 16: fft_pass_fixed<16, SGN, SKEW>(s, pitch, TC, N, M, tw, T, direct_tw); break;
    default: fft_pass_generic<SGN>(s, pitch, TC, N, M, R, tw, T, skewmask); break;
  }
}

#ifdef __CUDACC__
// all passes, CTA-wide (every thread of the block must call it; tile must be complete and synced before).
// `tw` is the twiddle table the passes read: the plan's global table or a shared-memory copy of it.
template <int SGN, bool SKEW> __device__ void fft_tile(const FftPlan &P, cd *s, int pitch, int TC, const cd *tw, bool direct_tw) {
  const TileThread T = tile_thread(threadIdx.x, blockDim.x, TC);
  for (int p = 0; p < P.npass; ++p) {
    fft_pass<SGN, SKEW>(P, p, s, pitch, TC, tw, T, direct_tw);
    __syncthreads();
  }
}
#endif

// ---------------------------------------------------------------------------------------------
// host side: factorisation + master twiddle table
// ---------------------------------------------------------------------------------------------
struct FftPlanHost {
  FftPlan plan{};          // plan.tw / plan.pos are device pointers
  std::vector<cd> tw_host; // host copy (tests / hostsim)
  std::vector<int> pos_host;
  bool supported = true;
  ~FftPlanHost();
};

// fills radix[] / M[]; returns false if a prime factor exceeds TQ_MAX_GENERIC_RADIX
bool tq_fft_factorize(int N, FftPlan &P);
// exp(+2 pi i j / N) with long-double evaluation and octant symmetry
void tq_fft_twiddles_host(int N, std::vector<cd> &tw);
// cached plan with the table uploaded to the device
tq_status tq_get_fft_plan(tq_ctx *ctx, long N, std::shared_ptr<FftPlanHost> *out);
