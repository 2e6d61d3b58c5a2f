// Register-resident small complex linear algebra (host/device so the CPU build box can unit-test it).
#pragma once
#include "tq_common.cuh"

// swap a <-> b when `sw` (compiles to selects; keeps both operands in registers)
TQ_HD void cswap_if(bool sw, cd &a, cd &b) {
  cd ta = a, tb = b;
  a.x = sw ? tb.x : ta.x;
  a.y = sw ? tb.y : ta.y;
  b.x = sw ? ta.x : tb.x;
  b.y = sw ? ta.y : tb.y;
}

// In-place inverse of an N x N complex matrix: Gauss-Jordan with partial (row) pivoting.
// All loops are fully unrolled and the pivot row is moved with predicated swaps, so M stays in registers.
template <int N> TQ_HD void tq_invert_inplace(cd (&M)[N][N]) {
  int piv[N];
#pragma unroll
  for (int p = 0; p < N; ++p) {
    // pivot search in column p, rows p..N-1
    double best = cabs2(M[p][p]);
    int r = p;
#pragma unroll
    for (int i = p + 1; i < N; ++i) {
      double a = cabs2(M[i][p]);
      bool gt = a > best;
      best = gt ? a : best;
      r = gt ? i : r;
    }
    piv[p] = r;
#pragma unroll
    for (int i = p + 1; i < N; ++i) {
      bool sw = (r == i);
#pragma unroll
      for (int j = 0; j < N; ++j) cswap_if(sw, M[p][j], M[i][j]);
    }
    // 1 / pivot
    cd d = M[p][p];
    double inv = 1.0 / cabs2(d);
    cd ip = cmk(d.x * inv, -d.y * inv);
    M[p][p] = cmk(1.0, 0.0);
#pragma unroll
    for (int j = 0; j < N; ++j) M[p][j] = cmul(M[p][j], ip);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i == p) continue;
      cd f = M[i][p];
      M[i][p] = cmk(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < N; ++j) M[i][j] = cfma(cneg(f), M[p][j], M[i][j]);
    }
  }
  // undo the row interchanges as column interchanges, in reverse order
#pragma unroll
  for (int p = N - 1; p >= 0; --p) {
#pragma unroll
    for (int c = p + 1; c < N; ++c) {
      bool sw = (piv[p] == c);
#pragma unroll
      for (int i = 0; i < N; ++i) cswap_if(sw, M[i][p], M[i][c]);
    }
  }
}

// Gauss-Jordan WITHOUT row interchanges, guarded by threshold pivoting: step p keeps the natural pivot only if
// |a_pp|^2 >= TQ_PIVOT_TAU max_i |a_ip|^2 (element growth <= 1/sqrt(TAU) per step, like partial pivoting with a relaxed
// choice).  Returns false -- with M destroyed -- as soon as a pivot fails the test (or is 0 / NaN); the caller then redoes
// the matrix with tq_invert_inplace.  For the Dyson matrices (i w + mu) - eps_k - Sigma(i w) the anti-Hermitian part is
// definite, the natural pivots are essentially always accepted, and the ~800 predicated register moves of the pivoted
// version (more instructions than its floating-point work) disappear from the hot loop.
#define TQ_PIVOT_TAU 0.015625
template <int N> TQ_HD bool tq_invert_inplace_nopivot(cd (&M)[N][N]) {
  bool ok = true;
#pragma unroll
  for (int p = 0; p < N; ++p) {
    const cd d = M[p][p];
    const double d2 = cabs2(d);
    double mx = 0.0;
#pragma unroll
    for (int i = p + 1; i < N; ++i) mx = fmax(mx, cabs2(M[i][p]));
    ok = ok && (d2 >= TQ_PIVOT_TAU * mx) && (d2 > 0.0);
    const double inv = 1.0 / d2;
    const cd ip = cmk(d.x * inv, -d.y * inv);
    M[p][p] = cmk(1.0, 0.0);
#pragma unroll
    for (int j = 0; j < N; ++j) M[p][j] = cmul(M[p][j], ip);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i == p) continue;
      const cd f = M[i][p];
      M[i][p] = cmk(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < N; ++j) M[i][j] = cfma(cneg(f), M[p][j], M[i][j]);
    }
  }
  return ok;
}
