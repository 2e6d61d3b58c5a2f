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
