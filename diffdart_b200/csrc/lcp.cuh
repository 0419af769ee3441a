// lcp.cuh — batched boxed-LCP solvers, one world per thread, FP64, sm_100a.
//
//   dantzigSolve()  restates the ODE Dantzig solver the reference vendors and calls through
//                   DantzigBoxedLcpSolver::solve (dart/constraint/DantzigBoxedLcpSolver.cpp:55-77 ->
//                   dart/external/odelcpsolver/lcp.cpp:780-1114 dSolveLCP, struct dLCP :358-772,
//                   swapRowsAndCols/swapProblem :205-300, matrix.cpp:286-470 dLDLTAddTL/dLDLTRemove/
//                   dRemoveRowCol, fastlsolve.cpp / fastltsolve.cpp / fastdot.cpp).
//   pgsSolve()      PgsBoxedLcpSolver::solve (dart/constraint/PgsBoxedLcpSolver.cpp:79-262).
//   lcpValid()      LCPUtils::isLCPSolutionValid (dart/constraint/LCPUtils.cpp:12-81).
//
// The contact active set has to match the reference bit-exactly, and Dantzig's pivot choices hang
// on comparisons between computed doubles, so every floating-point operation below is issued in
// the reference's order: dot products accumulate left to right, the 4-row blocked triangular solves
// keep their association (x1 = (b1 - Z) - l10*x0), and this file is compiled with -fmad=false so
// that no multiply-add is contracted (the reference's release build has no FMA either).
//
// Storage: every world owns one slice of a scratch buffer, interleaved across the 32 lanes of its
// warp (element e of lane l at base[e*32 + l]) so that lanes that walk the algorithm in lock step
// touch one 256-byte segment per access. The pivoting is data dependent, so worlds in a warp do
// diverge; that costs issue slots, never correctness.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ddb {

#define DDB_LCP_DI __device__ __forceinline__

DDB_LCP_DI int dpad(int a) { return (a > 1) ? (((a - 1) | 3) + 1) : a; }  // common.h:141 dPAD

// doubles / ints of scratch per world for problems up to n rows
__host__ __device__ inline size_t lcpScratchDoubles(int n) {
  int ns = (n > 1) ? (((n - 1) | 3) + 1) : n;
  return (size_t)2 * n * ns + 10 * (size_t)n + 2 * (size_t)ns + n;
}
__host__ __device__ inline size_t lcpScratchInts(int n) { return 5 * (size_t)n; }

struct LcpView {
  double* db;  // this world's double scratch, lane-interleaved
  int* ib;     // this world's int scratch, lane-interleaved
  int n, nskip;
  int st;  // element stride: 32 = interleaved across the worlds of a warp (one world per thread), 1 = one world per warp
  int lane, nl;  // one world per warp: this thread's lane and 32, so that loops over independent columns / rows are split
                 // across the warp (each column is still summed by one lane in the scalar order: results do not change)
  // offsets (elements)
  int oA, oL, ox, ob, ow, olo, ohi, od, odw, odx, oDell, oell, otmp;
  int op, oC, ofi, oar, ost;
  DDB_LCP_DI void init(double* d, int* i, int n_, int nmax, int stride = 32) {
    db = d; ib = i; n = n_; nskip = dpad(n_); st = stride;
    lane = 0; nl = 1;
#ifdef __CUDA_ARCH__
    if (stride == 1) { lane = threadIdx.x & 31; nl = 32; }
#endif
    int nsm = dpad(nmax);
    oA = 0; oL = nmax * nsm; ox = 2 * nmax * nsm; ob = ox + nmax; ow = ob + nmax; olo = ow + nmax; ohi = olo + nmax;
    od = ohi + nmax; odw = od + nmax; odx = odw + nmax; oDell = odx + nmax; oell = oDell + nmax; otmp = oell + nmax;
    op = 0; oC = nmax; ofi = 2 * nmax; oar = 3 * nmax; ost = 4 * nmax;
  }
  DDB_LCP_DI double& D(int off, int i) const { return db[(size_t)(off + i) * st]; }
  DDB_LCP_DI int& I(int off, int i) const { return ib[(size_t)(off + i) * st]; }
  DDB_LCP_DI double& A(int i, int j) const { return db[(size_t)(oA + I(oar, i) * nskip + j) * st]; }  // row pointers
  DDB_LCP_DI double& Araw(int r, int j) const { return db[(size_t)(oA + r * nskip + j) * st]; }
  DDB_LCP_DI double& L(int i, int j) const { return db[(size_t)(oL + i * nskip + j) * st]; }
  DDB_LCP_DI double& x(int i) const { return D(ox, i); }
  DDB_LCP_DI double& b(int i) const { return D(ob, i); }
  DDB_LCP_DI double& w(int i) const { return D(ow, i); }
  DDB_LCP_DI double& lo(int i) const { return D(olo, i); }
  DDB_LCP_DI double& hi(int i) const { return D(ohi, i); }
  DDB_LCP_DI double& d(int i) const { return D(od, i); }
  DDB_LCP_DI double& dw(int i) const { return D(odw, i); }
  DDB_LCP_DI double& dx(int i) const { return D(odx, i); }
  DDB_LCP_DI double& Dell(int i) const { return D(oDell, i); }
  DDB_LCP_DI double& ell(int i) const { return D(oell, i); }
  DDB_LCP_DI double& tmp(int i) const { return D(otmp, i); }
  DDB_LCP_DI int& p(int i) const { return I(op, i); }
  DDB_LCP_DI int& C(int i) const { return I(oC, i); }
  DDB_LCP_DI int& findex(int i) const { return I(ofi, i); }
  DDB_LCP_DI int& state(int i) const { return I(ost, i); }
};

// ---- lcp.cpp:205-243 swapRowsAndCols (ROWPTRS variant) ----
DDB_LCP_DI void lcpSwapRowsAndCols(const LcpView& v, int i1, int i2, int fast) {
  const int r1 = v.I(v.oar, i1), r2 = v.I(v.oar, i2);
  for (int i = i1 + 1; i < i2; ++i) {
    double t = v.A(i, i1);
    v.Araw(r1, i) = t;
    v.A(i, i1) = v.Araw(r2, i);
  }
  v.Araw(r1, i2) = v.Araw(r1, i1);
  v.Araw(r1, i1) = v.Araw(r2, i1);
  v.Araw(r2, i1) = v.Araw(r2, i2);
  if (fast) {
    v.I(v.oar, i1) = r2;
    v.I(v.oar, i2) = r1;
  } else {
    for (int k = 0; k <= i2; ++k) {
      double t = v.Araw(r1, k);
      v.Araw(r1, k) = v.Araw(r2, k);
      v.Araw(r2, k) = t;
    }
  }
  for (int j = i2 + 1; j < v.n; ++j) {
    double t = v.A(j, i1);
    v.A(j, i1) = v.A(j, i2);
    v.A(j, i2) = t;
  }
}

// ---- lcp.cpp:245-300 swapProblem ----
DDB_LCP_DI void lcpSwapProblem(const LcpView& v, int i1, int i2, int fast) {
  if (i1 == i2) return;
  lcpSwapRowsAndCols(v, i1, i2, fast);
  double t;
  t = v.x(i1); v.x(i1) = v.x(i2); v.x(i2) = t;
  t = v.b(i1); v.b(i1) = v.b(i2); v.b(i2) = t;
  t = v.w(i1); v.w(i1) = v.w(i2); v.w(i2) = t;
  t = v.lo(i1); v.lo(i1) = v.lo(i2); v.lo(i2) = t;
  t = v.hi(i1); v.hi(i1) = v.hi(i2); v.hi(i2) = t;
  int ti;
  ti = v.p(i1); v.p(i1) = v.p(i2); v.p(i2) = ti;
  ti = v.state(i1); v.state(i1) = v.state(i2); v.state(i2) = ti;
  ti = v.findex(i1); v.findex(i1) = v.findex(i2); v.findex(i2) = ti;
}

// ---- fastlsolve.cpp: solve L x = b, unit lower-triangular L at (row0, col0) of the L storage, blocks of 4 ----
// rev = false: dSolveL1. rev = true: dSolveL1T (fastltsolve.cpp), the same blocked recurrence on the
// index-reversed system L'(a,b) = L(n-1-b, n-1-a), x'(k) = x(n-1-k).
template <bool REV, class LF, class BF>
DDB_LCP_DI void lcpSolveL1Generic(int n, LF Lf, BF Bf) {
  // Lf(a,b): a > b element of the (possibly reversed) triangular matrix; Bf(k): rhs/solution reference
  int i = 0;
  for (; i <= n - 4; i += 4) {
    double Z1 = 0, Z2 = 0, Z3 = 0, Z4 = 0;
#pragma unroll 4
    for (int j = 0; j < i; ++j) {
      const double q = Bf(j);
      Z1 += Lf(i, j) * q;
      Z2 += Lf(i + 1, j) * q;
      Z3 += Lf(i + 2, j) * q;
      Z4 += Lf(i + 3, j) * q;
    }
    Z1 = Bf(i) - Z1;
    Bf(i) = Z1;
    Z2 = Bf(i + 1) - Z2 - Lf(i + 1, i) * Z1;
    Bf(i + 1) = Z2;
    Z3 = Bf(i + 2) - Z3 - Lf(i + 2, i) * Z1 - Lf(i + 2, i + 1) * Z2;
    Bf(i + 2) = Z3;
    Z4 = Bf(i + 3) - Z4 - Lf(i + 3, i) * Z1 - Lf(i + 3, i + 1) * Z2 - Lf(i + 3, i + 2) * Z3;
    Bf(i + 3) = Z4;
  }
  for (; i < n; ++i) {
    double Z = 0;
#pragma unroll 8
    for (int j = 0; j < i; ++j) Z += Lf(i, j) * Bf(j);
    Bf(i) = Bf(i) - Z;
  }
}
// dSolveL1(L, B, n): B is one of the view's vectors (offset boff)
DDB_LCP_DI void lcpSolveL1(const LcpView& v, int boff, int n) {
  lcpSolveL1Generic<false>(n, [&](int a, int b) -> double { return v.L(a, b); }, [&](int k) -> double& { return v.D(boff, k); });
}
DDB_LCP_DI void lcpSolveL1T(const LcpView& v, int boff, int n) {
  lcpSolveL1Generic<true>(n, [&](int a, int b) -> double { return v.L(n - 1 - b, n - 1 - a); },
                          [&](int k) -> double& { return v.D(boff, n - 1 - k); });
}
// fastdot.cpp: sum of products, left to right
DDB_LCP_DI double lcpDot(const LcpView& v, int aoff, int boff, int n) {
  double s = 0;
#pragma unroll 8
  for (int k = 0; k < n; ++k) s += v.D(aoff, k) * v.D(boff, k);
  return s;
}

// ---- matrix.cpp:286-366 dLDLTAddTL on the sub-factor starting at (r0, r0) ----
DDB_LCP_DI void lcpLDLTAddTL(const LcpView& v, int r0, int aoff /*a in tmp region*/, int n, int w1off, int w2off) {
  if (n < 2) return;
  const double SQ = 0.70710678118654752440;  // M_SQRT1_2
  v.D(w1off, 0) = 0.0;
  v.D(w2off, 0) = 0.0;
  for (int j = 1; j < n; ++j) {
    double t = v.D(aoff, j) * SQ;
    v.D(w1off, j) = t;
    v.D(w2off, j) = t;
  }
  const double a0 = v.D(aoff, 0);
  const double W11 = (0.5 * a0 + 1) * SQ;
  const double W21 = (0.5 * a0 - 1) * SQ;
  double alpha1 = 1.0, alpha2 = 1.0;
  {
    double dee = v.d(r0);
    double alphanew = alpha1 + (W11 * W11) * dee;
    dee /= alphanew;
    double gamma1 = W11 * dee;
    dee *= alpha1;
    alpha1 = alphanew;
    alphanew = alpha2 - (W21 * W21) * dee;
    dee /= alphanew;
    alpha2 = alphanew;
    double k1 = 1.0 - W21 * gamma1;
    double k2 = W21 * gamma1 * W11 - W21;
    for (int p = 1; p < n; ++p) {
      double Wp = v.D(w1off, p);
      double ell = v.L(r0 + p, r0);
      v.D(w1off, p) = Wp - W11 * ell;
      v.D(w2off, p) = k1 * Wp + k2 * ell;
    }
  }
  for (int j = 1; j < n; ++j) {
    double k1 = v.D(w1off, j);
    double k2 = v.D(w2off, j);
    double dee = v.d(r0 + j);
    double alphanew = alpha1 + (k1 * k1) * dee;
    dee /= alphanew;
    double gamma1 = k1 * dee;
    dee *= alpha1;
    alpha1 = alphanew;
    alphanew = alpha2 - (k2 * k2) * dee;
    dee /= alphanew;
    double gamma2 = k2 * dee;
    dee *= alpha2;
    v.d(r0 + j) = dee;
    alpha2 = alphanew;
    for (int p = j + 1; p < n; ++p) {
      double ell = v.L(r0 + p, r0 + j);
      double Wp = v.D(w1off, p) - k1 * ell;
      ell += gamma1 * Wp;
      v.D(w1off, p) = Wp;
      Wp = v.D(w2off, p) - k2 * ell;
      ell -= gamma2 * Wp;
      v.D(w2off, p) = Wp;
      v.L(r0 + p, r0 + j) = ell;
    }
  }
}

// ---- matrix.cpp:440-470 dRemoveRowCol on L ----
DDB_LCP_DI void lcpRemoveRowCol(const LcpView& v, int n, int r) {
  if (r >= n - 1) return;
  if (r > 0) {
    for (int i = 0; i < r; ++i)
      for (int k = 0; k < n - r - 1; ++k) v.L(i, r + k) = v.L(i, r + 1 + k);
    for (int i = r; i < n - 1; ++i)
      for (int k = 0; k < r; ++k) v.L(i, k) = v.L(i + 1, k);
  }
  for (int i = r; i < n - 1; ++i)
    for (int k = 0; k < n - r - 1; ++k) v.L(i, r + k) = v.L(i + 1, r + 1 + k);
}

// ---- matrix.cpp:380-437 dLDLTRemove(A, p = C, L, d, n1, n2 = nC, r) ----
DDB_LCP_DI double lcpGetA(const LcpView& v, int i, int j) { return (i > j) ? v.A(i, j) : v.A(j, i); }
DDB_LCP_DI void lcpLDLTRemove(const LcpView& v, int n2, int r) {
  if (r == n2 - 1) return;  // deleting the last row/col is easy
  const int w1 = v.otmp, w2 = v.otmp + v.nskip, ta = v.otmp + 2 * v.nskip;  // W1, W2, then t|a
  if (r == 0) {
    const int p0 = v.C(0);
    for (int i = 0; i < n2; ++i) v.D(ta, i) = -lcpGetA(v, v.C(i), p0);
    v.D(ta, 0) += 1.0;
    lcpLDLTAddTL(v, 0, ta, n2, w1, w2);
  } else {
    for (int i = 0; i < r; ++i) v.D(ta, i) = v.L(r, i) / v.d(i);
    const int a = ta + r;
    const int pr = v.C(r);
    for (int i = 0; i < n2 - r; ++i) {
      double s = 0;
      for (int k = 0; k < r; ++k) s += v.L(r + i, k) * v.D(ta, k);
      v.D(a, i) = s - lcpGetA(v, v.C(r + i), pr);
    }
    v.D(a, 0) += 1.0;
    lcpLDLTAddTL(v, r, a, n2 - r, w1, w2);
  }
  lcpRemoveRowCol(v, n2, r);
  for (int k = r; k < n2 - 1; ++k) v.d(k) = v.d(k + 1);
}

// plain (not order-exact) LDL^T of the leading nub x nub block for the unbounded variables;
// contact LCPs have nub = 0 and never reach it (lcp.cpp:470-487 uses dFactorLDLT/dSolveLDLT)
DDB_LCP_DI void lcpFactorNub(const LcpView& v, int nub) {
  for (int j = 0; j < nub; ++j)
    for (int k = 0; k <= j; ++k) v.L(j, k) = v.A(j, k);
  for (int j = 0; j < nub; ++j) {
    double dj = v.L(j, j);
    for (int k = 0; k < j; ++k) dj -= v.L(j, k) * v.L(j, k) / v.d(k);
    v.d(j) = 1.0 / dj;
    for (int i = j + 1; i < nub; ++i) {
      double s = v.L(i, j);
      for (int k = 0; k < j; ++k) s -= v.L(i, k) * v.L(j, k) / v.d(k);
      v.L(i, j) = s * v.d(j);
    }
  }
}

struct DantzigState {
  int nC, nN, nub;
};

// dLCP::solve1 (lcp.cpp:703-752)
DDB_LCP_DI void lcpSolve1(const LcpView& v, const DantzigState& st, int i, int dir, int only_transfer) {
  if (st.nC > 0) {
    int j = 0;
    for (; j < st.nub; ++j) v.Dell(j) = v.A(i, j);
    for (; j < st.nC; ++j) v.Dell(j) = v.A(i, v.C(j));
    lcpSolveL1(v, v.oDell, st.nC);
    for (j = 0; j < st.nC; ++j) v.ell(j) = v.Dell(j) * v.d(j);
    if (!only_transfer) {
      for (j = 0; j < st.nC; ++j) v.dw(j) = v.ell(j);  // tmp aliases delta_w (lcp.cpp:825-827)
      lcpSolveL1T(v, v.odw, st.nC);
      if (dir > 0) {
        for (j = 0; j < st.nC; ++j) v.dx(v.C(j)) = -v.dw(j);
      } else {
        for (j = 0; j < st.nC; ++j) v.dx(v.C(j)) = v.dw(j);
      }
    }
  }
}
// dLCP::transfer_i_to_C (lcp.cpp:517-545)
DDB_LCP_DI void lcpTransferToC(const LcpView& v, DantzigState& st, int i) {
  if (st.nC > 0) {
    for (int j = 0; j < st.nC; ++j) v.L(st.nC, j) = v.ell(j);
    v.d(st.nC) = 1.0 / (v.A(i, i) - lcpDot(v, v.oell, v.oDell, st.nC));
  } else {
    v.d(0) = 1.0 / v.A(i, i);
  }
  lcpSwapProblem(v, st.nC, i, 1);
  v.C(st.nC) = st.nC;
  st.nC++;
}
// dLCP::transfer_i_from_N_to_C (lcp.cpp:548-596)
DDB_LCP_DI void lcpTransferNtoC(const LcpView& v, DantzigState& st, int i) {
  if (st.nC > 0) {
    int j = 0;
    for (; j < st.nub; ++j) v.Dell(j) = v.A(i, j);
    for (; j < st.nC; ++j) v.Dell(j) = v.A(i, v.C(j));
    lcpSolveL1(v, v.oDell, st.nC);
    for (j = 0; j < st.nC; ++j) {
      double e = v.Dell(j) * v.d(j);
      v.ell(j) = e;
      v.L(st.nC, j) = e;
    }
    v.d(st.nC) = 1.0 / (v.A(i, i) - lcpDot(v, v.oell, v.oDell, st.nC));
  } else {
    v.d(0) = 1.0 / v.A(i, i);
  }
  lcpSwapProblem(v, st.nC, i, 1);
  v.C(st.nC) = st.nC;
  st.nN--;
  st.nC++;
}
// dLCP::transfer_i_from_C_to_N (lcp.cpp:599-646)
DDB_LCP_DI void lcpTransferCtoN(const LcpView& v, DantzigState& st, int i) {
  const int nC = st.nC;
  int last_idx = -1;
  int j = 0;
  for (; j < nC; ++j) {
    if (v.C(j) == nC - 1) last_idx = j;
    if (v.C(j) == i) {
      lcpLDLTRemove(v, nC, j);
      int k;
      if (last_idx == -1) {
        for (k = j + 1; k < nC; ++k)
          if (v.C(k) == nC - 1) break;
      } else {
        k = last_idx;
      }
      v.C(k) = v.C(j);
      for (int m = j; m < nC - 1; ++m) v.C(m) = v.C(m + 1);
      break;
    }
  }
  lcpSwapProblem(v, i, nC - 1, 1);
  st.nN++;
  st.nC = nC - 1;
}

// dSolveLCP (lcp.cpp:780-1114) with w = nullptr as DantzigBoxedLcpSolver::solve calls it. The view holds
// A (row i at physical row i), x (ignored: zeroed, lcp.cpp:411), b, lo, hi, findex on entry and the
// solution in x on exit. Returns false only for earlyTermination && s <= 0 (lcp.cpp:1007-1027).
DDB_LCP_DI bool dantzigSolve(const LcpView& v, int nub_in, bool earlyTermination) {
  const int n = v.n;
  const double INF = INFINITY;
  DantzigState st;
  st.nC = 0; st.nN = 0; st.nub = nub_in;
  for (int k = 0; k < n; ++k) { v.x(k) = 0.0; v.I(v.oar, k) = k; v.p(k) = k; v.state(k) = 0; v.w(k) = 0.0; }
  // all unbounded variables to the front (lcp.cpp:456-468); findex >= 0 rows are never "unbounded"
  for (int k = st.nub; k < n; ++k) {
    if (v.findex(k) >= 0) continue;
    if (v.lo(k) == -INF && v.hi(k) == INF) {
      lcpSwapProblem(v, st.nub, k, 0);
      st.nub++;
    }
  }
  if (st.nub > 0) {
    lcpFactorNub(v, st.nub);
    for (int k = 0; k < st.nub; ++k) v.x(k) = v.b(k);
    lcpSolveL1(v, v.ox, st.nub);
    for (int k = 0; k < st.nub; ++k) v.x(k) *= v.d(k);
    lcpSolveL1T(v, v.ox, st.nub);
    for (int k = 0; k < st.nub; ++k) { v.w(k) = 0.0; v.C(k) = k; }
    st.nC = st.nub;
  }
  // friction rows to the end (lcp.cpp:490-502)
  {
    int num_at_end = 0;
    for (int k = n - 1; k >= st.nub; k--) {
      if (v.findex(k) >= 0) {
        lcpSwapProblem(v, k, n - 1 - num_at_end, 1);
        num_at_end++;
      }
    }
  }
  const int adj_nub = st.nub;
  bool hit_first_friction_index = false;
  for (int i = adj_nub; i < n; ++i) {
    bool s_error = false;
    if (!hit_first_friction_index && v.findex(i) >= 0) {
      // freeze the friction bounds at |hi * x[findex]| (lcp.cpp:856-873)
      for (int j = 0; j < n; ++j) v.dw(v.p(j)) = v.x(j);
      for (int k = i; k < n; ++k) {
        double wfk = v.dw(v.findex(k));
        if (wfk == 0) {
          v.hi(k) = 0;
          v.lo(k) = 0;
        } else {
          v.hi(k) = fabs(v.hi(k) * wfk);
          v.lo(k) = -v.hi(k);
        }
      }
      hit_first_friction_index = true;
    }
    // w[i] = A(i,C) x(C) + A(i,N) x(N) - b[i]
    {
      double s1 = 0, s2 = 0;
#pragma unroll 8
      for (int k = 0; k < st.nC; ++k) s1 += v.A(i, k) * v.x(k);
#pragma unroll 8
      for (int k = 0; k < st.nN; ++k) s2 += v.A(i, st.nC + k) * v.x(st.nC + k);
      v.w(i) = s1 + s2 - v.b(i);
    }
    if (v.lo(i) == 0 && v.w(i) >= 0) {
      st.nN++;
      v.state(i) = 0;
    } else if (v.hi(i) == 0 && v.w(i) <= 0) {
      st.nN++;
      v.state(i) = 1;
    } else if (v.w(i) == 0) {
      lcpSolve1(v, st, i, 0, 1);
      lcpTransferToC(v, st, i);
    } else {
      for (;;) {
        int dir;
        double dirf;
        if (v.w(i) <= 0) { dir = 1; dirf = 1.0; } else { dir = -1; dirf = -1.0; }
        lcpSolve1(v, st, i, dir, 0);
        // delta_w(N) = A(N,C) delta_x(C) + dir * A(N,i)
        for (int k = 0; k < st.nN; ++k) {
          double s = 0;
#pragma unroll 8
          for (int c = 0; c < st.nC; ++c) s += v.A(st.nC + k, c) * v.dx(c);
          v.dw(st.nC + k) = s;
        }
        if (dir > 0) {
          for (int k = 0; k < st.nN; ++k) v.dw(st.nC + k) += v.A(i, st.nC + k);
        } else {
          for (int k = 0; k < st.nN; ++k) v.dw(st.nC + k) -= v.A(i, st.nC + k);
        }
        {
          double s = 0;
          for (int c = 0; c < st.nC; ++c) s += v.A(i, c) * v.dx(c);
          v.dw(i) = s + v.A(i, i) * dirf;
        }
        int cmd = 1, si = 0;
        double s = -v.w(i) / v.dw(i);
        if (dir > 0) {
          if (v.hi(i) < INF) {
            double s2 = (v.hi(i) - v.x(i)) * dirf;
            if (s2 < s) { s = s2; cmd = 3; }
          }
        } else {
          if (v.lo(i) > -INF) {
            double s2 = (v.lo(i) - v.x(i)) * dirf;
            if (s2 < s) { s = s2; cmd = 2; }
          }
        }
        for (int k = 0; k < st.nN; ++k) {
          const int ik = st.nC + k;
          const double dwk = v.dw(ik);
          if (!v.state(ik) ? dwk < 0 : dwk > 0) {
            if (v.lo(ik) == 0 && v.hi(ik) == 0) continue;
            double s2 = -v.w(ik) / dwk;
            if (s2 < s) { s = s2; cmd = 4; si = ik; }
          }
        }
        for (int k = adj_nub; k < st.nC; ++k) {
          const double dxk = v.dx(k);
          if (dxk < 0 && v.lo(k) > -INF) {
            double s2 = (v.lo(k) - v.x(k)) / dxk;
            if (s2 < s) { s = s2; cmd = 5; si = k; }
          }
          if (dxk > 0 && v.hi(k) < INF) {
            double s2 = (v.hi(k) - v.x(k)) / dxk;
            if (s2 < s) { s = s2; cmd = 6; si = k; }
          }
        }
        if (s <= 0.0) {
          if (earlyTermination) return false;
          for (int k = i; k < n; ++k) { v.x(k) = 0.0; v.w(k) = 0.0; }
          s_error = true;
          break;
        }
        for (int c = 0; c < st.nC; ++c) v.x(c) += s * v.dx(c);
        v.x(i) += s * dirf;
        for (int k = 0; k < st.nN; ++k) v.w(st.nC + k) += s * v.dw(st.nC + k);
        v.w(i) += s * v.dw(i);
        switch (cmd) {
          case 1: v.w(i) = 0; lcpTransferToC(v, st, i); break;
          case 2: v.x(i) = v.lo(i); v.state(i) = 0; st.nN++; break;
          case 3: v.x(i) = v.hi(i); v.state(i) = 1; st.nN++; break;
          case 4: v.w(si) = 0; lcpTransferNtoC(v, st, si); break;
          case 5: v.x(si) = v.lo(si); v.state(si) = 0; lcpTransferCtoN(v, st, si); break;
          case 6: v.x(si) = v.hi(si); v.state(si) = 1; lcpTransferCtoN(v, st, si); break;
        }
        if (cmd <= 3) break;
      }
    }
    if (s_error) break;
  }
  // unpermute x (and w) (lcp.cpp:755-772)
  for (int j = 0; j < n; ++j) v.dw(j) = v.x(j);
  for (int j = 0; j < n; ++j) v.x(v.p(j)) = v.dw(j);
  for (int j = 0; j < n; ++j) v.dw(j) = v.w(j);
  for (int j = 0; j < n; ++j) v.w(v.p(j)) = v.dw(j);
  return true;
}

// PgsBoxedLcpSolver::solve (PgsBoxedLcpSolver.cpp:79-262) with the options BoxedLcpConstraintSolver sets
// (:81-83): maxIteration, deltaXThreshold, relativeDeltaXTolerance, epsilonForDivision, no shuffling.
// A (physical rows, destroyed by the row normalisation), b, lo, hi, findex in the view; x is the warm start.
// order[] lives in the view's C array.
struct PgsOptions {
  int maxIteration;
  double deltaXThreshold, relativeDeltaXTolerance, epsilonForDivision;
};
DDB_LCP_DI bool pgsSolve(const LcpView& v, const PgsOptions& opt) {
  const int n = v.n;
  int norder = 0;
  bool possibleToTerminate = true;
  for (int i = 0; i < n; ++i) {
    if (v.Araw(i, i) < opt.epsilonForDivision) {
      v.x(i) = 0.0;
      continue;
    }
    v.C(norder++) = i;
    const double old_x = v.x(i);
    double new_x = v.b(i);
#pragma unroll 8
    for (int j = 0; j < i; ++j) new_x -= v.Araw(i, j) * v.x(j);
#pragma unroll 8
    for (int j = i + 1; j < n; ++j) new_x -= v.Araw(i, j) * v.x(j);
    new_x /= v.Araw(i, i);
    const int fi = v.findex(i);
    if (fi >= 0) {
      const double hi_tmp = v.hi(i) * v.x(fi), lo_tmp = -hi_tmp;
      v.x(i) = new_x > hi_tmp ? hi_tmp : (new_x < lo_tmp ? lo_tmp : new_x);
    } else {
      v.x(i) = new_x > v.hi(i) ? v.hi(i) : (new_x < v.lo(i) ? v.lo(i) : new_x);
    }
    if (possibleToTerminate && fabs(v.x(i) - old_x) > opt.deltaXThreshold) possibleToTerminate = false;
  }
  if (possibleToTerminate) return true;
  for (int o = 0; o < norder; ++o) {
    const int index = v.C(o);
    const double dummy = 1.0 / v.Araw(index, index);
    v.b(index) *= dummy;
    for (int j = 0; j < n; ++j) v.Araw(index, j) *= dummy;
  }
  for (int iter = 1; iter < opt.maxIteration; ++iter) {
    possibleToTerminate = true;
    for (int o = 0; o < norder; ++o) {
      const int index = v.C(o);
      double new_x = v.b(index);
      const double old_x = v.x(index);
#pragma unroll 8
      for (int j = 0; j < index; ++j) new_x -= v.Araw(index, j) * v.x(j);
#pragma unroll 8
      for (int j = index + 1; j < n; ++j) new_x -= v.Araw(index, j) * v.x(j);
      const int fi = v.findex(index);
      double xi;
      if (fi >= 0) {
        const double hi_tmp = v.hi(index) * v.x(fi), lo_tmp = -hi_tmp;
        xi = new_x > hi_tmp ? hi_tmp : (new_x < lo_tmp ? lo_tmp : new_x);
      } else {
        xi = new_x > v.hi(index) ? v.hi(index) : (new_x < v.lo(index) ? v.lo(index) : new_x);
      }
      v.x(index) = xi;
      if (possibleToTerminate && fabs(xi) > opt.epsilonForDivision) {
        const double rel = fabs((xi - old_x) / xi);
        if (rel > opt.relativeDeltaXTolerance) possibleToTerminate = false;
      }
    }
    if (possibleToTerminate) break;
  }
  return possibleToTerminate;
}

// Minimum-norm least squares  min |x| s.t. |Q x - c| minimal  for an n x n system, standing in for Eigen's
// completeOrthogonalDecomposition().solve() (ConstrainedGroupGradientMatrices.cpp:265, LCPUtils.cpp:114,134):
// Householder QR with column pivoting and LAPACK-style norm downdating, rank = #{|R_ii| > eps n |R_max|}, then
// reflections from the right that annihilate R12 (stored in place), back-substitution, un-permutation.
// Storage: Q in the view's L area (leading dimension n), rhs in dw (overwritten), result in dx; Dell/ell/d/tmp and the
// int array p are scratch.
DDB_LCP_DI void coopSync(const LcpView& v) {
#ifdef __CUDA_ARCH__
  if (v.nl > 1) __syncwarp();
#endif
}
// base + (first lane whose flag is set), -1 when none: the candidates base .. base + nl - 1 are tested side by side
DDB_LCP_DI int coopFirstTrue(const LcpView& v, bool flag, int base) {
#ifdef __CUDA_ARCH__
  if (v.nl > 1) {
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    return m ? base + __ffs(m) - 1 : -1;
  }
#endif
  return flag ? base : -1;
}
// first index of the largest value over j in [from, n), evaluated by the whole warp (or by the one thread)
template <class F>
DDB_LCP_DI int coopArgMaxFirst(const LcpView& v, int from, int n, F val) {
  int best = -1;
  double bv = 0.0;
  for (int j = from + v.lane; j < n; j += v.nl) { const double x = val(j); if (best < 0 || x > bv) { best = j; bv = x; } }
#ifdef __CUDA_ARCH__
  if (v.nl > 1) {
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int ob = __shfl_xor_sync(0xffffffffu, best, off);
      if (ob >= 0 && (best < 0 || ov > bv || (ov == bv && ob < best))) { best = ob; bv = ov; }
    }
  }
#endif
  return best;
}

// Householder QR with column pivoting + the right-hand reflections of the complete orthogonal decomposition.
// WITH_RHS: the left reflections are applied to dw on the fly (codSolveDev); otherwise their scales are kept in ell
// (codFactorDev). Loops over independent columns / rows are split across the lanes when one warp owns the world.
template <bool WITH_RHS>
DDB_LCP_DI int codFactorImpl(const LcpView& v, int n) {
  auto Q = [&](int i, int j) -> double& { return v.db[(size_t)(v.oL + i * n + j) * v.st]; };
  const double eps = 2.220446049250313e-16, tiny = 2.2250738585072014e-308;
  const int lane = v.lane, nl = v.nl;
  coopSync(v);  // the lanes filled the matrix redundantly: all of those stores land before any lane starts to transform it
  for (int j = lane; j < n; j += nl) {
    v.p(j) = j;
    double s = 0;
    for (int i = 0; i < n; i++) s += Q(i, j) * Q(i, j);
    v.Dell(j) = v.ell(j) = sqrt(s);
  }
  coopSync(v);
  double maxNorm = 0;
  for (int j = 0; j < n; j++) maxNorm = fmax(maxNorm, v.Dell(j));
  const double thresholdHelper = (maxNorm * eps) * (maxNorm * eps) / double(n);
  const double downdateThreshold = sqrt(eps);
  int nonzeroPivots = n;
  double maxPivot = 0;
  for (int k = 0; k < n; k++) {
    const int big = coopArgMaxFirst(v, k, n, [&](int j) { return v.Dell(j); });
    const double bigSq = v.Dell(big) * v.Dell(big);
    if (nonzeroPivots == n && bigSq < thresholdHelper * double(n - k)) nonzeroPivots = k;
    coopSync(v);
    if (big != k) {
      for (int i = lane; i < n; i += nl) { double t = Q(i, k); Q(i, k) = Q(i, big); Q(i, big) = t; }
      if (lane == 0) {
        double t = v.Dell(k); v.Dell(k) = v.Dell(big); v.Dell(big) = t;
        t = v.ell(k); v.ell(k) = v.ell(big); v.ell(big) = t;
        int ti = v.p(k); v.p(k) = v.p(big); v.p(big) = ti;
      }
      coopSync(v);
    }
    double c0 = Q(k, k), tailSq = 0;
    for (int i = k + 1; i < n; i++) tailSq += Q(i, k) * Q(i, k);
    double beta, t;
    coopSync(v);
    if (tailSq <= tiny) {
      t = 0; beta = c0;
      for (int i = k + 1 + lane; i < n; i += nl) Q(i, k) = 0;
    } else {
      beta = sqrt(c0 * c0 + tailSq);
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1 + lane; i < n; i += nl) Q(i, k) /= (c0 - beta);
      t = (beta - c0) / beta;
    }
    if (lane == 0) Q(k, k) = beta;
    if (fabs(beta) > maxPivot) maxPivot = fabs(beta);
    coopSync(v);
    for (int j = k + 1 + lane; j < n; j += nl) {
      double s = Q(k, j);
      for (int i = k + 1; i < n; i++) s += Q(i, k) * Q(i, j);
      s *= t;
      Q(k, j) -= s;
      for (int i = k + 1; i < n; i++) Q(i, j) -= s * Q(i, k);
      // LAPACK-style column norm downdate of the same column
      if (v.Dell(j) != 0) {
        double temp = fabs(Q(k, j)) / v.Dell(j);
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0 ? 0 : temp;
        const double r = v.Dell(j) / v.ell(j);
        const double temp2 = temp * r * r;
        if (temp2 <= downdateThreshold) {
          double s2 = 0;
          for (int i = k + 1; i < n; i++) s2 += Q(i, j) * Q(i, j);
          v.ell(j) = v.Dell(j) = sqrt(s2);
        } else {
          v.Dell(j) *= sqrt(temp);
        }
      }
    }
    if (WITH_RHS) {
      double s = v.dw(k);
      for (int i = k + 1; i < n; i++) s += Q(i, k) * v.dw(i);
      s *= t;
      coopSync(v);
      if (lane == 0) v.dw(k) -= s;
      for (int i = k + 1 + lane; i < n; i += nl) v.dw(i) -= s * Q(i, k);
    } else {
      if (lane == 0) v.ell(k) = t;  // column k is final: its norm slots are free
    }
    coopSync(v);
  }
  int rank = 0;
  const double premult = maxPivot * eps * double(n);
  for (int i = 0; i < nonzeroPivots; i++) if (fabs(Q(i, i)) > premult) rank++;
  if (rank > 0 && rank < n) {
    for (int k = rank - 1; k >= 0; k--) {
      double c0 = Q(k, k), tailSq = 0;
      for (int j = rank; j < n; j++) tailSq += Q(k, j) * Q(k, j);
      double beta, t;
      coopSync(v);
      if (tailSq <= tiny) { t = 0; beta = c0; for (int j = rank + lane; j < n; j += nl) Q(k, j) = 0; }
      else {
        beta = sqrt(c0 * c0 + tailSq);
        if (c0 >= 0) beta = -beta;
        for (int j = rank + lane; j < n; j += nl) Q(k, j) /= (c0 - beta);  // reflector vector kept in place of the annihilated row
        t = (beta - c0) / beta;
      }
      coopSync(v);
      for (int i = lane; i < k; i += nl) {
        double s = Q(i, k);
        for (int j = rank; j < n; j++) s += Q(i, j) * Q(k, j);
        s *= t;
        Q(i, k) -= s;
        for (int j = rank; j < n; j++) Q(i, j) -= s * Q(k, j);
      }
      if (lane == 0) { Q(k, k) = beta; v.d(k) = t; }
      coopSync(v);
    }
  }
  return rank;
}
// tmp <- solution of the triangular system + the right-hand reflections; every lane runs it (O(n^2), sequential sums)
DDB_LCP_DI void codBackSolve(const LcpView& v, int n, int rank) {
  auto Q = [&](int i, int j) -> double& { return v.db[(size_t)(v.oL + i * n + j) * v.st]; };
  coopSync(v);
  for (int i = rank - 1; i >= 0; i--) {
    double s = v.dw(i);
    for (int j = i + 1; j < rank; j++) s -= Q(i, j) * v.tmp(j);
    coopSync(v);
    if (v.lane == 0) v.tmp(i) = s / Q(i, i);
    coopSync(v);
  }
  for (int j = rank + v.lane; j < n; j += v.nl) v.tmp(j) = 0.0;
  coopSync(v);
  if (rank < n) {
    for (int k = 0; k < rank; k++) {
      double s = v.tmp(k);
      for (int j = rank; j < n; j++) s += Q(k, j) * v.tmp(j);
      s *= v.d(k);
      coopSync(v);
      if (v.lane == 0) v.tmp(k) -= s;
      for (int j = rank + v.lane; j < n; j += v.nl) v.tmp(j) -= s * Q(k, j);
      coopSync(v);
    }
  }
}
DDB_LCP_DI int codSolveDev(const LcpView& v, int n) {
  const int rank = codFactorImpl<true>(v, n);
  for (int j = v.lane; j < n; j += v.nl) v.dx(j) = 0.0;
  coopSync(v);
  if (rank == 0) return 0;
  codBackSolve(v, n, rank);
  for (int j = v.lane; j < n; j += v.nl) v.dx(v.p(j)) = v.tmp(j);
  coopSync(v);
  return rank;
}

// The same complete orthogonal decomposition kept as a factorisation so that P b and P^T a (P = Q^+) come out of ONE
// decomposition, as Eigen's `Qfactored.solve(b)` / `Qfactored.pseudoInverse()` do at BackpropSnapshot.cpp:3875-3876,
// 3925-3932. Storage as codSolveDev; the QR reflector scales are kept in ell, those of the right-hand reflections in d.
DDB_LCP_DI int codFactorDev(const LcpView& v, int n) { return codFactorImpl<false>(v, n); }
// out = P rhs (transposed = false) or P^T rhs (transposed = true) from codFactorDev's factorisation; dw / tmp are scratch
template <class RhsF, class OutF>
DDB_LCP_DI void codApplyDev(const LcpView& v, int n, int rank, bool transposed, RhsF rhs, OutF out) {
  auto Q = [&](int i, int j) -> double& { return v.db[(size_t)(v.oL + i * n + j) * v.st]; };
  const int lane = v.lane, nl = v.nl;
  if (rank == 0) { for (int j = 0; j < n; j++) out(j, 0.0); return; }
  coopSync(v);
  if (!transposed) {
    for (int i = lane; i < n; i += nl) v.dw(i) = rhs(i);
    coopSync(v);
    for (int k = 0; k < n; k++) {
      double s = v.dw(k);
      for (int i = k + 1; i < n; i++) s += Q(i, k) * v.dw(i);
      s *= v.ell(k);
      coopSync(v);
      if (lane == 0) v.dw(k) -= s;
      for (int i = k + 1 + lane; i < n; i += nl) v.dw(i) -= s * Q(i, k);
      coopSync(v);
    }
    codBackSolve(v, n, rank);
    for (int j = lane; j < n; j += nl) v.dw(v.p(j)) = v.tmp(j);
    coopSync(v);
    for (int j = 0; j < n; j++) out(j, v.dw(j));
  } else {
    for (int j = lane; j < n; j += nl) v.tmp(j) = rhs(v.p(j));
    coopSync(v);
    if (rank < n) {
      for (int k = rank - 1; k >= 0; k--) {
        double s = v.tmp(k);
        for (int j = rank; j < n; j++) s += Q(k, j) * v.tmp(j);
        s *= v.d(k);
        coopSync(v);
        if (lane == 0) v.tmp(k) -= s;
        for (int j = rank + lane; j < n; j += nl) v.tmp(j) -= s * Q(k, j);
        coopSync(v);
      }
    }
    for (int i = 0; i < rank; i++) {  // T^T y = z
      double s = v.tmp(i);
      for (int j = 0; j < i; j++) s -= Q(j, i) * v.dw(j);
      coopSync(v);
      if (lane == 0) v.dw(i) = s / Q(i, i);
      coopSync(v);
    }
    for (int j = rank + lane; j < n; j += nl) v.dw(j) = 0.0;
    coopSync(v);
    for (int k = n - 1; k >= 0; k--) {
      double s = v.dw(k);
      for (int i = k + 1; i < n; i++) s += Q(i, k) * v.dw(i);
      s *= v.ell(k);
      coopSync(v);
      if (lane == 0) v.dw(k) -= s;
      for (int i = k + 1 + lane; i < n; i += nl) v.dw(i) -= s * Q(i, k);
      coopSync(v);
    }
    for (int j = 0; j < n; j++) out(j, v.dw(j));
  }
  coopSync(v);
}

// LCPUtils::isLCPSolutionValid (LCPUtils.cpp:12-81), tol 1e-5. Af(i,j), xf, bf, hif, lof, fif are accessors
// on the ORIGINAL problem (the solvers destroy their copies).
template <class AF, class XF, class BF, class HF, class LF, class IF>
DDB_LCP_DI bool lcpValid(int n, AF Af, XF xf, BF bf, HF hif, LF lof, IF fif, bool ignoreFrictionIndices) {
  const double tol = 1e-5;
  for (int i = 0; i < n; i++) {
    double vi = 0;
    for (int j = 0; j < n; j++) vi += Af(i, j) * xf(j);
    vi -= bf(i);
    double upperLimit = hif(i), lowerLimit = lof(i);
    const int fi = fif(i);
    if (fi != -1) {
      if (ignoreFrictionIndices) {
        if (xf(i) != 0) return false;
        continue;
      }
      upperLimit *= xf(fi);
      lowerLimit *= xf(fi);
    }
    const double xi = xf(i);
    if (fabs(lowerLimit) < tol && fabs(upperLimit) < tol && fabs(xi) < tol) {
    } else if (fabs(xi - lowerLimit) < tol) {
      if (vi < -tol) return false;
    } else if (fabs(xi - upperLimit) < tol) {
      if (vi > tol) return false;
    } else if (xi > lowerLimit && xi < upperLimit) {
      if (fabs(vi) > tol) return false;
    } else {
      return false;
    }
  }
  return true;
}

}  // namespace ddb
