// TEST INFRASTRUCTURE ONLY.
// Thin C shim over the REFERENCE's own, unmodified ODE Dantzig solver
// (/root/reference/dart/external/odelcpsolver/lcp.cpp:780-1114). This file is ours;
// the solver sources are compiled from where they lie under /root/reference by
// oracle/Makefile into oracle/_ref/ (git-ignored) and are never copied into the repo.
// Calling convention mirrors DantzigBoxedLcpSolver::solve
// (dart/constraint/DantzigBoxedLcpSolver.cpp:55-77): w = nullptr, nub = 0,
// exceptions swallowed -> false.
#include "dart/external/odelcpsolver/lcp.h"

extern "C" {
int ddref_dpad(int n) { return dPAD(n); }
// A is n x dPAD(n) row-major and is destroyed; x,b,lo,hi length n; findex length n.
int ddref_dantzig_solve(int n, double* A, double* x, double* b, double* lo, double* hi, int* findex, int earlyTermination) {
  try {
    return dSolveLCP(n, A, x, b, nullptr, 0, lo, hi, findex, earlyTermination != 0) ? 1 : 0;
  } catch (...) {
    return 0;
  }
}
}
