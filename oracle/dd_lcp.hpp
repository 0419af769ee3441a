// TEST INFRASTRUCTURE ONLY — see dd_spatial.hpp.
//
// dd_lcp.hpp — CPU restatement of the reference's boxed-LCP helpers that are NOT the Dantzig solver
// (the Dantzig oracle is the reference's own code, oracle/_ref):
//   pgs()         PgsBoxedLcpSolver::solve           dart/constraint/PgsBoxedLcpSolver.cpp:79-262
//   lcpValid()    LCPUtils::isLCPSolutionValid       dart/constraint/LCPUtils.cpp:12-81
#pragma once
#include <cmath>
#include <vector>

namespace ddo {

inline int dPad(int a) { return (a > 1) ? (((a - 1) | 3) + 1) : a; }

// A: n x nskip row-major (destroyed), x warm start -> solution. Options as BoxedLcpConstraintSolver.cpp:81-83.
inline bool pgs(int n, double* A, double* x, double* b, const double* lo, const double* hi, const int* findex, int maxIteration = 10000,
                double deltaXThreshold = 1e-10, double relativeDeltaXTolerance = 1e-8, double epsilonForDivision = 1e-8) {
  const int nskip = dPad(n);
  std::vector<int> order;
  bool possibleToTerminate = true;
  for (int i = 0; i < n; ++i) {
    if (A[nskip * i + i] < epsilonForDivision) { x[i] = 0.0; continue; }
    order.push_back(i);
    const double* Ap = A + nskip * i;
    const double old_x = x[i];
    double new_x = b[i];
    for (int j = 0; j < i; ++j) new_x -= Ap[j] * x[j];
    for (int j = i + 1; j < n; ++j) new_x -= Ap[j] * x[j];
    new_x /= A[nskip * i + i];
    if (findex[i] >= 0) {
      const double hi_tmp = hi[i] * x[findex[i]], lo_tmp = -hi_tmp;
      x[i] = new_x > hi_tmp ? hi_tmp : (new_x < lo_tmp ? lo_tmp : new_x);
    } else {
      x[i] = new_x > hi[i] ? hi[i] : (new_x < lo[i] ? lo[i] : new_x);
    }
    if (possibleToTerminate && std::abs(x[i] - old_x) > deltaXThreshold) possibleToTerminate = false;
  }
  if (possibleToTerminate) return true;
  for (int index : order) {
    const double dummy = 1.0 / A[nskip * index + index];
    b[index] *= dummy;
    for (int j = 0; j < n; ++j) A[nskip * index + j] *= dummy;
  }
  for (int iter = 1; iter < maxIteration; ++iter) {
    possibleToTerminate = true;
    for (int index : order) {
      const double* Ap = A + nskip * index;
      double new_x = b[index];
      const double old_x = x[index];
      for (int j = 0; j < index; j++) new_x -= Ap[j] * x[j];
      for (int j = index + 1; j < n; j++) new_x -= Ap[j] * x[j];
      if (findex[index] >= 0) {
        const double hi_tmp = hi[index] * x[findex[index]], lo_tmp = -hi_tmp;
        x[index] = new_x > hi_tmp ? hi_tmp : (new_x < lo_tmp ? lo_tmp : new_x);
      } else {
        x[index] = new_x > hi[index] ? hi[index] : (new_x < lo[index] ? lo[index] : new_x);
      }
      if (possibleToTerminate && std::abs(x[index]) > epsilonForDivision) {
        const double rel = std::abs((x[index] - old_x) / x[index]);
        if (rel > relativeDeltaXTolerance) possibleToTerminate = false;
      }
    }
    if (possibleToTerminate) break;
  }
  return possibleToTerminate;
}

// A: n x n row-major
inline bool lcpValid(int n, const double* A, const double* x, const double* b, const double* hi, const double* lo, const int* findex,
                     bool ignoreFrictionIndices) {
  const double tol = 1e-5;
  for (int i = 0; i < n; i++) {
    double v = 0;
    for (int j = 0; j < n; j++) v += A[i * n + j] * x[j];
    v -= b[i];
    double upperLimit = hi[i], lowerLimit = lo[i];
    if (findex[i] != -1) {
      if (ignoreFrictionIndices) { if (x[i] != 0) return false; continue; }
      upperLimit *= x[findex[i]];
      lowerLimit *= x[findex[i]];
    }
    if (std::abs(lowerLimit) < tol && std::abs(upperLimit) < tol && std::abs(x[i]) < tol) {
    } else if (std::abs(x[i] - lowerLimit) < tol) { if (v < -tol) return false; }
    else if (std::abs(x[i] - upperLimit) < tol) { if (v > tol) return false; }
    else if (x[i] > lowerLimit && x[i] < upperLimit) { if (std::abs(v) > tol) return false; }
    else return false;
  }
  return true;
}

}  // namespace ddo
