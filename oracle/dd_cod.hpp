// TEST INFRASTRUCTURE ONLY — see dd_spatial.hpp.
//
// dd_cod.hpp — minimum-norm least squares through a complete orthogonal decomposition, standing in for Eigen's
// `completeOrthogonalDecomposition().solve(b)` at the reference's call sites on this path
// (dart/neural/ConstrainedGroupGradientMatrices.cpp:265, dart/constraint/LCPUtils.cpp:114,134,
// dart/neural/BackpropSnapshot.cpp:3875-3876). Eigen (3.3.7 in the reference's CI) is not vendored, so this restates
// its published algorithm: Householder QR with column pivoting (LAPACK-style column-norm downdating, rank = number of
// diagonal entries above eps * size * |max pivot|), then Householder reflections from the right that annihilate the
// trailing block so that A P = Q [T 0; 0 0] Z; x = P Z^T [T^-1 (Q^T b)_1..r; 0].
// PARITY UNPINNED against Eigen itself; pinned here by properties (tests/test_oracle_cpu.py): A^T(Ax - b) = 0 and
// x orthogonal to the null space (minimum norm), on full-rank and rank-deficient matrices.
#pragma once
#include <cmath>
#include <limits>
#include <vector>

namespace ddo {

// A: m x n row-major (copied), b: m. Returns the minimum-norm least-squares solution (n) and the detected rank.
inline std::vector<double> codSolve(int m, int n, const std::vector<double>& Ain, const std::vector<double>& bin, int* rankOut = nullptr) {
  std::vector<double> A = Ain, c = bin, x(n, 0.0);
  if (m == 0 || n == 0) { if (rankOut) *rankOut = 0; return x; }
  const double eps = std::numeric_limits<double>::epsilon();
  const int size = std::min(m, n);
  std::vector<int> perm(n);
  std::vector<double> normsU(n), normsD(n), tau(size, 0.0);
  for (int j = 0; j < n; j++) {
    perm[j] = j;
    double s = 0;
    for (int i = 0; i < m; i++) s += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    normsU[j] = normsD[j] = std::sqrt(s);
  }
  double maxNorm = 0;
  for (int j = 0; j < n; j++) maxNorm = std::max(maxNorm, normsU[j]);
  const double thresholdHelper = (maxNorm * eps) * (maxNorm * eps) / double(m);
  const double downdateThreshold = std::sqrt(eps);
  int nonzeroPivots = size;
  double maxPivot = 0;
  for (int k = 0; k < size; k++) {
    int big = k;
    for (int j = k + 1; j < n; j++) if (normsU[j] > normsU[big]) big = j;
    const double bigSq = normsU[big] * normsU[big];
    if (nonzeroPivots == size && bigSq < thresholdHelper * double(m - k)) nonzeroPivots = k;
    if (big != k) {
      for (int i = 0; i < m; i++) std::swap(A[(size_t)i * n + k], A[(size_t)i * n + big]);
      std::swap(normsU[k], normsU[big]); std::swap(normsD[k], normsD[big]); std::swap(perm[k], perm[big]);
    }
    // Householder reflector for A(k:m, k)
    double c0 = A[(size_t)k * n + k], tailSq = 0;
    for (int i = k + 1; i < m; i++) tailSq += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    double beta, t;
    if (tailSq <= std::numeric_limits<double>::min()) {
      t = 0; beta = c0;
      for (int i = k + 1; i < m; i++) A[(size_t)i * n + k] = 0;
    } else {
      beta = std::sqrt(c0 * c0 + tailSq);
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < m; i++) A[(size_t)i * n + k] /= (c0 - beta);
      t = (beta - c0) / beta;
    }
    tau[k] = t;
    A[(size_t)k * n + k] = beta;
    if (std::abs(beta) > maxPivot) maxPivot = std::abs(beta);
    // apply H = I - t v v^T (v = [1; essential]) to the remaining columns and to the right-hand side
    for (int j = k + 1; j < n; j++) {
      double s = A[(size_t)k * n + j];
      for (int i = k + 1; i < m; i++) s += A[(size_t)i * n + k] * A[(size_t)i * n + j];
      s *= t;
      A[(size_t)k * n + j] -= s;
      for (int i = k + 1; i < m; i++) A[(size_t)i * n + j] -= s * A[(size_t)i * n + k];
    }
    {
      double s = c[k];
      for (int i = k + 1; i < m; i++) s += A[(size_t)i * n + k] * c[i];
      s *= t;
      c[k] -= s;
      for (int i = k + 1; i < m; i++) c[i] -= s * A[(size_t)i * n + k];
    }
    // column-norm downdate
    for (int j = k + 1; j < n; j++) {
      if (normsU[j] != 0) {
        double temp = std::abs(A[(size_t)k * n + j]) / normsU[j];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0 ? 0 : temp;
        double r = normsU[j] / normsD[j];
        double temp2 = temp * r * r;
        if (temp2 <= downdateThreshold) {
          double s = 0;
          for (int i = k + 1; i < m; i++) s += A[(size_t)i * n + j] * A[(size_t)i * n + j];
          normsD[j] = normsU[j] = std::sqrt(s);
        } else {
          normsU[j] *= std::sqrt(temp);
        }
      }
    }
  }
  int rank = 0;
  const double premult = maxPivot * eps * double(size);
  for (int i = 0; i < nonzeroPivots; i++) if (std::abs(A[(size_t)i * n + i]) > premult) rank++;
  if (rankOut) *rankOut = rank;
  if (rank == 0) return x;
  // reflections from the right: [R11 R12] Z = [T 0]
  std::vector<double> ztau(rank, 0.0);
  std::vector<std::vector<double>> zv(rank);
  if (rank < n) {
    for (int k = rank - 1; k >= 0; k--) {
      double c0 = A[(size_t)k * n + k], tailSq = 0;
      for (int j = rank; j < n; j++) tailSq += A[(size_t)k * n + j] * A[(size_t)k * n + j];
      double beta, t;
      std::vector<double> v(n - rank, 0.0);
      if (tailSq <= std::numeric_limits<double>::min()) { t = 0; beta = c0; }
      else {
        beta = std::sqrt(c0 * c0 + tailSq);
        if (c0 >= 0) beta = -beta;
        for (int j = rank; j < n; j++) v[j - rank] = A[(size_t)k * n + j] / (c0 - beta);
        t = (beta - c0) / beta;
      }
      // apply to rows 0..k-1 (columns k and rank..n-1)
      for (int i = 0; i < k; i++) {
        double s = A[(size_t)i * n + k];
        for (int j = rank; j < n; j++) s += A[(size_t)i * n + j] * v[j - rank];
        s *= t;
        A[(size_t)i * n + k] -= s;
        for (int j = rank; j < n; j++) A[(size_t)i * n + j] -= s * v[j - rank];
      }
      A[(size_t)k * n + k] = beta;
      for (int j = rank; j < n; j++) A[(size_t)k * n + j] = 0;
      ztau[k] = t; zv[k] = v;
    }
  }
  // y = T^-1 c(0:rank)
  std::vector<double> y(n, 0.0);
  for (int i = rank - 1; i >= 0; i--) {
    double s = c[i];
    for (int j = i + 1; j < rank; j++) s -= A[(size_t)i * n + j] * y[j];
    y[i] = s / A[(size_t)i * n + i];
  }
  // x_perm = Z [y; 0]: apply Z_0 first ... Z_{rank-1} last
  if (rank < n) {
    for (int k = 0; k < rank; k++) {
      double s = y[k];
      for (int j = rank; j < n; j++) s += zv[k][j - rank] * y[j];
      s *= ztau[k];
      y[k] -= s;
      for (int j = rank; j < n; j++) y[j] -= s * zv[k][j - rank];
    }
  }
  for (int j = 0; j < n; j++) x[perm[j]] = y[j];
  return x;
}

}  // namespace ddo
