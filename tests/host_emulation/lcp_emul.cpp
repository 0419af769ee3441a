// TEST INFRASTRUCTURE ONLY — see cuda_shim.h. Host build of diffdart_b200/csrc/lcp.cuh (no FMA contraction:
// -ffp-contract=off) exposing the device routines through a C entry for the CPU tests.
#include "cuda_shim.h"
#define cuda_runtime_h
#include <vector>
namespace std {}
#include "lcp_nocuda.cuh"
using namespace ddb;

extern "C" {
// solver 0 = Dantzig, 1 = PGS. A is n x n row-major. returns ok flag; x_out length n.
int emul_lcp_solve(int n, const double* A, const double* b, const double* lo, const double* hi, const int* findex, const double* x0,
                   int solver, int early, double* x_out) {
  std::vector<double> d(lcpScratchDoubles(n) * 32, 0.0);
  std::vector<int> ib(lcpScratchInts(n) * 32, 0);
  LcpView v;
  v.init(d.data() + 5, ib.data() + 5, n, n);  // lane 5 of the interleaved slice
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) v.Araw(i, j) = A[i * n + j];
    v.b(i) = b[i]; v.lo(i) = lo[i]; v.hi(i) = hi[i]; v.findex(i) = findex[i];
    v.x(i) = x0 ? x0[i] : 0.0;
  }
  bool ok;
  if (solver == 0) ok = dantzigSolve(v, 0, early != 0);
  else { PgsOptions o{10000, 1e-10, 1e-8, 1e-8}; ok = pgsSolve(v, o); }
  for (int i = 0; i < n; i++) x_out[i] = v.x(i);
  return ok ? 1 : 0;
}
int emul_cod_solve(int n, const double* Q, const double* c, double* x) {
  std::vector<double> d(lcpScratchDoubles(n) * 32, 0.0);
  std::vector<int> ib(lcpScratchInts(n) * 32, 0);
  LcpView v;
  v.init(d.data() + 7, ib.data() + 7, n, n);
  for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) v.db[(size_t)(v.oL + i * n + j) * 32] = Q[i * n + j]; v.dw(i) = c[i]; }
  int rank = codSolveDev(v, n);
  for (int i = 0; i < n; i++) x[i] = v.dx(i);
  return rank;
}
int emul_lcp_valid(int n, const double* A, const double* x, const double* b, const double* hi, const double* lo, const int* findex, int ignoreFriction) {
  return lcpValid(n, [&](int i, int j) { return A[i * n + j]; }, [&](int i) { return x[i]; }, [&](int i) { return b[i]; },
                  [&](int i) { return hi[i]; }, [&](int i) { return lo[i]; }, [&](int i) { return findex[i]; }, ignoreFriction != 0) ? 1 : 0;
}
}
