// lemke_api.cu — batched Lemke solver for standard LCPs  w = M z + q >= 0, z >= 0, z.w = 0 :
// replaces dart::lcpsolver::Lemke(M, q, &z) (dart/lcpsolver/Lemke.cpp:61-337, Lemke.hpp) for n_problems independent problems.
//
// One problem per warp (CTA of 32 threads). The reference re-solves B d = Be by an unpivoted Householder QR at every pivot
// (Lemke.cpp:192); here the working copy [B | Be] lives in shared memory, column-major with an odd leading dimension, and the
// lanes split the independent columns of every reflector application and the independent rows of every substitution step, so
// each number sees the same operations in the same order as in a sequential run. The pivot rules (ratio test with the 1e-5
// slack, "artificial variable first", the reference's scan for the largest d) follow Lemke.cpp:195-262 literally; they are
// evaluated redundantly by all lanes from shared memory. B itself (n x n per problem) stays in global scratch: one column is
// replaced per pivot.
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>
#include "../../include/ddb.h"

namespace {

constexpr double kZerTol = 1e-5, kPivTol = 1e-8;  // Lemke.cpp:66-67
constexpr int kMaxIter = 1000;                    // Lemke.cpp:68

struct LemkeSmem {
  double* W;  // [n + 1][ld] column-major: columns 0..n-1 = B, column n = right-hand side
  double *x, *d, *Be, *v;
  int* bas;
  int ld;
  __device__ LemkeSmem(unsigned char* raw, int n) {
    ld = n | 1;
    W = reinterpret_cast<double*>(raw);
    x = W + (size_t)(n + 1) * ld;
    d = x + n; Be = d + n; v = Be + n;
    bas = reinterpret_cast<int*>(v + n);
  }
  static size_t bytes(int n) { return ((size_t)(n + 1) * (n | 1) + 4 * (size_t)n) * sizeof(double) + (size_t)n * sizeof(int); }
};

// d = B^-1 Be through Householder QR of the shared-memory copy (Eigen's householderQr().solve: makeHouseholderInPlace per column,
// applyHouseholderOnTheLeft on the trailing columns and the right-hand side, back substitution on R)
__device__ void qrSolve(const LemkeSmem& s, int n, int lane) {
  const int ld = s.ld;
  double* W = s.W;
  for (int k = 0; k < n; k++) {
    double* ck = W + (size_t)k * ld;
    const double c0 = ck[k];
    double tailSq = 0.0;
    for (int r = k + 1; r < n; r++) tailSq += ck[r] * ck[r];  // every lane: broadcast reads, identical result
    double beta, tau;
    if (tailSq <= DBL_MIN) {
      tau = 0.0; beta = c0;
      for (int r = k + 1 + lane; r < n; r += 32) s.v[r] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tailSq);
      if (c0 >= 0.0) beta = -beta;
      const double inv = c0 - beta;
      for (int r = k + 1 + lane; r < n; r += 32) s.v[r] = ck[r] / inv;
      tau = (beta - c0) / beta;
    }
    __syncwarp();
    for (int j = k + 1 + lane; j <= n; j += 32) {  // trailing columns and the right-hand side, one per lane
      double* cj = W + (size_t)j * ld;
      double acc = cj[k];
      for (int r = k + 1; r < n; r++) acc += s.v[r] * cj[r];
      acc *= tau;
      cj[k] -= acc;
      for (int r = k + 1; r < n; r++) cj[r] -= acc * s.v[r];
    }
    __syncwarp();
    if (lane == 0) ck[k] = beta;
    __syncwarp();
  }
  // R d = Q^T Be (column n), column-oriented back substitution: after d_i is known every row above is updated in parallel
  double* y = W + (size_t)n * ld;
  for (int i = n - 1; i >= 0; i--) {
    const double di = y[i] / W[(size_t)i * ld + i];
    __syncwarp();
    if (lane == 0) s.d[i] = di;
    for (int r = lane; r < i; r += 32) y[r] -= W[(size_t)i * ld + r] * di;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32) k_lemke(int n_problems, int n_max, const int* __restrict__ n_rows, const double* __restrict__ M_all,
                                              const double* __restrict__ q_all, double* __restrict__ B_all, double* __restrict__ z_all,
                                              int* __restrict__ err_all, int* __restrict__ pivots_all) {
  extern __shared__ __align__(16) unsigned char lemke_raw[];
  const int p = blockIdx.x, lane = threadIdx.x;
  if (p >= n_problems) return;
  const int n = n_rows ? n_rows[p] : n_max;
  const double* M = M_all + (size_t)p * n_max * n_max;  // row-major, leading dimension n_max
  const double* q = q_all + (size_t)p * n_max;
  double* B = B_all + (size_t)p * n_max * n_max;        // column-major n x n, leading dimension n
  double* z = z_all + (size_t)p * n_max;
  LemkeSmem s(lemke_raw, n_max);
  for (int i = lane; i < n_max; i += 32) z[i] = 0.0;
  int err = 0, pivots = 0;
  double qmin = INFINITY;
  for (int i = 0; i < n; i++) qmin = fmin(qmin, q[i]);
  if (n == 0 || qmin >= 0.0) {  // trivial solution (Lemke.cpp:70-75)
    if (lane == 0) { err_all[p] = 0; if (pivots_all) pivots_all[p] = 0; }
    return;
  }
  const int t = 2 * n;
  // x = q, B = -I, bas = slack variables; the most negative x leaves for the artificial variable (Lemke.cpp:90-171)
  int lvindex = 0;
  double tval = -q[0];
  for (int i = 1; i < n; i++) if (-q[i] > tval) { tval = -q[i]; lvindex = i; }  // maxCoeff: first maximum
  for (int i = lane; i < n; i += 32) {
    const double xi = q[i], U = xi < 0.0 ? 1.0 : 0.0;
    s.bas[i] = n + i;
    s.x[i] = xi + tval * U;
    s.Be[i] = U;  // -(B U) with B = -I
    for (int r = 0; r < n; r++) B[(size_t)i * n + r] = (r == i) ? -1.0 : 0.0;
  }
  __syncwarp();
  int leaving = n + lvindex, entering = t;
  if (lane == 0) { s.bas[lvindex] = t; s.x[lvindex] = tval; }
  for (int r = lane; r < n; r += 32) B[(size_t)lvindex * n + r] = s.Be[r];
  __syncwarp();
  int iter = 0;
  for (; iter < kMaxIter; iter++) {
    if (leaving == t) break;
    if (leaving < n) {
      entering = n + leaving;
      for (int r = lane; r < n; r += 32) s.Be[r] = (r == leaving) ? -1.0 : 0.0;
    } else {
      entering = leaving - n;
      for (int r = lane; r < n; r += 32) s.Be[r] = M[(size_t)r * n_max + entering];
    }
    __syncwarp();
    // working copy [B | Be]
    for (int c = 0; c < n; c++)
      for (int r = lane; r < n; r += 32) s.W[(size_t)c * s.ld + r] = B[(size_t)c * n + r];
    for (int r = lane; r < n; r += 32) s.W[(size_t)n * s.ld + r] = s.Be[r];
    __syncwarp();
    qrSolve(s, n, lane);
    // ratio test (Lemke.cpp:195-262), evaluated by every lane on the same shared data
    double theta = INFINITY;
    int cand = 0;
    for (int i = 0; i < n; i++)
      if (s.d[i] > kPivTol) { cand++; theta = fmin(theta, (s.x[i] + kZerTol) / s.d[i]); }
    if (cand == 0) { err = 2; break; }  // ray termination
    int lv = -1, first = -1;
    double best = 0.0;
    bool art = false;
    for (int i = 0; i < n; i++) {
      if (!(s.d[i] > kPivTol) || !(s.x[i] / s.d[i] <= theta)) continue;
      if (s.bas[i] == t) { lv = i; art = true; }  // always use the artificial variable if possible
      if (art) continue;
      if (first < 0) { first = i; best = s.d[i]; lv = i; }
      else if (s.d[i] - best > kPivTol) { best = s.d[i]; lv = i; }
    }
    if (lv < 0) { err = 4; break; }
    lvindex = lv;
    leaving = s.bas[lvindex];
    const double ratio = s.x[lvindex] / s.d[lvindex];
    __syncwarp();
    for (int i = lane; i < n; i += 32) s.x[i] = (i == lvindex) ? ratio : s.x[i] - ratio * s.d[i];
    for (int r = lane; r < n; r += 32) B[(size_t)lvindex * n + r] = s.Be[r];
    if (lane == 0) s.bas[lvindex] = entering;
    pivots++;
    __syncwarp();
  }
  if (iter >= kMaxIter && leaving != t) err = 1;
  if (err == 0) {
    __syncwarp();
    for (int i = lane; i < n; i += 32)
      if (s.bas[i] < n) z[s.bas[i]] = s.x[i];
    __syncwarp();
    __threadfence_block();
    // validate (Lemke.cpp:340-357, threshold 1e-4)
    bool ok = true;
    for (int i = 0; i < n && ok; i++) {
      double w = q[i];
      for (int j = 0; j < n; j++) w += M[(size_t)i * n_max + j] * z[j];
      if (w < -1e-4 || z[i] < -1e-4 || fabs(w * z[i]) > 1e-4) ok = false;
    }
    if (!ok) err = 3;
  } else {
    for (int i = lane; i < n_max; i += 32) z[i] = 0.0;  // solve failed: zero vector (Lemke.cpp:318-321)
  }
  if (lane == 0) { err_all[p] = err; if (pivots_all) pivots_all[p] = pivots; }
}

}  // namespace

extern "C" int ddb_lemke_solve(int32_t device, void* stream, int32_t n_problems, int32_t n_max, const int32_t* n_rows, const double* M,
                               const double* q, double* z, int32_t* err, int32_t* pivots) {
  if (n_problems < 1 || n_max < 1 || n_max > 128 || !M || !q || !z || !err) return DDB_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DDB_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return DDB_ERR_CUDA;
  cudaStream_t s = (cudaStream_t)stream;
  double* B = nullptr;
  if (cudaMallocAsync(&B, (size_t)n_problems * n_max * n_max * sizeof(double), s) != cudaSuccess) return DDB_ERR_CUDA;
  const size_t smem = LemkeSmem::bytes(n_max);
  if (smem > 48 * 1024 && cudaFuncSetAttribute(k_lemke, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaFreeAsync(B, s);
    return DDB_ERR_CUDA;
  }
  k_lemke<<<n_problems, 32, smem, s>>>(n_problems, n_max, n_rows, M, q, B, z, err, pivots);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(B, s);
  return e == cudaSuccess ? DDB_OK : DDB_ERR_CUDA;
}
