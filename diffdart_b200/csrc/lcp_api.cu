// lcp_api.cu — batched boxed-LCP entry points of the C ABI (compiled with -fmad=false, see lcp.cuh).
#include <cstdio>
#include <cstdlib>
#include "../../include/ddb.h"
#include "contact.cuh"
#include "contact_bwd.cuh"
#include "contact_api.h"

// resident CTAs per SM the contact kernels are compiled for (register cap 255 / 168 / 128 at 1-2 / 3 / 4 CTAs of 128 threads)
#ifndef DDB_KC_MINB
#define DDB_KC_MINB 1
#endif
#ifndef DDB_KCB_BLOCK
#define DDB_KCB_BLOCK 512  // threads per CTA of k_contact_backward (one world per warp; see k_contact about the CTA size): measured 64 -> 256
#endif                     // walker 25.5 -> 16.9 ms, half-cheetah 6.2 -> 3.9 ms, box stacking 50.7 -> 43.9 ms per backward pass; 512 with the
                           // meeting points: 8.5 / 3.3 / 19.7 ms
#ifndef DDB_KCB_CONVOY
#define DDB_KCB_CONVOY 1  // likewise in k_contact_backward
#endif
#ifndef DDB_KC_CONVOY
#define DDB_KC_CONVOY 1  // 0: the warps of a wide k_contact CTA do not wait for each other at the ends of the stage's sections
#endif
#ifndef DDB_KC_BLOCK
#define DDB_KC_BLOCK 128  // threads per CTA of k_contact (one world per warp) for worlds of more than kContactWideRows rows; DDB_KC_WIDE_BLOCK below
#endif
#ifndef DDB_KCB_MINB
#define DDB_KCB_MINB 1
#endif
using namespace ddb;
#ifndef DDB_KC_WIDE_ROWS
#define DDB_KC_WIDE_ROWS 64
#endif
#ifndef DDB_KC_WIDE_BLOCK
#define DDB_KC_WIDE_BLOCK 512  // threads per CTA of k_contact for worlds of up to DDB_KC_WIDE_ROWS rows: 16 warps at 128 registers (1.5 KB of
#endif                         // spills) beat 8 at 252 - walker 22.1 -> 17.0 ms, half-cheetah 4.0 -> 3.1 ms; 24 and 32 warps spill too much
#ifndef DDB_KC_WIDE_BODIES
#define DDB_KC_WIDE_BODIES 64
#endif
#ifndef DDB_KC_CONVOY_MIN_BLOCK
#define DDB_KC_CONVOY_MIN_BLOCK 256
#endif
constexpr int kContactWideRows = DDB_KC_WIDE_ROWS;

namespace {

// one world per thread; inputs world-major [N][n_max]([n_max]); scratch lane-interleaved per warp
__global__ void __launch_bounds__(64)
k_lcp_solve(int N, int nmax, const int* __restrict__ n_rows, const double* __restrict__ A, const double* __restrict__ b,
            const double* __restrict__ lo, const double* __restrict__ hi, const int* __restrict__ findex, int solver, int early,
            PgsOptions pgs, double* __restrict__ x, int* __restrict__ ok, double* __restrict__ sd, int* __restrict__ si) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= N) return;
  const int n = n_rows ? n_rows[w] : nmax;
  if (n <= 0) { if (ok) ok[w] = 1; return; }
  LcpView v;
  const size_t warp = w >> 5, lane = w & 31;
  v.init(sd + warp * lcpScratchDoubles(nmax) * 32 + lane, si + warp * lcpScratchInts(nmax) * 32 + lane, n, nmax);
  const double* Aw = A + (size_t)w * nmax * nmax;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) v.Araw(i, j) = Aw[(size_t)i * nmax + j];
    const size_t ix = (size_t)w * nmax + i;
    v.b(i) = b[ix]; v.lo(i) = lo[ix]; v.hi(i) = hi[ix]; v.findex(i) = findex[ix];
    v.x(i) = x[ix];  // PGS warm start (Dantzig zeroes it, lcp.cpp:411)
  }
  bool good = (solver == DDB_LCP_DANTZIG) ? dantzigSolve(v, 0, early != 0) : pgsSolve(v, pgs);
  for (int i = 0; i < n; ++i) x[(size_t)w * nmax + i] = v.x(i);
  if (ok) ok[w] = good ? 1 : 0;
}

__global__ void __launch_bounds__(128)
k_lcp_valid(int N, int nmax, const int* __restrict__ n_rows, const double* __restrict__ A, const double* __restrict__ x,
            const double* __restrict__ b, const double* __restrict__ lo, const double* __restrict__ hi, const int* __restrict__ findex,
            int ignoreFriction, int* __restrict__ valid) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= N) return;
  const int n = n_rows ? n_rows[w] : nmax;
  const double* Aw = A + (size_t)w * nmax * nmax;
  const size_t o = (size_t)w * nmax;
  valid[w] = lcpValid(n, [&](int i, int j) { return Aw[(size_t)i * nmax + j]; }, [&](int i) { return x[o + i]; },
                      [&](int i) { return b[o + i]; }, [&](int i) { return hi[o + i]; }, [&](int i) { return lo[o + i]; },
                      [&](int i) { return findex[o + i]; }, ignoreFriction != 0) ? 1 : 0;
}

// one shape pair per thread; out: [N][8][26] doubles (8 vectors, depth, type), counts [N]
__global__ void __launch_bounds__(128)
k_collide_pairs(int N, const int* __restrict__ shape1, const double* __restrict__ size1, const double* __restrict__ T1,
                const int* __restrict__ shape2, const double* __restrict__ size2, const double* __restrict__ T2, double* __restrict__ out,
                int* __restrict__ counts) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= N) return;
  Pose a, b;
  for (int i = 0; i < 9; i++) { a.R[i] = T1[(size_t)w * 12 + i]; b.R[i] = T2[(size_t)w * 12 + i]; }
  for (int i = 0; i < 3; i++) { a.p[i] = T1[(size_t)w * 12 + 9 + i]; b.p[i] = T2[(size_t)w * 12 + 9 + i]; }
  ContactPoint c[8];
  const int n = collidePair(shape1[w], size1 + (size_t)w * 3, a, shape2[w], size2 + (size_t)w * 3, b, c);
  counts[w] = n;
  for (int k = 0; k < n; k++) {
    double* o = out + ((size_t)w * 8 + k) * 26;
    const double* vs[8] = {c[k].point, c[k].normal, c[k].edgeAClosest, c[k].edgeAFixed, c[k].edgeADir, c[k].edgeBClosest, c[k].edgeBFixed, c[k].edgeBDir};
    for (int v = 0; v < 8; v++) for (int i = 0; i < 3; i++) o[3 * v + i] = vs[v][i];
    o[24] = c[k].depth;
    o[25] = (double)c[k].type;
  }
}

constexpr int kPgsRequests = 4;  // PGS solves a world can hand over per step (2 tiers x 2 constrained groups); more run inline

// The 32 sequential steps of one row block (phase 2 of k_pgs_warp). Every lane projects its own candidate each step and the
// owner's is broadcast, so the critical path of a step is one shuffle, one DFMA and the min / max of the projection; the
// matrix element of the next step is fetched (and scaled) ahead of the shuffle. Rows that are not iterated (zero
// diagonal) carry lo = hi = their value, so the projection returns it unchanged.
template <bool FIRST>
__device__ __forceinline__ void pgsBlockSteps(const double* __restrict__ S, const double* __restrict__ sc, double* __restrict__ xsb, int blk0, int r,
                                              int cnt, int lane, int fiLocal, double bh, double bl, double xn, double diag, double f_r, double acc) {
  // S[r][blk0 + jj] is only used by the lanes below the diagonal (jj < lane), where it sits in this lane's packed row: one base
  // pointer per block instead of index arithmetic per step. A lane's residual matters until its own step (its x is taken there):
  // from then on it keeps subtracting whatever follows its row in shared memory (in bounds: the triangle is followed by the vectors)
  // instead of paying a compare and two selects per step to stand still, and the element of the next step is fetched without
  // clamping the index at the block's end for the same reason. ncu on a heavy launch: 67 % of the solver's instructions are these
  // steps, 44 instructions per step before this change (profiles/README.md, round 2).
  // The bounds of a row are (bl, bh) * xn: xn is the live normal impulse of a friction row (bh = mu, bl = -mu; taken from the
  // broadcast in the step of its normal row when that row is in this block) and 1 for every other row (bl = lo, bh = hi), so a
  // step pays one compare and select for xn and two independent multiplies instead of selects on both bounds behind a multiply and
  // a negation. The owner of step jj writes its x straight to shared memory (one predicated store instead of two selects per step).
  const double* __restrict__ rowS = S + (r * (r + 1)) / 2 + blk0;
  const double* __restrict__ scb = sc + blk0;
  auto elem = [&](int jj) { return rowS[jj] * scb[jj] * f_r; };
  double a_next = elem(0);
#pragma unroll 8
  for (int jj = 0; jj < cnt; jj++) {
    const double a = a_next;
    a_next = elem(jj + 1);
    const double cand = FIRST ? acc / diag : acc;
    const double hi_e = bh * xn, lo_e = bl * xn;
    // projection onto [lo_e, hi_e] (lo_e <= hi_e) with two independent comparisons of the candidate, as the reference writes it
    // (PgsBoxedLcpSolver.cpp:139-150), instead of fmin / fmax, whose NaN handling costs 6-7 instructions each on the critical path
    double nx = cand < lo_e ? lo_e : cand;
    nx = cand > hi_e ? hi_e : nx;  // (the upper bound is tested first there)
    if (lane == jj) xsb[jj] = nx;
    const double xj = __shfl_sync(0xffffffffu, nx, jj);
    acc = __fma_rn(-a, xj, acc);  // (this translation unit is built with -fmad=false for the Dantzig solver)
    xn = fiLocal == jj ? xj : xn;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// PgsBoxedLcpSolver::solve (PgsBoxedLcpSolver.cpp:79-262), one warp per world, the matrix in shared memory.
// Gauss-Seidel is sequential in the rows, so the warp works on blocks of 32 consecutive rows with lane r owning row r:
//   phase 1 (parallel)    acc_r = b_r - sum_{j before the block} A_rj x_j(new) - sum_{j > r} A_rj x_j(old)
//   phase 2 (32 steps)    lane j projects acc_j into x_j and broadcasts it (one shuffle); lanes r > j subtract A_rj x_j
// The sweep order, the projections (friction bounds from the live normal impulse) and the stopping rules are the
// reference's; the summation order inside a row is not (the reference accumulates left to right), so the result agrees
// with a scalar run to rounding, not bit for bit. Rows whose diagonal is below epsilonForDivision are left at 0.
// The reduced matrix is A[keep][keep] * diag(scale) with A symmetric (LCPUtils::reduce merges duplicate columns by
// doubling one of them): only the packed lower triangle of the symmetric part is kept, which halves the shared-memory
// footprint and doubles the worlds resident on an SM. Triangular offsets of 16 consecutive rows are distinct modulo
// 16, so the lane-owns-row reads are free of bank conflicts in both halves of the triangle.
// ---------------------------------------------------------------------------------------------------------------
// Problems of more than one row block run as a PAIR of warps: while warp 0 walks the 32 dependent steps of block b, warp 1 sums
// phase 1 of block b + 1 over every column except those of block b (whose new values do not exist yet) and leaves the partial sum
// in shared memory; warp 0 then only adds the 32 columns of block b before it starts on block b + 1. A world is one dependent
// chain, and the SM holds as many worlds as their matrices fit: the helper shortens the chain (phase 1 was 40 % of a sweep's
// latency at 84 rows) with issue slots that were idle. One CTA barrier per block; the helper runs one block ahead, across the end
// of a sweep as well (its sum is dropped when the sweep turns out to be the last).
template <bool PAIR>
__global__ void __launch_bounds__(PAIR ? 64 : 32)
k_pgs_warp(int N, int kmax, int nLow, int CAP, int* __restrict__ ctrl, double* __restrict__ lcp_d, int* __restrict__ lcp_i,
           double* __restrict__ ans_x, int* __restrict__ ans_ok, PgsOptions opt) {
  const int w = blockIdx.x, lane = threadIdx.x & 31, warp = PAIR ? threadIdx.x >> 5 : 0;
  const int nthreads = PAIR ? 64 : 32;
  const int n = ctrl[8 * w + 2];
  if (ctrl[8 * w] || n <= nLow || n > CAP) return;
  extern __shared__ double sm[];
  double* S = sm;                          // packed lower triangle, row i at i (i + 1) / 2
  double* bs = S + (CAP * (CAP + 1)) / 2;  // [CAP] each:
  double* los = bs + CAP;
  double* his = los + CAP;
  double* xs = his + CAP;   // x
  double* ys = xs + CAP;    // scale * x
  double* sc = ys + CAP;    // column scale of the reduced problem
  double* fs = sc + CAP;    // 1 / diagonal of the scaled matrix; 0 for rows that are not iterated (diagonal below epsilonForDivision)
  int* fis = reinterpret_cast<int*>(fs + CAP);
  double* part = fs + CAP + (CAP + 1) / 2;             // PAIR: [2][32] partial sums of the helper
  int* stop = reinterpret_cast<int*>(part + 64);      // PAIR: the sweep just finished was the last
  auto ctaSync = [&]() { if (PAIR) __syncthreads(); else __syncwarp(); };
  LcpView v;
  v.init(lcp_d + (size_t)w * lcpScratchDoubles(kmax), lcp_i + (size_t)w * lcpScratchInts(kmax), n, kmax, 1);
  for (int i = threadIdx.x; i < n; i += nthreads) {
    bs[i] = v.b(i); los[i] = v.lo(i); his[i] = v.hi(i); xs[i] = v.x(i); fis[i] = v.findex(i);
    sc[i] = v.tmp(i);  // buildReduced leaves the column scales here (powers of two)
  }
  ctaSync();
  for (int r = warp; r < n; r += PAIR ? 2 : 1)
    for (int j = lane; j <= r; j += 32) S[(r * (r + 1)) / 2 + j] = v.Araw(r, j) / sc[j];
  for (int i = threadIdx.x; i < n; i += nthreads) ys[i] = sc[i] * xs[i];
  ctaSync();
  // (the division of the row normalisation once, not in every sweep)
  for (int i = threadIdx.x; i < n; i += nthreads) {
    const double d = S[(i * (i + 1)) / 2 + i] * sc[i];
    fs[i] = d < opt.epsilonForDivision ? 0.0 : 1.0 / d;
  }
  ctaSync();
  const int nblk = (n + 31) >> 5;
  // the three column ranges of phase 1 for row r (tr = its offset in the triangle), four partial sums
  struct Sum4 { double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0; __device__ double total() const { return (a0 + a1) + (a2 + a3); } };
  auto rowPart = [&](Sum4& a, int tr, int j0, int j1) {  // columns before the block: this row of the triangle
    int j = j0;
    for (; j + 3 < j1; j += 4) {
      a.a0 = __fma_rn(S[tr + j], ys[j], a.a0); a.a1 = __fma_rn(S[tr + j + 1], ys[j + 1], a.a1);
      a.a2 = __fma_rn(S[tr + j + 2], ys[j + 2], a.a2); a.a3 = __fma_rn(S[tr + j + 3], ys[j + 3], a.a3);
    }
    for (; j < j1; j++) a.a0 += S[tr + j] * ys[j];
  };
  // the rows of this block behind row r (old values): rows at or before r are masked out through their x (the matrix element
  // read instead is some other finite entry of the triangle)
  auto inBlock = [&](Sum4& a, int r, int blk0, int blkEnd) {
    int j = blk0;
    for (; j + 3 < blkEnd; j += 4) {
      const double y0 = j > r ? ys[j] : 0.0, y1 = j + 1 > r ? ys[j + 1] : 0.0, y2 = j + 2 > r ? ys[j + 2] : 0.0, y3 = j + 3 > r ? ys[j + 3] : 0.0;
      a.a0 = __fma_rn(S[(j * (j + 1)) / 2 + r], y0, a.a0); a.a1 = __fma_rn(S[((j + 1) * (j + 2)) / 2 + r], y1, a.a1);
      a.a2 = __fma_rn(S[((j + 2) * (j + 3)) / 2 + r], y2, a.a2); a.a3 = __fma_rn(S[((j + 3) * (j + 4)) / 2 + r], y3, a.a3);
    }
    for (; j < blkEnd; j++) a.a1 = __fma_rn(S[(j * (j + 1)) / 2 + r], j > r ? ys[j] : 0.0, a.a1);
  };
  auto colPart = [&](Sum4& a, int r, int j0, int j1) {  // columns behind the block: column r of the triangle
    int j = j0;
    for (; j + 3 < j1; j += 4) {
      a.a0 = __fma_rn(S[(j * (j + 1)) / 2 + r], ys[j], a.a0); a.a1 = __fma_rn(S[((j + 1) * (j + 2)) / 2 + r], ys[j + 1], a.a1);
      a.a2 = __fma_rn(S[((j + 2) * (j + 3)) / 2 + r], ys[j + 2], a.a2); a.a3 = __fma_rn(S[((j + 3) * (j + 4)) / 2 + r], ys[j + 3], a.a3);
    }
    for (; j < j1; j++) a.a0 += S[(j * (j + 1)) / 2 + r] * ys[j];
  };
  auto rowOf = [&](int blk) { const int blk0 = blk << 5, blkEnd = min(blk0 + 32, n); return blk0 + lane < n ? blk0 + lane : blkEnd - 1; };
  // helper (PAIR): phase 1 of block blk without the columns of the block before it (in sweep order: the last block for block 0)
  auto helperSum = [&](int blk, double* out) {
    const int blk0 = blk << 5, blkEnd = min(blk0 + 32, n), r = rowOf(blk), tr = (r * (r + 1)) / 2;
    Sum4 a;
    if (blk > 0) { rowPart(a, tr, 0, blk0 - 32); inBlock(a, r, blk0, blkEnd); colPart(a, r, blkEnd, n); }
    else { inBlock(a, r, blk0, blkEnd); colPart(a, r, blkEnd, (nblk - 1) << 5); }
    out[lane] = a.total();
  };
  bool converged = false;
  int sweeps = 0, t = 0;
  if (PAIR) {
    if (warp == 1) helperSum(0, part);
    if (threadIdx.x == 0) *stop = 0;
    __syncthreads();
  }
  // sweep 0 works on the raw matrix (divides by the diagonal, absolute stopping rule); later sweeps on the normalised rows
  for (int iter = 0; iter < opt.maxIteration; iter++) {
    bool mine = true;  // possibleToTerminate restricted to this lane's rows
    sweeps++;
    for (int blk = 0; blk < nblk; blk++, t++) {
      if (warp == 1) {
        helperSum(blk + 1 < nblk ? blk + 1 : 0, part + 32 * ((t + 1) & 1));
        __syncthreads();
        continue;
      }
      const int blk0 = blk << 5, blkEnd = min(blk0 + 32, n);
      const bool valid = blk0 + lane < n;
      const int r = valid ? blk0 + lane : blkEnd - 1;  // idle lanes shadow the last row (their results are dropped)
      const int tr = (r * (r + 1)) / 2;
      const double inv = fs[r];
      const bool active = inv != 0.0;
      const double old_x = xs[r], hi_r = his[r];
      const int fi = fis[r];
      // row normalisation (PgsBoxedLcpSolver.cpp:171-178); rows that are not iterated get 0 so that their residual stays finite
      // (the projection below selects with comparisons: a NaN would pass through it)
      const double f_r = !active ? 0.0 : (iter == 0 ? 1.0 : inv);
      // phase 1 (PAIR: what the helper left out, on top of its sum)
      Sum4 a;
      double sum;
      if (PAIR) {
        if (blk > 0) rowPart(a, tr, blk0 - 32, blk0);
        else colPart(a, r, (nblk - 1) << 5, n);
        sum = part[32 * (t & 1) + lane] + a.total();
      } else {
        rowPart(a, tr, 0, blk0); inBlock(a, r, blk0, blkEnd); colPart(a, r, blkEnd, n);
        sum = a.total();
      }
      double acc = (bs[r] - sum) * f_r;
      // phase 2: bounds (bl, bh) * xn (pgsBlockSteps)
      double bh = hi_r, bl = los[r], xn = 1.0;
      if (fi >= 0) { bl = -hi_r; xn = xs[fi]; }  // a normal row of an earlier block is final; one of this block is replaced in its step
      if (!active) { bh = bl = iter == 0 ? 0.0 : old_x; xn = 1.0; }
      __syncwarp();  // (every lane has read what it needs of xs before the owners of this block's rows overwrite them)
      if (iter == 0) pgsBlockSteps<true>(S, sc, xs + blk0, blk0, r, blkEnd - blk0, lane, active ? fi - blk0 : -1, bh, bl, xn, active ? S[tr + r] * sc[r] : 1.0, f_r, acc);
      else pgsBlockSteps<false>(S, sc, xs + blk0, blk0, r, blkEnd - blk0, lane, active ? fi - blk0 : -1, bh, bl, xn, 1.0, f_r, acc);
      __syncwarp();
      if (valid) {
        const double x_new = xs[r];
        ys[r] = sc[r] * x_new;
        if (active) {
          if (iter == 0) { if (fabs(x_new - old_x) > opt.deltaXThreshold) mine = false; }
          // |dx / x| > tolerance (PgsBoxedLcpSolver.cpp:238-244) without the division: |x| > epsilonForDivision > 0 here
          else if (fabs(x_new) > opt.epsilonForDivision) { if (fabs(x_new - old_x) > opt.relativeDeltaXTolerance * fabs(x_new)) mine = false; }
        }
      }
      if (PAIR) {
        if (blk == nblk - 1) { const bool all = __all_sync(0xffffffffu, mine); if (lane == 0) *stop = all ? 1 : 0; }
        __syncthreads();
      } else {
        __syncwarp();
      }
    }
    converged = PAIR ? *stop != 0 : __all_sync(0xffffffffu, mine);
    if (converged) break;
  }
  if (warp == 1) return;
  const int slot = ctrl[8 * w + 1];
  double* ax = ans_x + ((size_t)w * kPgsRequests + slot) * kmax;
  for (int i = lane; i < n; i += 32) ax[i] = xs[i];
  if (lane == 0) {
    ans_ok[(size_t)w * kPgsRequests + slot] = converged ? 1 : 0;
    ctrl[8 * w + 1] = slot + 1;
    ctrl[8 * w + 2] = 0;
    ctrl[8 * w + 3] += sweeps;  // statistics: Gauss-Seidel sweeps spent on this world in this step
    if (n > ctrl[8 * w + 5]) ctrl[8 * w + 5] = n;  // ... and the largest reduced problem it handed to the solver
  }
}

// Problems of at most 32 rows (one row block: Atlas' feet, the walkers): lane r keeps row r of the matrix in REGISTERS, normalised
// once after sweep 0 as the reference does (A_ij * (1 / A_ii), PgsBoxedLcpSolver.cpp:171-178), and one residual that serves two
// sweeps: up to its own step a lane subtracts the new x_j (j < r) of this sweep, from its own step on it rebuilds b_r - sum_{j > r}
// A_rj x_j for the next one. A step is {project, broadcast by one shuffle, one DFMA}: no phase 1, no matrix in shared memory, no
// element loads on the way. Sweep order, projections and stopping rules are the reference's; the summation order inside a row is not.
template <bool FIRST>
__device__ __forceinline__ void pgsSmallSweep(int n, int lane, const double (&a)[32], double start, int fi, double bh, double bl, double diag,
                                              double& acc, double& x_r) {
  double xn = 1.0;  // live normal impulse of a friction row (its normal row precedes it: replaced before this lane's own step)
#pragma unroll
  for (int jj = 0; jj < 32; jj++) {
    if (jj >= n) break;
    const double cand = FIRST ? acc / diag : acc;
    const double hi_e = bh * xn, lo_e = bl * xn;
    double nx = cand < lo_e ? lo_e : cand;
    nx = cand > hi_e ? hi_e : nx;  // (PgsBoxedLcpSolver.cpp:139-150: the upper bound is tested first)
    const double xj = __shfl_sync(0xffffffffu, nx, jj);
    const bool own = lane == jj;
    x_r = own ? xj : x_r;
    acc = own ? start : __fma_rn(-a[jj], xj, acc);
    xn = fi == jj ? xj : xn;
  }
}

__global__ void __launch_bounds__(32)
k_pgs_small(int N, int kmax, int* __restrict__ ctrl, double* __restrict__ lcp_d, int* __restrict__ lcp_i, double* __restrict__ ans_x,
            int* __restrict__ ans_ok, PgsOptions opt) {
  const int w = blockIdx.x, lane = threadIdx.x;
  const int n = ctrl[8 * w + 2];
  if (ctrl[8 * w] || n <= 0 || n > 32) return;
  LcpView v;
  v.init(lcp_d + (size_t)w * lcpScratchDoubles(kmax), lcp_i + (size_t)w * lcpScratchInts(kmax), n, kmax, 1);
  const bool valid = lane < n;
  const int r = valid ? lane : n - 1;  // idle lanes shadow the last row (their results are dropped; they never own a step)
  // row r of the reduced matrix A[keep][keep] * diag(scale), its upper part from the symmetric lower one as k_pgs_warp reads it
  const double sc_r = v.tmp(r);  // buildReduced leaves the column scales here (powers of two)
  double a[32];
#pragma unroll
  for (int j = 0; j < 32; j++) a[j] = j >= n ? 0.0 : (j <= r ? v.Araw(r, j) : v.Araw(j, r) / sc_r * v.tmp(j));
  double diag = 0.0;
#pragma unroll
  for (int j = 0; j < 32; j++) diag = j == r ? a[j] : diag;
  const bool active = !(diag < opt.epsilonForDivision);
  const double inv = active ? 1.0 / diag : 0.0;
  const double b_r = v.b(r), hi_r = v.hi(r);
  const int fi = active ? v.findex(r) : -1;
  double x_r = v.x(r);
  double bh = hi_r, bl = v.lo(r);
  if (fi >= 0) bl = -hi_r;
  if (!active) {  // rows that are not iterated: 0 in sweep 0, unchanged afterwards; a zero row keeps their residual finite
    bh = bl = 0.0;
#pragma unroll
    for (int j = 0; j < 32; j++) a[j] = 0.0;
  }
  // residual of sweep 0: b_r - sum_{j > r} A_rj x_j over the initial x
  double acc = active ? b_r : 0.0;
#pragma unroll
  for (int jj = 0; jj < 32; jj++) {
    if (jj >= n) break;
    const double xj = __shfl_sync(0xffffffffu, x_r, jj);
    acc = jj > lane ? __fma_rn(-a[jj], xj, acc) : acc;
  }
  bool converged = false;
  int sweeps = 1;
  {
    const double old_x = x_r;
    pgsSmallSweep<true>(n, lane, a, active ? b_r : 0.0, fi, bh, bl, active ? diag : 1.0, acc, x_r);
    const bool mine = !(valid && active && fabs(x_r - old_x) > opt.deltaXThreshold);
    converged = __all_sync(0xffffffffu, mine);
  }
  if (!converged) {
    // row normalisation; the residual under way (raw so far) with it
    acc *= inv;
#pragma unroll
    for (int j = 0; j < 32; j++) a[j] *= inv;
    const double start = b_r * inv;
    if (!active) bh = bl = x_r;  // (0 after sweep 0)
    for (int iter = 1; iter < opt.maxIteration; iter++) {
      sweeps++;
      const double old_x = x_r;
      pgsSmallSweep<false>(n, lane, a, start, fi, bh, bl, 1.0, acc, x_r);
      // |dx / x| > tolerance (PgsBoxedLcpSolver.cpp:238-244) without the division: |x| > epsilonForDivision > 0 here
      const bool mine = !(valid && active && fabs(x_r) > opt.epsilonForDivision && fabs(x_r - old_x) > opt.relativeDeltaXTolerance * fabs(x_r));
      converged = __all_sync(0xffffffffu, mine);
      if (converged) break;
    }
  }
  const int slot = ctrl[8 * w + 1];
  if (valid) ans_x[((size_t)w * kPgsRequests + slot) * kmax + r] = x_r;
  if (lane == 0) {
    ans_ok[(size_t)w * kPgsRequests + slot] = converged ? 1 : 0;
    ctrl[8 * w + 1] = slot + 1;
    ctrl[8 * w + 2] = 0;
    ctrl[8 * w + 3] += sweeps;
    if (n > ctrl[8 * w + 5]) ctrl[8 * w + 5] = n;
  }
}

// problems too large for shared memory: the scalar solver, one world per thread
__global__ void __launch_bounds__(64)
k_pgs_thread(int N, int kmax, int nLow, int* __restrict__ ctrl, double* __restrict__ lcp_d, int* __restrict__ lcp_i, double* __restrict__ ans_x,
             int* __restrict__ ans_ok, PgsOptions opt) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= N) return;
  const int n = ctrl[8 * w + 2];
  if (ctrl[8 * w] || n <= nLow) return;
  LcpView v;
  v.init(lcp_d + (size_t)w * lcpScratchDoubles(kmax), lcp_i + (size_t)w * lcpScratchInts(kmax), n, kmax, 1);
  const bool ok = pgsSolve(v, opt);
  const int slot = ctrl[8 * w + 1];
  for (int i = 0; i < n; i++) ans_x[((size_t)w * kPgsRequests + slot) * kmax + i] = v.x(i);
  ans_ok[(size_t)w * kPgsRequests + slot] = ok ? 1 : 0;
  ctrl[8 * w + 1] = slot + 1;
  ctrl[8 * w + 2] = 0;
  if (n > ctrl[8 * w + 5]) ctrl[8 * w + 5] = n;
}

// The contact stage of World::step, one world per WARP: the 32 lanes run the same scalar code on the same data (loads are
// broadcasts, stores coincide), so a batch of N worlds keeps N warps in flight instead of N / 32. The stage is a chain of
// dependent memory accesses (pivoting solvers, Householder sweeps); what bounds it is how many of those chains the SMs
// can overlap, not arithmetic, and the redundant lanes cost nothing that matters (FP64 issue stays far below the pipe).
// CTA size: the stage is 54 000 instructions of which a world touches 24 000 once or twice - more than half of a warp's stall
// cycles are instruction fetches (ncu: stall_no_instruction 8.7 of 16 cycles per issue, profiles/README.md). Warps of one CTA
// start together and run the same code at about the same time, so they share the fetched lines: 8 warps per CTA instead of 4
// was 10-19 % faster for the walkers, meeting points at the ends of the stage's sections (ContactWork::mark) another 10 %, and with
// the fetches shared 16 warps at 128 registers beat 8 at 252 (DDB_KC_WIDE_BLOCK). Box stacking, whose worlds differ by 10x in
// solver time, loses to the CTA's wait for its slowest warp and keeps 4 warps at 252 registers (contactLaunch chooses by the row
// capacity). Measurements: DESIGN.md 4b.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, DDB_KC_MINB)
k_contact(int N, int layout, ContactLayout L, double* __restrict__ snap, double* __restrict__ qd, double* __restrict__ lcp_cache,
          int* __restrict__ n_cache, int* __restrict__ counts, int* __restrict__ active_out, double* __restrict__ work_d,
          int* __restrict__ work_i, double* __restrict__ lcp_d, int* __restrict__ lcp_i, PgsOptions pgs, double* __restrict__ keep_d,
          int* __restrict__ keep_i, int* __restrict__ ctrl, const double* __restrict__ ans_x, const int* __restrict__ ans_ok, long long* __restrict__ prof) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  constexpr bool kConvoy = BLOCK >= DDB_KC_CONVOY_MIN_BLOCK && DDB_KC_CONVOY;
  __shared__ int reached[BLOCK / 32];  // kConvoy: ContactWork::mark
  const bool live = w < N && !(ctrl && ctrl[8 * w]);  // (else: past the batch, or committed by an earlier pass)
  if (kConvoy) {
    reached[threadIdx.x >> 5] = live ? 0 : 0x7fffffff;
    __syncthreads();
  }
  if (!live) return;
  Snap sn{reinterpret_cast<double2*>(snap), N, w};
  ContactWork cw{work_d + (size_t)w * L.doubles(), work_i + (size_t)w * L.ints(), L, prof};
  if (kConvoy) { cw.convoy = reached; cw.convoyWarps = BLOCK / 32; }
  const ContactSnapLayout KL{L.nd, L.cmax, L.kmax};
  double* ld = lcp_d + (size_t)w * lcpScratchDoubles(L.kmax);
  int* li = lcp_i + (size_t)w * lcpScratchInts(L.kmax);
  int nc = n_cache[w];
  PgsDefer df{ctrl ? ctrl[8 * w + 1] : 0, 0, L.kmax, 0, kPgsRequests, ctrl ? ctrl[8 * w + 4] : 0, ans_x + (size_t)w * kPgsRequests * L.kmax,
              ans_ok + (size_t)w * kPgsRequests};
  ContactStepOut o = contactStage(sn, cw, ld, li, qd, qd, layout, N, w, lcp_cache + (size_t)w * L.kmax, &nc, pgs,
                                  keep_d ? keep_d + (size_t)w * KL.doubles() : nullptr, keep_i ? keep_i + (size_t)w * KL.ints() : nullptr,
                                  ctrl ? &df : nullptr, 1);
  __syncwarp();  // every lane has read the control words before any lane rewrites them
  if (kConvoy) reached[threadIdx.x >> 5] = 0x7fffffff;
  if (o.pending) { ctrl[8 * w + 2] = df.pendingN; ctrl[8 * w + 4] = df.resumeKind; return; }  // the reduced problem waits in this world's LCP view
  if (ctrl) { ctrl[8 * w] = 1; ctrl[8 * w + 2] = 0; }
  n_cache[w] = nc;
  counts[4 * w] = o.n_contacts; counts[4 * w + 1] = o.n_rows; counts[4 * w + 2] = o.n_groups; counts[4 * w + 3] = o.tier_mask;
  if (active_out) {
    // ConstraintMapping per LCP row (ConstrainedGroupGradientMatrices.hpp:32-38); rows beyond n_rows are IRRELEVANT (-4)
    for (int k = 0; k < L.kmax; k++) active_out[(size_t)w * L.kmax + k] = (k < o.n_rows && c_ctopo.gradients) ? cw.activeAll()[k] : -4;
  }
}

// BackpropSnapshot::backprop through the contact solution of the step retained in this slot, one world per thread
template <int NB>
__global__ void __launch_bounds__(DDB_KCB_BLOCK, DDB_KCB_MINB)
k_contact_backward(int N, int layout, ContactSnapLayout KL, ContactBwdLayout BL, double* __restrict__ snap, const double* __restrict__ keep_d,
                   const int* __restrict__ keep_i, const double* __restrict__ gqn, const double* __restrict__ gqdn, double* __restrict__ gq,
                   double* __restrict__ gqd, double* __restrict__ gtau, double* __restrict__ work_d, int* __restrict__ work_i,
                   double* __restrict__ lcp_d, int* __restrict__ lcp_i, Limits lim, int* __restrict__ status) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one world per warp, lanes redundant (see k_contact)
  constexpr bool kConvoy = DDB_KCB_BLOCK >= 256 && DDB_KCB_CONVOY;
  __shared__ int reached[DDB_KCB_BLOCK / 32];
  if (kConvoy) {
    reached[threadIdx.x >> 5] = w < N ? 0 : 0x7fffffff;
    __syncthreads();
  }
  if (w >= N) return;
  BwdConvoy convoy;
  if (kConvoy) { convoy.reached = reached; convoy.warps = DDB_KCB_BLOCK / 32; }
  Snap sn{reinterpret_cast<double2*>(snap), N, w};
  ContactBwdWork bw{work_d + (size_t)w * BL.doubles(), work_i + (size_t)w * BL.ints(), BL};
  double* ld = lcp_d + (size_t)w * lcpScratchDoubles(BL.kmax);
  int* li = lcp_i + (size_t)w * lcpScratchInts(BL.kmax);
  const int st = contactBackwardWorld<NB>(sn, KL, keep_d + (size_t)w * KL.doubles(), keep_i + (size_t)w * KL.ints(), bw, ld, li, gqn, gqdn, gq,
                                          gqd, gtau, layout, N, w, 1, convoy);
  if (kConvoy) reached[threadIdx.x >> 5] = 0x7fffffff;
  if (status) status[w] = st;
  // clipLossGradientsToBounds belongs to backprop() only (BackpropSnapshot.cpp:374-428); the Jacobian getters pass no limits
  if (c_topo.has_limits && lim.pos_lo) {
    const int n = c_topo.nd;
    const double* clip = clipSection(snap, c_topo.nb, N);
    for (int d = 0; d < n; d++) {
      const size_t ix = vidx(layout, N, n, w, d);
      gq[ix] = clipGrad(gq[ix], clip[(size_t)d * N + w], lim.pos_lo[d], lim.pos_hi[d]);
      gqd[ix] = clipGrad(gqd[ix], clip[(size_t)(n + d) * N + w], lim.vel_lo[d], lim.vel_hi[d]);
      gtau[ix] = clipGrad(gtau[ix], clip[(size_t)(2 * n + d) * N + w], lim.frc_lo[d], lim.frc_hi[d]);
    }
  }
}

// helpers of ddb_jacobians with contacts: unit loss gradients in, one Jacobian row out
__global__ void k_unit_vector(int N, int n, int layout, int j, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * n) return;
  const int w = i / n, d = i % n;
  out[vidx(layout, N, n, w, d)] = d == j ? 1.0 : 0.0;
}
__global__ void k_scatter_row(int N, int n, int layout, int j, const double* __restrict__ src, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * n) return;
  const int w = i / n, d = i % n;
  dst[(size_t)w * n * n + (size_t)d * n + j] = src[vidx(layout, N, n, w, d)];
}

}  // namespace

namespace ddb {
cudaError_t contactBindTopology(const DevTopo* topo, const void* ctopo_host, cudaStream_t s) {
  cudaError_t e = cudaMemcpyToSymbolAsync(c_topo, topo, sizeof(DevTopo), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbolAsync(c_ctopo, ctopo_host, sizeof(ContactTopo), 0, cudaMemcpyHostToDevice, s);
}
size_t contactTopoBytes() { return sizeof(ContactTopo); }
void contactScratchSizes(int N, int nb, int nd, int cmax, int kmax, size_t* wd, size_t* wi, size_t* ld, size_t* li) {
  ContactLayout L{nb, nd, cmax, kmax};
  const size_t warps = ((size_t)N + 31) / 32;
  ContactBwdLayout BL{nb, nd, kmax};
  *wd = (size_t)N * (L.doubles() > BL.doubles() ? L.doubles() : BL.doubles());
  *wi = (size_t)N * (L.ints() > BL.ints() ? L.ints() : BL.ints());
  *ld = warps * lcpScratchDoubles(kmax) * 32;  // one slice per world (contact kernels: stride 1; the rounding keeps the
  *li = warps * lcpScratchInts(kmax) * 32;     // interleaved layout of k_lcp_solve's callers valid too)
}
// DDB_CONTACT_PROFILE=1: clock64 ticks summed over worlds per section of the contact stage (see ContactWork::mark)
static long long* contactProfileBuffer() {  // one buffer per device (the current one)
  static long long* buf[64] = {nullptr};
  static bool init[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!init[dev]) {
    init[dev] = true;
    const char* e = getenv("DDB_CONTACT_PROFILE");
    if (e && e[0] == '1' && cudaMalloc(&buf[dev], 16 * sizeof(long long)) == cudaSuccess) cudaMemset(buf[dev], 0, 16 * sizeof(long long));
    else buf[dev] = nullptr;
  }
  return buf[dev];
}
// Shared memory of one world of at most cap rows: packed triangle, seven vectors, the friction indices
// (+ the helper's two partial-sum buffers and the stop flag of the warp pairs)
static size_t pgsWarpBytes(int cap) {
  return (size_t)((cap * (cap + 1)) / 2 + 7 * cap + (cap + 1) / 2 + (cap > 32 ? 65 : 0)) * sizeof(double);
}
// The size classes (nLow, cap] of the warp solver. A world's shared memory is sized by its class, and the resident worlds per SM
// (all the latency hiding this solver has: Gauss-Seidel is one dependent chain per world) by that: box stacking hands over problems of
// 60 ... 108 rows in steps of 6 (two contacts), and a class that ends just above a frequent size holds 6 or 7 worlds per SM where
// a coarser one held 5 (measured).
// Classes above 32 rows (more than one row block) run as warp pairs.
static const int kPgsCaps[] = {16, 24, 32, 40, 48, 60, 66, 72, 78, 84, 90, 96, 102, 108, 116, 128, 144, 160, 192};
constexpr int kPgsClasses = sizeof(kPgsCaps) / sizeof(kPgsCaps[0]);
constexpr int kPgsWarpMax = 192;
static cudaError_t launchPgsWarp(cudaStream_t s, int N, int kmax, int nLow, int cap, int* ctrl, double* lcp_d, int* lcp_i, double* ans_x, int* ans_ok,
                                 const PgsOptions& pgs) {
  static bool configured[64] = {false};  // per device: the attribute belongs to the function in that context
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!configured[dev & 63]) {
    e = cudaFuncSetAttribute(k_pgs_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pgsWarpBytes(kPgsWarpMax));
    if (e != cudaSuccess) return e;
    configured[dev & 63] = true;
  }
  if (nLow >= 32) k_pgs_warp<true><<<N, 64, pgsWarpBytes(cap), s>>>(N, kmax, nLow, cap, ctrl, lcp_d, lcp_i, ans_x, ans_ok, pgs);
  else k_pgs_warp<false><<<N, 32, pgsWarpBytes(cap), s>>>(N, kmax, nLow, cap, ctrl, lcp_d, lcp_i, ans_x, ans_ok, pgs);
  return cudaGetLastError();
}
// helper streams of the concurrent PGS size classes, one set per device
constexpr int kPgsStreams = 8;
struct PgsStreams {
  cudaStream_t st[kPgsStreams];
  cudaEvent_t fork, join[kPgsStreams];
};
static PgsStreams* pgsStreams() {
  static PgsStreams sets[64];
  static bool ready[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!ready[dev]) {
    PgsStreams& p = sets[dev];
    for (int i = 0; i < kPgsStreams; i++) {
      if (cudaStreamCreateWithFlags(&p.st[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ready[dev] = true;
  }
  return &sets[dev];
}
void contactPgsSizes(int N, int kmax, size_t* ctrl_ints, size_t* ans_doubles, size_t* ans_ints) {
  *ctrl_ints = (size_t)N * 8;  // done, answered, pending rows, sweeps (statistics), resume tier, largest problem (statistics), 2 spare
  *ans_doubles = (size_t)N * kPgsRequests * kmax;
  *ans_ints = (size_t)N * kPgsRequests;
}
cudaError_t contactLaunch(cudaStream_t s, int N, int nb, int nd, int cmax, int kmax, int layout, double* snap, double* qd, double* lcp_cache,
                          int* n_cache, int* counts, int* active_out, double* work_d, int* work_i, double* lcp_d, int* lcp_i, double* keep_d,
                          int* keep_i, int* ctrl, double* ans_x, int* ans_ok, int* launches) {
  ContactLayout L{nb, nd, cmax, kmax};
  PgsOptions pgs{10000, 1e-10, 1e-8, 1e-8};
  cudaError_t e = cudaMemsetAsync(ctrl, 0, (size_t)N * 8 * sizeof(int), s);
  if (e != cudaSuccess) return e;
  // pass p answers the PGS solves asked for in passes < p; a world is committed in the first pass that runs through
  for (int pass = 0; pass <= kPgsRequests; pass++) {
    if (kmax <= kContactWideRows && nb <= DDB_KC_WIDE_BODIES) k_contact<DDB_KC_WIDE_BLOCK><<<(int)(((size_t)N * 32 + DDB_KC_WIDE_BLOCK - 1) / DDB_KC_WIDE_BLOCK), DDB_KC_WIDE_BLOCK, 0, s>>>(N, layout, L, snap, qd, lcp_cache, n_cache, counts, active_out, work_d, work_i, lcp_d, lcp_i, pgs,
                                           keep_d, keep_i, ctrl, ans_x, ans_ok, contactProfileBuffer());
    else k_contact<DDB_KC_BLOCK><<<(int)(((size_t)N * 32 + DDB_KC_BLOCK - 1) / DDB_KC_BLOCK), DDB_KC_BLOCK, 0, s>>>(N, layout, L, snap, qd, lcp_cache, n_cache, counts, active_out, work_d, work_i, lcp_d, lcp_i, pgs,
                                           keep_d, keep_i, ctrl, ans_x, ans_ok, contactProfileBuffer());
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) (*launches)++;
    if (pass == kPgsRequests) break;
    // The size classes touch disjoint worlds, so they run side by side on helper streams (fork / join on events): the small
    // problems fill the SMs while the few worlds that need all 10000 sweeps finish.
    PgsStreams* ps = pgsStreams();
    if (!ps) return cudaErrorUnknown;
    if ((e = cudaEventRecord(ps->fork, s)) != cudaSuccess) return e;
    bool used[kPgsStreams] = {false};
    int nLaunched = 0;
    auto onHelper = [&](auto launch) -> cudaError_t {
      const int i = nLaunched++ % kPgsStreams;
      cudaError_t r = cudaSuccess;
      if (!used[i]) { used[i] = true; if ((r = cudaStreamWaitEvent(ps->st[i], ps->fork, 0)) != cudaSuccess) return r; }
      return launch(ps->st[i]);
    };
    // The largest problems first (fewest resident worlds per SM, last to finish): measured on box stacking, 386 ms per step in this
    // order, 398 ms from the middle class outwards, 438 ms smallest first; the single-block solver's launch therefore comes last.
    // DDB_PGS_WARP_MAX=<rows> in the environment sends smaller problems
    // to the scalar kernel as well: that path is otherwise only reached by worlds of more than 64 contacts (tests)
    int warpMax = kPgsWarpMax;
    if (const char* env = getenv("DDB_PGS_WARP_MAX")) { const int m = atoi(env); if (m >= 0 && m < warpMax) warpMax = m; }
    if (kmax > warpMax) {
      if ((e = onHelper([&](cudaStream_t hs) {
             k_pgs_thread<<<(N + 63) / 64, 64, 0, hs>>>(N, kmax, warpMax, ctrl, lcp_d, lcp_i, ans_x, ans_ok, pgs);
             return cudaGetLastError();
           })) != cudaSuccess) return e;
    }
    for (int c = kPgsClasses - 1; c >= 0; c--) {
      const int nLow = c ? kPgsCaps[c - 1] : 0;
      if (nLow >= kmax || nLow >= warpMax || (warpMax >= 32 && kPgsCaps[c] <= 32)) continue;
      const int cap = kPgsCaps[c] < warpMax ? kPgsCaps[c] : warpMax;
      if ((e = onHelper([&](cudaStream_t hs) { return launchPgsWarp(hs, N, kmax, nLow, cap, ctrl, lcp_d, lcp_i, ans_x, ans_ok, pgs); })) != cudaSuccess) return e;
    }
    if (warpMax >= 32) {  // one block of rows: the register-resident solver
      if ((e = onHelper([&](cudaStream_t hs) {
             k_pgs_small<<<N, 32, 0, hs>>>(N, kmax, ctrl, lcp_d, lcp_i, ans_x, ans_ok, pgs);
             return cudaGetLastError();
           })) != cudaSuccess) return e;
    }
    for (int i = 0; i < kPgsStreams; i++) {
      if (!used[i]) continue;
      if ((e = cudaEventRecord(ps->join[i], ps->st[i])) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(s, ps->join[i], 0)) != cudaSuccess) return e;
    }
    if (launches) (*launches) += nLaunched;
  }
  return cudaSuccess;
}
int contactProfileRead(long long* out16) {
  long long* b = contactProfileBuffer();
  if (!b) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(out16, b, 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemset(b, 0, 16 * sizeof(long long));
  return 1;
}
void contactKeepSizes(int N, int nd, int cmax, int kmax, size_t* kd, size_t* ki) {
  ContactSnapLayout KL{nd, cmax, kmax};
  *kd = (size_t)N * KL.doubles();
  *ki = (size_t)N * KL.ints();
}
cudaError_t contactBackwardLaunch(cudaStream_t s, int N, int nb, int nd, int cmax, int kmax, int layout, double* snap, const double* keep_d,
                                  const int* keep_i, const double* gqn, const double* gqdn, double* gq, double* gqd, double* gtau, double* work_d,
                                  int* work_i, double* lcp_d, int* lcp_i, const double* lim6, int* status) {
  ContactSnapLayout KL{nd, cmax, kmax};
  ContactBwdLayout BL{nb, nd, kmax};
  Limits lim{lim6, lim6 ? lim6 + nd : nullptr, lim6 ? lim6 + 2 * nd : nullptr, lim6 ? lim6 + 3 * nd : nullptr, lim6 ? lim6 + 4 * nd : nullptr,
             lim6 ? lim6 + 5 * nd : nullptr};
  const int grid = (int)(((size_t)N * 32 + DDB_KCB_BLOCK - 1) / DDB_KCB_BLOCK);
#define DDB_CB(NBV) k_contact_backward<NBV><<<grid, DDB_KCB_BLOCK, 0, s>>>(N, layout, KL, BL, snap, keep_d, keep_i, gqn, gqdn, gq, gqd, gtau, work_d, work_i, lcp_d, lcp_i, lim, status)
  if (nb <= 8) DDB_CB(8);
  else if (nb <= 16) DDB_CB(16);
  else if (nb <= 32) DDB_CB(32);
  else DDB_CB(DDB_MAX_BODIES);
#undef DDB_CB
  return cudaGetLastError();
}
cudaError_t contactUnitVector(cudaStream_t s, int N, int n, int layout, int j, double* out) {
  k_unit_vector<<<(N * n + 255) / 256, 256, 0, s>>>(N, n, layout, j, out);
  return cudaGetLastError();
}
cudaError_t contactScatterRow(cudaStream_t s, int N, int n, int layout, int j, const double* src, double* dst) {
  k_scatter_row<<<(N * n + 255) / 256, 256, 0, s>>>(N, n, layout, j, src, dst);
  return cudaGetLastError();
}
}  // namespace ddb

extern "C" {

int ddb_collide_pairs(int32_t device, void* stream, int32_t n_pairs, const int32_t* shape1, const double* size1, const double* T1,
                      const int32_t* shape2, const double* size2, const double* T2, double* contacts_out, int32_t* counts_out) {
  if (n_pairs < 1 || !shape1 || !size1 || !T1 || !shape2 || !size2 || !T2 || !contacts_out || !counts_out) return DDB_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DDB_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return DDB_ERR_CUDA;
  k_collide_pairs<<<(n_pairs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_pairs, shape1, size1, T1, shape2, size2, T2, contacts_out, counts_out);
  return cudaGetLastError() == cudaSuccess ? DDB_OK : DDB_ERR_CUDA;
}

int ddb_lcp_solve(int32_t device, void* stream, int32_t n_problems, int32_t n_max, const int32_t* n_rows, const double* A, const double* b,
                  const double* lo, const double* hi, const int32_t* findex, int32_t solver, int32_t early_termination, double* x,
                  int32_t* ok) {
  if (n_problems < 1 || n_max < 1 || !A || !b || !lo || !hi || !findex || !x) return DDB_ERR_INVALID;
  if (solver != DDB_LCP_DANTZIG && solver != DDB_LCP_PGS) return DDB_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DDB_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return DDB_ERR_CUDA;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t warps = ((size_t)n_problems + 31) / 32;
  double* sd = nullptr;
  int* si = nullptr;
  if (cudaMallocAsync(&sd, warps * lcpScratchDoubles(n_max) * 32 * sizeof(double), s) != cudaSuccess) return DDB_ERR_CUDA;
  if (cudaMallocAsync(&si, warps * lcpScratchInts(n_max) * 32 * sizeof(int), s) != cudaSuccess) { cudaFreeAsync(sd, s); return DDB_ERR_CUDA; }
  // PgsBoxedLcpSolver options as BoxedLcpConstraintSolver sets them (BoxedLcpConstraintSolver.cpp:81-83)
  PgsOptions pgs{10000, 1e-10, 1e-8, 1e-8};
  k_lcp_solve<<<(n_problems + 63) / 64, 64, 0, s>>>(n_problems, n_max, n_rows, A, b, lo, hi, findex, solver, early_termination, pgs, x, ok, sd, si);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(sd, s);
  cudaFreeAsync(si, s);
  return e == cudaSuccess ? DDB_OK : DDB_ERR_CUDA;
}

int ddb_lcp_valid(int32_t device, void* stream, int32_t n_problems, int32_t n_max, const int32_t* n_rows, const double* A, const double* x,
                  const double* b, const double* lo, const double* hi, const int32_t* findex, int32_t ignore_friction_indices,
                  int32_t* valid_out) {
  if (n_problems < 1 || n_max < 1 || !A || !x || !b || !lo || !hi || !findex || !valid_out) return DDB_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DDB_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return DDB_ERR_CUDA;
  k_lcp_valid<<<(n_problems + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_problems, n_max, n_rows, A, x, b, lo, hi, findex,
                                                                          ignore_friction_indices, valid_out);
  return cudaGetLastError() == cudaSuccess ? DDB_OK : DDB_ERR_CUDA;
}

}  // extern "C"
