// kernels.cuh — __global__ entry points of the contact-free path (one world per thread).
//   k_forward    neural::forwardPass -> World::step for N worlds (NeuralUtils.cpp:26-66, World.cpp:206-279)
//   k_backward   BackpropSnapshot::backprop as a vector-Jacobian product (BackpropSnapshot.cpp:115-186,374-428)
//   k_jacobians  BackpropSnapshot::get{PosPos,PosVel,VelPos,VelVel,ForceVel}Jacobian (:432-735,1237-1360)
//   k_inv_mass   Skeleton::getInvMassMatrix (Skeleton.cpp:2899-2905,4035-4098)
// Template parameter CHAIN: 0 = any tree (dynamics.cuh, run-time loops over the topology), 1 = serial chain of 1-DOF joints
// (chain.cuh, unrolled, canonical joint frames), 2 = such a chain with revolute joints only.
// The snapshot kept between forward and backward (per slot) is described in dynamics.cuh (struct Snap);
// q, qd, tau are appended only when joint limits are finite (clipLossGradientsToBounds needs them).
#pragma once
#include "dynamics.cuh"
#include "chain.cuh"

#ifndef DDB_JAC_RESIDENT
#define DDB_JAC_RESIDENT 1  // 0: k_jacobians / k_inv_mass read the snapshot with plain loads (A/B timing only)
#endif
#ifndef DDB_BWD_STAGE
#ifdef DDB_HOST_EMULATION
#define DDB_BWD_STAGE 0  // (no cp.async on the host)
#else
#define DDB_BWD_STAGE 1  // 0: the chain backward reads the snapshot with plain loads (A/B timing only)
#endif
#endif

// register cap of the chain forward / backward kernels: __launch_bounds__(BLOCK, MINB) by default; DDB_MAXNREG_FWD / _BWD = N caps them
// with __maxnreg__(N) instead (ptxas only offers 255 / 168 / 128 ... through the min-blocks argument; A/B timing of 9 and 10 warps per SM)
#ifdef DDB_MAXNREG_FWD
#define DDB_FWD_BOUNDS __maxnreg__(DDB_MAXNREG_FWD)
#else
#define DDB_FWD_BOUNDS __launch_bounds__(BLOCK, MINB)
#endif
#ifdef DDB_MAXNREG_BWD
#define DDB_BWD_BOUNDS __maxnreg__(DDB_MAXNREG_BWD)
#else
#define DDB_BWD_BOUNDS __launch_bounds__(BLOCK, MINB)
#endif

namespace ddb {

struct Limits {  // device arrays of length n (valid only when c_topo.has_limits)
  const double *pos_lo, *pos_hi, *vel_lo, *vel_hi, *frc_lo, *frc_hi;
};

// doubles per world in one snapshot slot
__host__ __device__ inline size_t snapDoublesPerWorld(int snapPairs, int n, bool limits) { return (size_t)2 * snapPairs + (limits ? 3 * n : 0); }
// worlds are padded to whole tiles of 32 in the snapshot
__host__ __device__ inline size_t snapWorldsPadded(int N) { return ((size_t)N + 31) / 32 * 32; }
// the clip section (q, qd, tau as [3n][N]) sits behind the per-body pairs
DDB_DI double* clipSection(double* snap, int /*nb*/, int N) { return snap + (size_t)2 * c_topo.snapPairs * snapWorldsPadded(N); }

// ---- host-buffer variants (HOSTIO) of the chain kernels: the e2e path of a binding that holds its state in host memory ----
// q, qd, tau, ... then point into PINNED HOST memory (mapped into the device's address space). A CTA moves the rows of its
// worlds (world-major: BLOCK * NB consecutive doubles per vector) between host memory and shared memory with coalesced
// 16-byte accesses, so PCIe sees full-width bursts; threads take their own values from shared memory. The copy engines are
// not involved: reads of later CTAs, arithmetic and result writes of earlier CTAs overlap on the full-duplex link without
// chunking the call, and there is no per-copy set-up cost (5 copies x chunks, ~5 us each, on the cudaMemcpyAsync path).
DDB_DI void ioCopyIn(double* __restrict__ dst, const double* __restrict__ src, int cnt, int tid, int nthreads) {
  // src is 16-byte aligned: CTA boundaries are multiples of 32 worlds
  const int pairs = cnt >> 1;
  for (int i = tid; i < pairs; i += nthreads) reinterpret_cast<double2*>(dst)[i] = reinterpret_cast<const double2*>(src)[i];
  if ((cnt & 1) && tid == 0) dst[cnt - 1] = src[cnt - 1];
}
DDB_DI void ioCopyOut(double* __restrict__ dst, const double* __restrict__ src, int cnt, int tid, int nthreads) {
  const int pairs = cnt >> 1;
  for (int i = tid; i < pairs; i += nthreads) reinterpret_cast<double2*>(dst)[i] = reinterpret_cast<const double2*>(src)[i];
  if ((cnt & 1) && tid == 0) dst[cnt - 1] = src[cnt - 1];
}
#ifdef DDB_HOST_EMULATION
DDB_DI void ctaSync() {}
DDB_DI double* ioShared() { return nullptr; }
#else
DDB_DI void ctaSync() { __syncthreads(); }
DDB_DI double* ioShared() { extern __shared__ double2 ddb_io_shared[]; return reinterpret_cast<double*>(ddb_io_shared); }
#endif

template <int NB, int CHAIN, int BLOCK, int MINB, bool HOSTIO = false>
__global__ void DDB_FWD_BOUNDS
k_forward(int N, int w0, int w1, const double* __restrict__ q, const double* __restrict__ qd, const double* __restrict__ tau,
          double* __restrict__ qn, double* __restrict__ qdn, double* __restrict__ snap) {
  // worlds [w0, w1) of the N-world batch (the host splits a batch into chunks that run on different streams)
  const int w = w0 + blockIdx.x * BLOCK + threadIdx.x;
  if (!HOSTIO && w >= w1) return;
  const int n = c_topo.nd, nb = c_topo.nb, layout = c_topo.layout;
  const double dt = c_topo.dt;
  Snap sn{reinterpret_cast<double2*>(snap), N, w, CHAIN ? kChainPairs : 0};
  double* clip = clipSection(snap, nb, N);
  if constexpr (CHAIN) {
    double ql[NB], qdl[NB], tl[NB], acc[NB];
    double* io = nullptr;
    const int wb = w0 + blockIdx.x * BLOCK;                // first world of this CTA
    const int cnt = ((wb + BLOCK < w1 ? wb + BLOCK : w1) - wb) * NB;       // doubles per vector of this CTA
    if (HOSTIO) {  // (world-major vectors)
      io = ioShared();
      ioCopyIn(io, q + (size_t)wb * NB, cnt, threadIdx.x, BLOCK);
      ioCopyIn(io + BLOCK * NB, qd + (size_t)wb * NB, cnt, threadIdx.x, BLOCK);
      ioCopyIn(io + 2 * BLOCK * NB, tau + (size_t)wb * NB, cnt, threadIdx.x, BLOCK);
      ctaSync();
      const int t = w < w1 ? threadIdx.x : 0;  // threads past the end shadow world wb (results dropped)
#pragma unroll
      for (int d = 0; d < NB; d++) { ql[d] = io[t * NB + d]; qdl[d] = io[(BLOCK + t) * NB + d]; tl[d] = io[(2 * BLOCK + t) * NB + d]; }
      ctaSync();
      if (w >= w1) sn.w = wb;
    } else {
#pragma unroll
      for (int d = 0; d < NB; d++) {
        size_t ix = vidx(layout, N, n, w, d);
        ql[d] = q[ix]; qdl[d] = qd[ix]; tl[d] = tau[ix];
      }
    }
    chainForwardDynamics<NB, CHAIN == 2>(
        sn, [&](int d) { return ql[d]; }, [&](int d) { return qdl[d]; }, [&](int d) { return tl[d]; },
        [&](int d, double a) { acc[d] = a; });
    if (HOSTIO) {
#pragma unroll
      for (int d = 0; d < NB; d++) {
        io[threadIdx.x * NB + d] = fma(dt, qdl[d], ql[d]);
        io[(BLOCK + threadIdx.x) * NB + d] = fma(dt, acc[d], qdl[d]);
        if (c_topo.has_limits && w < w1) {
          clip[(size_t)d * N + w] = ql[d];
          clip[(size_t)(n + d) * N + w] = qdl[d];
          clip[(size_t)(2 * n + d) * N + w] = tl[d];
        }
      }
      ctaSync();
      ioCopyOut(qn + (size_t)wb * NB, io, cnt, threadIdx.x, BLOCK);
      ioCopyOut(qdn + (size_t)wb * NB, io + BLOCK * NB, cnt, threadIdx.x, BLOCK);
      return;
    }
#pragma unroll
    for (int d = 0; d < NB; d++) {
      size_t ix = vidx(layout, N, n, w, d);
      // explicit position update with the pre-step velocity (World.cpp:262-274); qd += dt*qdd
      qn[ix] = fma(dt, qdl[d], ql[d]);
      qdn[ix] = fma(dt, acc[d], qdl[d]);
      if (c_topo.has_limits) {
        clip[(size_t)d * N + w] = ql[d];
        clip[(size_t)(n + d) * N + w] = qdl[d];
        clip[(size_t)(2 * n + d) * N + w] = tl[d];
      }
    }
  } else {
    forwardDynamics<NB, false>(
        sn, nb, [&](int d) { return q[vidx(layout, N, n, w, d)]; }, [&](int d) { return qd[vidx(layout, N, n, w, d)]; },
        [&](int d) { return tau[vidx(layout, N, n, w, d)]; },
        [&](int d, double a) {
          size_t ix = vidx(layout, N, n, w, d);
          qdn[ix] = fma(dt, a, qd[ix]);
        });
#pragma unroll 1
    for (int d = 0; d < n; d++) {
      size_t ix = vidx(layout, N, n, w, d);
      double qv = q[ix], qdv = qd[ix];
      qn[ix] = fma(dt, qdv, qv);
      if (c_topo.has_limits) {
        clip[(size_t)d * N + w] = qv;
        clip[(size_t)(n + d) * N + w] = qdv;
        clip[(size_t)(2 * n + d) * N + w] = tau[ix];
      }
    }
    if (c_topo.has_free) {
      // FreeJoint::integratePositionsExplicit (FreeJoint.cpp:596-601): Q+ = Q * [expMapRot(w dt) | v dt], q+ = [logMap(R+); p+]
#pragma unroll 1
      for (int i = 0; i < nb; i++) {
        const DevBody& b = c_topo.b[i];
        if (b.type != DDB_JOINT_FREE) continue;
        double p6[6], v6[6];
        for (int k = 0; k < 6; k++) { size_t ix = vidx(layout, N, n, w, b.dof + k); p6[k] = q[ix]; v6[k] = qd[ix]; }
        Xf Q, D;
        Q.R = expMapRot(mk(p6[0], p6[1], p6[2])); Q.p = mk(p6[3], p6[4], p6[5]);
        D.R = expMapRot(mk(v6[0] * dt, v6[1] * dt, v6[2] * dt)); D.p = mk(v6[3] * dt, v6[4] * dt, v6[5] * dt);
        Xf Qn = mulXf(Q, D);
        V3 lw = logMapRot(Qn.R);
        const double o6[6] = {lw.x, lw.y, lw.z, Qn.p.x, Qn.p.y, Qn.p.z};
        for (int k = 0; k < 6; k++) qn[vidx(layout, N, n, w, b.dof + k)] = o6[k];
      }
    }
  }
}

// clipLossGradientsToBounds (BackpropSnapshot.cpp:374-428): exact-equality tests against the limits
DDB_DI double clipGrad(double g, double x, double lo, double hi) {
  if (x == lo && g > 0) return 0.0;
  if (x == hi && g < 0) return 0.0;
  return g;
}

template <int NB, int CHAIN, int BLOCK, int MINB, bool HOSTIO = false>
__global__ void DDB_BWD_BOUNDS
k_backward(int N, int w0, int w1, double* __restrict__ snap, const double* __restrict__ gqn, const double* __restrict__ gqdn,
           const double* __restrict__ gqnAdd, const double* __restrict__ gqdnAdd,
           double* __restrict__ gq, double* __restrict__ gqd, double* __restrict__ gtau, Limits lim) {
  // loss gradient w.r.t. the post-step state = gqn (+ gqnAdd), gqdn (+ gqdnAdd): in a rollout the second term is what the
  // later steps pulled back (SingleShot.cpp:667-670); NULL otherwise
  int w = w0 + blockIdx.x * BLOCK + threadIdx.x;
  if (!HOSTIO && w >= w1) return;
  const int n = c_topo.nd, nb = c_topo.nb, layout = c_topo.layout;
  const double dt = c_topo.dt;
  // HOSTIO (chains, world-major, pinned host vectors; see k_forward): the CTA's rows travel through shared memory
  const int wb = w0 + blockIdx.x * BLOCK;
  const int cnt = ((wb + BLOCK < w1 ? wb + BLOCK : w1) - wb) * NB;
  const bool live = w < w1;
  if (HOSTIO && !live) w = wb;  // threads past the end shadow world wb (results dropped)
  Snap sn{reinterpret_cast<double2*>(snap), N, w, CHAIN ? kChainPairs : 0};
  BwdScratch<NB> ws;
  const double* mu = ws.mu;
  // chains read the snapshot through the cp.async staging area (chain.cuh, ChainStageA / ChainStageB); trees read global memory
#ifdef DDB_HOST_EMULATION
  double2* ddb_stage = nullptr;
#else
  extern __shared__ double2 ddb_stage[];
#endif
  ChainStageA<NB, BLOCK> sa;
  ChainStageB<NB, BLOCK> sb;
  constexpr bool STAGED = CHAIN != 0 && DDB_BWD_STAGE;
  // the I/O rows of the host-buffer variant sit behind the staging area: [3][BLOCK * NB] doubles
  double* io = HOSTIO ? reinterpret_cast<double*>(ddb_stage) + (STAGED ? ChainStageGeom<NB, BLOCK>::kBytes / sizeof(double) : 0) : nullptr;
  // the loss gradients are the first thing the up sweep consumes and the last thing the adjoint sweep adds: chains ask for them
  // before the staging copies instead of at the point of use
  double lamv[NB], gqv[NB];
  if (CHAIN && !HOSTIO) {
#pragma unroll
    for (int d = 0; d < NB; d++) {
      const size_t ix = vidx(layout, N, n, w, d);
      lamv[d] = gqdnAdd ? gqdn[ix] + gqdnAdd[ix] : gqdn[ix];
      gqv[d] = gqnAdd ? gqn[ix] + gqnAdd[ix] : gqn[ix];
    }
  }
  if (STAGED) {
    sa.init(sn, ddb_stage);
    sb.init(sn, ddb_stage);
    sa.start();
  }
  if (CHAIN && HOSTIO) {
    ioCopyIn(io, gqn + (size_t)wb * NB, cnt, threadIdx.x, BLOCK);
    ioCopyIn(io + BLOCK * NB, gqdn + (size_t)wb * NB, cnt, threadIdx.x, BLOCK);
    ctaSync();
    const int t = live ? threadIdx.x : 0;
#pragma unroll
    for (int d = 0; d < NB; d++) { gqv[d] = io[t * NB + d]; lamv[d] = io[(BLOCK + t) * NB + d]; }
    ctaSync();
  }
  // lambda = dL/dqd'
  auto gqnOf = [&](size_t ix) { return gqnAdd ? gqn[ix] + gqnAdd[ix] : gqn[ix]; };
  auto gqdnOf = [&](size_t ix) { return gqdnAdd ? gqdn[ix] + gqdnAdd[ix] : gqdn[ix]; };
  auto lamOf = [&](int d) { return CHAIN ? lamv[d < NB ? d : 0] : gqdnOf(vidx(layout, N, n, w, d)); };
  auto muFree = [&](int d, double m) { gtau[vidx(layout, N, n, w, d)] = dt * m; };
  if constexpr (CHAIN) {
    if (STAGED) chainSolveInvMass<NB, CHAIN == 2>(sa, lamOf, ws);
    else chainSolveInvMass<NB, CHAIN == 2>(ChainPlain{{}, sn}, lamOf, ws);
  } else {
    solveInvMass<NB, false>(sn, nb, lamOf, ws, muFree);
  }
  if (CHAIN) {
#pragma unroll
    for (int i = 0; i < NB; i++) {
      if (HOSTIO) io[(2 * BLOCK + threadIdx.x) * NB + i] = dt * mu[i];
      else gtau[vidx(layout, N, n, w, i)] = dt * mu[i];
    }
  } else {
#pragma unroll 1
    for (int i = 0; i < nb; i++)
      if (c_topo.b[i].ndof == 1) gtau[vidx(layout, N, n, w, c_topo.b[i].dof)] = dt * mu[i];
  }
  auto out1 = [&](int d, double dq, double dqd) {
    size_t ix = vidx(layout, N, n, w, d);
    double a = CHAIN ? gqv[d < NB ? d : 0] : gqnOf(ix), l = CHAIN ? lamv[d < NB ? d : 0] : gqdnOf(ix);
    // pos-pos = I, vel-pos = dt I for R^1 joints; pos-vel^T l = -dt dPhi/dq; vel-vel^T l = l - dt dPhi/dqd
    if (HOSTIO) {
      io[threadIdx.x * NB + d] = fma(-dt, dq, a);
      io[(BLOCK + threadIdx.x) * NB + d] = fma(dt, a, fma(-dt, dqd, l));
    } else {
      gq[ix] = fma(-dt, dq, a);
      gqd[ix] = fma(dt, a, fma(-dt, dqd, l));
    }
  };
  auto out6 = [&](int body, const double* dq6, const double* dqd6) {
    // FreeJoint block: dL/dq = posPos^T g_q' - dt dPhi/dq, dL/dqd = velPos^T g_q' + l - dt dPhi/dqd
    const DevBody& b = c_topo.b[body];
    double q6[6], qd6[6], pp[36], vp[36], a6[6];
    ld6(sn, body, 56, q6); ld6(sn, body, 48, qd6);
    freeJointPosJacobians(q6, qd6, dt, pp, vp);
    for (int k = 0; k < 6; k++) a6[k] = gqnOf(vidx(layout, N, n, w, b.dof + k));
    for (int k = 0; k < 6; k++) {
      double sp = 0.0, sv2 = 0.0;
      for (int r = 0; r < 6; r++) { sp += pp[r * 6 + k] * a6[r]; sv2 += vp[r * 6 + k] * a6[r]; }
      const size_t ix = vidx(layout, N, n, w, b.dof + k);
      gq[ix] = sp - dt * dq6[k];
      gqd[ix] = sv2 + gqdnOf(ix) - dt * dqd6[k];
    }
  };
  if constexpr (CHAIN) {
    if (STAGED) chainIdAdjoint<NB, CHAIN == 2>(sb, ws, out1);
    else chainIdAdjoint<NB, CHAIN == 2>(ChainPlain{{}, sn}, ws, out1);
  } else {
    idAdjoint<NB, false>(sn, nb, ws, out1, out6);
  }
  if (CHAIN && HOSTIO) {
    if (c_topo.has_limits) {
      const double* clip = clipSection(snap, nb, N);
#pragma unroll 1
      for (int d = 0; d < NB; d++) {
        double* o = io + threadIdx.x * NB + d;
        o[0] = clipGrad(o[0], clip[(size_t)d * N + w], lim.pos_lo[d], lim.pos_hi[d]);
        o[BLOCK * NB] = clipGrad(o[BLOCK * NB], clip[(size_t)(n + d) * N + w], lim.vel_lo[d], lim.vel_hi[d]);
        o[2 * BLOCK * NB] = clipGrad(o[2 * BLOCK * NB], clip[(size_t)(2 * n + d) * N + w], lim.frc_lo[d], lim.frc_hi[d]);
      }
    }
    ctaSync();
    ioCopyOut(gq + (size_t)wb * NB, io, cnt, threadIdx.x, BLOCK);
    ioCopyOut(gqd + (size_t)wb * NB, io + BLOCK * NB, cnt, threadIdx.x, BLOCK);
    ioCopyOut(gtau + (size_t)wb * NB, io + 2 * BLOCK * NB, cnt, threadIdx.x, BLOCK);
    return;
  }
  if (c_topo.has_limits) {
    const double* clip = clipSection(snap, nb, N);
#pragma unroll 1
    for (int d = 0; d < n; d++) {
      size_t ix = vidx(layout, N, n, w, d);
      gq[ix] = clipGrad(gq[ix], clip[(size_t)d * N + w], lim.pos_lo[d], lim.pos_hi[d]);
      gqd[ix] = clipGrad(gqd[ix], clip[(size_t)(n + d) * N + w], lim.vel_lo[d], lim.vel_hi[d]);
      gtau[ix] = clipGrad(gtau[ix], clip[(size_t)(2 * n + d) * N + w], lim.frc_lo[d], lim.frc_hi[d]);
    }
  }
}

// Each unit vector e_j pulled back through the step gives row j of vel-vel, pos-vel, force-vel.
// Outputs are [world][col][row] (column-major per world); NULL pointers are skipped.
#ifdef DDB_HOST_EMULATION
template <int NB, int CHAIN, int BLOCK> constexpr bool kJacResident = false;
#else
template <int NB, int CHAIN, int BLOCK> constexpr bool kJacResident = CHAIN != 0 && DDB_JAC_RESIDENT && ChainResident<NB, BLOCK>::kFits;
#endif
template <int NB, int CHAIN, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_jacobians(int N, double* __restrict__ snap, double* __restrict__ posPos, double* __restrict__ posVel,
            double* __restrict__ velPos, double* __restrict__ velVel, double* __restrict__ forceVel) {
  const int w = blockIdx.x * BLOCK + threadIdx.x;
  if (w >= N) return;
  const int n = c_topo.nd, nb = c_topo.nb;
  const double dt = c_topo.dt;
  Snap sn{reinterpret_cast<double2*>(snap), N, w, CHAIN ? kChainPairs : 0};
  BwdScratch<NB> ws;
  const double* mu = ws.mu;
  const size_t base = (size_t)w * n * n;
  auto reader = [&]() {  // chains small enough keep the snapshot in shared memory for the n pull-backs (ChainResident)
#ifndef DDB_HOST_EMULATION
    if constexpr (CHAIN != 0 && kJacResident<NB, CHAIN, BLOCK>) {
      extern __shared__ double2 ddb_stage[];
      ChainResident<NB, BLOCK> r;
      r.load(sn, ddb_stage);
      return r;
    } else
#endif
    {
      return ChainPlain{{}, sn};
    }
  };
  const auto rd = reader();
#pragma unroll 1
  for (int j = 0; j < n; j++) {
    if constexpr (CHAIN) chainSolveInvMass<NB, CHAIN == 2>(rd, [&](int d) { return d == j ? 1.0 : 0.0; }, ws);
    else solveInvMass<NB, false>(sn, nb, [&](int d) { return d == j ? 1.0 : 0.0; }, ws,
                                 [&](int d, double m) { if (forceVel) forceVel[base + (size_t)d * n + j] = dt * m; });
    if (forceVel) {
      if (CHAIN) {
#pragma unroll
        for (int i = 0; i < NB; i++) forceVel[base + (size_t)i * n + j] = dt * mu[i];
      } else {
#pragma unroll 1
        for (int i = 0; i < nb; i++)
          if (c_topo.b[i].ndof == 1) forceVel[base + (size_t)c_topo.b[i].dof * n + j] = dt * mu[i];
      }
    }
    if (posVel || velVel) {
      auto outRow = [&](int d, double dq, double dqd) {
        if (posVel) posVel[base + (size_t)d * n + j] = -dt * dq;
        if (velVel) velVel[base + (size_t)d * n + j] = (d == j ? 1.0 : 0.0) - dt * dqd;
      };
      if constexpr (CHAIN) chainIdAdjoint<NB, CHAIN == 2>(rd, ws, outRow);
      else idAdjoint<NB, false>(sn, nb, ws, outRow, [&](int body, const double* dq6, const double* dqd6) {
        const int d0 = c_topo.b[body].dof;
        for (int k = 0; k < 6; k++) {
          if (posVel) posVel[base + (size_t)(d0 + k) * n + j] = -dt * dq6[k];
          if (velVel) velVel[base + (size_t)(d0 + k) * n + j] = (d0 + k == j ? 1.0 : 0.0) - dt * dqd6[k];
        }
      });
    }
    if (posPos || velPos) {
#pragma unroll 1
      for (int d = 0; d < n; d++) {
        if (posPos) posPos[base + (size_t)d * n + j] = d == j ? 1.0 : 0.0;
        if (velPos) velPos[base + (size_t)d * n + j] = d == j ? dt : 0.0;
      }
    }
  }
  if (c_topo.has_free && (posPos || velPos)) {
    // FreeJoint blocks of pos-pos / vel-pos (the identity / dt entries written above are overwritten)
#pragma unroll 1
    for (int i = 0; i < nb; i++) {
      const DevBody& b = c_topo.b[i];
      if (b.type != DDB_JOINT_FREE) continue;
      double q6[6], qd6[6], pp[36], vp[36];
      ld6(sn, i, 56, q6); ld6(sn, i, 48, qd6);
      freeJointPosJacobians(q6, qd6, dt, pp, vp);
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
          if (posPos) posPos[base + (size_t)(b.dof + c) * n + b.dof + r] = pp[r * 6 + c];
          if (velPos) velPos[base + (size_t)(b.dof + c) * n + b.dof + r] = vp[r * 6 + c];
        }
    }
  }
}

template <int NB, int CHAIN, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_inv_mass(int N, double* __restrict__ snap, double* __restrict__ minv) {
  const int w = blockIdx.x * BLOCK + threadIdx.x;
  if (w >= N) return;
  const int n = c_topo.nd, nb = c_topo.nb;
  Snap sn{reinterpret_cast<double2*>(snap), N, w, CHAIN ? kChainPairs : 0};
  BwdScratch<NB> ws;
  const double* mu = ws.mu;
  const size_t base = (size_t)w * n * n;
#pragma unroll 1
  for (int j = 0; j < n; j++) {
    if constexpr (CHAIN) chainSolveInvMass<NB, CHAIN == 2>(ChainPlain{{}, sn}, [&](int d) { return d == j ? 1.0 : 0.0; }, ws);
    else solveInvMass<NB, false>(sn, nb, [&](int d) { return d == j ? 1.0 : 0.0; }, ws,
                                 [&](int d, double m) { minv[base + (size_t)j * n + d] = m; });
    if (CHAIN) {
#pragma unroll
      for (int i = 0; i < NB; i++) minv[base + (size_t)j * n + i] = mu[i];
    } else {
#pragma unroll 1
      for (int i = 0; i < nb; i++)
        if (c_topo.b[i].ndof == 1) minv[base + (size_t)j * n + c_topo.b[i].dof] = mu[i];
    }
  }
}

}  // namespace ddb
