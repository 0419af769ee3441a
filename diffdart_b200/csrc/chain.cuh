// chain.cuh — serial chains of 1-DOF joints (Revolute / Prismatic, dof i on body i) in CANONICAL JOINT FRAMES, one world per
// thread, FP64, sm_100a. Same recursions as dynamics.cuh (which cites the reference file:line for each of them); what changes
// is the choice of body frames, which is private to these kernels (nothing else reads a chain's snapshot):
//
//   body frame i' = the joint frame of joint i on the child side, turned so that the joint axis is its z axis.
//   With X_i = T_childBodyToJoint,i * (A_i, 0), A_i e_z = axis_i, the pose of the new frame in the reference's body frame:
//     T'_i(q) = X_p^-1 T_i(q) X_i = F_i * Rz(q)      (revolute)        F_i = X_p^-1 T_parentBodyToJoint,i (A_i, 0)  constant
//             = F_i * Trans(q e_z)                   (prismatic)
//     S'_i    = (e_z; 0) or (0; e_z)                  (Joint::getRelativeJacobian in the new frame)
//     I'_i    = the rigid-body inertia about the new origin in the new axes
//   Joint accelerations, the generalized coordinates and every gradient w.r.t. them do not depend on the body frames.
//
// What it buys (SASS of the 5-link kernels): products with S become component picks (IA S is a column of IA,
// S^T x a component, ad(V, S qd) and dad(S, F) four multiplications or none), the relative transform needs 12 numbers of
// constants instead of 36 and 15 FP64 instructions instead of 24, and the snapshot keeps (cos q, sin q) instead of the 12
// numbers of T_i: 13 pairs per body instead of 18.
#pragma once
#include "dynamics.cuh"

namespace ddb {

// Snapshot pairs of one chain body (struct Snap with chainPairs = kChainPairs):
//   0      (cos q, sin q)  [prismatic: (q, 0)]        7-9    AIS_i = IA_i S_i
//   1-3    spatial velocity V_i                       10     (psi_i, qdd_i)
//   4-6    spatial acceleration A_i                   11-12  (G_i = R_world^T g, qd_i)
constexpr int kChainPairs = 13;
constexpr int cCS = 0, cV = 1, cA = 4, cAIS = 7, cPQ = 10, cG = 11;

// ---- products with S = (e_z; 0) [rev] or (0; e_z) [prismatic] ----
DDB_DI double pickS(bool rev, SV x) { return rev ? x.w.z : x.v.z; }                     // S^T x
DDB_DI void addS(bool rev, SV& x, double a) { if (rev) x.w.z += a; else x.v.z += a; }   // x += a S
DDB_DI SV colS(bool rev, const ABI& I) {                                                // IA S: a column of [[A, B], [B^T, D]]
  return rev ? sv(mk(I.A[4], I.A[5], I.A[2]), mk(I.B[6], I.B[7], I.B[8])) : sv(mk(I.B[2], I.B[5], I.B[8]), mk(I.D[4], I.D[5], I.D[2]));
}
// ad(V, S qd): rev qd (w x e_z; v x e_z), prismatic (0; qd w x e_z)
DDB_DI SV adS(bool rev, SV V, double qd) {
  const double a = qd * V.w.y, b = -(qd * V.w.x);
  if (rev) return sv(mk(a, b, 0.0), mk(qd * V.v.y, -(qd * V.v.x), 0.0));
  return sv(mk(0.0, 0.0, 0.0), mk(a, b, 0.0));
}
// dad(S, F).X for a wrench F and a twist X: rev (F.w x e_z).X.w + (F.v x e_z).X.v, prismatic (F.v x e_z).X.w
DDB_DI double dadSdot(bool rev, SV F, SV X) {
  if (rev) return fma(F.w.y, X.w.x, fma(-F.w.x, X.w.y, fma(F.v.y, X.v.x, -(F.v.x * X.v.y))));
  return fma(F.v.y, X.w.x, -(F.v.x * X.w.y));
}
// qd dad(S, F)
DDB_DI SV dadS(bool rev, SV F, double qd) {
  if (rev) return sv(mk(qd * F.w.y, -(qd * F.w.x), 0.0), mk(qd * F.v.y, -(qd * F.v.x), 0.0));
  return sv(mk(qd * F.v.y, -(qd * F.v.x), 0.0), mk(0.0, 0.0, 0.0));
}
// component S^T dad(V, F) of dad(s, t) = [t.w x s.w + t.v x s.v ; t.v x s.w]
DDB_DI double pickSdad(bool rev, SV V, SV F) {
  if (rev) return fma(F.w.x, V.w.y, -(F.w.y * V.w.x)) + fma(F.v.x, V.v.y, -(F.v.y * V.v.x));
  return fma(F.v.x, V.w.y, -(F.v.y * V.w.x));
}

// ALLREV: every joint of the chain is revolute (pendulums, arms), known when the handle is created: no selects, no branches
template <bool ALLREV>
DDB_DI bool chainRev(int i) { return ALLREV || c_topo.b[i].type == DDB_JOINT_REVOLUTE; }
DDB_DI void chainCS(bool rev, double qi, double& c, double& s) {
  if (rev) sincos(qi, &s, &c);
  else { c = qi; s = 0.0; }
}
// T'_i = F_i Rz(q) / F_i Trans(q e_z) with F_i = (N0, P0)
DDB_DI Xf chainT(const DevBody& b, bool rev, double c, double s) {
  Xf t;
  if (rev) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
      t.R.m[3 * r] = fma(s, b.N0[3 * r + 1], c * b.N0[3 * r]);
      t.R.m[3 * r + 1] = fma(c, b.N0[3 * r + 1], -(s * b.N0[3 * r]));
      t.R.m[3 * r + 2] = b.N0[3 * r + 2];
    }
    t.p = mk(b.P0[0], b.P0[1], b.P0[2]);
  } else {
#pragma unroll
    for (int k = 0; k < 9; k++) t.R.m[k] = b.N0[k];
    t.p = mk(fma(c, b.N0[2], b.P0[0]), fma(c, b.N0[5], b.P0[1]), fma(c, b.N0[8], b.P0[2]));
  }
  return t;
}

// ---------------------------------------------------------------------------------------------
// Staged snapshot rows of the chain backward (see the note above StageGeom in dynamics.cuh): each thread copies its own rows
// global -> shared with cp.async far ahead of their use.
//   phase A (M^-1 solve, up + down): rows 0, 7-10 of every body = 5 slots per body, requested kAhead bodies ahead of the up sweep
//   phase B (ID adjoint, leaves -> root): rows 0-6, 11-12 of one body = 9 slots, kBufs buffers; they reuse the slots of phase A
//     bodies the down sweep has finished with (chains long enough), so a thread owns max(5 NB, 18) slots of 16 bytes.
// ---------------------------------------------------------------------------------------------
template <int NB, int BLOCK>
struct ChainStageGeom {
  static constexpr int kRowsA = 5, kRowsB = 9, kBufs = DDB_STAGE_NBUF;
  static constexpr bool kAlias = kRowsA * NB >= kBufs * kRowsB && NB >= 3;
  static constexpr int kBufBase = kAlias ? 0 : kRowsA * NB;
  static constexpr int kSlots = (kRowsA * NB > kBufBase + kBufs * kRowsB) ? kRowsA * NB : kBufBase + kBufs * kRowsB;
  static constexpr size_t kBytes = (size_t)kSlots * BLOCK * sizeof(double2);
};

template <int NB, int BLOCK>
struct ChainStageBase {
  typedef ChainStageGeom<NB, BLOCK> G;
  const double2* g;   // this world's column of the snapshot in global memory
  size_t strideN;     // pairs between consecutive rows
  double2* sm;        // this thread's column of the CTA's staging area
  unsigned smAddr;
  DDB_DI void init(const Snap& sn, double2* smemBase) {
    g = sn.base + sn.origin();
    strideN = sn.rowStride();
    sm = smemBase + threadIdx.x;
    smAddr = sharedAddr(sm);
  }
  DDB_DI void copyRow(int slot, int body, int k) const {
    cpAsync16(smAddr + (unsigned)(slot * BLOCK * (int)sizeof(double2)), g + (size_t)(kChainPairs * body + k) * strideN);
  }
  DDB_DI static int slotA(int body, int k) { return G::kRowsA * body + (k == 0 ? 0 : k - 6); }
  DDB_DI static int bufOf(int body) { return G::kBufBase + ((NB - 1 - body) % G::kBufs) * G::kRowsB; }
  DDB_DI static int slotB(int body, int k) { return bufOf(body) + (k < 7 ? k : k - 4); }
  DDB_DI void requestA(int body) const {
    copyRow(slotA(body, 0), body, 0);
#pragma unroll
    for (int k = 7; k < 11; k++) copyRow(slotA(body, k), body, k);
    cpAsyncCommit();
  }
  DDB_DI void requestB(int body) const {  // body < 0: an empty group keeps the group count uniform
    if (body >= 0) {
#pragma unroll
      for (int k = 0; k < 7; k++) copyRow(slotB(body, k), body, k);
#pragma unroll
      for (int k = 11; k < 13; k++) copyRow(slotB(body, k), body, k);
    }
    cpAsyncCommit();
  }
};
template <int NB, int BLOCK>
struct ChainStageA : ChainStageBase<NB, BLOCK>, SnapReaders<ChainStageA<NB, BLOCK>> {
  typedef ChainStageBase<NB, BLOCK> B;
  typedef ChainStageGeom<NB, BLOCK> G;
  static constexpr int kAhead = DDB_STAGE_AHEAD;
  DDB_DI void requestAOrEmpty(int body) const { if (body >= 0) B::requestA(body); else cpAsyncCommit(); }
  DDB_DI void start() const {
    if (!G::kAlias) {
#pragma unroll
      for (int b = 0; b < G::kBufs; b++) B::requestB(NB - 1 - b);
    }
#pragma unroll
    for (int a = 0; a < kAhead; a++) requestAOrEmpty(NB - 1 - a);
  }
  DDB_DI double2 ld(int body, int k) const { return B::sm[B::slotA(body, k) * BLOCK]; }
  DDB_DI void beginUp(int i) const {
    cpAsyncWait<kAhead - 1>();
    requestAOrEmpty(i - kAhead);
  }
  // the down sweep is done with bodies 0..i: buffer b (slots [9 b, 9 b + 9)) is free once 5 (i + 1) >= 9 (b + 1)
  DDB_DI void endDown(int i) const {
    if (G::kAlias) {
#pragma unroll
      for (int b = 0; b < G::kBufs; b++) {
        const int need = (G::kRowsB * (b + 1) + G::kRowsA - 1) / G::kRowsA - 1;  // first i that frees buffer b
        if (i == need) B::requestB(NB - 1 - b);
      }
    }
  }
  DDB_DI void beginAdj(int) const {}
  DDB_DI void endAdj(int) const {}
};
template <int NB, int BLOCK>
struct ChainStageB : ChainStageBase<NB, BLOCK>, SnapReaders<ChainStageB<NB, BLOCK>> {
  typedef ChainStageBase<NB, BLOCK> B;
  DDB_DI double2 ld(int body, int k) const { return B::sm[B::slotB(body, k) * BLOCK]; }
  DDB_DI void beginUp(int) const {}
  DDB_DI void endDown(int) const {}
  DDB_DI void beginAdj(int) const { cpAsyncWait<ChainStageGeom<NB, BLOCK>::kBufs - 1>(); }
  DDB_DI void endAdj(int i) const { B::requestB(i - ChainStageGeom<NB, BLOCK>::kBufs); }
};
// The whole snapshot of this thread's world copied to shared memory once (cp.async, 13 pairs per body): for the kernels that walk
// it n times (k_jacobians, k_inv_mass: one pull-back per DOF), whose plain reads were three of every six cycles of a warp (long
// scoreboard; ncu, profiles/README.md round 2). A thread reads only its own column, so no CTA barrier is needed.
template <int NB, int BLOCK>
struct ChainResident : SnapReaders<ChainResident<NB, BLOCK>> {
  static constexpr int kRows = kChainPairs * NB;
  static constexpr size_t kBytes = (size_t)kRows * BLOCK * sizeof(double2);
  static constexpr bool kFits = kBytes <= 72 * 1024;  // three CTAs per SM
  const double2* sm;
  DDB_DI void load(const Snap& sn, double2* smemBase) {
    double2* col = smemBase + threadIdx.x;
    const double2* g = sn.base + sn.origin();
    const size_t strideN = sn.rowStride();
    const unsigned a = sharedAddr(col);
#pragma unroll
    for (int r = 0; r < kRows; r++) cpAsync16(a + (unsigned)(r * BLOCK * (int)sizeof(double2)), g + (size_t)r * strideN);
    cpAsyncCommit();
    cpAsyncWait<0>();
    sm = col;
  }
  DDB_DI double2 ld(int body, int k) const { return sm[(kChainPairs * body + k) * BLOCK]; }
  DDB_DI void beginUp(int) const {}
  DDB_DI void endDown(int) const {}
  DDB_DI void beginAdj(int) const {}
  DDB_DI void endAdj(int) const {}
};
// plain reads of the snapshot with the same interface (A/B timing with DDB_BWD_STAGE=0, k_jacobians, k_inv_mass, host emulation)
struct ChainPlain : SnapReaders<ChainPlain> {
  Snap sn;
  DDB_DI double2 ld(int body, int k) const { return sn.ld(body, k); }
  DDB_DI void beginUp(int) const {}
  DDB_DI void endDown(int) const {}
  DDB_DI void beginAdj(int) const {}
  DDB_DI void endAdj(int) const {}
};

// ---------------------------------------------------------------------------------------------
// Forward dynamics (Featherstone ABA, d = k = 0) + snapshot; sweeps as forwardDynamics() in dynamics.cuh.
// ---------------------------------------------------------------------------------------------
template <int NB, bool ALLREV, class QF, class QDF, class TauF, class Store>
DDB_DI void chainForwardDynamics(const Snap& sn, QF qOf, QDF qdOf, TauF tauOf, Store store) {
  struct {  // one object (see BwdScratch)
    SV V[NB], AIS[NB], A[NB], bias[NB], eta[NB];
    V3 G[NB];
    ABI IA[NB];
    double psi[NB], u[NB], qdv[NB], cth[NB], sth[NB];
  } ws;
  SV *V = ws.V, *AIS = ws.AIS, *A = ws.A, *bias = ws.bias, *eta = ws.eta;
  V3* G = ws.G;
  ABI* IA = ws.IA;
  double *psi = ws.psi, *u = ws.u, *qdv = ws.qdv, *cth = ws.cth, *sth = ws.sth;
  // T_i is rebuilt from (cos q_i, sin q_i) in sweeps 2 and 3 instead of occupying 24 registers per body across the whole
  // kernel (the empty asm makes the value opaque: otherwise the compiler merges the rebuild with the first evaluation)
#ifdef DDB_HOST_EMULATION
  auto opaque = [](double x) { return x; };
#else
  auto opaque = [](double x) { asm volatile("" : "+d"(x)); return x; };
#endif
  auto Tof = [&](int i) -> Xf { return chainT(c_topo.b[i], chainRev<ALLREV>(i), opaque(cth[i]), opaque(sth[i])); };
  auto down1 = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const bool rev = chainRev<ALLREV>(i);
    const double qdi = qdOf(i);
    chainCS(rev, qOf(i), cth[i], sth[i]);
    const Xf t = chainT(b, rev, cth[i], sth[i]);
    qdv[i] = qdi;
    SV Vi = (i > 0) ? AdInvT(t, V[i - 1]) : svzero();
    addS(rev, Vi, qdi);
    V[i] = Vi;
    const V3 gp = (i > 0) ? G[i - 1] : mk(c_topo.g[0], c_topo.g[1], c_topo.g[2]);
    G[i] = mulT(t.R, gp);
    sn.st(i, cCS, cth[i], sth[i]);
    sn.stSV(i, cV, Vi);
    sn.st(i, cG, G[i].x, G[i].y);
    sn.st(i, cG + 1, G[i].z, qdi);
  };
  auto rigid = [&](int i) {
    const DevBody& b = c_topo.b[i];
    IA[i] = toABI(b.I);
    // bias = -dad(V, I V) - I [0; R_w^T g]   (BodyNode.cpp:1692-1702)
    const SV IV = mul(b.I, V[i]);
    const SV fg = b.gravity ? mul(b.I, sv(mk(0, 0, 0), G[i])) : svzero();
    bias[i] = -dad(V[i], IV) - fg;
    // partial acceleration ad(V, S qd) (GenericJoint.hpp:1740-1750, dS = 0): all that sweeps 2 and 3 need of V_i
    eta[i] = adS(chainRev<ALLREV>(i), V[i], qdv[i]);
  };
  auto up = [&](int i) {
    const bool rev = chainRev<ALLREV>(i);
    const SV ais = colS(rev, IA[i]);
    const double ps = 1.0 / pickS(rev, ais);
    // total force u = tau - S^T (IA eta + bias)   (GenericJoint.hpp:2208-2229 with d = k = 0)
    const SV w = mul(IA[i], eta[i]) + bias[i];
    const double ui = tauOf(i) - pickS(rev, w);
    AIS[i] = ais;
    psi[i] = ps;
    u[i] = ui;
    sn.stSV(i, cAIS, ais);
    if (i > 0) {
      // beta = bias + IA (eta + S psi u)   (GenericJoint.hpp:2049-2074)
      const SV beta = w + (ps * ui) * ais;
      const Xf Ti = Tof(i);
      bias[i - 1] = bias[i - 1] + dAdInvT(Ti, beta);
      ABI Pi = IA[i];
      rank1Downdate(Pi, ais, ps);
      addTransformedInertia(IA[i - 1], Ti, Pi);
    }
  };
  auto down2 = [&](int i) {
    const bool rev = chainRev<ALLREV>(i);
    SV Ai = (i > 0) ? AdInvT(Tof(i), A[i - 1]) : svzero();
    // qdd = psi (u - (IA S)^T AdInvT(T, A_parent))   (GenericJoint.hpp:2316-2327)
    const double acc = psi[i] * (u[i] - dot(AIS[i], Ai));
    store(i, acc);
    Ai = Ai + eta[i];
    addS(rev, Ai, acc);
    A[i] = Ai;
    sn.stSV(i, cA, Ai);
    sn.st(i, cPQ, psi[i], acc);
  };
#pragma unroll
  for (int i = 0; i < NB; i++) { down1(i); rigid(i); }
#pragma unroll
  for (int i = NB - 1; i >= 0; i--) up(i);
#pragma unroll
  for (int i = 0; i < NB; i++) down2(i);
}

// ---------------------------------------------------------------------------------------------
// mu = M^-1 lambda from the snapshot (solveInvMass() in dynamics.cuh): leaves U_i, the spatial velocity body i would have
// under joint rates mu, and mu_i.
// ---------------------------------------------------------------------------------------------
template <int NB, bool ALLREV, class SN, class LamF>
DDB_DI void chainSolveInvMass(const SN& sn, LamF lamOf, BwdScratch<NB>& ws) {
  SV* c = ws.c;
  double* a = ws.a;
  double* mu = ws.mu;
  SV* U = ws.U;
  auto Tof = [&](int i) -> Xf { const double2 cs = sn.ld(i, cCS); return chainT(c_topo.b[i], chainRev<ALLREV>(i), cs.x, cs.y); };
#pragma unroll
  for (int i = 0; i < NB; i++) c[i] = svzero();
#pragma unroll
  for (int i = NB - 1; i >= 0; i--) {
    const bool rev = chainRev<ALLREV>(i);
    sn.beginUp(i);
    const double ai = lamOf(i) - pickS(rev, c[i]);
    a[i] = ai;
    if (i > 0) {
      const SV ais = sn.ldSV(i, cAIS);
      const double ps = sn.ld(i, cPQ).x;
      c[i - 1] = c[i - 1] + dAdInvT(Tof(i), c[i] + (ps * ai) * ais);
    }
  }
#pragma unroll
  for (int i = 0; i < NB; i++) {
    const bool rev = chainRev<ALLREV>(i);
    SV Ui = svzero();
    const double ps = sn.ld(i, cPQ).x;
    double m = ps * a[i];
    if (i > 0) {
      Ui = AdInvT(Tof(i), U[i - 1]);
      m = ps * (a[i] - dot(sn.ldSV(i, cAIS), Ui));
    }
    mu[i] = m;
    addS(rev, Ui, m);
    U[i] = Ui;
    sn.endDown(i);
  }
}

// ---------------------------------------------------------------------------------------------
// Reverse sweep over the inverse-dynamics recursion (idAdjointEx<MODE 0> in dynamics.cuh, which states the adjoint recursions):
// out(dof, dPhi/dq, dPhi/dqd) for Phi(q, qd) = mu^T ID(q, qd, qdd). The terms -mu S, -qd S, -qdd S of X_i U_p, X_i V_p, X_i A_p
// only touch z components, which dad(S, .) does not look at.
// ---------------------------------------------------------------------------------------------
template <int NB, bool ALLREV, class SN, class Out>
DDB_DI void chainIdAdjoint(const SN& sn, BwdScratch<NB>& ws, Out out) {
  const SV* U = ws.U;
  SV *pU = ws.pU, *pA = ws.pA, *pV = ws.pV;
  V3* pG = ws.pG;
#pragma unroll
  for (int i = 0; i < NB; i++) { pU[i] = svzero(); pA[i] = svzero(); pV[i] = svzero(); pG[i] = mk(0, 0, 0); }
#pragma unroll
  for (int i = NB - 1; i >= 0; i--) {
    const DevBody& b = c_topo.b[i];
    const bool rev = chainRev<ALLREV>(i);
    sn.beginAdj(i);
    const SV Vi = sn.ldSV(i, cV), Ai = sn.ldSV(i, cA);
    const double2 g0 = sn.ld(i, cG), g1 = sn.ld(i, cG + 1);
    const double qdi = g1.y;
    const V3 Gi = mk(g0.x, g0.y, g1.x);
    const SV Ui = U[i];
    const SV IV = mul(b.I, Vi);
    SV Aeff = Ai;
    if (b.gravity) Aeff.v = Aeff.v - Gi;
    const SV P = mul(b.I, Ui);
    const SV Ub = mul(b.I, Aeff) - dad(Vi, IV) + pU[i];
    const SV Ab = P + pA[i];
    const SV Vb = dad(Ui, IV) + mul(b.I, ad(Ui, Vi)) - dadS(rev, Ab, qdi) + pV[i];
    const V3 Gb = b.gravity ? pG[i] - P.v : pG[i];
    const double dqd = pickS(rev, Vb) + pickSdad(rev, Vi, Ab);
    const SV XA = Ai - adS(rev, Vi, qdi);
    double dq = -(dadSdot(rev, Ub, Ui) + dadSdot(rev, Ab, XA) + dadSdot(rev, Vb, Vi));
    if (rev) dq -= fma(Gb.y, Gi.x, -(Gb.x * Gi.y));  // Gbar . (e_z x G)
    out(i, dq, dqd);
    if (i > 0) {
      const double2 cs = sn.ld(i, cCS);
      const Xf t = chainT(b, rev, cs.x, cs.y);
      pU[i - 1] = pU[i - 1] + dAdInvT(t, Ub);
      pA[i - 1] = pA[i - 1] + dAdInvT(t, Ab);
      pV[i - 1] = pV[i - 1] + dAdInvT(t, Vb);
      pG[i - 1] = pG[i - 1] + mul(t.R, Gb);
    }
    sn.endAdj(i);
  }
}

}  // namespace ddb
