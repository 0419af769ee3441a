// dynamics.cuh — per-thread (one world per thread) tree recursions for skeletons whose
// joints are Revolute / Prismatic / Weld, FP64, sm_100a.
//
// What each routine replaces in the reference:
//   kinematics()          Frame.cpp:144-160; RevoluteJoint.cpp:177-185; PrismaticJoint.cpp:179-187;
//                         GenericJoint.hpp:1740-1750 (partial acceleration)
//   articulated()         BodyNode::updateArtInertia BodyNode.cpp:1659-1686,
//                         GenericJoint.hpp:1825-1839,1936-1948
//   forwardDynamics()     Skeleton::computeForwardDynamics Skeleton.cpp:4734-4752
//                         (BodyNode.cpp:1689-1727,1772-1790; GenericJoint.hpp:2049-2074,2208-2229,2316-2327)
//   solveInvMass()        Skeleton::multiplyByImplicitInvMassMatrix Skeleton.cpp:2990-3060
//   vjp()                 the no-contact chain of BackpropSnapshot::backprop
//                         (BackpropSnapshot.cpp:115-186 with :464-467, :611-615, :1002-1076) —
//                         evaluated in closed form by a reverse sweep over the inverse-dynamics
//                         recursion instead of the reference's central differences
//                         (BackpropSnapshot.cpp:5990-6027,6121-6157; Skeleton.cpp:1912-1935).
//
// Template parameters: NB = capacity of the per-thread body arrays; CHAIN = the skeleton is a
// serial chain of 1-DOF joints with dof i on body i, so every array index is a compile-time
// constant after unrolling and the arrays live in registers.
#pragma once
#include "spatial.cuh"
#include "../../include/ddb.h"

namespace ddb {

struct DevBody {
  int parent, type, dof, gravity;
  // relative transform T(q) = (N0 + c N1 + s N2, P0 + c P1 + s P2) with (c,s) = (cos q, sin q)
  // for a revolute joint, (q, 0) for a prismatic joint, and (0, 0) for a weld (host-folded
  // T_parentToJoint * J(q) * T_childToJoint^-1).
  double N0[9], N1[9], N2[9];
  double P0[3], P1[3], P2[3];
  SV S;   // Joint::getRelativeJacobian() column (body frame, constant)
  RBI I;  // spatial inertia
};

struct DevTopo {
  int nb, nd, layout, has_limits;
  double g[3];
  double dt;
  DevBody b[DDB_MAX_BODIES];
};

// single translation unit: the topology of the handle that launched last lives here
__constant__ DevTopo c_topo;

// batched vector element (world w, dof d)
DDB_DI size_t vidx(int layout, int N, int n, int w, int d) {
  return layout == DDB_LAYOUT_WORLD_MAJOR ? (size_t)w * n + d : (size_t)d * N + w;
}

template <int NB, bool CHAIN>
struct TreeState {
  Xf T[NB];      // relative transforms
  SV V[NB];      // spatial velocities
  SV AIS[NB];    // IA * S
  double psi[NB];// 1 / (S^T IA S)   (0 for a weld)
  double qd[NB]; // joint velocity of the body's joint
  SV A[NB];      // spatial accelerations

  DDB_DI static int parentOf(int i) { return CHAIN ? i - 1 : c_topo.b[i].parent; }
  DDB_DI static int dofOf(int i) { return CHAIN ? i : c_topo.b[i].dof; }
  DDB_DI static bool hasDof(int i) { return CHAIN ? true : c_topo.b[i].type != DDB_JOINT_WELD; }

  DDB_DI void kinematicsBody(int i, double qi, double qdi) {
    const DevBody& b = c_topo.b[i];
    double c = 0.0, s = 0.0;
    if (b.type == DDB_JOINT_REVOLUTE) sincos(qi, &s, &c);
    else if (b.type == DDB_JOINT_PRISMATIC) c = qi;
    Xf t;
#pragma unroll
    for (int k = 0; k < 9; k++) t.R.m[k] = fma(s, b.N2[k], fma(c, b.N1[k], b.N0[k]));
    t.p.x = fma(s, b.P2[0], fma(c, b.P1[0], b.P0[0]));
    t.p.y = fma(s, b.P2[1], fma(c, b.P1[1], b.P0[1]));
    t.p.z = fma(s, b.P2[2], fma(c, b.P1[2], b.P0[2]));
    T[i] = t;
    qd[i] = qdi;
    SV Sq = qdi * b.S;
    const int p = parentOf(i);
    V[i] = (p >= 0) ? AdInvT(t, V[p]) + Sq : Sq;
  }
  DDB_DI SV eta(int i) const { return ad(V[i], qd[i] * c_topo.b[i].S); }
};

// ---------------------------------------------------------------------------------------------
// Forward dynamics. Fills st (T, V, AIS, psi, A) and writes qdd through the callback store(dof, v).
// tauOf(dof) supplies the joint force.
// ---------------------------------------------------------------------------------------------
template <int NB, bool CHAIN, class QF, class QDF, class TauF, class Store>
DDB_DI void forwardDynamics(TreeState<NB, CHAIN>& st, int nb, QF qOf, QDF qdOf, TauF tauOf, Store store) {
  typedef TreeState<NB, CHAIN> TS;
  // sweep 1: root -> leaves
  if (CHAIN) {
#pragma unroll
    for (int i = 0; i < NB; i++) st.kinematicsBody(i, qOf(i), qdOf(i));
  } else {
#pragma unroll 1
    for (int i = 0; i < nb; i++) {
      bool d = TS::hasDof(i);
      st.kinematicsBody(i, d ? qOf(TS::dofOf(i)) : 0.0, d ? qdOf(TS::dofOf(i)) : 0.0);
    }
  }
  // gravity in each body frame: G_i = R_i^T G_parent
  V3 G[NB];
  // sweep 2: leaves -> root. IA/bias accumulators per body.
  ABI IA[NB];
  SV bias[NB];
  double u[NB];
  auto initBody = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    V3 gp = (p >= 0) ? G[p] : mk(c_topo.g[0], c_topo.g[1], c_topo.g[2]);
    G[i] = mulT(st.T[i].R, gp);
    IA[i] = toABI(b.I);
    // bias = -dad(V, I V) - I * [0; R_w^T g]   (BodyNode.cpp:1692-1702)
    SV IV = mul(b.I, st.V[i]);
    SV fg = b.gravity ? mul(b.I, sv(mk(0, 0, 0), G[i])) : svzero();
    bias[i] = -dad(st.V[i], IV) - fg;
  };
  auto upBody = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    SV eta = st.eta(i);
    SV ais = svzero();
    double psi = 0.0, ui = 0.0;
    if (TS::hasDof(i)) {
      ais = mul(IA[i], b.S);
      psi = 1.0 / dot(b.S, ais);
      // total force u = tau - S^T (IA eta + bias)   (GenericJoint.hpp:2208-2229 with d = k = 0)
      ui = tauOf(TS::dofOf(i)) - dot(b.S, mul(IA[i], eta) + bias[i]);
    }
    st.AIS[i] = ais;
    st.psi[i] = psi;
    u[i] = ui;
    if (p >= 0) {
      // beta = bias + IA (eta + S psi u)   (GenericJoint.hpp:2049-2074)
      SV beta = bias[i] + mul(IA[i], eta) + (psi * ui) * ais;
      bias[p] = bias[p] + dAdInvT(st.T[i], beta);
      ABI Pi = IA[i];
      if (TS::hasDof(i)) rank1Downdate(Pi, ais, psi);
      addTransformedInertia(IA[p], st.T[i], Pi);
    }
  };
  auto downBody = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    SV Ap = (p >= 0) ? AdInvT(st.T[i], st.A[p]) : svzero();
    double acc = 0.0;
    if (TS::hasDof(i)) {
      // qdd = psi (u - (IA S)^T AdInvT(T, A_parent))   (GenericJoint.hpp:2316-2327)
      acc = st.psi[i] * (u[i] - dot(st.AIS[i], Ap));
      store(TS::dofOf(i), acc);
    }
    st.A[i] = Ap + acc * b.S + st.eta(i);
  };
  if (CHAIN) {
#pragma unroll
    for (int i = 0; i < NB; i++) initBody(i);
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) upBody(i);
#pragma unroll
    for (int i = 0; i < NB; i++) downBody(i);
  } else {
#pragma unroll 1
    for (int i = 0; i < nb; i++) initBody(i);
#pragma unroll 1
    for (int i = nb - 1; i >= 0; i--) upBody(i);
#pragma unroll 1
    for (int i = 0; i < nb; i++) downBody(i);
  }
}

// ---------------------------------------------------------------------------------------------
// Kinematics + articulated inertias only (no bias pass): what the backward needs when the
// accelerations are already known. accOf(dof) supplies qdd.
// ---------------------------------------------------------------------------------------------
template <int NB, bool CHAIN, class QF, class QDF, class AccF>
DDB_DI void recomputeState(TreeState<NB, CHAIN>& st, int nb, QF qOf, QDF qdOf, AccF accOf) {
  typedef TreeState<NB, CHAIN> TS;
  ABI IA[NB];
  auto down = [&](int i) {
    bool d = TS::hasDof(i);
    st.kinematicsBody(i, d ? qOf(TS::dofOf(i)) : 0.0, d ? qdOf(TS::dofOf(i)) : 0.0);
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    SV Ap = (p >= 0) ? AdInvT(st.T[i], st.A[p]) : svzero();
    double acc = d ? accOf(TS::dofOf(i)) : 0.0;
    st.A[i] = Ap + acc * b.S + st.eta(i);
    IA[i] = toABI(b.I);
  };
  auto up = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    SV ais = svzero();
    double psi = 0.0;
    if (TS::hasDof(i)) {
      ais = mul(IA[i], b.S);
      psi = 1.0 / dot(b.S, ais);
    }
    st.AIS[i] = ais;
    st.psi[i] = psi;
    if (p >= 0) {
      ABI Pi = IA[i];
      if (TS::hasDof(i)) rank1Downdate(Pi, ais, psi);
      addTransformedInertia(IA[p], st.T[i], Pi);
    }
  };
  if (CHAIN) {
#pragma unroll
    for (int i = 0; i < NB; i++) down(i);
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) up(i);
  } else {
#pragma unroll 1
    for (int i = 0; i < nb; i++) down(i);
#pragma unroll 1
    for (int i = nb - 1; i >= 0; i--) up(i);
  }
}

// ---------------------------------------------------------------------------------------------
// mu = M^-1 lambda with the cached articulated quantities; also leaves U_i (the spatial
// "velocity" of body i under joint rates mu) in U[].
// ---------------------------------------------------------------------------------------------
template <int NB, bool CHAIN, class LamF>
DDB_DI void solveInvMass(const TreeState<NB, CHAIN>& st, int nb, LamF lamOf, double* mu /*[NB] per body*/, SV* U) {
  typedef TreeState<NB, CHAIN> TS;
  SV c[NB];
  double a[NB];
  auto init = [&](int i) { c[i] = svzero(); };
  auto up = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    double ai = 0.0;
    if (TS::hasDof(i)) ai = lamOf(TS::dofOf(i)) - dot(b.S, c[i]);
    a[i] = ai;
    if (p >= 0) c[p] = c[p] + dAdInvT(st.T[i], c[i] + (st.psi[i] * ai) * st.AIS[i]);
  };
  auto down = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    SV Up = (p >= 0) ? AdInvT(st.T[i], U[p]) : svzero();
    double m = TS::hasDof(i) ? st.psi[i] * (a[i] - dot(st.AIS[i], Up)) : 0.0;
    mu[i] = m;
    U[i] = Up + m * b.S;
  };
  if (CHAIN) {
#pragma unroll
    for (int i = 0; i < NB; i++) init(i);
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) up(i);
#pragma unroll
    for (int i = 0; i < NB; i++) down(i);
  } else {
#pragma unroll 1
    for (int i = 0; i < nb; i++) init(i);
#pragma unroll 1
    for (int i = nb - 1; i >= 0; i--) up(i);
#pragma unroll 1
    for (int i = 0; i < nb; i++) down(i);
  }
}

// ---------------------------------------------------------------------------------------------
// Reverse sweep over the inverse-dynamics recursion: with Phi(q, qd) = mu^T ID(q, qd, qdd)
//   = sum_i U_i^T [ I_i (A_i - [0;G_i]) - dad(V_i, I_i V_i) ],
// returns dPhi/dq_i and dPhi/dqd_i for every joint through out(dof, dPhidq, dPhidqd).
// Adjoint recursions (leaves -> root), X_c^T = dAdInvT(T_c, .):
//   Ubar_i = F_i + sum X_c^T Ubar_c           F_i = I_i(A_i - g_i) - dad(V_i, I_i V_i)
//   Abar_i = I_i U_i + sum X_c^T Abar_c
//   Vbar_i = dad(U_i, I_i V_i) + I_i ad(U_i, V_i) - dad(S_i qd_i, Abar_i) + sum X_c^T Vbar_c
//   Gbar_i = -(I_i U_i).lin + sum R_c Gbar_c
//   dPhi/dqd_i = S_i . (Vbar_i + dad(V_i, Abar_i))
//   dPhi/dq_i  = -[ dad(S_i,Ubar_i).(X_i U_p) + dad(S_i,Abar_i).(X_i A_p) + dad(S_i,Vbar_i).(X_i V_p) ]
//                - Gbar_i . (w_S x G_i)
// ---------------------------------------------------------------------------------------------
template <int NB, bool CHAIN, class Out>
DDB_DI void idAdjoint(const TreeState<NB, CHAIN>& st, int nb, const SV* U, Out out) {
  typedef TreeState<NB, CHAIN> TS;
  V3 G[NB];
  SV Ub[NB], Ab[NB], Vb[NB];
  V3 Gb[NB];
  auto init = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    V3 gp = (p >= 0) ? G[p] : mk(c_topo.g[0], c_topo.g[1], c_topo.g[2]);
    G[i] = mulT(st.T[i].R, gp);
    SV IV = mul(b.I, st.V[i]);
    SV Aeff = st.A[i];
    if (b.gravity) Aeff.v = Aeff.v - G[i];
    SV P = mul(b.I, U[i]);
    Ub[i] = mul(b.I, Aeff) - dad(st.V[i], IV);
    Ab[i] = P;
    Vb[i] = dad(U[i], IV) + mul(b.I, ad(U[i], st.V[i]));
    Gb[i] = b.gravity ? -P.v : mk(0, 0, 0);
  };
  auto up = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    // Abar_i is complete here (children already added): finish Vbar_i
    SV Sq = st.qd[i] * b.S;
    Vb[i] = Vb[i] - dad(Sq, Ab[i]);
    if (TS::hasDof(i)) {
      double dqd = dot(b.S, Vb[i] + dad(st.V[i], Ab[i]));
      SV XU = (p >= 0) ? AdInvT(st.T[i], U[p]) : svzero();
      SV XA = (p >= 0) ? AdInvT(st.T[i], st.A[p]) : svzero();
      SV XV = (p >= 0) ? AdInvT(st.T[i], st.V[p]) : svzero();
      double dq = -(dot(dad(b.S, Ub[i]), XU) + dot(dad(b.S, Ab[i]), XA) + dot(dad(b.S, Vb[i]), XV)) - dot(Gb[i], cross(b.S.w, G[i]));
      out(TS::dofOf(i), dq, dqd);
    }
    if (p >= 0) {
      Ub[p] = Ub[p] + dAdInvT(st.T[i], Ub[i]);
      Ab[p] = Ab[p] + dAdInvT(st.T[i], Ab[i]);
      Vb[p] = Vb[p] + dAdInvT(st.T[i], Vb[i]);
      Gb[p] = Gb[p] + mul(st.T[i].R, Gb[i]);
    }
  };
  if (CHAIN) {
#pragma unroll
    for (int i = 0; i < NB; i++) init(i);
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) up(i);
  } else {
#pragma unroll 1
    for (int i = 0; i < nb; i++) init(i);
#pragma unroll 1
    for (int i = nb - 1; i >= 0; i--) up(i);
  }
}

}  // namespace ddb
