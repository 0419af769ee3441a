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
//   solveInvMass()+idAdjoint()  the no-contact chain of BackpropSnapshot::backprop
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

// ---------------------------------------------------------------------------------------------
// Snapshot of one step, written by the forward kernel and read by the backward / Jacobian
// kernels. 18 double2 per body, pair-major ([pair][world]) so that every access of a warp is
// one contiguous 512-byte segment:
//   pairs 0-5  relative transform T_i (R row-major, p)      pairs 12-14 AIS_i = IA_i S_i
//   pairs 6-8  spatial velocity V_i                         pair  15    (psi_i, qd_i)
//   pairs 9-11 spatial acceleration A_i                     pairs 16-17 (G_i = R_world^T g, qdd_i)
// Keeping these instead of recomputing them moves work from the FP64 pipe (the binding unit)
// to L2/HBM, which the path barely uses otherwise.
// ---------------------------------------------------------------------------------------------
constexpr int kSnapPairs = 18;

struct Snap {
  double2* base;  // [nb*18][N]
  int N, w;
  DDB_DI void st(int body, int k, double a, double b) const { base[(size_t)(body * kSnapPairs + k) * N + w] = make_double2(a, b); }
  DDB_DI double2 ld(int body, int k) const { return __ldg(base + (size_t)(body * kSnapPairs + k) * N + w); }
  DDB_DI void stSV(int body, int k, SV x) const {
    st(body, k, x.w.x, x.w.y); st(body, k + 1, x.w.z, x.v.x); st(body, k + 2, x.v.y, x.v.z);
  }
  DDB_DI SV ldSV(int body, int k) const {
    double2 a = ld(body, k), b = ld(body, k + 1), c = ld(body, k + 2);
    return sv(mk(a.x, a.y, b.x), mk(b.y, c.x, c.y));
  }
  DDB_DI void stT(int body, const Xf& t) const {
    st(body, 0, t.R.m[0], t.R.m[1]); st(body, 1, t.R.m[2], t.R.m[3]); st(body, 2, t.R.m[4], t.R.m[5]);
    st(body, 3, t.R.m[6], t.R.m[7]); st(body, 4, t.R.m[8], t.p.x); st(body, 5, t.p.y, t.p.z);
  }
  DDB_DI Xf ldT(int body) const {
    double2 a = ld(body, 0), b = ld(body, 1), c = ld(body, 2), d = ld(body, 3), e = ld(body, 4), f = ld(body, 5);
    Xf t;
    t.R.m[0] = a.x; t.R.m[1] = a.y; t.R.m[2] = b.x; t.R.m[3] = b.y; t.R.m[4] = c.x; t.R.m[5] = c.y;
    t.R.m[6] = d.x; t.R.m[7] = d.y; t.R.m[8] = e.x; t.p = mk(e.y, f.x, f.y);
    return t;
  }
};

template <int NB, bool CHAIN>
struct Tree {
  DDB_DI static int parentOf(int i) { return CHAIN ? i - 1 : c_topo.b[i].parent; }
  DDB_DI static int dofOf(int i) { return CHAIN ? i : c_topo.b[i].dof; }
  DDB_DI static bool hasDof(int i) { return CHAIN ? true : c_topo.b[i].type != DDB_JOINT_WELD; }
};

DDB_DI Xf jointTransform(const DevBody& b, double qi) {
  double c = 0.0, s = 0.0;
  if (b.type == DDB_JOINT_REVOLUTE) sincos(qi, &s, &c);
  else if (b.type == DDB_JOINT_PRISMATIC) c = qi;
  Xf t;
#pragma unroll
  for (int k = 0; k < 9; k++) t.R.m[k] = fma(s, b.N2[k], fma(c, b.N1[k], b.N0[k]));
  t.p.x = fma(s, b.P2[0], fma(c, b.P1[0], b.P0[0]));
  t.p.y = fma(s, b.P2[1], fma(c, b.P1[1], b.P0[1]));
  t.p.z = fma(s, b.P2[2], fma(c, b.P1[2], b.P0[2]));
  return t;
}

// ---------------------------------------------------------------------------------------------
// Forward dynamics (Featherstone ABA in body frames, d = k = 0) + snapshot.
//   sweep 1 root->leaves: T_i, V_i, G_i
//   sweep 2 leaves->root: IA_i, bias_i -> AIS_i, psi_i, u_i
//   sweep 3 root->leaves: qdd_i, A_i
// qOf/qdOf/tauOf(dof) supply the inputs, store(dof, qdd) receives the accelerations.
// ---------------------------------------------------------------------------------------------
template <int NB, bool CHAIN, class QF, class QDF, class TauF, class Store>
DDB_DI void forwardDynamics(const Snap& sn, int nb, QF qOf, QDF qdOf, TauF tauOf, Store store) {
  typedef Tree<NB, CHAIN> TS;
  struct {  // one object, see BwdScratch
    Xf T[NB];
    SV V[NB], AIS[NB], A[NB], bias[NB];
    V3 G[NB];
    ABI IA[NB];
    double psi[NB], u[NB], qdv[NB];
  } ws;
  Xf* T = ws.T;
  SV *V = ws.V, *AIS = ws.AIS, *A = ws.A, *bias = ws.bias;
  V3* G = ws.G;
  ABI* IA = ws.IA;
  double *psi = ws.psi, *u = ws.u, *qdv = ws.qdv;
  auto down1 = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    const bool d = TS::hasDof(i);
    const double qi = d ? qOf(TS::dofOf(i)) : 0.0, qdi = d ? qdOf(TS::dofOf(i)) : 0.0;
    Xf t = jointTransform(b, qi);
    T[i] = t;
    qdv[i] = qdi;
    SV Sq = qdi * b.S;
    V[i] = (p >= 0) ? AdInvT(t, V[p]) + Sq : Sq;
    V3 gp = (p >= 0) ? G[p] : mk(c_topo.g[0], c_topo.g[1], c_topo.g[2]);
    G[i] = mulT(t.R, gp);
    sn.stT(i, t);
    sn.stSV(i, 6, V[i]);
    IA[i] = toABI(b.I);
    // bias = -dad(V, I V) - I [0; R_w^T g]   (BodyNode.cpp:1692-1702)
    SV IV = mul(b.I, V[i]);
    SV fg = b.gravity ? mul(b.I, sv(mk(0, 0, 0), G[i])) : svzero();
    bias[i] = -dad(V[i], IV) - fg;
  };
  auto up = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    SV eta = ad(V[i], qdv[i] * b.S);  // partial acceleration (GenericJoint.hpp:1740-1750, dS = 0)
    SV ais = svzero();
    double ps = 0.0, ui = 0.0;
    if (TS::hasDof(i)) {
      ais = mul(IA[i], b.S);
      ps = 1.0 / dot(b.S, ais);
      // total force u = tau - S^T (IA eta + bias)   (GenericJoint.hpp:2208-2229 with d = k = 0)
      ui = tauOf(TS::dofOf(i)) - dot(b.S, mul(IA[i], eta) + bias[i]);
    }
    AIS[i] = ais;
    psi[i] = ps;
    u[i] = ui;
    sn.stSV(i, 12, ais);
    sn.st(i, 15, ps, qdv[i]);
    if (p >= 0) {
      // beta = bias + IA (eta + S psi u)   (GenericJoint.hpp:2049-2074)
      SV beta = bias[i] + mul(IA[i], eta) + (ps * ui) * ais;
      bias[p] = bias[p] + dAdInvT(T[i], beta);
      ABI Pi = IA[i];
      if (TS::hasDof(i)) rank1Downdate(Pi, ais, ps);
      addTransformedInertia(IA[p], T[i], Pi);
    }
  };
  auto down2 = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    SV Ap = (p >= 0) ? AdInvT(T[i], A[p]) : svzero();
    double acc = 0.0;
    if (TS::hasDof(i)) {
      // qdd = psi (u - (IA S)^T AdInvT(T, A_parent))   (GenericJoint.hpp:2316-2327)
      acc = psi[i] * (u[i] - dot(AIS[i], Ap));
      store(TS::dofOf(i), acc);
    }
    A[i] = Ap + acc * b.S + ad(V[i], qdv[i] * b.S);
    sn.stSV(i, 9, A[i]);
    sn.st(i, 16, G[i].x, G[i].y);
    sn.st(i, 17, G[i].z, acc);
  };
  if (CHAIN) {
#pragma unroll
    for (int i = 0; i < NB; i++) down1(i);
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) up(i);
#pragma unroll
    for (int i = 0; i < NB; i++) down2(i);
  } else {
#pragma unroll 1
    for (int i = 0; i < nb; i++) down1(i);
#pragma unroll 1
    for (int i = nb - 1; i >= 0; i--) up(i);
#pragma unroll 1
    for (int i = 0; i < nb; i++) down2(i);
  }
}

// ---------------------------------------------------------------------------------------------
// mu = M^-1 lambda from the snapshot (Skeleton::multiplyByImplicitInvMassMatrix,
// Skeleton.cpp:2990-3060): leaves->root bias pass, root->leaves solve. Leaves mu_i per body and
// U_i, the spatial velocity body i would have under joint rates mu.
// ---------------------------------------------------------------------------------------------
// All per-thread arrays of the backward live in ONE object: with separate local arrays nvcc 12.9's
// stack colouring overlapped U with pA in the generic (dynamically indexed) instantiation although
// both are live (seen in the PTX: both at __local_depot0+0), which zeroed U before the adjoint sweep.
template <int NB>
struct BwdScratch {
  SV U[NB];
  double mu[NB];
  SV c[NB];
  double a[NB];
  SV pU[NB], pA[NB], pV[NB];  // adjoint contributions pending from the children
  V3 pG[NB];
};

template <int NB, bool CHAIN, class LamF>
DDB_DI void solveInvMass(const Snap& sn, int nb, LamF lamOf, BwdScratch<NB>& ws) {
  typedef Tree<NB, CHAIN> TS;
  SV* c = ws.c;
  double* a = ws.a;
  double* mu = ws.mu;
  SV* U = ws.U;
  auto init = [&](int i) { c[i] = svzero(); };
  auto up = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    double ai = 0.0;
    if (TS::hasDof(i)) ai = lamOf(TS::dofOf(i)) - dot(b.S, c[i]);
    a[i] = ai;
    if (p >= 0) {
      SV ais = sn.ldSV(i, 12);
      double ps = sn.ld(i, 15).x;
      c[p] = c[p] + dAdInvT(sn.ldT(i), c[i] + (ps * ai) * ais);
    }
  };
  auto down = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    double m = 0.0;
    SV Up = svzero();
    if (p >= 0) Up = AdInvT(sn.ldT(i), U[p]);
    if (TS::hasDof(i)) {
      SV ais = sn.ldSV(i, 12);
      double ps = sn.ld(i, 15).x;
      m = ps * (a[i] - dot(ais, Up));
    }
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
// with X_i U_p = U_i - mu_i S_i, X_i V_p = V_i - qd_i S_i, X_i A_p = A_i - qdd_i S_i - ad(V_i, S_i qd_i).
// ---------------------------------------------------------------------------------------------
template <int NB, bool CHAIN, class Out>
DDB_DI void idAdjoint(const Snap& sn, int nb, BwdScratch<NB>& ws, Out out) {
  typedef Tree<NB, CHAIN> TS;
  const double* mu = ws.mu;
  const SV* U = ws.U;
  SV *pU = ws.pU, *pA = ws.pA, *pV = ws.pV;
  V3* pG = ws.pG;
  auto init = [&](int i) { pU[i] = svzero(); pA[i] = svzero(); pV[i] = svzero(); pG[i] = mk(0, 0, 0); };
  auto up = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    const SV Vi = sn.ldSV(i, 6), Ai = sn.ldSV(i, 9);
    const double qdi = sn.ld(i, 15).y;
    const double2 g0 = sn.ld(i, 16), g1 = sn.ld(i, 17);
    const V3 Gi = mk(g0.x, g0.y, g1.x);
    const double qddi = g1.y;
    const SV Ui = U[i];
    SV IV = mul(b.I, Vi);
    SV Aeff = Ai;
    if (b.gravity) Aeff.v = Aeff.v - Gi;
    SV P = mul(b.I, Ui);
    SV Sq = qdi * b.S;
    SV Ub = mul(b.I, Aeff) - dad(Vi, IV) + pU[i];
    SV Ab = P + pA[i];
    SV Vb = dad(Ui, IV) + mul(b.I, ad(Ui, Vi)) - dad(Sq, Ab) + pV[i];
    V3 Gb = b.gravity ? pG[i] - P.v : pG[i];
    if (TS::hasDof(i)) {
      double dqd = dot(b.S, Vb + dad(Vi, Ab));
      SV XU = Ui - mu[i] * b.S;
      SV XV = Vi - Sq;
      SV XA = Ai - qddi * b.S - ad(Vi, Sq);
      double dq = -(dot(dad(b.S, Ub), XU) + dot(dad(b.S, Ab), XA) + dot(dad(b.S, Vb), XV)) - dot(Gb, cross(b.S.w, Gi));
      out(TS::dofOf(i), dq, dqd);
    }
    if (p >= 0) {
      Xf t = sn.ldT(i);
      pU[p] = pU[p] + dAdInvT(t, Ub);
      pA[p] = pA[p] + dAdInvT(t, Ab);
      pV[p] = pV[p] + dAdInvT(t, Vb);
      pG[p] = pG[p] + mul(t.R, Gb);
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
