// TEST INFRASTRUCTURE ONLY — see dd_spatial.hpp.
//
// dd_dynamics.hpp — CPU restatement of the reference's tree recursions for a
// world described by a ddb_topology:
//   kinematics           Frame.cpp:144-160,261-265; GenericJoint.hpp:1740-1750;
//                        RevoluteJoint.cpp:140-189; PrismaticJoint.cpp:139-187;
//                        FreeJoint.cpp:64-81,673-693
//   articulated inertia  BodyNode.cpp:1659-1686; GenericJoint.hpp:1825-1839,1936-1948
//   forward dynamics     Skeleton.cpp:4734-4752; BodyNode.cpp:1689-1727,1772-1790;
//                        GenericJoint.hpp:2049-2074,2208-2229,2316-2327
//   M^-1 (columns)       Skeleton.cpp:4035-4098; BodyNode.cpp:2253-2322;
//                        GenericJoint.hpp:2527-2593
//   M^-1 x (implicit)    Skeleton.cpp:2990-3060
//   C + g                Skeleton.cpp:4352-4381; BodyNode.cpp:2060-2106
//   impulse dynamics     Skeleton.cpp:4859-4894,4990-5033; BodyNode.cpp:1730-1751,1793-1820;
//                        GenericJoint.hpp:2135-2152,2266-2272,2365-2377
//   position integration GenericJoint.hpp:1414-1421; FreeJoint.cpp:587-601
// Springs and damping are identically zero on this path (World.cpp:698-732).
#pragma once
#include <vector>
#include <cassert>
#include "../include/ddb.h"
#include "dd_spatial.hpp"

namespace ddo {

struct Model {
  struct Body {
    int parent, type, dof, ndof, skel;
    bool gravity;
    Iso<double> Tp2j, Tc2j, Tc2jInv;
    Vec3<double> axis;
    Mat6<double> I;
    double S[6][6];  // S[row][k], k < ndof: Joint::getRelativeJacobian()
    double mass, friction, restitution;
    int shape;
    Vec3<double> shapeSize;
    Iso<double> shapeOffset;
    std::vector<int> children;
  };
  int nb = 0, nd = 0;
  std::vector<Body> b;
  Vec3<double> gravity;
  double dt = 1e-3;
  double clipDepth = 0.03;
  std::vector<double> posLo, posHi, velLo, velHi, frcLo, frcHi;
  std::vector<int> dofBody;  // body owning each dof
};

inline Iso<double> fromC(const ddb_transform& t) {
  Iso<double> r;
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) r.R.m[i][j] = t.r[3 * i + j]; r.p[i] = t.p[i]; }
  return r;
}
inline int jointDofs(int type) { return type == DDB_JOINT_FREE ? 6 : (type == DDB_JOINT_WELD ? 0 : 1); }

inline Model buildModel(const ddb_topology& t, const ddb_params& p) {
  Model M;
  M.nb = t.n_bodies; M.nd = t.n_dofs;
  M.gravity = v3<double>(p.gravity[0], p.gravity[1], p.gravity[2]);
  M.dt = p.time_step; M.clipDepth = p.contact_clipping_depth;
  M.b.resize(M.nb);
  M.dofBody.assign(M.nd, -1);
  for (int i = 0; i < M.nb; i++) {
    const ddb_body& s = t.bodies[i];
    Model::Body& b = M.b[i];
    b.parent = s.parent; b.type = s.joint_type; b.dof = s.dof_offset; b.ndof = jointDofs(s.joint_type);
    b.skel = s.skeleton; b.gravity = s.gravity_mode != 0;
    b.Tp2j = fromC(s.parent_to_joint); b.Tc2j = fromC(s.child_to_joint); b.Tc2jInv = inverse(b.Tc2j);
    b.axis = v3<double>(s.axis[0], s.axis[1], s.axis[2]);
    b.mass = s.mass; b.friction = s.friction; b.restitution = s.restitution;
    b.shape = s.shape_type; b.shapeSize = v3<double>(s.shape_size[0], s.shape_size[1], s.shape_size[2]);
    b.shapeOffset = fromC(s.shape_offset);
    // Inertia::computeSpatialTensor (Inertia.cpp:394-409)
    Mat3<double> C = skew(v3<double>(s.com[0], s.com[1], s.com[2]));
    Mat3<double> mom;
    mom.m[0][0] = s.moment[0]; mom.m[1][1] = s.moment[1]; mom.m[2][2] = s.moment[2];
    mom.m[0][1] = mom.m[1][0] = s.moment[3]; mom.m[0][2] = mom.m[2][0] = s.moment[4]; mom.m[1][2] = mom.m[2][1] = s.moment[5];
    Mat3<double> CCt = mul(C, transpose(C));
    b.I = zero66<double>();
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        b.I.m[r][c] = mom.m[r][c] + s.mass * CCt.m[r][c];
        b.I.m[r][3 + c] = s.mass * C.m[r][c];
        b.I.m[3 + r][c] = s.mass * C.m[c][r];
        b.I.m[3 + r][3 + c] = r == c ? s.mass : 0.0;
      }
    for (int r = 0; r < 6; r++) for (int k = 0; k < 6; k++) b.S[r][k] = 0;
    if (b.type == DDB_JOINT_REVOLUTE) {  // AdTAngular(T_c2j, axis)
      Vec3<double> w = mul(b.Tc2j.R, b.axis), v = cross(b.Tc2j.p, w);
      for (int r = 0; r < 3; r++) { b.S[r][0] = w[r]; b.S[3 + r][0] = v[r]; }
    } else if (b.type == DDB_JOINT_PRISMATIC) {  // AdTLinear(T_c2j, axis)
      Vec3<double> v = mul(b.Tc2j.R, b.axis);
      for (int r = 0; r < 3; r++) b.S[3 + r][0] = v[r];
    } else if (b.type == DDB_JOINT_FREE) {  // getAdTMatrix(T_c2j)
      Mat6<double> X = AdTMatrix(b.Tc2j);
      for (int r = 0; r < 6; r++) for (int k = 0; k < 6; k++) b.S[r][k] = X.m[r][k];
    }
    if (b.parent >= 0) M.b[b.parent].children.push_back(i);
    for (int k = 0; k < b.ndof; k++) M.dofBody[b.dof + k] = i;
  }
  auto lim = [&](const double* src, double def) {
    std::vector<double> v(M.nd, def);
    if (src) for (int i = 0; i < M.nd; i++) v[i] = src[i];
    return v;
  };
  M.posLo = lim(t.position_lower, -INFINITY); M.posHi = lim(t.position_upper, INFINITY);
  M.velLo = lim(t.velocity_lower, -INFINITY); M.velHi = lim(t.velocity_upper, INFINITY);
  M.frcLo = lim(t.force_lower, -INFINITY); M.frcHi = lim(t.force_upper, INFINITY);
  return M;
}

template <class S> inline Iso<S> cast(const Iso<double>& a) {
  Iso<S> r;
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) r.R.m[i][j] = S(a.R.m[i][j]); r.p[i] = S(a.p[i]); }
  return r;
}
template <class S> inline Mat6<S> cast(const Mat6<double>& a) {
  Mat6<S> r;
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) r.m[i][j] = S(a.m[i][j]);
  return r;
}

// Inverse of a small SPD matrix (stands in for math::inverse<ConfigSpace>:
// closed form for <=4, LDLT for 6; dart/math/detail/ConfigurationSpace.hpp:47-65).
template <class S> inline void spdInverse(int n, const S* A /*n*n row-major*/, S* out) {
  S L[36], D[6];
  for (int j = 0; j < n; j++) {
    S d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k] * D[k];
    D[j] = d;
    for (int i = j + 1; i < n; i++) {
      S s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k] * D[k];
      L[i * n + j] = s / d;
    }
  }
  for (int c = 0; c < n; c++) {
    S y[6];
    for (int i = 0; i < n; i++) { S s = S(i == c); for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k]; y[i] = s; }
    for (int i = 0; i < n; i++) y[i] /= D[i];
    for (int i = n - 1; i >= 0; i--) { S s = y[i]; for (int k = i + 1; k < n; k++) s -= L[k * n + i] * y[k]; y[i] = s; }
    for (int i = 0; i < n; i++) out[i * n + c] = y[i];
  }
}

// Per-world dynamic state in scalar type S.
template <class S> struct Dyn {
  const Model& M;
  std::vector<S> q, qd;
  struct B {
    Iso<S> T, W;            // relative / world transform
    Vec6<S> V, eta;         // spatial velocity, partial acceleration
    Mat6<S> IA;             // articulated inertia
    S AIS[6][6];            // IA * S (columns)
    S psi[36];              // (S^T IA S)^-1, ndof x ndof row-major
    Vec6<S> bias, A;        // ABA bias force, spatial acceleration
    S u[6];                 // total joint force / impulse
    Vec6<S> cgdV, cgF;      // C+g recursion
    Vec6<S> invMc, invMU;   // M^-1 recursions
    S invMa[6], seg[6];
    Vec6<S> impBias, delV;  // impulse recursions
    Vec6<S> consImp;        // constraint impulse on the body (body frame)
    S dqd[6];               // velocity change
  };
  std::vector<B> b;
  bool artDirty = true;

  explicit Dyn(const Model& m) : M(m), q(m.nd, S(0)), qd(m.nd, S(0)), b(m.nb) {}

  Vec6<S> Scol(int i, int k) const { Vec6<S> r; for (int x = 0; x < 6; x++) r[x] = S(M.b[i].S[x][k]); return r; }
  Vec6<S> Sdot(int i, const S* x) const {  // S * x
    Vec6<S> r = zero6<S>();
    for (int k = 0; k < M.b[i].ndof; k++) for (int a = 0; a < 6; a++) r[a] += S(M.b[i].S[a][k]) * x[k];
    return r;
  }

  void setState(const S* q_, const S* qd_) {
    for (int i = 0; i < M.nd; i++) { q[i] = q_[i]; qd[i] = qd_[i]; }
    kinematics();
  }
  void setVelocities(const S* qd_) { for (int i = 0; i < M.nd; i++) qd[i] = qd_[i]; kinematics(); }

  Iso<S> jointTransform(int i) const {
    const Model::Body& mb = M.b[i];
    Iso<S> J = isoIdentity<S>();
    if (mb.type == DDB_JOINT_REVOLUTE) {
      J.R = expAngular(v3<S>(S(mb.axis[0]) * q[mb.dof], S(mb.axis[1]) * q[mb.dof], S(mb.axis[2]) * q[mb.dof]));
    } else if (mb.type == DDB_JOINT_PRISMATIC) {
      J.p = v3<S>(S(mb.axis[0]) * q[mb.dof], S(mb.axis[1]) * q[mb.dof], S(mb.axis[2]) * q[mb.dof]);
    } else if (mb.type == DDB_JOINT_FREE) {
      J.R = expMapRot(v3<S>(q[mb.dof], q[mb.dof + 1], q[mb.dof + 2]));
      J.p = v3<S>(q[mb.dof + 3], q[mb.dof + 4], q[mb.dof + 5]);
    }
    return mul(mul(cast<S>(mb.Tp2j), J), cast<S>(mb.Tc2jInv));
  }

  void kinematics() {
    artDirty = true;
    for (int i = 0; i < M.nb; i++) {
      const Model::Body& mb = M.b[i];
      B& x = b[i];
      x.T = jointTransform(i);
      Vec6<S> Sqd = Sdot(i, &qd[mb.ndof ? mb.dof : 0]);
      if (mb.parent >= 0) {
        x.W = mul(b[mb.parent].W, x.T);
        x.V = AdInvT(x.T, b[mb.parent].V) + Sqd;
      } else {
        x.W = x.T;
        x.V = Sqd;
      }
      x.eta = ad(x.V, Sqd);  // dS = 0 for every joint kind here
    }
  }

  void artInertia() {
    if (!artDirty) return;
    for (int i = M.nb - 1; i >= 0; i--) {
      const Model::Body& mb = M.b[i];
      B& x = b[i];
      x.IA = cast<S>(mb.I);
      for (int c : mb.children) {
        const Model::Body& mc = M.b[c];
        B& y = b[c];
        Mat6<S> PI = y.IA;
        for (int r = 0; r < 6; r++)
          for (int s = 0; s < 6; s++) {
            S acc = 0;
            for (int k = 0; k < mc.ndof; k++) for (int l = 0; l < mc.ndof; l++) acc += y.AIS[r][k] * y.psi[k * mc.ndof + l] * y.AIS[s][l];
            PI.m[r][s] -= acc;
          }
        x.IA = x.IA + transformInertia(inverse(y.T), PI);
      }
      S proj[36];
      for (int k = 0; k < mb.ndof; k++) {
        Vec6<S> col = mul(x.IA, Scol(i, k));
        for (int r = 0; r < 6; r++) x.AIS[r][k] = col[r];
      }
      for (int k = 0; k < mb.ndof; k++)
        for (int l = 0; l < mb.ndof; l++) { S s = 0; for (int r = 0; r < 6; r++) s += S(mb.S[r][k]) * x.AIS[r][l]; proj[k * mb.ndof + l] = s; }
      if (mb.ndof == 1) x.psi[0] = S(1) / proj[0];
      else if (mb.ndof > 1) spdInverse<S>(mb.ndof, proj, x.psi);
    }
    artDirty = false;
  }

  Vec6<S> gravityForce(int i) const {
    if (!M.b[i].gravity) return zero6<S>();
    return mul(cast<S>(M.b[i].I), AdInvRLinear(b[i].W, v3<S>(S(M.gravity[0]), S(M.gravity[1]), S(M.gravity[2]))));
  }

  // Skeleton::computeForwardDynamics
  void forwardDynamics(const S* tau, S* ddq) {
    artInertia();
    for (int i = M.nb - 1; i >= 0; i--) {
      const Model::Body& mb = M.b[i];
      B& x = b[i];
      Mat6<S> I = cast<S>(mb.I);
      x.bias = -dad(x.V, mul(I, x.V)) - gravityForce(i);
      for (int c : mb.children) {
        const Model::Body& mc = M.b[c];
        B& y = b[c];
        S pu[6];
        for (int k = 0; k < mc.ndof; k++) { S s = 0; for (int l = 0; l < mc.ndof; l++) s += y.psi[k * mc.ndof + l] * y.u[l]; pu[k] = s; }
        Vec6<S> beta = y.bias + mul(y.IA, y.eta + Sdot(c, pu));
        x.bias = x.bias + dAdInvT(y.T, beta);
      }
      Vec6<S> bf = mul(x.IA, x.eta) + x.bias;
      for (int k = 0; k < mb.ndof; k++) x.u[k] = tau[mb.dof + k] - dot(Scol(i, k), bf);
    }
    for (int i = 0; i < M.nb; i++) {
      const Model::Body& mb = M.b[i];
      B& x = b[i];
      Vec6<S> Ap = mb.parent >= 0 ? AdInvT(x.T, b[mb.parent].A) : zero6<S>();
      Vec6<S> IAp = mul(x.IA, Ap);
      S r[6], acc[6];
      for (int k = 0; k < mb.ndof; k++) r[k] = x.u[k] - dot(Scol(i, k), IAp);
      for (int k = 0; k < mb.ndof; k++) { S s = 0; for (int l = 0; l < mb.ndof; l++) s += x.psi[k * mb.ndof + l] * r[l]; acc[k] = s; ddq[mb.dof + k] = s; }
      x.A = Ap + Sdot(i, acc) + x.eta;
    }
  }

  // Skeleton::updateCoriolisAndGravityForces
  void coriolisGravity(S* C) {
    for (int i = 0; i < M.nb; i++) {
      const Model::Body& mb = M.b[i];
      b[i].cgdV = mb.parent >= 0 ? AdInvT(b[i].T, b[mb.parent].cgdV) + b[i].eta : b[i].eta;
    }
    for (int i = M.nb - 1; i >= 0; i--) {
      const Model::Body& mb = M.b[i];
      B& x = b[i];
      Mat6<S> I = cast<S>(mb.I);
      x.cgF = mul(I, x.cgdV);
      x.cgF = x.cgF - gravityForce(i);
      x.cgF = x.cgF - dad(x.V, mul(I, x.V));
      for (int c : mb.children) x.cgF = x.cgF + dAdInvT(b[c].T, b[c].cgF);
      for (int k = 0; k < mb.ndof; k++) C[mb.dof + k] = dot(Scol(i, k), x.cgF);
    }
  }

  // One M^-1 pass with generalized force f; if stopAtDof >= 0 the forward sweep
  // stops after the joint owning that dof (Skeleton.cpp:4071-4086).
  void invMassPass(const S* f, S* out, int stopAtDof) {
    artInertia();
    for (int i = M.nb - 1; i >= 0; i--) {
      const Model::Body& mb = M.b[i];
      B& x = b[i];
      x.invMc = zero6<S>();
      for (int c : mb.children) {
        const Model::Body& mc = M.b[c];
        B& y = b[c];
        S pa[6];
        for (int k = 0; k < mc.ndof; k++) { S s = 0; for (int l = 0; l < mc.ndof; l++) s += y.psi[k * mc.ndof + l] * y.invMa[l]; pa[k] = s; }
        Vec6<S> beta = y.invMc;
        for (int r = 0; r < 6; r++) for (int k = 0; k < mc.ndof; k++) beta[r] += y.AIS[r][k] * pa[k];
        x.invMc = x.invMc + dAdInvT(y.T, beta);
      }
      for (int k = 0; k < mb.ndof; k++) x.invMa[k] = f[mb.dof + k] - dot(Scol(i, k), x.invMc);
    }
    for (int i = 0; i < M.nb; i++) {
      const Model::Body& mb = M.b[i];
      B& x = b[i];
      Vec6<S> Up = mb.parent >= 0 ? AdInvT(x.T, b[mb.parent].invMU) : zero6<S>();
      Vec6<S> IUp = mul(x.IA, Up);
      S r[6];
      for (int k = 0; k < mb.ndof; k++) r[k] = x.invMa[k] - dot(Scol(i, k), IUp);
      for (int k = 0; k < mb.ndof; k++) { S s = 0; for (int l = 0; l < mb.ndof; l++) s += x.psi[k * mb.ndof + l] * r[l]; x.seg[k] = s; out[mb.dof + k] = s; }
      x.invMU = Up + Sdot(i, x.seg);
      if (stopAtDof >= 0 && mb.ndof > 0 && mb.dof + mb.ndof > stopAtDof) break;
    }
  }

  // Skeleton::multiplyByImplicitInvMassMatrix
  void mulInvMass(const S* f, S* out) { invMassPass(f, out, -1); }

  // Skeleton::updateInvMassMatrix: upper triangle by columns, then mirrored. Column-major n x n.
  // (Skeletons are handled per tree in the reference; with one flattened body list a
  // unit force on skeleton a produces exact zeros on skeleton b, so the result is the
  // same block-diagonal matrix.)
  void invMassMatrix(S* Minv) {
    int n = M.nd;
    std::vector<S> f(n, S(0)), col(n, S(0));
    for (int i = 0; i < n * n; i++) Minv[i] = S(0);
    for (int j = 0; j < n; j++) {
      f[j] = S(1);
      for (int i = 0; i < n; i++) col[i] = S(0);
      invMassPass(f.data(), col.data(), j);
      for (int i = 0; i < n; i++) Minv[j * n + i] = col[i];
      f[j] = S(0);
    }
    // triangularView<StrictlyLower> = transpose
    for (int c = 0; c < n; c++) for (int r = c + 1; r < n; r++) Minv[c * n + r] = Minv[r * n + c];
  }

  // Impulse dynamics: body-frame constraint impulses in b[i].consImp -> dqd
  // (Skeleton::updateBiasImpulse + updateVelocityChange).
  void impulsePass(S* dqdOut) {
    artInertia();
    for (int i = M.nb - 1; i >= 0; i--) {
      const Model::Body& mb = M.b[i];
      B& x = b[i];
      x.impBias = -x.consImp;
      for (int c : mb.children) {
        const Model::Body& mc = M.b[c];
        B& y = b[c];
        S pu[6];
        for (int k = 0; k < mc.ndof; k++) { S s = 0; for (int l = 0; l < mc.ndof; l++) s += y.psi[k * mc.ndof + l] * y.u[l]; pu[k] = s; }
        Vec6<S> beta = y.impBias;
        for (int r = 0; r < 6; r++) for (int k = 0; k < mc.ndof; k++) beta[r] += y.AIS[r][k] * pu[k];
        x.impBias = x.impBias + dAdInvT(y.T, beta);
      }
      for (int k = 0; k < mb.ndof; k++) x.u[k] = -dot(Scol(i, k), x.impBias);
    }
    for (int i = 0; i < M.nb; i++) {
      const Model::Body& mb = M.b[i];
      B& x = b[i];
      Vec6<S> Vp = mb.parent >= 0 ? AdInvT(x.T, b[mb.parent].delV) : zero6<S>();
      Vec6<S> IVp = mul(x.IA, Vp);
      S r[6];
      for (int k = 0; k < mb.ndof; k++) r[k] = x.u[k] - dot(Scol(i, k), IVp);
      for (int k = 0; k < mb.ndof; k++) { S s = 0; for (int l = 0; l < mb.ndof; l++) s += x.psi[k * mb.ndof + l] * r[l]; x.dqd[k] = s; dqdOut[mb.dof + k] = s; }
      x.delV = Vp + Sdot(i, x.dqd);
    }
  }
  void clearImpulses() { for (auto& x : b) x.consImp = zero6<S>(); }

  // Skeleton::integratePositionsExplicit (Skeleton.cpp:2369-2382)
  void integratePositionsExplicit(const S* pos, const S* vel, S dt, S* out) const {
    for (int i = 0; i < M.nb; i++) {
      const Model::Body& mb = M.b[i];
      if (mb.type == DDB_JOINT_FREE) {
        Iso<S> Q{expMapRot(v3<S>(pos[mb.dof], pos[mb.dof + 1], pos[mb.dof + 2])), v3<S>(pos[mb.dof + 3], pos[mb.dof + 4], pos[mb.dof + 5])};
        Iso<S> D{expMapRot(v3<S>(vel[mb.dof] * dt, vel[mb.dof + 1] * dt, vel[mb.dof + 2] * dt)),
                 v3<S>(vel[mb.dof + 3] * dt, vel[mb.dof + 4] * dt, vel[mb.dof + 5] * dt)};
        Iso<S> Qn = mul(Q, D);
        Vec3<S> w = logMapRot(Qn.R);
        for (int k = 0; k < 3; k++) { out[mb.dof + k] = w[k]; out[mb.dof + 3 + k] = Qn.p[k]; }
      } else {
        for (int k = 0; k < mb.ndof; k++) out[mb.dof + k] = pos[mb.dof + k] + vel[mb.dof + k] * dt;
      }
    }
  }
};

}  // namespace ddo
