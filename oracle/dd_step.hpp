// TEST INFRASTRUCTURE ONLY — see dd_spatial.hpp.
//
// dd_step.hpp — CPU restatement of one contact-free differentiable timestep:
//   World::step                          dart/simulation/World.cpp:206-279
//   BackpropSnapshot Jacobians           dart/neural/BackpropSnapshot.cpp:432-735,1002-1076,1237-1360
//     force-vel = dt M^-1 (:464-467), vel-vel = I - forceVel * velC (:611-615),
//     pos-vel = dM + M^-1(-dt dC) (:1066-1070), pos-pos / vel-pos from
//     World::getPosPosJacobian / getVelPosJacobian (World.cpp:1469-1500)
//   the reference's finite-difference recipes inside those "analytic" Jacobians:
//     d(M^-1 z)/dq  central, eps 5e-7   BackpropSnapshot.cpp:5990-6027
//     dC/dq         central, eps 1e-7   BackpropSnapshot.cpp:6121-6157
//     dC/dqd        central, eps 1e-6   Skeleton.cpp:1912-1935
//     FreeJoint pos-pos eps 1e-6, vel-pos eps 1e-7   FreeJoint.cpp:604-653
//   BackpropSnapshot::backprop + clipLossGradientsToBounds   BackpropSnapshot.cpp:115-186,374-428
// "Reference mode" reproduces those recipes in double; "exact mode" differentiates
// the step in long double with Richardson-extrapolated central differences and is
// used to check that the GPU's closed-form derivatives are right to ~1e-11.
#pragma once
#include "dd_dynamics.hpp"

namespace ddo {

// q' , qd' of one step without contacts. Also returns qdd.
template <class S> inline void stepNoContact(const Model& M, const S* q, const S* qd, const S* tau, S* qNext, S* qdNext, S* qdd = nullptr) {
  Dyn<S> d(M);
  d.setState(q, qd);
  std::vector<S> acc(M.nd);
  d.forwardDynamics(tau, acc.data());
  // Skeleton::integrateVelocities: qd += dt*qdd (GenericJoint.hpp:1404-1409)
  for (int i = 0; i < M.nd; i++) qdNext[i] = qd[i] + acc[i] * S(M.dt);
  // explicit position update with the pre-step velocity (World.cpp:262-274)
  d.integratePositionsExplicit(q, qd, S(M.dt), qNext);
  if (qdd) for (int i = 0; i < M.nd; i++) qdd[i] = acc[i];
}

struct Jacobians {
  int n = 0;
  std::vector<double> posPos, posVel, velPos, velVel, forceVel;  // n x n column-major
  void resize(int n_) {
    n = n_;
    posPos.assign(n * n, 0); posVel.assign(n * n, 0); velPos.assign(n * n, 0); velVel.assign(n * n, 0); forceVel.assign(n * n, 0);
  }
};

inline void matmulCM(int n, const double* A, const double* B, double* C) {  // C = A*B, col-major
  for (int c = 0; c < n; c++)
    for (int r = 0; r < n; r++) { double s = 0; for (int k = 0; k < n; k++) s += A[k * n + r] * B[c * n + k]; C[c * n + r] = s; }
}

// Reference recipe (double precision, the reference's own eps and evaluation order).
inline void jacobiansReferenceNoContact(const Model& M, const double* q, const double* qd, const double* tau, Jacobians& J) {
  int n = M.nd;
  J.resize(n);
  double dt = M.dt;
  Dyn<double> d(M);
  d.setState(q, qd);
  std::vector<double> Minv(n * n), C(n);
  d.invMassMatrix(Minv.data());
  d.coriolisGravity(C.data());
  // force-vel
  for (int i = 0; i < n * n; i++) J.forceVel[i] = dt * Minv[i];
  // velC: Skeleton::finiteDifferenceVelCJacobian, EPS = 1e-6
  std::vector<double> velC(n * n), tw(n), cp(n), cn(n);
  {
    const double EPS = 1e-6;
    for (int i = 0; i < n; i++) {
      for (int k = 0; k < n; k++) tw[k] = qd[k];
      tw[i] += EPS;
      d.setVelocities(tw.data()); d.coriolisGravity(cp.data());
      for (int k = 0; k < n; k++) tw[k] = qd[k];
      tw[i] -= EPS;
      d.setVelocities(tw.data()); d.coriolisGravity(cn.data());
      for (int r = 0; r < n; r++) velC[i * n + r] = (cp[r] - cn[r]) / (2 * EPS);
    }
    d.setVelocities(qd);
  }
  // vel-vel = I - forceVel * velC
  std::vector<double> tmp(n * n);
  matmulCM(n, J.forceVel.data(), velC.data(), tmp.data());
  for (int c = 0; c < n; c++) for (int r = 0; r < n; r++) J.velVel[c * n + r] = (r == c ? 1.0 : 0.0) - tmp[c * n + r];
  // pos-vel = dM + Minv * (-dt * dC)
  std::vector<double> z(n), dM(n * n), dC(n * n), vp(n), vm(n);
  for (int i = 0; i < n; i++) z[i] = dt * (tau[i] - C[i]);
  {
    const double EPS = 5e-7;
    for (int i = 0; i < n; i++) {
      for (int k = 0; k < n; k++) tw[k] = q[k];
      tw[i] += EPS;
      d.setState(tw.data(), qd); d.mulInvMass(z.data(), vp.data());
      for (int k = 0; k < n; k++) tw[k] = q[k];
      tw[i] -= EPS;
      d.setState(tw.data(), qd); d.mulInvMass(z.data(), vm.data());
      for (int r = 0; r < n; r++) dM[i * n + r] = (vp[r] - vm[r]) / (2 * EPS);
    }
  }
  {
    const double EPS = 1e-7;
    for (int i = 0; i < n; i++) {
      for (int k = 0; k < n; k++) tw[k] = q[k];
      tw[i] += EPS;
      d.setState(tw.data(), qd); d.coriolisGravity(cp.data());
      for (int k = 0; k < n; k++) tw[k] = q[k];
      tw[i] -= EPS;
      d.setState(tw.data(), qd); d.coriolisGravity(cn.data());
      for (int r = 0; r < n; r++) dC[i * n + r] = (cp[r] - cn[r]) / (2 * EPS);
    }
    d.setState(q, qd);
  }
  for (int i = 0; i < n * n; i++) dC[i] *= -dt;
  matmulCM(n, Minv.data(), dC.data(), tmp.data());
  for (int i = 0; i < n * n; i++) J.posVel[i] = dM[i] + tmp[i];
  // pos-pos / vel-pos: identity and dt*I for R^1 joints (GenericJoint.hpp:1426-1445),
  // central differences of integratePositionsExplicit for FreeJoint.
  for (int i = 0; i < n; i++) { J.posPos[i * n + i] = 1.0; J.velPos[i * n + i] = dt; }
  for (int bi = 0; bi < M.nb; bi++) {
    const Model::Body& mb = M.b[bi];
    if (mb.type != DDB_JOINT_FREE) continue;
    std::vector<double> p(n), v(n), op(n), om(n);
    for (int pass = 0; pass < 2; pass++) {
      const double EPS = pass == 0 ? 1e-6 : 1e-7;
      for (int i = 0; i < 6; i++) {
        for (int k = 0; k < n; k++) { p[k] = q[k]; v[k] = qd[k]; }
        (pass == 0 ? p : v)[mb.dof + i] += EPS;
        d.integratePositionsExplicit(p.data(), v.data(), dt, op.data());
        for (int k = 0; k < n; k++) { p[k] = q[k]; v[k] = qd[k]; }
        (pass == 0 ? p : v)[mb.dof + i] -= EPS;
        d.integratePositionsExplicit(p.data(), v.data(), dt, om.data());
        std::vector<double>& out = pass == 0 ? J.posPos : J.velPos;
        for (int r = 0; r < 6; r++) out[(mb.dof + i) * n + mb.dof + r] = (op[mb.dof + r] - om[mb.dof + r]) / (2 * EPS);
      }
    }
  }
}

// Exact mode: Richardson-extrapolated central differences of the whole step in long double.
inline void jacobiansExactNoContact(const Model& M, const double* q, const double* qd, const double* tau, Jacobians& J) {
  typedef long double L;
  int n = M.nd;
  J.resize(n);
  std::vector<L> x(3 * n), xp(3 * n);
  for (int i = 0; i < n; i++) { x[i] = q[i]; x[n + i] = qd[i]; x[2 * n + i] = tau[i]; }
  auto eval = [&](const std::vector<L>& in, std::vector<L>& out) {
    out.resize(2 * n);
    stepNoContact<L>(M, &in[0], &in[n], &in[2 * n], &out[0], &out[n]);
  };
  std::vector<L> a, b, c, d2;
  for (int i = 0; i < 3 * n; i++) {
    const L h = 2e-5L;
    xp = x; xp[i] = x[i] + h; eval(xp, a);
    xp = x; xp[i] = x[i] - h; eval(xp, b);
    xp = x; xp[i] = x[i] + h / 2; eval(xp, c);
    xp = x; xp[i] = x[i] - h / 2; eval(xp, d2);
    for (int r = 0; r < 2 * n; r++) {
      L D1 = (a[r] - b[r]) / (2 * h), D2 = (c[r] - d2[r]) / h;
      double v = (double)((4 * D2 - D1) / 3);
      int blk = i / n, col = i % n;
      if (blk == 0) (r < n ? J.posPos : J.posVel)[col * n + (r % n)] = v;
      else if (blk == 1) (r < n ? J.velPos : J.velVel)[col * n + (r % n)] = v;
      else if (r >= n) J.forceVel[col * n + (r - n)] = v;
    }
  }
}

// BackpropSnapshot::backprop (:115-186) with clipLossGradientsToBounds (:374-428).
inline void backprop(const Model& M, const Jacobians& J, const double* q, const double* qd, const double* tau,
                     const double* gqN, const double* gqdN, double* gq, double* gqd, double* gtau) {
  int n = M.nd;
  for (int c = 0; c < n; c++) {
    double a = 0, b = 0, t = 0;
    for (int r = 0; r < n; r++) {
      a += J.posPos[c * n + r] * gqN[r] + J.posVel[c * n + r] * gqdN[r];
      b += J.velPos[c * n + r] * gqN[r] + J.velVel[c * n + r] * gqdN[r];
      t += J.forceVel[c * n + r] * gqdN[r];
    }
    gq[c] = a; gqd[c] = b; gtau[c] = t;
  }
  for (int i = 0; i < n; i++) {
    if (q[i] == M.posLo[i] && gq[i] > 0) gq[i] = 0;
    if (q[i] == M.posHi[i] && gq[i] < 0) gq[i] = 0;
    if (qd[i] == M.velLo[i] && gqd[i] > 0) gqd[i] = 0;
    if (qd[i] == M.velHi[i] && gqd[i] < 0) gqd[i] = 0;
    if (tau[i] == M.frcLo[i] && gtau[i] > 0) gtau[i] = 0;
    if (tau[i] == M.frcHi[i] && gtau[i] < 0) gtau[i] = 0;
  }
}

}  // namespace ddo
