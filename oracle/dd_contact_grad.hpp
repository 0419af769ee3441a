// TEST INFRASTRUCTURE ONLY — see dd_spatial.hpp.
//
// dd_contact_grad.hpp — CPU restatement of the backward pass through a timestep WITH contacts:
//   world screw axes                Joint::getWorldAxisScrewForPosition / ForVelocity (dynamics/Joint.cpp:389-410),
//                                   FreeJoint::getWorldAxisScrewAt (dynamics/FreeJoint.cpp:703-733)
//   contact geometry gradients      DifferentiableContactConstraint::{getDofContactType, getContactPositionGradient,
//                                   getContactNormalGradient, getContactForceGradient, getContactWorldForceGradient,
//                                   getScrewAxisForForceGradient, getConstraintForces, getConstraintForcesJacobian, getForceMultiple}
//                                   (neural/DifferentiableContactConstraint.cpp:113-267,307-568,573-1107,1205-1293,1433-1576,2881-2908);
//                                   math::gradientWrtTheta / gradientWrtThetaPureRotation / getContactPointGradient
//                                   (math/Geometry.cpp:727-760,855-960); ContactConstraint::getTangentBasisMatrixODEGradient
//                                   (constraint/ContactConstraint.cpp:783-858)
//   world-level matrices            A_c, A_ub, E, f_c, b_c from the per-group ConstraintMapping
//                                   (neural/ConstrainedGroupGradientMatrices.cpp:697-830)
//   Jacobians                       BackpropSnapshot::getVelJacobianWrt (:1002-1076), getJacobianOfConstraintForce (:3853-3894),
//                                   getJacobianOfLCPConstraintMatrixClampingSubset (:3898-4055, incl. the pseudo-inverse
//                                   derivative), getJacobianOfLCPOffsetClampingSubset (:4252-4299),
//                                   getJacobianOfClampingConstraints(Transpose) / UpperBound (:5005-5091), with the reference's
//                                   finite-difference recipes for d(M^-1 x)/dq (eps 5e-7, :5990-6027) and dC (eps 1e-7, :6121-6157)
// The derivative of a FreeJoint's force screw with respect to the joint's own coordinates is taken as ad(position screw,
// force screw), the closed form of what FreeJoint::getScrewAxisGradientForForce (FreeJoint.cpp:779-814) evaluates.
// Restitution enters through getBounceDiagonals (1 + e on clamping rows that bounced) in the LCP offset b_c and its Jacobians;
// BackpropSnapshot::getBounceApproximationJacobian is a stand-alone getter the backprop path does not call (pos-pos is the joints' own).
#pragma once
#include "dd_contact.hpp"
#include "dd_step.hpp"
#include <cstdio>
#include <cstdlib>

namespace ddo {

struct Screws {
  std::vector<Vec6<double>> force, pos;  // world screw per dof: for forces / velocities, and for positions
};

inline Screws worldScrews(const Model& M, const Dyn<double>& d) {
  Screws s;
  s.force.resize(M.nd); s.pos.resize(M.nd);
  for (int i = 0; i < M.nb; i++) {
    const Model::Body& mb = M.b[i];
    for (int k = 0; k < mb.ndof; k++) {
      s.force[mb.dof + k] = AdT(d.b[i].W, d.Scol(i, k));
      s.pos[mb.dof + k] = s.force[mb.dof + k];
    }
    if (mb.type == DDB_JOINT_FREE) {
      const Iso<double> Wp = mb.parent >= 0 ? d.b[mb.parent].W : isoIdentity<double>();
      const Mat3<double> J = expMapJac(v3<double>(d.q[mb.dof], d.q[mb.dof + 1], d.q[mb.dof + 2]));
      Iso<double> tr = isoIdentity<double>();
      tr.p = v3<double>(d.q[mb.dof + 3], d.q[mb.dof + 4], d.q[mb.dof + 5]);
      for (int k = 0; k < 6; k++) {
        Vec6<double> g = zero6<double>();
        if (k < 3) { for (int r = 0; r < 3; r++) g[r] = J.m[r][k]; g = AdT(tr, g); }
        else g[k] = 1.0;
        s.pos[mb.dof + k] = AdT(Wp, AdT(mb.Tp2j, g));
      }
    }
  }
  return s;
}

// DegreeOfFreedom::isParentOf(BodyNode*): the dof's joint lies on the path from the body to its root
inline bool dofMovesBody(const Model& M, int dof, int body) {
  const int jb = M.dofBody[dof];
  for (int b = body; b >= 0; b = M.b[b].parent) if (b == jb) return true;
  return false;
}
// DegreeOfFreedom::isParentOf(DegreeOfFreedom*) as getConstraintForcesJacobian's loop uses it: every dof of an ancestor joint,
// and every dof of the same joint when that joint is a FreeJoint (DifferentiableContactConstraint.cpp:1501-1541)
inline bool dofRotatesScrew(const Model& M, int rotateDof, int screwDof) {
  const int jr = M.dofBody[rotateDof], js = M.dofBody[screwDof];
  if (jr == js) return M.b[js].type == DDB_JOINT_FREE;
  for (int b = M.b[js].parent; b >= 0; b = M.b[b].parent) if (b == jr) return true;
  return false;
}

enum DofRole { R_NONE, R_VERTEX, R_FACE, R_EDGE_A, R_EDGE_B, R_SELF, R_SPHERE_A, R_SPHERE_B, R_SPHERE_TO_BOX, R_BOX_TO_SPHERE };

inline DofRole dofRole(const Model& M, const Constraint& c, int dof) {
  const bool pa = dofMovesBody(M, dof, c.bodyA), pb = dofMovesBody(M, dof, c.bodyB);
  if (pa && pb) return R_SELF;
  if (!pa && !pb) return R_NONE;
  switch (c.contact.type) {
    case 1: return pa ? R_VERTEX : R_FACE;             // VERTEX_FACE
    case 2: return pa ? R_FACE : R_VERTEX;             // FACE_VERTEX
    case 3: return pa ? R_EDGE_A : R_EDGE_B;           // EDGE_EDGE
    case 4: return pa ? R_SPHERE_TO_BOX : R_BOX_TO_SPHERE;  // SPHERE_BOX
    case 5: return pa ? R_BOX_TO_SPHERE : R_SPHERE_TO_BOX;  // BOX_SPHERE
    case 6: return pa ? R_SPHERE_A : R_SPHERE_B;       // SPHERE_SPHERE
    default: return R_NONE;
  }
}

inline Vec3<double> thetaGrad(const Vec6<double>& s, const Vec3<double>& p) {  // math::gradientWrtTheta(s, p, 0)
  const Vec3<double> w = ang(s), v = lin(s);
  if (norm(w) > 1e-6) {  // DART_EPSILON
    const Vec3<double> wv = cross(w, v), wwv = cross(w, wv);
    return (cross(w, p) - wwv) + v + wwv;
  }
  return v;
}

inline Vec3<double> edgePointGradient(const Vec3<double>& pa, const Vec3<double>& dpa, const Vec3<double>& ua, const Vec3<double>& dua,
                                      const Vec3<double>& pb, const Vec3<double>& dpb, const Vec3<double>& ub, const Vec3<double>& dub) {
  // math::getContactPointGradient with radiusA = radiusB = 1
  const Vec3<double> p = pb - pa, dp = dpb - dpa;
  const double uaub = dot(ua, ub), d_uaub = dot(dua, ub) + dot(ua, dub);
  const double q1 = dot(ua, p), d_q1 = dot(dua, p) + dot(ua, dp);
  const double q2 = -dot(ub, p), d_q2 = -dot(dub, p) - dot(ub, dp);
  const double d = 1 - uaub * uaub, d_d = -2 * d_uaub * uaub;
  if (d <= 0) return 0.5 * (dpa + dpb);
  const double e = 1.0 / d, d_e = -(1.0 / (d * d)) * d_d;
  const double alpha = (q1 + uaub * q2) * e, d_alpha = (q1 + uaub * q2) * d_e + (d_q1 + d_uaub * q2 + uaub * d_q2) * e;
  const double beta = (uaub * q1 + q2) * e, d_beta = (uaub * q1 + q2) * d_e + (d_uaub * q1 + uaub * d_q1 + d_q2) * e;
  return 0.5 * ((dpa + alpha * dua + d_alpha * ua) + (dpb + beta * dub + d_beta * ub));
}

struct ContactMeta {
  Vec3<double> eAFixed, eADir, eBClosest, eBFixed, eBDir;
  explicit ContactMeta(const Contact& c) {
    for (int i = 0; i < 3; i++) { eAFixed[i] = c.meta[3 + i]; eADir[i] = c.meta[6 + i]; eBClosest[i] = c.meta[9 + i]; eBFixed[i] = c.meta[12 + i]; eBDir[i] = c.meta[15 + i]; }
  }
};

inline Vec3<double> projectLockedFaces(const Contact& c, Vec3<double> g) {
  const ContactMeta m(c);
  const int locks = (int)c.meta[7];
  const Vec3<double> fn[3] = {m.eBClosest, m.eBFixed, m.eBDir};
  for (int k = 0; k < 3; k++) if (locks & (1 << k)) g = g - dot(fn[k], g) * fn[k];
  return g;
}

inline Vec3<double> contactPositionGradient(const Constraint& c, DofRole role, const Vec6<double>& sp) {
  const Contact& ct = c.contact;
  const ContactMeta m(ct);
  switch (role) {
    case R_VERTEX: case R_SELF: return thetaGrad(sp, ct.point);
    case R_EDGE_A: return edgePointGradient(m.eAFixed, thetaGrad(sp, m.eAFixed), m.eADir, cross(ang(sp), m.eADir), m.eBFixed, zero3<double>(), m.eBDir, zero3<double>());
    case R_EDGE_B: return edgePointGradient(m.eAFixed, zero3<double>(), m.eADir, zero3<double>(), m.eBFixed, thetaGrad(sp, m.eBFixed), m.eBDir, cross(ang(sp), m.eBDir));
    case R_SPHERE_A: { const double rA = ct.meta[6], rB = ct.meta[15]; return (rB / (rA + rB)) * thetaGrad(sp, m.eAFixed); }
    case R_SPHERE_B: { const double rA = ct.meta[6], rB = ct.meta[15]; return (rA / (rA + rB)) * thetaGrad(sp, m.eBFixed); }
    case R_SPHERE_TO_BOX: return projectLockedFaces(ct, thetaGrad(sp, m.eAFixed));
    case R_BOX_TO_SPHERE: return thetaGrad(sp, ct.point) + projectLockedFaces(ct, -thetaGrad(sp, m.eAFixed));
    default: return zero3<double>();
  }
}

inline Vec3<double> contactNormalGradient(const Constraint& c, DofRole role, const Vec6<double>& sp) {
  const Contact& ct = c.contact;
  const ContactMeta m(ct);
  const Vec3<double> n = ct.normal, w = ang(sp);
  switch (role) {
    case R_FACE: case R_SELF: return cross(w, n);
    case R_EDGE_A: { const double sign = dot(cross(m.eBDir, m.eADir), n) < 0 ? -1.0 : 1.0; return sign * cross(m.eBDir, cross(w, m.eADir)); }
    case R_EDGE_B: { const double sign = dot(cross(m.eBDir, m.eADir), n) < 0 ? -1.0 : 1.0; return sign * cross(cross(w, m.eBDir), m.eADir); }
    case R_SPHERE_A: case R_SPHERE_B: {
      const double nn = norm(m.eAFixed - m.eBFixed);
      Vec3<double> g = (1.0 / nn) * thetaGrad(sp, role == R_SPHERE_A ? m.eAFixed : m.eBFixed);
      g = g - dot(n, g) * n;
      return role == R_SPHERE_A ? g : -g;
    }
    case R_SPHERE_TO_BOX: case R_BOX_TO_SPHERE: {
      const double nn = norm(m.eAFixed - ct.point);
      Vec3<double> cp = contactPositionGradient(c, role, sp), sg = role == R_SPHERE_TO_BOX ? thetaGrad(sp, m.eAFixed) : zero3<double>();
      if (nn > 1e-5) { cp = (1.0 / nn) * cp; sg = (1.0 / nn) * sg; }
      Vec3<double> t = ct.type == 5 ? cp - sg : sg - cp;  // BOX_SPHERE: normal = point - centre; SPHERE_BOX: centre - point
      return t - dot(t, n) * n;
    }
    default: return zero3<double>();
  }
}

// ContactConstraint::getTangentBasisMatrixODEGradient, column `col`
inline Vec3<double> tangentBasisGradient(const Vec3<double>& n, const Vec3<double>& g, int col) {
  Vec3<double> cr = v3<double>(0, 0, 1), t = cross(cr, n);
  if (dot(t, t) < 1e-12) {
    cr = v3<double>(1, 0, 0); t = cross(cr, n);
    if (dot(t, t) < 1e-12) {
      cr = v3<double>(0, 1, 0); t = cross(cr, n);
      if (dot(t, t) < 1e-12) { cr = v3<double>(0, 0, 1); t = cross(cr, n); }
    }
  }
  const double tn = norm(t);
  t = (1.0 / tn) * t;
  Vec3<double> gc = (1.0 / tn) * cross(cr, g), gt;
  if (std::abs(tn - 1.0) > 1e-6) gt = gc - dot(gc, t) * t; else gt = gc;
  if (col == 0) return gt;
  return cross(g, t) + cross(n, gt);
}

// one LCP row of the world: constraint, direction index, its world force [p x d; d]
struct GradRow {
  int cons, k;
  Vec3<double> dir;
  Vec6<double> F;
};

inline Vec3<double> rowDirection(const Constraint& c, int k) {
  if (k == 0) return c.contact.normal;
  Vec3<double> t1, t2;
  tangentBasis(c.contact.normal, t1, t2);
  return k == 1 ? t1 : t2;
}

inline Vec6<double> worldForceGradient(const Model& M, const Constraint& c, const GradRow& row, int wrt, const Vec6<double>& sp) {
  const DofRole role = dofRole(M, c, wrt);
  if (role == R_NONE) return zero6<double>();
  Vec3<double> fg = zero3<double>();
  if (role != R_VERTEX) {
    const Vec3<double> ng = contactNormalGradient(c, role, sp);
    if (row.k == 0 || dot(ng, ng) <= 1e-12) fg = ng;
    else fg = tangentBasisGradient(c.contact.normal, ng, row.k - 1);
  }
  const Vec3<double> pg = contactPositionGradient(c, role, sp);
  return join(cross(c.contact.point, fg) + cross(pg, row.dir), fg);
}

inline double forceMultiple(const Model& M, const Constraint& c, int dof) {
  const bool pa = dofMovesBody(M, dof, c.bodyA), pb = dofMovesBody(M, dof, c.bodyB);
  if (pa && pb) return 0.0;
  return pa ? 1.0 : (pb ? -1.0 : 0.0);
}

// getConstraintForces: column of A (nd)
inline std::vector<double> constraintForces(const Model& M, const Screws& sc, const Constraint& c, const GradRow& row) {
  std::vector<double> t(M.nd, 0.0);
  for (int i = 0; i < M.nd; i++) { const double m = forceMultiple(M, c, i); if (m != 0) t[i] = dot(sc.force[i], row.F) * m; }
  return t;
}
// getConstraintForcesJacobian: nd x nd column-major, J(row i, col wrt)
inline std::vector<double> constraintForcesJacobian(const Model& M, const Screws& sc, const Constraint& c, const GradRow& row) {
  const int n = M.nd;
  std::vector<double> J((size_t)n * n, 0.0);
  std::vector<Vec6<double>> dF(n);
  for (int w = 0; w < n; w++) dF[w] = worldForceGradient(M, c, row, w, sc.pos[w]);
  for (int i = 0; i < n; i++) {
    const double m = forceMultiple(M, c, i);
    if (m == 0) continue;
    for (int w = 0; w < n; w++) {
      double v = dot(sc.force[i], dF[w]);
      if (dofRotatesScrew(M, w, i)) v += dot(ad(sc.pos[w], sc.force[i]), row.F);
      J[(size_t)w * n + i] = m * v;
    }
  }
  return J;
}

// ---- small dense helpers (column-major n x m unless stated) ----
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return a[(size_t)j * r + i]; }
};
inline Mat mm(const Mat& A, const Mat& B) {
  Mat C(A.r, B.c);
  for (int j = 0; j < B.c; j++) for (int k = 0; k < A.c; k++) { const double b = B(k, j); if (b == 0) continue; for (int i = 0; i < A.r; i++) C(i, j) += A(i, k) * b; }
  return C;
}
inline Mat tr(const Mat& A) { Mat T(A.c, A.r); for (int i = 0; i < A.r; i++) for (int j = 0; j < A.c; j++) T(j, i) = A(i, j); return T; }
inline Mat add(const Mat& A, const Mat& B, double s = 1.0) { Mat C = A; for (size_t i = 0; i < C.a.size(); i++) C.a[i] += s * B.a[i]; return C; }
inline Mat scale(const Mat& A, double s) { Mat C = A; for (double& x : C.a) x *= s; return C; }
inline Mat eye(int n) { Mat I(n, n); for (int i = 0; i < n; i++) I(i, i) = 1; return I; }
inline Mat colvec(const std::vector<double>& v) { Mat C((int)v.size(), 1); C.a = v; return C; }
// Eigen's completeOrthogonalDecomposition().solve(B) / pseudoInverse()
inline Mat codSolveMat(const Mat& Q, const Mat& B) {
  std::vector<double> rm((size_t)Q.r * Q.c);
  for (int i = 0; i < Q.r; i++) for (int j = 0; j < Q.c; j++) rm[(size_t)i * Q.c + j] = Q(i, j);
  Mat X(Q.c, B.c);
  for (int j = 0; j < B.c; j++) {
    std::vector<double> b(B.r);
    for (int i = 0; i < B.r; i++) b[i] = B(i, j);
    std::vector<double> x = codSolve(Q.r, Q.c, rm, b);
    for (int i = 0; i < Q.c; i++) X(i, j) = x[i];
  }
  return X;
}

// Conditioning probe: relative noise (a few ulps) applied to Q = A_c^T M^-1 A before it is decomposed. Where the numerical
// rank of Q sits on the edge of the decomposition's threshold, the reference's pseudo-inverse formulas amplify rounding
// differences by ||Q^+||^2; the tests measure that amplification with this knob and widen their tolerance by it.
inline double& qNoise() { static double v = 0.0; return v; }

struct ContactGradInfo {  // what the test compares / reports
  int numClamping = 0, numUpperBound = 0;
  bool pseudoInverseBranch = false;
  double fcVsLeastSquares = 0.0;  // max |f_c - Q^+ b|: 0 when the multipliers were standardised
};

// BackpropSnapshot Jacobians of one step with contacts, from the result of stepWithContacts(gradients = true).
inline void jacobiansReferenceContact(const Model& M, const double* q, const double* qd, const double* tau, const ContactStepResult& R,
                                      Jacobians& J, ContactGradInfo* info = nullptr) {
  const int n = M.nd;
  const double dt = M.dt;
  // world-level row classification
  std::vector<GradRow> clamp, ub;
  std::vector<double> fc, bc, bounce;  // bounce: 1 + restitution coefficient of a clamping row that bounced
  std::vector<int> ubNormalClamp;  // clamping index of the normal row an upper-bound row points at
  std::vector<double> ubE;
  for (const GroupLCP& g : R.groups) {
    std::vector<int> clampIdxOfRow(g.n, -1);
    for (int pass = 0; pass < 2; pass++)
      for (size_t i = 0; i < g.cons.size(); i++) {
        const Constraint& c = R.constraints[g.cons[i]];
        for (int k = 0; k < c.dim; k++) {
          const int r = g.offset[i] + k;
          GradRow row{g.cons[i], k, rowDirection(c, k), zero6<double>()};
          row.F = join(cross(c.contact.point, row.dir), row.dir);
          if (pass == 0 && g.mapping[r] == -1) {
            clampIdxOfRow[r] = (int)clamp.size();
            clamp.push_back(row); fc.push_back(g.x[r]); bc.push_back(g.b[r]);
            bounce.push_back(1.0 + ((size_t)r < g.restitution.size() ? g.restitution[r] : 0.0));  // getBounceDiagonals (:753)
          } else if (pass == 1 && g.mapping[r] >= 0) {
            const int fp = g.mapping[r];
            const double upper = g.x[fp] * g.hi[r], lower = g.x[fp] * g.lo[r];
            ub.push_back(row);
            ubNormalClamp.push_back(clampIdxOfRow[fp]);
            ubE.push_back(std::abs(g.x[r] - upper) < std::abs(g.x[r] - lower) ? g.hi[r] : g.lo[r]);
          }
        }
      }
  }
  const int nc = (int)clamp.size(), nu = (int)ub.size();
  if (info) { info->numClamping = nc; info->numUpperBound = nu; info->pseudoInverseBranch = false; }
  if (nc == 0) { jacobiansReferenceNoContact(M, q, qd, tau, J); return; }
  J.resize(n);
  Dyn<double> d(M);
  d.setState(q, qd);
  const Screws sc = worldScrews(M, d);
  Mat Minv(n, n);
  d.invMassMatrix(Minv.a.data());
  std::vector<double> C(n);
  d.coriolisGravity(C.data());
  Mat Ac(n, nc), Aub(n, nu), E(nu, nc);
  std::vector<std::vector<double>> Jc(nc), Ju(nu);
  for (int j = 0; j < nc; j++) {
    const Constraint& c = R.constraints[clamp[j].cons];
    std::vector<double> col = constraintForces(M, sc, c, clamp[j]);
    for (int i = 0; i < n; i++) Ac(i, j) = col[i];
    Jc[j] = constraintForcesJacobian(M, sc, c, clamp[j]);
  }
  for (int j = 0; j < nu; j++) {
    const Constraint& c = R.constraints[ub[j].cons];
    std::vector<double> col = constraintForces(M, sc, c, ub[j]);
    for (int i = 0; i < n; i++) Aub(i, j) = col[i];
    Ju[j] = constraintForcesJacobian(M, sc, c, ub[j]);
    E(j, ubNormalClamp[j]) = ubE[j];
  }
  const Mat A = nu ? add(Ac, mm(Aub, E)) : Ac;
  const Mat AcT = tr(Ac);
  Mat Qm = mm(AcT, mm(Minv, A));
  if (qNoise() != 0.0) { unsigned st = 12345u; for (double& x : Qm.a) { st = st * 1664525u + 1013904223u; x *= 1.0 + qNoise() * ((st >> 8) / 16777216.0 - 0.5); } }
  const Mat Q = Qm;
  const Mat I_n = eye(n), I_c = eye(nc);
  // finite-difference helpers (the reference's recipes)
  std::vector<double> tw(n), vp(n), vm(n);
  auto dMinv = [&](const std::vector<double>& x) {  // d(M^-1 x)/dq, eps 5e-7
    Mat out(n, n);
    const double EPS = 5e-7;
    for (int i = 0; i < n; i++) {
      for (int k = 0; k < n; k++) tw[k] = q[k];
      tw[i] += EPS; d.setState(tw.data(), qd); d.mulInvMass(x.data(), vp.data());
      for (int k = 0; k < n; k++) tw[k] = q[k];
      tw[i] -= EPS; d.setState(tw.data(), qd); d.mulInvMass(x.data(), vm.data());
      for (int r = 0; r < n; r++) out(r, i) = (vp[r] - vm[r]) / (2 * EPS);
    }
    d.setState(q, qd);
    return out;
  };
  auto dCof = [&](bool wrtVel) {  // dC/dq or dC/dqd, eps 1e-7
    Mat out(n, n);
    const double EPS = 1e-7;
    for (int i = 0; i < n; i++) {
      for (int s = 0; s < 2; s++) {
        for (int k = 0; k < n; k++) tw[k] = wrtVel ? qd[k] : q[k];
        tw[i] += s == 0 ? EPS : -EPS;
        if (wrtVel) d.setState(q, tw.data()); else d.setState(tw.data(), qd);
        d.coriolisGravity(s == 0 ? vp.data() : vm.data());
      }
      for (int r = 0; r < n; r++) out(r, i) = (vp[r] - vm[r]) / (2 * EPS);
    }
    d.setState(q, qd);
    return out;
  };
  auto toVec = [](const Mat& m) { return m.a; };
  auto matvec = [&](const Mat& m, const std::vector<double>& v) { return mm(m, colvec(v)).a; };
  // sum_j w_j * J_j  (getJacobianOfClampingConstraints / UpperBound) and rows J_j^T v (…Transpose)
  auto dAc = [&](const std::vector<double>& w) { Mat o(n, n); for (int j = 0; j < nc; j++) for (size_t i = 0; i < o.a.size(); i++) o.a[i] += w[j] * Jc[j][i]; return o; };
  auto dAub = [&](const std::vector<double>& w) { Mat o(n, n); for (int j = 0; j < nu; j++) for (size_t i = 0; i < o.a.size(); i++) o.a[i] += w[j] * Ju[j][i]; return o; };
  auto dAcT = [&](const std::vector<double>& v) { Mat o(nc, n); for (int j = 0; j < nc; j++) for (int w = 0; w < n; w++) { double s = 0; for (int i = 0; i < n; i++) s += Jc[j][(size_t)w * n + i] * v[i]; o(j, w) = s; } return o; };
  auto dAubT = [&](const std::vector<double>& v) { Mat o(nu, n); for (int j = 0; j < nu; j++) for (int w = 0; w < n; w++) { double s = 0; for (int i = 0; i < n; i++) s += Ju[j][(size_t)w * n + i] * v[i]; o(j, w) = s; } return o; };

  // getJacobianOfLCPOffsetClampingSubset (BackpropSnapshot.cpp:4252-4299): every variant is getBounceDiagonals().asDiagonal() * -(...)
  auto bounced = [&](Mat m) { for (int j = 0; j < nc; j++) for (int w = 0; w < m.c; w++) m(j, w) *= bounce[j]; return m; };
  // ---- force-vel and vel-vel ----
  {
    const Mat dB = bounced(scale(mm(AcT, Minv), -dt));
    const Mat dF = codSolveMat(Q, dB);
    const Mat fv = mm(Minv, add(mm(A, dF), I_n, dt));
    J.forceVel = fv.a;
  }
  const Mat dCv = dCof(true);
  {
    const Mat dB = bounced(scale(mm(AcT, add(I_n, mm(Minv, dCv), -dt)), -1.0));
    const Mat dF = codSolveMat(Q, dB);
    const Mat vv = add(I_n, mm(Minv, add(mm(A, dF), dCv, -dt)));
    J.velVel = vv.a;
  }
  // ---- pos-vel ----
  {
    const Mat dCq = dCof(false);
    std::vector<double> f(n), z(n), vf = R.preConstraintVel;
    const std::vector<double> Afc = matvec(A, fc);
    for (int i = 0; i < n; i++) { f[i] = tau[i] - C[i]; z[i] = dt * f[i] + Afc[i]; }
    const Mat dM = dMinv(z);
    // dB
    const Mat dB = bounced(scale(add(dAcT(vf), scale(mm(AcT, add(dMinv(f), mm(Minv, dCq), -1.0)), dt)), -1.0));
    // dQ_b
    const Mat P = codSolveMat(Q, I_c);
    const std::vector<double> Eempty;
    auto Ev = [&](const std::vector<double>& v) { return nu ? matvec(E, v) : std::vector<double>(); };
    auto dQ = [&](const std::vector<double>& rhs) {
      const std::vector<double> Ar = matvec(A, rhs);
      Mat inner = add(dMinv(Ar), mm(Minv, nu ? add(dAc(rhs), dAub(Ev(rhs))) : dAc(rhs)));
      return add(dAcT(matvec(Minv, Ar)), mm(AcT, inner));
    };
    auto dQT = [&](const std::vector<double>& rhs) {
      if (!nu) return dQ(rhs);
      const std::vector<double> Acr = matvec(Ac, rhs), MAcr = matvec(Minv, Acr);
      const Mat inner = add(dMinv(Acr), mm(Minv, dAc(rhs)));
      Mat o = add(dAcT(MAcr), mm(AcT, inner));
      return add(o, mm(tr(E), add(dAubT(MAcr), mm(tr(Aub), inner))));
    };
    const std::vector<double> r1 = matvec(P, bc);
    if (info) for (int i = 0; i < nc; i++) info->fcVsLeastSquares = std::max(info->fcVsLeastSquares, std::abs(r1[i] - fc[i]));
    const Mat imprecision = add(I_c, mm(Q, P), -1.0);
    double sq = 0;
    for (double v : imprecision.a) sq += v * v;
    Mat dQb = scale(mm(P, dQ(r1)), -1.0);
    if (!(sq < 1e-18)) {
      if (info) info->pseudoInverseBranch = true;
      const std::vector<double> r2 = matvec(imprecision, bc), r3 = matvec(mm(tr(P), P), bc);
      dQb = add(dQb, mm(mm(P, tr(P)), dQT(r2)));
      dQb = add(dQb, mm(add(I_c, mm(P, Q), -1.0), dQT(r3)));
    }
    const Mat dF = add(dQb, codSolveMat(Q, dB));
    Mat inner = add(mm(A, dF), dAc(fc));
    if (nu) inner = add(inner, dAub(Ev(fc)));
    inner = add(inner, dCq, -dt);
    const Mat pv = add(dM, mm(Minv, inner));
    J.posVel = pv.a;
  }
  // ---- pos-pos / vel-pos: World::getPosPosJacobian / getVelPosJacobian (BackpropSnapshot.cpp:1237-1360) ----
  {
    Jacobians K;
    K.resize(n);
    for (int i = 0; i < n; i++) { K.posPos[i * n + i] = 1.0; K.velPos[i * n + i] = dt; }
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
          std::vector<double>& out = pass == 0 ? K.posPos : K.velPos;
          for (int r = 0; r < 6; r++) out[(mb.dof + i) * n + mb.dof + r] = (op[mb.dof + r] - om[mb.dof + r]) / (2 * EPS);
        }
      }
    }
    J.posPos = K.posPos; J.velPos = K.velPos;
  }
  (void)toVec;
}

// Frozen-active-set velocity map used to check the analytic Jacobians against finite differences that do re-run the
// collision detector: qd' = v_f + M^-1 A f_c with f_c = Q^+ b_c, rows classified as in `ref`; returns false when the
// contact structure (count, bodies, type, row count) differs from `ref`.
inline bool frozenSetVelocity(const Model& M, const double* q, const double* qd, const double* tau, const ContactStepResult& ref,
                              std::vector<double>& qdNext) {
  const int n = M.nd;
  Dyn<double> d(M);
  d.setState(q, qd);
  std::vector<double> acc(n), vel(n);
  d.forwardDynamics(tau, acc.data());
  for (int i = 0; i < n; i++) vel[i] = qd[i] + acc[i] * M.dt;
  d.setVelocities(vel.data());
  d.artInertia();
  std::vector<Contact> contacts = collideWorld(M, d);
  std::vector<Constraint> cons = buildConstraints(M, d, contacts);
  if (cons.size() != ref.constraints.size()) return false;
  {
    // the narrow phase may emit the points of one manifold in another order at a perturbed state: match every reference
    // constraint with the nearest perturbed one of the same bodies / type
    std::vector<Constraint> ordered(cons.size());
    std::vector<char> used(cons.size(), 0);
    for (size_t i = 0; i < cons.size(); i++) {
      const Constraint& r = ref.constraints[i];
      int best = -1;
      double bd = 1e-4;
      for (size_t j = 0; j < cons.size(); j++) {
        if (used[j] || cons[j].bodyA != r.bodyA || cons[j].bodyB != r.bodyB || cons[j].contact.type != r.contact.type || cons[j].dim != r.dim) continue;
        const double dd = norm(cons[j].contact.point - r.contact.point);
        if (dd < bd) { bd = dd; best = (int)j; }
      }
      if (best < 0) return false;
      used[best] = 1;
      ordered[i] = cons[best];
    }
    cons.swap(ordered);
  }
  const Screws sc = worldScrews(M, d);
  std::vector<std::vector<double>> colsC, colsU;
  std::vector<int> ubN;
  std::vector<double> ubE, bounceC;  // bounceC: 1 + restitution of the clamping rows that bounced in `ref` (frozen with the active set)
  for (const GroupLCP& g : ref.groups) {
    std::vector<int> clampIdxOfRow(g.n, -1);
    for (int pass = 0; pass < 2; pass++)
      for (size_t i = 0; i < g.cons.size(); i++) {
        const Constraint& c = cons[g.cons[i]];
        for (int k = 0; k < c.dim; k++) {
          const int r = g.offset[i] + k;
          GradRow row{g.cons[i], k, rowDirection(c, k), zero6<double>()};
          row.F = join(cross(c.contact.point, row.dir), row.dir);
          if (pass == 0 && g.mapping[r] == -1) {
            clampIdxOfRow[r] = (int)colsC.size(); colsC.push_back(constraintForces(M, sc, c, row));
            bounceC.push_back(1.0 + ((size_t)r < g.restitution.size() ? g.restitution[r] : 0.0));
          }
          else if (pass == 1 && g.mapping[r] >= 0) {
            const int fp = g.mapping[r];
            const double upper = g.x[fp] * g.hi[r], lower = g.x[fp] * g.lo[r];
            colsU.push_back(constraintForces(M, sc, c, row));
            ubN.push_back(clampIdxOfRow[fp]);
            ubE.push_back(std::abs(g.x[r] - upper) < std::abs(g.x[r] - lower) ? g.hi[r] : g.lo[r]);
          }
        }
      }
  }
  const int nc = (int)colsC.size(), nu = (int)colsU.size();
  qdNext = vel;
  if (nc == 0) return true;
  Mat Ac(n, nc), A(n, nc);
  for (int j = 0; j < nc; j++) for (int i = 0; i < n; i++) { Ac(i, j) = colsC[j][i]; A(i, j) = colsC[j][i]; }
  for (int j = 0; j < nu; j++) for (int i = 0; i < n; i++) A(i, ubN[j]) += colsU[j][i] * ubE[j];
  Mat Minv(n, n);
  d.setState(q, qd);
  d.invMassMatrix(Minv.a.data());
  const Mat MA = mm(Minv, A);
  const Mat Q = mm(tr(Ac), MA);
  Mat b(nc, 1);
  for (int j = 0; j < nc; j++) { double s = 0; for (int i = 0; i < n; i++) s += Ac(i, j) * vel[i]; b(j, 0) = -bounceC[j] * s; }
  const Mat f = codSolveMat(Q, b);
  const Mat dv = mm(MA, f);
  for (int i = 0; i < n; i++) qdNext[i] = vel[i] + dv(i, 0);
  return true;
}

}  // namespace ddo
