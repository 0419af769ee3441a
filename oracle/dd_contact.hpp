// TEST INFRASTRUCTURE ONLY — see dd_spatial.hpp.
//
// dd_contact.hpp — CPU restatement of the contact half of World::step:
//   collision pass            DARTCollisionDetector::collide (collision/dart/DARTCollisionDetector.cpp:118-179), the
//                             BodyNodeCollisionFilter (collision/CollisionFilter.cpp:105-153), postProcess de-duplication
//                             at 3e-12 (:349-386); box-box is the REFERENCE's own dBoxBox (oracle/_ref/libboxbox_ref.so),
//                             sphere cases restate DARTCollide.cpp:1429-1800
//   contact constraints       ConstraintSolver::updateConstraints (constraint/ConstraintSolver.cpp:537-610: depth in
//                             [0, clipping depth]); ContactConstraint ctor / getInformation / getRelVelocity /
//                             getTangentBasisMatrixODE (constraint/ContactConstraint.cpp:66-229,349-455,675-683,716-746)
//   constrained groups        uniteSkeletons + buildConstrainedGroups (ContactConstraint.cpp:861-889, ConstraintSolver.cpp:718-788)
//   LCP assembly              BoxedLcpConstraintSolver::solveConstrainedGroup impulse tests (BoxedLcpConstraintSolver.cpp:227-328)
//   solver chain (no grads)   reduce -> Dantzig (the reference's own, oracle/_ref) -> isLCPSolutionValid -> PGS ->
//                             removeFriction + PGS (BoxedLcpConstraintSolver.cpp:425-616; LCPUtils.cpp:145-248,347-534)
//   applying the impulses     ContactConstraint::applyImpulse (:618-672) + Skeleton::computeImpulseForwardDynamics
//                             (Skeleton.cpp:5009-5033), then the explicit position update (World.cpp:262-274)
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include "dd_dynamics.hpp"
#include "dd_lcp.hpp"
#include "dd_cod.hpp"

namespace ddo {

typedef int (*RefBoxBoxFn)(const double*, const double*, const double*, const double*, double*, int);
typedef int (*RefDantzigFn)(int, double*, double*, double*, double*, double*, int*, int);
struct RefLibs {
  RefBoxBoxFn boxbox = nullptr;
  RefDantzigFn dantzig = nullptr;
};
inline RefLibs& refLibs() { static RefLibs r; return r; }

struct Contact {
  int bodyA, bodyB;
  Vec3<double> point, normal;
  double depth;
  int type;
  double meta[18];  // edgeAClosest, edgeAFixed, edgeADir, edgeBClosest, edgeBFixed, edgeBDir
};

inline void poseOf(const Iso<double>& W, const Iso<double>& off, double* T12) {
  Iso<double> S = mul(W, off);
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T12[3 * r + c] = S.R.m[r][c]; T12[9 + r] = S.p[r]; }
}

// sphere cases (DARTCollide.cpp:1429-1800); returns 0/1 contact
inline int boxSphere(const double* size, const double* Tb, double r, const double* c0, bool sphereFirst, Contact& c) {
  double half[3] = {0.5 * size[0], 0.5 * size[1], 0.5 * size[2]};
  double d[3] = {c0[0] - Tb[9], c0[1] - Tb[10], c0[2] - Tb[11]}, p[3];
  for (int k = 0; k < 3; k++) p[k] = Tb[k] * d[0] + Tb[3 + k] * d[1] + Tb[6 + k] * d[2];
  bool inside = true;
  int locks = 0;  // face{1,2,3}Locked (DARTCollide.cpp:1452-1493)
  for (int k = 0; k < 3; k++) {
    if (p[k] < -half[k]) { p[k] = -half[k]; inside = false; locks |= 1 << k; }
    if (p[k] > half[k]) { p[k] = half[k]; inside = false; locks |= 1 << k; }
  }
  auto side = [&](int& idx) { double mn = half[0] - std::abs(p[0]); double t = half[1] - std::abs(p[1]); idx = 0; if (t < mn) { mn = t; idx = 1; } t = half[2] - std::abs(p[2]); if (t < mn) { mn = t; idx = 2; } return mn; };
  auto axisNormal = [&](int idx) {
    double e[3] = {0, 0, 0};
    e[idx] = sphereFirst ? (p[idx] > 0.0 ? 1.0 : -1.0) : (p[idx] > 0.0 ? -1.0 : 1.0);
    for (int i = 0; i < 3; i++) c.normal[i] = Tb[3 * i] * e[0] + Tb[3 * i + 1] * e[1] + Tb[3 * i + 2] * e[2];
  };
  for (double& m : c.meta) m = 0;
  // Contact::sphereCenter / sphereRadius / face{1,2,3}Locked / face{1,2,3}Normal travel in the edge slots:
  // edgeAFixed = centre, edgeADir = (radius, lock mask, 0), edgeBClosest / edgeBFixed / edgeBDir = box axes
  for (int i = 0; i < 3; i++) { c.meta[3 + i] = c0[i]; c.meta[9 + i] = Tb[3 * i]; c.meta[12 + i] = Tb[3 * i + 1]; c.meta[15 + i] = Tb[3 * i + 2]; }
  c.meta[6] = r; c.meta[7] = (double)locks;
  c.type = sphereFirst ? 4 : 5;
  if (inside) {
    int idx; double mn = side(idx);
    axisNormal(idx);
    c.depth = mn + r; c.type = sphereFirst ? 1 : 2;
    for (int i = 0; i < 3; i++) c.point[i] = c0[i];
    return 1;
  }
  double cp[3], n[3];
  for (int i = 0; i < 3; i++) cp[i] = Tb[3 * i] * p[0] + Tb[3 * i + 1] * p[1] + Tb[3 * i + 2] * p[2] + Tb[9 + i];
  for (int i = 0; i < 3; i++) n[i] = sphereFirst ? c0[i] - cp[i] : cp[i] - c0[i];
  double mag = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), pen = r - mag;
  if (pen < 0.0) return 0;
  if (mag > 1e-6) { double inv = 1.0 / mag; for (int i = 0; i < 3; i++) c.normal[i] = n[i] * inv; }
  else { int idx; side(idx); axisNormal(idx); }
  for (int i = 0; i < 3; i++) c.point[i] = cp[i];
  c.depth = pen;
  return 1;
}
inline int sphereSphere(double r0, const double* c0, double r1, const double* c1, Contact& c) {
  double rsum = r0 + r1, n[3] = {c0[0] - c1[0], c0[1] - c1[1], c0[2] - c1[2]};
  double nsq = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  if (nsq > rsum * rsum) return 0;
  r0 /= rsum; r1 /= rsum;
  for (double& m : c.meta) m = 0;
  for (int i = 0; i < 3; i++) { c.point[i] = r1 * c0[i] + r0 * c1[i]; c.meta[3 + i] = c0[i]; c.meta[12 + i] = c1[i]; }  // centerA, centerB
  c.meta[6] = r0 * rsum; c.meta[15] = r1 * rsum;  // radiusA, radiusB
  c.type = 6;
  if (nsq < 1e-6) { c.normal = zero3<double>(); c.depth = rsum; return 1; }
  nsq = std::sqrt(nsq);
  double inv = 1.0 / nsq;
  for (int i = 0; i < 3; i++) c.normal[i] = n[i] * inv;
  c.depth = rsum - nsq;
  return 1;
}

// BodyNode::isReactive (BodyNode.cpp:1999-2018): the skeleton is mobile and the body has dependent generalized
// coordinates, i.e. the body or one of its ancestors hangs on a joint with DOFs (per body, not per skeleton)
inline bool isReactive(const Model& M, int body) {
  for (int i = body; i >= 0; i = M.b[i].parent) if (M.b[i].ndof > 0) return true;
  return false;
}

// collideCapsuleBox / collideBoxCapsule (DARTCollide.cpp:4217-4361): the end sphere the contact lies beyond goes to the sphere-box routine.
// The reference takes the contact position from libccd's MPR (not in the reference tree: parity unpinned there); here the end is the one
// from which the squared distance of the capsule's axis to the box grows strictly (see collide.cuh: capsuleBox, the same arithmetic).
// Returns the contact count, or -1 where the closest feature is the cylinder part (the reference's pipe contacts, not carried).
inline int capsuleBox(const double* cap, const double* Tc, const double* boxSize, const double* Tb, bool capsuleFirst, Contact& c) {
  const double r = cap[0], hh = 0.5 * cap[1];
  const double half[3] = {0.5 * boxSize[0], 0.5 * boxSize[1], 0.5 * boxSize[2]};
  double e[2][3], pl[2][3];
  for (int s = 0; s < 2; s++) {
    const double z = s == 0 ? -hh : hh;
    for (int i = 0; i < 3; i++) e[s][i] = Tc[3 * i + 2] * z + Tc[9 + i];
    const double dd[3] = {e[s][0] - Tb[9], e[s][1] - Tb[10], e[s][2] - Tb[11]};
    for (int k = 0; k < 3; k++) pl[s][k] = Tb[k] * dd[0] + Tb[3 + k] * dd[1] + Tb[6 + k] * dd[2];
  }
  auto gap = [&](double x, int k) { return x > half[k] ? x - half[k] : (x < -half[k] ? x + half[k] : 0.0); };
  double slope[2] = {0.0, 0.0}, dist2[2] = {0.0, 0.0};
  for (int s = 0; s < 2; s++)
    for (int k = 0; k < 3; k++) { const double g = gap(pl[s][k], k); slope[s] += g * (pl[1][k] - pl[0][k]); dist2[s] += g * g; }
  int end = -1;
  if (dist2[0] == 0.0 && dist2[1] == 0.0) return -1;
  if (dist2[0] == 0.0) end = 0;
  else if (dist2[1] == 0.0) end = 1;
  else if (slope[0] > 0.0) end = 0;
  else if (slope[1] < 0.0) end = 1;
  if (end < 0) {
    double lo = 0.0, hi = 1.0;
    for (int it = 0; it < 60; it++) {
      const double t = 0.5 * (lo + hi);
      double sl = 0.0;
      for (int k = 0; k < 3; k++) sl += gap(pl[0][k] + t * (pl[1][k] - pl[0][k]), k) * (pl[1][k] - pl[0][k]);
      if (sl > 0.0) hi = t; else lo = t;
    }
    const double t = 0.5 * (lo + hi);
    double f = 0.0;
    for (int k = 0; k < 3; k++) { const double g = gap(pl[0][k] + t * (pl[1][k] - pl[0][k]), k); f += g * g; }
    return f > r * r ? 0 : -1;
  }
  return boxSphere(boxSize, Tb, r, e[end], capsuleFirst, c);
}

// all-pairs narrow phase in collision-object order (one shape per body, flattened body order); *unsupported is set when a capsule
// pair is in a configuration this path does not carry
inline std::vector<Contact> collideWorld(const Model& M, const Dyn<double>& d, bool* unsupported = nullptr) {
  std::vector<Contact> total;
  std::vector<int> obj;
  for (int i = 0; i < M.nb; i++) if (M.b[i].shape != DDB_SHAPE_NONE) obj.push_back(i);
  for (size_t a = 0; a + 1 < obj.size(); a++)
    for (size_t b = a + 1; b < obj.size(); b++) {
      const int i = obj[a], j = obj[b];
      if (M.b[i].skel == M.b[j].skel) continue;  // self-collision checking is off by default (CollisionFilter.cpp:137-140)
      // a pair of two non-reactive bodies never yields an active constraint (ContactConstraint.cpp:342); it is left out of the
      // contact list here as in the CUDA pair table, so the reported contact counts agree
      if (!isReactive(M, i) && !isReactive(M, j)) continue;
      double T1[12], T2[12];
      poseOf(d.b[i].W, M.b[i].shapeOffset, T1);
      poseOf(d.b[j].W, M.b[j].shapeOffset, T2);
      std::vector<Contact> pair;
      const int s1 = M.b[i].shape, s2 = M.b[j].shape;
      if (s1 == DDB_SHAPE_BOX && s2 == DDB_SHAPE_BOX) {
        double out[26 * 16];
        int n = refLibs().boxbox(M.b[i].shapeSize.v, T1, M.b[j].shapeSize.v, T2, out, 16);
        for (int k = 0; k < n; k++) {
          Contact c;
          const double* o = out + 26 * k;
          for (int x = 0; x < 3; x++) { c.point[x] = o[x]; c.normal[x] = o[3 + x]; }
          for (int x = 0; x < 18; x++) c.meta[x] = o[6 + x];
          c.depth = o[24]; c.type = (int)o[25];
          pair.push_back(c);
        }
      } else {
        Contact c;
        int n = 0;
        if (s1 == DDB_SHAPE_CAPSULE && s2 == DDB_SHAPE_BOX) n = capsuleBox(M.b[i].shapeSize.v, T1, M.b[j].shapeSize.v, T2, true, c);
        else if (s1 == DDB_SHAPE_BOX && s2 == DDB_SHAPE_CAPSULE) n = capsuleBox(M.b[j].shapeSize.v, T2, M.b[i].shapeSize.v, T1, false, c);
        else if (s1 == DDB_SHAPE_CAPSULE || s2 == DDB_SHAPE_CAPSULE) n = -1;
        else if (s1 == DDB_SHAPE_SPHERE && s2 == DDB_SHAPE_SPHERE) n = sphereSphere(M.b[i].shapeSize[0], T1 + 9, M.b[j].shapeSize[0], T2 + 9, c);
        else if (s1 == DDB_SHAPE_SPHERE && s2 == DDB_SHAPE_BOX) n = boxSphere(M.b[j].shapeSize.v, T2, M.b[i].shapeSize[0], T1 + 9, true, c);
        else if (s1 == DDB_SHAPE_BOX && s2 == DDB_SHAPE_SPHERE) n = boxSphere(M.b[i].shapeSize.v, T1, M.b[j].shapeSize[0], T2 + 9, false, c);
        if (n < 0 && unsupported) *unsupported = true;
        if (n > 0) pair.push_back(c);
      }
      // postProcess (DARTCollisionDetector.cpp:349-386): drop contacts within 3e-12 of one already recorded
      for (Contact& c : pair) {
        bool close = false;
        for (const Contact& t : total) { Vec3<double> dd = c.point - t.point; if (norm(dd) < 3.0e-12) { close = true; break; } }
        if (close) continue;
        c.bodyA = i; c.bodyB = j;
        total.push_back(c);
      }
    }
  return total;
}

struct Constraint {  // ContactConstraint
  int bodyA, bodyB, dim;
  bool reactiveA, reactiveB;
  double mu, restitution;
  bool bounceOn;
  Vec6<double> NA[3], NB[3];  // mSpatialNormalA/B columns (body frame)
  Contact contact;
};


inline void tangentBasis(const Vec3<double>& n, Vec3<double>& t1, Vec3<double>& t2) {  // ContactConstraint.cpp:716-746, first direction = z
  Vec3<double> t = cross(v3<double>(0, 0, 1), n);
  if (dot(t, t) < 1e-12) {
    t = cross(v3<double>(1, 0, 0), n);
    if (dot(t, t) < 1e-12) {
      t = cross(v3<double>(0, 1, 0), n);
      if (dot(t, t) < 1e-12) t = cross(v3<double>(0, 0, 1), n);
    }
  }
  double z = dot(t, t);
  if (z > 0) { double nn = std::sqrt(z); t = v3<double>(t[0] / nn, t[1] / nn, t[2] / nn); }
  t1 = t;
  t2 = cross(n, t);
}

inline std::vector<Constraint> buildConstraints(const Model& M, const Dyn<double>& d, const std::vector<Contact>& contacts) {
  std::vector<Constraint> out;
  for (const Contact& ct : contacts) {
    if (dot(ct.normal, ct.normal) < 1e-12) continue;  // Contact::isZeroNormal
    if (ct.depth < 0.0) continue;
    if (ct.depth > M.clipDepth) continue;
    Constraint c;
    c.contact = ct;
    c.bodyA = ct.bodyA; c.bodyB = ct.bodyB;
    c.reactiveA = isReactive(M, ct.bodyA); c.reactiveB = isReactive(M, ct.bodyB);
    if (!c.reactiveA && !c.reactiveB) continue;  // update(): inactive
    c.restitution = M.b[ct.bodyA].restitution * M.b[ct.bodyB].restitution;
    c.bounceOn = c.restitution > 1e-3;
    c.mu = std::min(M.b[ct.bodyA].friction, M.b[ct.bodyB].friction);
    const bool frictionOn = c.mu > 1e-3;
    c.dim = frictionOn ? 3 : 1;
    Vec3<double> dirs[3];
    dirs[0] = ct.normal;
    if (frictionOn) tangentBasis(ct.normal, dirs[1], dirs[2]);
    const Iso<double>& WA = d.b[ct.bodyA].W;
    const Iso<double>& WB = d.b[ct.bodyB].W;
    Vec3<double> pA = apply(inverse(WA), ct.point), pB = apply(inverse(WB), ct.point);
    for (int k = 0; k < c.dim; k++) {
      Vec3<double> dA = mulT(WA.R, dirs[k]), dB = mulT(WB.R, -dirs[k]);
      c.NA[k] = join(cross(pA, dA), dA);
      c.NB[k] = join(cross(pB, dB), dB);
    }
    out.push_back(c);
  }
  return out;
}

// union-find over skeleton ids exactly as uniteSkeletons/compressPath do it; returns the group id of each constraint
// (groups numbered in order of first appearance) and the number of groups.
inline int buildGroups(const Model& M, const std::vector<Constraint>& cs, std::vector<int>& groupOf) {
  int nskel = 0;
  for (int i = 0; i < M.nb; i++) nskel = std::max(nskel, M.b[i].skel + 1);
  std::vector<int> root(nskel), size(nskel, 1);
  for (int i = 0; i < nskel; i++) root[i] = i;
  auto compress = [&](int s) { while (root[s] != s) { root[s] = root[root[s]]; s = root[s]; } return s; };
  auto getRoot = [&](int s) { while (root[s] != s) s = root[s]; return s; };
  for (const Constraint& c : cs) {
    if (!c.reactiveA || !c.reactiveB) continue;
    int sa = M.b[c.bodyA].skel, sb = M.b[c.bodyB].skel;
    if (sa == sb) continue;
    int ua = compress(sa), ub = compress(sb);
    if (ua == ub) continue;
    if (size[ua] < size[ub]) { root[ua] = ub; size[ub] += size[ua]; }
    else { root[ub] = ua; size[ua] += size[ub]; }
  }
  std::vector<int> groupRoots;
  groupOf.assign(cs.size(), -1);
  for (size_t k = 0; k < cs.size(); k++) {
    int s = cs[k].reactiveA ? M.b[cs[k].bodyA].skel : M.b[cs[k].bodyB].skel;
    int r = getRoot(s);
    int g = -1;
    for (size_t x = 0; x < groupRoots.size(); x++) if (groupRoots[x] == r) g = (int)x;
    if (g < 0) { g = (int)groupRoots.size(); groupRoots.push_back(r); }
    groupOf[k] = g;
  }
  return (int)groupRoots.size();
}

struct GroupLCP {
  int n = 0;
  std::vector<int> cons;    // constraint indices in group order
  std::vector<int> offset;  // first row of each constraint
  std::vector<double> A, b, lo, hi, x;  // A n x n row-major
  std::vector<int> findex;
  std::vector<double> massed;  // n columns of M^-1 J^T (nd each)
  std::vector<double> restitution;  // per row: the contact's coefficient on its normal row when it bounced, else 0
  int tier = 0;                // 0 short-circuit, 1 Dantzig, 2 PGS, 3 friction removed
  std::vector<int> mapping;    // ConstraintMapping per row: -1 CLAMPING, -2 NOT_CLAMPING, -3 ILLEGAL, >= 0 UPPER_BOUND -> normal row
  bool standardized = false;
};

// impulse response of a unit impulse along column k of constraint c: dqd (nd) and body velocity changes (in d.b[i].delV)
inline void impulseTest(const Model& M, Dyn<double>& d, const Constraint& c, int k, std::vector<double>& dqd) {
  d.clearImpulses();
  // each skeleton gets its own pass in the reference; one flattened pass is identical (other skeletons see zeros)
  if (c.reactiveA) d.b[c.bodyA].consImp = d.b[c.bodyA].consImp + c.NA[k];
  if (c.reactiveB) d.b[c.bodyB].consImp = d.b[c.bodyB].consImp + c.NB[k];
  dqd.assign(M.nd, 0.0);
  d.impulsePass(dqd.data());
}

inline void assembleGroup(const Model& M, Dyn<double>& d, const std::vector<Constraint>& cs, GroupLCP& g, double dt, bool penetrationCorrection) {
  g.offset.resize(g.cons.size());
  g.n = 0;
  for (size_t i = 0; i < g.cons.size(); i++) { g.offset[i] = g.n; g.n += cs[g.cons[i]].dim; }
  const int n = g.n;
  g.A.assign((size_t)n * n, 0.0); g.b.assign(n, 0.0); g.lo.assign(n, 0.0); g.hi.assign(n, 0.0); g.findex.assign(n, -1);
  g.massed.assign((size_t)n * M.nd, 0.0);
  std::vector<double> dqd;
  for (size_t i = 0; i < g.cons.size(); i++) {
    const Constraint& c = cs[g.cons[i]];
    const int o = g.offset[i];
    // getInformation: b = -(N_A^T V_A + N_B^T V_B) (+ bounce / ERP terms)
    for (int k = 0; k < c.dim; k++) {
      double rv = 0.0;
      rv -= dot(c.NA[k], d.b[c.bodyA].V);
      rv -= dot(c.NB[k], d.b[c.bodyB].V);
      g.b[o + k] = rv;
    }
    g.lo[o] = 0.0; g.hi[o] = INFINITY;
    if (c.dim == 3) { for (int t = 1; t < 3; t++) { g.lo[o + t] = -c.mu; g.hi[o + t] = c.mu; g.findex[o + t] = o; } }
    double bouncing = c.contact.depth - 0.0;  // error allowance 0
    if (bouncing < 0.0) bouncing = 0.0;
    else { bouncing *= 0.01 * (1.0 / dt); if (bouncing > 1e-3) bouncing = 1e-3; }
    if (!penetrationCorrection) bouncing = 0;
    if (c.bounceOn) {
      double restitutionVel = g.b[o] * c.restitution;
      if (restitutionVel > 1e-1) { if (restitutionVel > bouncing) { bouncing = restitutionVel; if (bouncing > 1e+2) bouncing = 1e+2; } }
    }
    g.b[o] += bouncing;
    {
      bool didBounce = c.bounceOn && (g.b[o] - bouncing) * c.restitution > 1e-1;
      g.restitution.resize(n, 0.0);
      g.restitution[o] = didBounce ? c.restitution : 0.0;
    }
    for (int k = 0; k < c.dim; k++) {
      impulseTest(M, d, c, k, dqd);
      for (int q = 0; q < M.nd; q++) g.massed[(size_t)(o + k) * M.nd + q] = dqd[q];
      // row (o+k): own block and the blocks of later constraints from the velocity changes; earlier ones by symmetry
      for (size_t m = i; m < g.cons.size(); m++) {
        const Constraint& cm = cs[g.cons[m]];
        for (int l = 0; l < cm.dim; l++) {
          double v = 0.0;
          if (cm.reactiveA) v += dot(cm.NA[l], d.b[cm.bodyA].delV);
          if (cm.reactiveB) v += dot(cm.NB[l], d.b[cm.bodyB].delV);
          g.A[(size_t)(o + k) * n + g.offset[m] + l] = v;
        }
      }
      for (size_t m = 0; m < i; m++)
        for (int l = 0; l < cs[g.cons[m]].dim; l++) g.A[(size_t)(o + k) * n + g.offset[m] + l] = g.A[(size_t)(g.offset[m] + l) * n + o + k];
    }
  }
}

// LCPUtils::reduce / mergeLCPColumns (LCPUtils.cpp:145-202,347-440) with mapOut kept as an index map
struct Reduced {
  int n;
  std::vector<double> A, x, b, hi, lo;
  std::vector<int> findex, mapTo;  // mapTo[original row] = reduced column (or -1 when dropped)
};
inline Reduced reduceLCP(const GroupLCP& g, const std::vector<double>& x0) {
  Reduced r;
  r.n = g.n; r.A = g.A; r.x = x0; r.b = g.b; r.hi = g.hi; r.lo = g.lo; r.findex = g.findex;
  r.mapTo.resize(g.n);
  for (int i = 0; i < g.n; i++) r.mapTo[i] = i;
  while (true) {
    const int n = r.n;
    bool found = false;
    for (int a = 0; a < n - 1 && !found; a++)
      for (int b = a + 1; b < n; b++) {
        double sq = 0;
        for (int k = 0; k < n; k++) { double dlt = r.A[(size_t)k * n + a] - r.A[(size_t)k * n + b]; sq += dlt * dlt; }
        if (sq < 1e-4 && std::abs(r.b[a] - r.b[b]) < 1e-4 && r.findex[a] == r.findex[b] && r.hi[a] == r.hi[b] && r.lo[a] == r.lo[b]) {
          // mergeLCPColumns(a, b): column a doubled, column/row b removed
          Reduced m;
          m.n = n - 1;
          m.A.assign((size_t)(n - 1) * (n - 1), 0.0); m.x.assign(n - 1, 0.0); m.b.assign(n - 1, 0.0); m.hi.assign(n - 1, 0.0); m.lo.assign(n - 1, 0.0);
          m.findex.assign(n - 1, 0);
          std::vector<double> cols((size_t)n * (n - 1), 0.0);
          for (int i = 0; i < n; i++) {
            if (i == b) continue;
            int ni = i > b ? i - 1 : i;
            for (int k = 0; k < n; k++) cols[(size_t)k * (n - 1) + ni] = r.A[(size_t)k * n + i] * (i == a ? 2.0 : 1.0);
            m.x[ni] = r.x[i]; m.b[ni] = r.b[i]; m.hi[ni] = r.hi[i]; m.lo[ni] = r.lo[i];
            if (r.findex[i] < b) m.findex[ni] = r.findex[i];
            else if (r.findex[i] == b) m.findex[ni] = a;
            else m.findex[ni] = r.findex[i] - 1;
          }
          for (int i = 0; i < n; i++) {
            if (i == b) continue;
            int ni = i > b ? i - 1 : i;
            for (int k = 0; k < n - 1; k++) m.A[(size_t)ni * (n - 1) + k] = cols[(size_t)i * (n - 1) + k];
          }
          m.mapTo = r.mapTo;
          for (int& t : m.mapTo) { if (t == b) t = a; else if (t > b) t--; }
          r = m;
          found = true;
          break;
        }
      }
    if (!found) break;
  }
  return r;
}
// LCPUtils::removeFriction (:209-248)
inline Reduced removeFrictionLCP(const GroupLCP& g, const std::vector<double>& x0) {
  Reduced r;
  std::vector<int> keep;
  for (int i = 0; i < g.n; i++) if (g.findex[i] == -1) keep.push_back(i);
  r.n = (int)keep.size();
  r.mapTo.assign(g.n, -1);
  r.A.assign((size_t)r.n * r.n, 0.0); r.x.resize(r.n); r.b.resize(r.n); r.hi.resize(r.n); r.lo.resize(r.n); r.findex.assign(r.n, -1);
  for (int a = 0; a < r.n; a++) {
    r.mapTo[keep[a]] = a;
    r.x[a] = x0[keep[a]]; r.b[a] = g.b[keep[a]]; r.hi[a] = g.hi[keep[a]]; r.lo[a] = g.lo[keep[a]];
    for (int b = 0; b < r.n; b++) r.A[(size_t)a * r.n + b] = g.A[(size_t)keep[a] * g.n + keep[b]];
  }
  return r;
}
inline void mapOut(const Reduced& r, std::vector<double>& x) {
  for (size_t i = 0; i < x.size(); i++) x[i] = r.mapTo[i] >= 0 ? r.x[r.mapTo[i]] : 0.0;
}
inline bool runDantzig(Reduced& r, bool early) {
  const int n = r.n, ns = dPad(n);
  std::vector<double> Ap((size_t)n * ns, 0.0);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ap[(size_t)i * ns + j] = r.A[(size_t)i * n + j];
  return refLibs().dantzig(n, Ap.data(), r.x.data(), r.b.data(), r.lo.data(), r.hi.data(), r.findex.data(), early ? 1 : 0) != 0;
}
inline bool runPgs(Reduced& r) {
  const int n = r.n, ns = dPad(n);
  std::vector<double> Ap((size_t)n * ns, 0.0);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ap[(size_t)i * ns + j] = r.A[(size_t)i * n + j];
  return pgs(n, Ap.data(), r.x.data(), r.b.data(), r.lo.data(), r.hi.data(), r.findex.data());
}
inline bool hasNaN(const std::vector<double>& v) { for (double x : v) if (std::isnan(x)) return true; return false; }

// BoxedLcpConstraintSolver::solveConstrainedGroup without gradient matrices (:425-616). x0 is the solver's mX on entry
// (zeros when the size changed).
inline void solveGroupNoGrad(GroupLCP& g, const std::vector<double>& x0) {
  const int n = g.n;
  g.x = x0;
  bool success = false;
  {
    Reduced r = reduceLCP(g, g.x);
    success = runDantzig(r, /*earlyTermination=*/true);
    if (success) {
      mapOut(r, g.x);
      if (!lcpValid(n, g.A.data(), g.x.data(), g.b.data(), g.hi.data(), g.lo.data(), g.findex.data(), false)) success = false;
    }
    g.tier = 1;
  }
  if (hasNaN(g.x)) { success = false; std::fill(g.x.begin(), g.x.end(), 0.0); }
  if (!success) {
    Reduced r = reduceLCP(g, x0);
    success = runPgs(r);
    if (success) {
      mapOut(r, g.x);
      if (!lcpValid(n, g.A.data(), g.x.data(), g.b.data(), g.hi.data(), g.lo.data(), g.findex.data(), false)) success = false;
    }
    g.tier = 2;
  }
  if (!success) {
    Reduced r = removeFrictionLCP(g, x0);
    runPgs(r);
    mapOut(r, g.x);
    g.tier = 3;
  }
  if (hasNaN(g.x)) std::fill(g.x.begin(), g.x.end(), 0.0);
}

// ---------------------------------------------------------------------------------------------------------
// Gradient-enabled path (BoxedLcpConstraintSolver.cpp:381-421,647-664; ConstrainedGroupGradientMatrices.cpp:167-324,467-855)
// ---------------------------------------------------------------------------------------------------------
// LCPUtils::guessSolution (LCPUtils.cpp:87-141)
inline std::vector<double> guessSolution(const GroupLCP& g) {
  const int n = g.n;
  std::vector<int> cl;
  for (int i = 0; i < n; i++) { if (g.findex[i] == -1) { if (g.b[i] > 0) cl.push_back(i); } else cl.push_back(i); }
  const int nc = (int)cl.size();
  if (nc == n) return codSolve(n, n, g.A, g.b);
  std::vector<double> full(n, 0.0);
  if (nc == 0) return full;
  std::vector<double> rA((size_t)nc * nc), rb(nc);
  for (int r = 0; r < nc; r++) { rb[r] = g.b[cl[r]]; for (int c = 0; c < nc; c++) rA[(size_t)r * nc + c] = g.A[(size_t)cl[r] * n + cl[c]]; }
  std::vector<double> rx = codSolve(nc, nc, rA, rb);
  for (int i = 0; i < nc; i++) full[cl[i]] = rx[i];
  return full;
}

struct Classification {
  std::vector<int> mapping, clampIdx, ubIdx;
  int numClamping = 0, numUpperBound = 0;
};
// constructMatrices, the classification part (ConstrainedGroupGradientMatrices.cpp:467-697)
inline Classification classify(const GroupLCP& g, const std::vector<double>& x, bool ignoreFriction) {
  const int n = g.n;
  Classification c;
  c.mapping = g.findex; c.clampIdx.assign(n, -1); c.ubIdx.assign(n, -1);
  const double CLAMPING_THRESHOLD = 1e-6;
  for (int j = 0; j < n; j++) {
    double colNorm = 0;  // aColNormGradientBackup: squared norm of column j of A
    for (int r = 0; r < n; r++) colNorm += g.A[(size_t)r * n + j] * g.A[(size_t)r * n + j];
    if (colNorm < 1e-9) { c.mapping[j] = -2; continue; }
    const double force = x[j];
    double upper = g.hi[j], lower = g.lo[j];
    const int fp = g.findex[j];
    if (fp != -1) { upper *= x[fp]; lower *= x[fp]; }
    if (std::abs(force) < CLAMPING_THRESHOLD) {
      if (fp != -1) {
        if (std::abs(x[fp]) < CLAMPING_THRESHOLD) c.mapping[j] = -2;
        else if (ignoreFriction) c.mapping[j] = -2;
        else { c.mapping[j] = -1; c.clampIdx[j] = c.numClamping++; }
      } else c.mapping[j] = -2;
      continue;
    }
    const double tie = 1e-5;
    if ((x[j] > lower + tie && x[j] < upper - tie) || (lower - x[j] > 1e-2 || x[j] - upper > 1e-2)) {
      c.mapping[j] = -1; c.clampIdx[j] = c.numClamping++;
    } else if (lower - x[j] > 1e-2 || x[j] - upper > 1e-2) {
      c.mapping[j] = -3;
    } else if (fp != -1 && std::abs(x[fp]) > 1e-9) {
      double fpNorm = 0;
      for (int r = 0; r < n; r++) fpNorm += g.A[(size_t)r * n + fp] * g.A[(size_t)r * n + fp];
      if (fpNorm > 1e-9 && (fp > j || c.mapping[fp] == -1)) { c.mapping[j] = fp; c.ubIdx[j] = c.numUpperBound++; }
      else c.mapping[j] = -2;
    } else c.mapping[j] = -2;
  }
  return c;
}
// constructMatrices + opportunisticallyStandardizeResults (:206-324); x is updated when the standardised solution is valid
inline bool constructAndStandardize(GroupLCP& g, std::vector<double>& x, bool ignoreFriction, Classification& cls) {
  const int n = g.n;
  for (int guard = 0; guard < 8; guard++) {
    cls = classify(g, x, ignoreFriction);
    if (n == 0) return true;
    const int nc = cls.numClamping;
    if (nc == 0) {
      std::vector<double> zero(n, 0.0);
      if (lcpValid(n, g.A.data(), zero.data(), g.b.data(), g.hi.data(), g.lo.data(), g.findex.data(), ignoreFriction)) { x = zero; return true; }
      return false;
    }
    // Q = A_c^T M^-1 (A_c + A_ub E) = A[c,c] + A[c,ub] E ; E(u, clampIdx[findex]) = hi or lo, whichever bound x sits on
    std::vector<int> rowOfClamp(nc);
    for (int j = 0; j < n; j++) if (cls.clampIdx[j] >= 0) rowOfClamp[cls.clampIdx[j]] = j;
    std::vector<double> Q((size_t)nc * nc, 0.0), bc(nc), fOrig(nc);
    for (int a = 0; a < nc; a++) {
      bc[a] = g.b[rowOfClamp[a]]; fOrig[a] = x[rowOfClamp[a]];
      for (int b2 = 0; b2 < nc; b2++) Q[(size_t)a * nc + b2] = g.A[(size_t)rowOfClamp[a] * n + rowOfClamp[b2]];
    }
    for (int j = 0; j < n; j++) {
      if (cls.mapping[j] < 0) continue;
      const int fp = cls.mapping[j];
      const double upper = x[fp] * g.hi[j], lower = x[fp] * g.lo[j];
      const double e = (std::abs(x[j] - upper) < std::abs(x[j] - lower)) ? g.hi[j] : g.lo[j];
      const int col = cls.clampIdx[fp];
      for (int a = 0; a < nc; a++) Q[(size_t)a * nc + col] += g.A[(size_t)rowOfClamp[a] * n + j] * e;
    }
    std::vector<double> fc = codSolve(nc, nc, Q, bc);
    std::vector<double> newX(n, 0.0);
    bool anyNewlyNotClamping = false;
    for (int i = 0; i < n; i++) {
      if (cls.clampIdx[i] != -1) {
        newX[i] = fc[cls.clampIdx[i]];
        if (std::abs(fc[cls.clampIdx[i]]) < 1e-6 && std::abs(x[i]) > 1e-6 && g.findex[i] == -1) anyNewlyNotClamping = true;
      }
      if (cls.ubIdx[i] != -1) {
        const int fi = g.findex[i];
        const double originalMultiple = fOrig[cls.clampIdx[fi]] / x[i];
        const double clean = (std::abs(originalMultiple - g.hi[i]) < std::abs(originalMultiple - g.lo[i])) ? g.hi[i] : g.lo[i];
        newX[i] = fc[cls.clampIdx[fi]] * clean;
      }
    }
    if (!lcpValid(n, g.A.data(), newX.data(), g.b.data(), g.hi.data(), g.lo.data(), g.findex.data(), ignoreFriction)) return false;
    x = newX;
    if (!anyNewlyNotClamping) return true;
  }
  return true;
}

// BoxedLcpConstraintSolver::solveConstrainedGroup with gradient matrices. mxResized: the solver's mX had another size.
inline void solveGroupGrad(GroupLCP& g, const std::vector<double>& cache, bool mxResized) {
  const int n = g.n;
  std::vector<double> x0 = mxResized ? guessSolution(g) : cache;
  Classification cls;
  g.x = x0;
  bool success = constructAndStandardize(g, g.x, false, cls);
  g.tier = 0;
  bool ignoreFriction = false;
  if (!success) {
    g.x = x0;
    {
      Reduced r = reduceLCP(g, g.x);
      success = runDantzig(r, true);
      if (success) { mapOut(r, g.x); if (!lcpValid(n, g.A.data(), g.x.data(), g.b.data(), g.hi.data(), g.lo.data(), g.findex.data(), false)) success = false; }
      g.tier = 1;
    }
    if (hasNaN(g.x)) { success = false; std::fill(g.x.begin(), g.x.end(), 0.0); }
    if (!success) {
      Reduced r = reduceLCP(g, x0);
      success = runPgs(r);
      if (success) { mapOut(r, g.x); if (!lcpValid(n, g.A.data(), g.x.data(), g.b.data(), g.hi.data(), g.lo.data(), g.findex.data(), false)) success = false; }
      g.tier = 2;
    }
    if (!success) {
      ignoreFriction = true;
      Reduced r = removeFrictionLCP(g, x0);
      runPgs(r);
      mapOut(r, g.x);
      g.tier = 3;
    }
    if (hasNaN(g.x)) std::fill(g.x.begin(), g.x.end(), 0.0);
    std::vector<double> xs = g.x;
    bool st = constructAndStandardize(g, xs, ignoreFriction, cls);
    if (st) g.x = xs;
    g.standardized = st;
  } else {
    g.standardized = true;
  }
  g.mapping = cls.mapping;
}

struct ContactStepResult {
  bool unsupportedPair = false;  // a capsule pair in a configuration this path does not carry (status bit 9)
  std::vector<double> qNext, qdNext, preConstraintVel;
  std::vector<Contact> contacts;
  std::vector<Constraint> constraints;
  std::vector<GroupLCP> groups;
};

// World::step with gradients disabled. lcpCache: the solver's mX (one group) in/out.
inline void stepWithContacts(const Model& M, const double* q, const double* qd, const double* tau, std::vector<double>& lcpCache,
                             ContactStepResult& R, bool gradients = false) {
  const int n = M.nd;
  Dyn<double> d(M);
  d.setState(q, qd);
  std::vector<double> acc(n), vel(n);
  d.forwardDynamics(tau, acc.data());
  for (int i = 0; i < n; i++) vel[i] = qd[i] + acc[i] * M.dt;
  R.preConstraintVel = vel;
  d.setVelocities(vel.data());  // body velocities now reflect the unconstrained velocity update
  d.artInertia();
  R.contacts = collideWorld(M, d, &R.unsupportedPair);
  R.constraints = buildConstraints(M, d, R.contacts);
  std::vector<int> groupOf;
  const int ng = buildGroups(M, R.constraints, groupOf);
  R.groups.assign(ng, GroupLCP());
  for (size_t k = 0; k < R.constraints.size(); k++) R.groups[groupOf[k]].cons.push_back((int)k);
  std::vector<Vec6<double>> imp(M.nb, zero6<double>());
  for (GroupLCP& g : R.groups) {
    assembleGroup(M, d, R.constraints, g, M.dt, false);
    std::vector<double> x0(g.n, 0.0);
    const bool resized = (int)lcpCache.size() != g.n;
    if (!resized) x0 = lcpCache;  // mX persists only while the size is unchanged (:205-210)
    if (gradients) solveGroupGrad(g, x0, resized);
    else solveGroupNoGrad(g, x0);
    lcpCache = g.x;
    for (size_t i = 0; i < g.cons.size(); i++) {
      const Constraint& c = R.constraints[g.cons[i]];
      for (int k = 0; k < c.dim; k++) {
        const double lam = g.x[g.offset[i] + k];
        if (c.reactiveA) imp[c.bodyA] = imp[c.bodyA] + lam * c.NA[k];
        if (c.reactiveB) imp[c.bodyB] = imp[c.bodyB] + lam * c.NB[k];
      }
    }
  }
  R.qdNext = vel;
  if (!R.constraints.empty()) {
    for (int i = 0; i < M.nb; i++) d.b[i].consImp = imp[i];
    std::vector<double> dqd(n, 0.0);
    d.impulsePass(dqd.data());
    for (int i = 0; i < n; i++) R.qdNext[i] = vel[i] + dqd[i];
  }
  R.qNext.resize(n);
  d.integratePositionsExplicit(q, qd, M.dt, R.qNext.data());
}

}  // namespace ddo
