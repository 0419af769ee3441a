// contact.cuh — the contact half of World::step for one world per thread (FP64, compiled with -fmad=false).
// Runs after k_forward has written the step snapshot and the unconstrained velocity update:
//   collision pass          DARTCollisionDetector::collide all-pairs loop + BodyNodeCollisionFilter + postProcess
//                           de-duplication at 3e-12 (collision/dart/DARTCollisionDetector.cpp:118-179,349-386;
//                           collision/CollisionFilter.cpp:105-153); narrow phase in collide.cuh
//   contact constraints     ConstraintSolver::updateConstraints (constraint/ConstraintSolver.cpp:537-610) and
//                           ContactConstraint (constraint/ContactConstraint.cpp:66-229,349-455,675-683,716-746)
//   constrained groups      uniteSkeletons / buildConstrainedGroups (ContactConstraint.cpp:861-889, ConstraintSolver.cpp:718-788)
//   LCP assembly            impulse tests through the articulated-body quantities of the snapshot
//                           (BoxedLcpConstraintSolver.cpp:227-328; Skeleton.cpp:4859-4894,4990-4994; BodyNode.cpp:1730-1751,1793-1820)
//   solver chain            LCPUtils::reduce -> Dantzig -> isLCPSolutionValid -> PGS -> removeFriction + PGS
//                           (BoxedLcpConstraintSolver.cpp:425-616; LCPUtils.cpp:145-248,347-534)
//   applying the impulses   ContactConstraint::applyImpulse (:618-672) + Skeleton::computeImpulseForwardDynamics (Skeleton.cpp:5009-5033)
#pragma once
#include "dynamics.cuh"
#include "lcp.cuh"
#include "collide.cuh"

namespace ddb {

struct ShapeDesc {  // per body collision shape (one ShapeNode per body)
  int shape, skel, reactive, pad;
  double friction, restitution;
  double size[3];
  double offR[9], offP[3];  // ShapeNode relative transform
};
struct ContactTopo {
  int n_pairs, cmax, kmax, nskel;
  int gradients, bouncing;  // bouncing: some collidable pair has restitution > 1e-3. gradients 1: the gradient-enabled solve of forwardPass (short-circuit, standardisation, classification)
  double clip_depth;
  int pairA[DDB_MAX_BODIES * 4], pairB[DDB_MAX_BODIES * 4];  // collidable pairs in collision-object order
  ShapeDesc shape[DDB_MAX_BODIES];
};
__constant__ ContactTopo c_ctopo;

// per-world scratch layout (doubles / ints), world-major
struct ContactLayout {
  int nb, nd, cmax, kmax;
  __host__ __device__ size_t oContacts() const { return 0; }                         // [cmax][26]
  __host__ __device__ size_t oNA() const { return oContacts() + (size_t)cmax * 26; }  // [cmax][3][6]
  __host__ __device__ size_t oNB() const { return oNA() + (size_t)cmax * 18; }
  __host__ __device__ size_t oMassed() const { return oNB() + (size_t)cmax * 18; }    // [kmax][nd]
  __host__ __device__ size_t oA() const { return oMassed() + (size_t)kmax * nd; }     // [kmax][kmax] (per group, n x n packed)
  __host__ __device__ size_t oB() const { return oA() + (size_t)kmax * kmax; }        // b lo hi x0 x xall xs xnew colnorm mx: 10 x [kmax]
  __host__ __device__ size_t oW() const { return oB() + (size_t)10 * kmax; }           // [nb][12] world transforms
  __host__ __device__ size_t oV() const { return oW() + (size_t)nb * 12; }            // [nb][6] body velocities (pre-constraint)
  __host__ __device__ size_t oImp() const { return oV() + (size_t)nb * 6; }           // [nb][6] constraint impulses
  __host__ __device__ size_t oBias() const { return oImp() + (size_t)nb * 6; }        // [nb][6]
  __host__ __device__ size_t oDelV() const { return oBias() + (size_t)nb * 6; }       // [nb][6]
  __host__ __device__ size_t oU() const { return oDelV() + (size_t)nb * 6; }          // [nb] + dqd [nd]
  __host__ __device__ size_t padDoubles() const { return (size_t)19 * nb + nd; }      // imp, bias, delV [nb][6] each, u [nb], dqd [nd]
  __host__ __device__ size_t oLanePads() const { return oU() + (size_t)nb + nd; }     // 32 private copies for the impulse tests run side by side
  __host__ __device__ size_t doubles() const { return oLanePads() + 32 * padDoubles(); }
  // ints: bodyA bodyB dim type group [cmax each], findex keep mapTo [kmax each], misc
  __host__ __device__ size_t oSaved() const { return (size_t)5 * cmax + (size_t)7 * kmax + 8; }  // loop state kept across a deferred PGS solve
  __host__ __device__ size_t ints() const { return oSaved() + 8 + DDB_MAX_BODIES; }
};

// scratch of one impulse-dynamics pass: constraint impulses, bias impulses, body velocity changes, joint impulses, dqd
struct ImpulsePad {
  double* p;
  int nb;
  DDB_LCP_DI double* imp(int body) const { return p + (size_t)body * 6; }
  DDB_LCP_DI double* bias(int body) const { return p + (size_t)(nb + body) * 6; }
  DDB_LCP_DI double* delV(int body) const { return p + (size_t)(2 * nb + body) * 6; }
  DDB_LCP_DI double* u() const { return p + (size_t)18 * nb; }
  DDB_LCP_DI double* dqd() const { return p + (size_t)19 * nb; }
};

struct ContactWork {
  double* d;
  int* i;
  ContactLayout L;
  long long* prof = nullptr;  // optional: clock64 ticks per section, summed over worlds (DDB_CONTACT_PROFILE=1)
  // optional (k_contact with 8 warps per CTA): the section each warp of the CTA has reached, in shared memory. The stage is bound
  // by instruction fetches and warps at the same place of the code share them, so at the end of a section a warp lets the others
  // catch up - for a bounded time: this is a scheduling hint, nothing depends on it (a warp that left the stage reads as "ahead").
  volatile int* convoy = nullptr;
  int convoyWarps = 0;
  mutable int convoyGroup = 0, convoySection = 0, convoyLast = 0;  // (sections completed in this group; last meeting point published)
  DDB_LCP_DI void convoyWait(int id) const {
#ifdef __CUDA_ARCH__
    if (id <= convoyLast) return;  // (meeting points are published in increasing order)
    convoyLast = id;
    convoy[threadIdx.x >> 5] = id;
    __syncwarp();
    for (int spins = 0; spins < 4000; spins++) {
      int behind = 0;
      for (int k = 0; k < convoyWarps; k++) behind |= convoy[k] < id;
      if (!behind) break;
      __nanosleep(100);
    }
    __syncwarp();
#endif
  }
  DDB_LCP_DI void mark(int section, long long& t) const {
#ifdef __CUDA_ARCH__
    if (prof) { const long long now = clock64(); atomicAdd(reinterpret_cast<unsigned long long*>(prof) + section, (unsigned long long)(now - t)); t = now; }
    if (convoy) { convoySection = section + 1; convoyWait(convoyGroup * 8 + convoySection); }
#endif
  }
  DDB_LCP_DI double* contact(int c) const { return d + L.oContacts() + (size_t)c * 26; }
  DDB_LCP_DI double* NA(int c, int k) const { return d + L.oNA() + ((size_t)c * 3 + k) * 6; }
  DDB_LCP_DI double* NB(int c, int k) const { return d + L.oNB() + ((size_t)c * 3 + k) * 6; }
  DDB_LCP_DI double* massed(int row) const { return d + L.oMassed() + (size_t)row * L.nd; }
  DDB_LCP_DI double* A() const { return d + L.oA(); }
  DDB_LCP_DI double* b() const { return d + L.oB(); }
  DDB_LCP_DI double* lo() const { return d + L.oB() + L.kmax; }
  DDB_LCP_DI double* hi() const { return d + L.oB() + 2 * (size_t)L.kmax; }
  DDB_LCP_DI double* x0() const { return d + L.oB() + 3 * (size_t)L.kmax; }
  DDB_LCP_DI double* x() const { return d + L.oB() + 4 * (size_t)L.kmax; }
  DDB_LCP_DI double* xall() const { return d + L.oB() + 5 * (size_t)L.kmax; }  // multipliers of every row of the world
  DDB_LCP_DI double* xs() const { return d + L.oB() + 6 * (size_t)L.kmax; }
  DDB_LCP_DI double* xnew() const { return d + L.oB() + 7 * (size_t)L.kmax; }
  DDB_LCP_DI double* colnorm() const { return d + L.oB() + 8 * (size_t)L.kmax; }
  DDB_LCP_DI double* mx() const { return d + L.oB() + 9 * (size_t)L.kmax; }  // the solver's cached solution while the groups are solved
  DDB_LCP_DI double* W(int body) const { return d + L.oW() + (size_t)body * 12; }
  DDB_LCP_DI double* V(int body) const { return d + L.oV() + (size_t)body * 6; }
  DDB_LCP_DI double* imp(int body) const { return d + L.oImp() + (size_t)body * 6; }
  DDB_LCP_DI double* bias(int body) const { return d + L.oBias() + (size_t)body * 6; }
  DDB_LCP_DI double* delV(int body) const { return d + L.oDelV() + (size_t)body * 6; }
  DDB_LCP_DI ImpulsePad pad0() const { return ImpulsePad{d + L.oImp(), L.nb}; }  // the world's own pad (imp/bias/delV/u/dqd above)
  DDB_LCP_DI ImpulsePad lanePad(int lane) const { return ImpulsePad{d + L.oLanePads() + (size_t)lane * L.padDoubles(), L.nb}; }
  DDB_LCP_DI double* u() const { return d + L.oU(); }
  DDB_LCP_DI double* dqd() const { return d + L.oU() + L.nb; }
  DDB_LCP_DI int* bodyA() const { return i; }
  DDB_LCP_DI int* bodyB() const { return i + L.cmax; }
  DDB_LCP_DI int* dim() const { return i + 2 * L.cmax; }
  DDB_LCP_DI int* groupOf() const { return i + 4 * L.cmax; }
  DDB_LCP_DI int* findex() const { return i + 5 * L.cmax; }
  DDB_LCP_DI int* keep() const { return i + 5 * L.cmax + L.kmax; }
  DDB_LCP_DI int* mapTo() const { return i + 5 * L.cmax + 2 * L.kmax; }
  DDB_LCP_DI int* mapping() const { return i + 5 * L.cmax + 3 * L.kmax; }   // ConstraintMapping of the current group
  DDB_LCP_DI int* clampIdx() const { return i + 5 * L.cmax + 4 * L.kmax; }
  DDB_LCP_DI int* ubIdx() const { return i + 5 * L.cmax + 5 * L.kmax; }
  DDB_LCP_DI int* activeAll() const { return i + 5 * L.cmax + 6 * L.kmax; }  // active set of every row of the world
  DDB_LCP_DI int* saved() const { return i + L.oSaved(); }
};

// What the backward pass needs of one step's contact stage, kept per snapshot slot (world-major):
//   doubles  contacts [cmax][26] | massed M^-1 J^T [kmax][nd] | multipliers x [kmax] | pre-constraint velocity v_f [nd]
//   ints     bodyA [cmax] | bodyB [cmax] | dim [cmax] | ConstraintMapping [kmax] | row info (constraint * 4 + direction) [kmax] | n_contacts, n_rows
struct ContactSnapLayout {
  int nd, cmax, kmax;
  __host__ __device__ size_t oContacts() const { return 0; }
  __host__ __device__ size_t oMassed() const { return (size_t)cmax * 26; }
  __host__ __device__ size_t oX() const { return oMassed() + (size_t)kmax * nd; }
  __host__ __device__ size_t oVf() const { return oX() + kmax; }
  __host__ __device__ size_t oRest() const { return oVf() + nd; }  // [kmax] restitution coefficient of a row that bounced (normal rows), else 0
  __host__ __device__ size_t doubles() const { return oRest() + kmax; }
  __host__ __device__ size_t oBodyA() const { return 0; }
  __host__ __device__ size_t oBodyB() const { return cmax; }
  __host__ __device__ size_t oDim() const { return 2 * (size_t)cmax; }
  __host__ __device__ size_t oMapping() const { return 3 * (size_t)cmax; }
  __host__ __device__ size_t oRowInfo() const { return 3 * (size_t)cmax + kmax; }
  __host__ __device__ size_t oCounts() const { return 3 * (size_t)cmax + 2 * (size_t)kmax; }
  __host__ __device__ size_t ints() const { return oCounts() + 2; }
};

DDB_LCP_DI void coopSyncWarp(int coopN) {
#ifdef __CUDA_ARCH__
  if (coopN > 1) __syncwarp();
#endif
}
DDB_LCP_DI SV ldSV6(const double* p) { return sv(mk(p[0], p[1], p[2]), mk(p[3], p[4], p[5])); }
DDB_LCP_DI void stSV6(double* p, SV a) { p[0] = a.w.x; p[1] = a.w.y; p[2] = a.w.z; p[3] = a.v.x; p[4] = a.v.y; p[5] = a.v.z; }
DDB_LCP_DI double dotSV6(const double* n, const double* v) { return n[0] * v[0] + n[1] * v[1] + n[2] * v[2] + n[3] * v[3] + n[4] * v[4] + n[5] * v[5]; }

// Impulse dynamics: body-frame constraint impulses in cw.imp -> dqd (per dof) and body velocity changes (cw.delV).
// (Skeleton::updateBiasImpulse + updateVelocityChange; GenericJoint.hpp:2135-2152,2266-2272,2365-2377)
DDB_LCP_DI void impulsePass(const Snap& sn, const ImpulsePad& cw, int nb) {
  for (int i = 0; i < nb; i++) {
    const double* im = cw.imp(i);
    double* bi = cw.bias(i);
    for (int k = 0; k < 6; k++) bi[k] = -im[k];
  }
  for (int i = nb - 1; i >= 0; i--) {
    const DevBody& b = c_topo.b[i];
    SV bi = ldSV6(cw.bias(i));
    if (b.type == DDB_JOINT_FREE) {
      // u = -S^T bias (6); beta = bias + AIS psi u goes to the parent
      double u6[6];
      mulS6T(b, bi, u6);
      for (int k = 0; k < 6; k++) u6[k] = -u6[k];
      // keep u in delV (it is rewritten by the downward sweep before anyone reads it)
      for (int k = 0; k < 6; k++) cw.delV(i)[k] = u6[k];
      cw.u()[i] = 0.0;
      if (b.parent >= 0) {
        double AISm[36], psi6[36], pu[6], av[6];
        ld36(sn, i, 12, AISm); ld36(sn, i, 30, psi6);
        for (int k = 0; k < 6; k++) { double acc = 0.0; for (int l = 0; l < 6; l++) acc += psi6[k * 6 + l] * u6[l]; pu[k] = acc; }
        for (int a = 0; a < 6; a++) { double acc = 0.0; for (int k = 0; k < 6; k++) acc += AISm[a * 6 + k] * pu[k]; av[a] = acc; }
        SV add = dAdInvT(sn.ldT(i), bi + sv6(av));
        SV pb = ldSV6(cw.bias(b.parent));
        stSV6(cw.bias(b.parent), pb + add);
      }
      continue;
    }
    double ui = 0.0;
    const bool hasDof = b.type != DDB_JOINT_WELD;
    if (hasDof) ui = -dot(b.S, bi);
    cw.u()[i] = ui;
    if (b.parent >= 0) {
      SV beta = bi;
      if (hasDof) {
        SV ais = sn.ldSV(i, 12);
        double ps = sn.ld(i, 15).x;
        beta = bi + (ps * ui) * ais;
      }
      SV add = dAdInvT(sn.ldT(i), beta);
      SV pb = ldSV6(cw.bias(b.parent));
      stSV6(cw.bias(b.parent), pb + add);
    }
  }
  for (int i = 0; i < nb; i++) {
    const DevBody& b = c_topo.b[i];
    SV Vp = svzero();
    if (b.parent >= 0) Vp = AdInvT(sn.ldT(i), ldSV6(cw.delV(b.parent)));
    if (b.type == DDB_JOINT_FREE) {
      double AISm[36], psi6[36], r6[6], dq6[6];
      ld36(sn, i, 12, AISm); ld36(sn, i, 30, psi6);
      const double vp[6] = {Vp.w.x, Vp.w.y, Vp.w.z, Vp.v.x, Vp.v.y, Vp.v.z};
      for (int k = 0; k < 6; k++) { double acc = 0.0; for (int a = 0; a < 6; a++) acc += AISm[a * 6 + k] * vp[a]; r6[k] = cw.delV(i)[k] - acc; }
      for (int k = 0; k < 6; k++) { double acc = 0.0; for (int l = 0; l < 6; l++) acc += psi6[k * 6 + l] * r6[l]; dq6[k] = acc; cw.dqd()[b.dof + k] = acc; }
      stSV6(cw.delV(i), Vp + mulS6(b, dq6));
      continue;
    }
    double dq = 0.0;
    if (b.type != DDB_JOINT_WELD) {
      SV ais = sn.ldSV(i, 12);
      double ps = sn.ld(i, 15).x;
      dq = ps * (cw.u()[i] - dot(ais, Vp));
      cw.dqd()[b.dof] = dq;
    }
    stSV6(cw.delV(i), Vp + dq * b.S);
  }
}

// ContactConstraint::getTangentBasisMatrixODE with mFirstFrictionalDirection = z (ContactConstraint.cpp:716-746)
DDB_LCP_DI void tangentBasis(const double* n, double* t1, double* t2) {
  auto crossInto = [](const double* a, const double* b, double* c) { c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0]; };
  const double ez[3] = {0, 0, 1}, ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0};
  double t[3];
  crossInto(ez, n, t);
  if (dot3(t, t) < 1e-12) {
    crossInto(ex, n, t);
    if (dot3(t, t) < 1e-12) {
      crossInto(ey, n, t);
      if (dot3(t, t) < 1e-12) crossInto(ez, n, t);
    }
  }
  normalize3(t);
  t1[0] = t[0]; t1[1] = t[1]; t1[2] = t[2];
  crossInto(n, t, t2);
}

struct ContactStepOut {
  int n_contacts, n_rows, n_groups, tier_mask;
  int pending;  // a PGS solve was handed to the warp-cooperative solver: nothing was committed, the stage is run again
};

// Deferred PGS: the projected Gauss-Seidel solves of the fallback chain run in their own kernel, one warp per world with
// the matrix in shared memory (k_pgs_warp in lcp_api.cu). The stage is deterministic, so it is simply run again with the
// answers of the solves it has already asked for: request k of a world is answered from ansX[k] / ansOk[k] when
// k < answered, otherwise the reduced problem is left in the world's LCP view and the stage returns `pending`.
// Everything the stage had computed up to the request lives in the world's scratch, so the next pass does not start over:
// it re-enters the group loop right behind the answered solve (resumeKind; loop scalars in ContactWork::saved()).
struct PgsDefer {
  int answered, seen, kmax, pendingN;
  int cap;             // answer slots per world; further solves of the same step run inline
  int resumeKind;      // in: 0 = run from the start, 2 / 3 = continue behind the answered PGS solve of that tier; out: tier of the solve just deferred
  const double* ansX;  // [requests][kmax]
  const int* ansOk;    // [requests]
};

// The whole contact stage for one world. qd_pre: velocities after the unconstrained update (in/out: becomes qd').
// lcp_cache/n_cache: the solver's cached solution mX (in/out). lv: this world's LCP scratch view storage.
DDB_LCP_DI ContactStepOut contactStage(const Snap& sn, const ContactWork& cw, double* lcpD, int* lcpI, const double* qd_pre_in, double* qd_out,
                                       int layout, int N, int w, double* lcp_cache, int* n_cache, const PgsOptions& pgs,
                                       double* keepD = nullptr, int* keepI = nullptr, PgsDefer* defer = nullptr, int lcpStride = 32) {
  const int nb = c_topo.nb, nd = c_topo.nd;
  const ContactTopo& ct = c_ctopo;
  ContactStepOut out = {0, 0, 0, 0, 0};
  int coopLane = 0, coopN = 1;  // one world per warp (lcpStride == 1): independent pieces of work are spread over the lanes
#ifdef __CUDA_ARCH__
  if (lcpStride == 1) { coopLane = threadIdx.x & 31; coopN = 32; }
#endif
  long long tmark = 0;
#ifdef __CUDA_ARCH__
  if (cw.prof) tmark = clock64();
#endif
  const bool resuming = defer && defer->resumeKind != 0;
  int* sv = cw.saved();
  int* offs = cw.i + 3 * cw.L.cmax;  // first LCP row of each constraint inside its group
  int nc = 0, ng = 0;
  if (resuming) { nc = sv[1]; ng = sv[2]; out.n_contacts = nc; out.n_groups = ng; }
  else {
  // ---- world transforms and body velocities under the unconstrained velocity update ----
  for (int i = 0; i < nb; i++) {
    const DevBody& b = c_topo.b[i];
    Xf t = sn.ldT(i);
    double* Wi = cw.W(i);
    if (b.parent >= 0) {
      const double* Wp = cw.W(b.parent);
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) Wi[3 * r + c] = Wp[3 * r] * t.R.m[c] + Wp[3 * r + 1] * t.R.m[3 + c] + Wp[3 * r + 2] * t.R.m[6 + c];
        Wi[9 + r] = Wp[3 * r] * t.p.x + Wp[3 * r + 1] * t.p.y + Wp[3 * r + 2] * t.p.z + Wp[9 + r];
      }
    } else {
      for (int k = 0; k < 9; k++) Wi[k] = t.R.m[k];
      Wi[9] = t.p.x; Wi[10] = t.p.y; Wi[11] = t.p.z;
    }
    SV Vp = svzero();
    if (b.parent >= 0) Vp = AdInvT(t, ldSV6(cw.V(b.parent)));
    if (b.type == DDB_JOINT_FREE) {
      double qd6[6];
      for (int k = 0; k < 6; k++) qd6[k] = qd_pre_in[vidx(layout, N, nd, w, b.dof + k)];
      stSV6(cw.V(i), Vp + mulS6(b, qd6));
    } else {
      double qdi = (b.type != DDB_JOINT_WELD) ? qd_pre_in[vidx(layout, N, nd, w, b.dof)] : 0.0;
      stSV6(cw.V(i), Vp + qdi * b.S);
    }
  }
  // ---- collision pass ----
  nc = 0;
  for (int pi = 0; pi < ct.n_pairs; pi++) {
    const int a = ct.pairA[pi], bb = ct.pairB[pi];
    const ShapeDesc& sa = ct.shape[a];
    const ShapeDesc& sb = ct.shape[bb];
    Pose Ta, Tb;
    auto shapePose = [&](const double* Wb, const ShapeDesc& s, Pose& T) {
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) T.R[3 * r + c] = Wb[3 * r] * s.offR[c] + Wb[3 * r + 1] * s.offR[3 + c] + Wb[3 * r + 2] * s.offR[6 + c];
        T.p[r] = Wb[3 * r] * s.offP[0] + Wb[3 * r + 1] * s.offP[1] + Wb[3 * r + 2] * s.offP[2] + Wb[9 + r];
      }
    };
    shapePose(cw.W(a), sa, Ta);
    shapePose(cw.W(bb), sb, Tb);
    ContactPoint pts[8];
    const int np = collidePair(sa.shape, sa.size, Ta, sb.shape, sb.size, Tb, pts);
    if (np < 0) out.tier_mask |= 1 << 9;  // a capsule pair in a configuration this path does not carry (collide.cuh): reported, never guessed
    for (int k = 0; k < np; k++) {
      const ContactPoint& c = pts[k];
      bool close = false;
      for (int t = 0; t < nc && !close; t++) {
        const double* o = cw.contact(t);
        double dx = c.point[0] - o[0], dy = c.point[1] - o[1], dz = c.point[2] - o[2];
        if (sqrt(dx * dx + dy * dy + dz * dz) < 3.0e-12) close = true;
      }
      if (close) continue;
      if (nc >= ct.cmax) continue;  // capacity reached (reported through n_contacts == cmax)
      double* o = cw.contact(nc);
      const double* vs[8] = {c.point, c.normal, c.edgeAClosest, c.edgeAFixed, c.edgeADir, c.edgeBClosest, c.edgeBFixed, c.edgeBDir};
      for (int v = 0; v < 8; v++) for (int x = 0; x < 3; x++) o[3 * v + x] = vs[v][x];
      o[24] = c.depth; o[25] = (double)c.type;
      cw.bodyA()[nc] = a; cw.bodyB()[nc] = bb;
      nc++;
    }
  }
  out.n_contacts = nc;
  // ---- constraints: contacts with a non-zero normal, depth in [0, clip], at least one reactive body ----
  for (int c = 0; c < nc; c++) {
    const double* o = cw.contact(c);
    const int a = cw.bodyA()[c], bb = cw.bodyB()[c];
    const bool ra = ct.shape[a].reactive != 0, rb = ct.shape[bb].reactive != 0;
    const double depth = o[24];
    cw.dim()[c] = 0;
    if (dot3(o + 3, o + 3) < 1e-12 || depth < 0.0 || depth > ct.clip_depth || (!ra && !rb)) continue;
    const double mu = fmin(ct.shape[a].friction, ct.shape[bb].friction);
    const int dim = mu > 1e-3 ? 3 : 1;
    cw.dim()[c] = dim;
    double dirs[3][3];
    for (int x = 0; x < 3; x++) dirs[0][x] = o[3 + x];
    if (dim == 3) tangentBasis(dirs[0], dirs[1], dirs[2]);
    const double* WA = cw.W(a);
    const double* WB = cw.W(bb);
    double pA[3], pB[3];
    {
      double dA[3] = {o[0] - WA[9], o[1] - WA[10], o[2] - WA[11]}, dB[3] = {o[0] - WB[9], o[1] - WB[10], o[2] - WB[11]};
      for (int x = 0; x < 3; x++) { pA[x] = colDot(WA, x, dA); pB[x] = colDot(WB, x, dB); }
    }
    for (int k = 0; k < dim; k++) {
      double dA[3], dB[3], nd3[3] = {-dirs[k][0], -dirs[k][1], -dirs[k][2]};
      for (int x = 0; x < 3; x++) { dA[x] = colDot(WA, x, dirs[k]); dB[x] = colDot(WB, x, nd3); }
      double* na = cw.NA(c, k);
      double* nbp = cw.NB(c, k);
      na[0] = pA[1] * dA[2] - pA[2] * dA[1]; na[1] = pA[2] * dA[0] - pA[0] * dA[2]; na[2] = pA[0] * dA[1] - pA[1] * dA[0];
      na[3] = dA[0]; na[4] = dA[1]; na[5] = dA[2];
      nbp[0] = pB[1] * dB[2] - pB[2] * dB[1]; nbp[1] = pB[2] * dB[0] - pB[0] * dB[2]; nbp[2] = pB[0] * dB[1] - pB[1] * dB[0];
      nbp[3] = dB[0]; nbp[4] = dB[1]; nbp[5] = dB[2];
    }
  }
  // ---- constrained groups: union-find over skeletons, groups numbered by first appearance ----
  int root[DDB_MAX_BODIES], usize[DDB_MAX_BODIES], groupRoot[DDB_MAX_BODIES];
  for (int s = 0; s < ct.nskel; s++) { root[s] = s; usize[s] = 1; }
  for (int c = 0; c < nc; c++) {
    if (!cw.dim()[c]) continue;
    const int a = cw.bodyA()[c], bb = cw.bodyB()[c];
    if (!ct.shape[a].reactive || !ct.shape[bb].reactive) continue;
    int sa = ct.shape[a].skel, sb = ct.shape[bb].skel;
    if (sa == sb) continue;
    while (root[sa] != sa) { root[sa] = root[root[sa]]; sa = root[sa]; }
    while (root[sb] != sb) { root[sb] = root[root[sb]]; sb = root[sb]; }
    if (sa == sb) continue;
    if (usize[sa] < usize[sb]) { root[sa] = sb; usize[sb] += usize[sa]; }
    else { root[sb] = sa; usize[sa] += usize[sb]; }
  }
  ng = 0;
  for (int c = 0; c < nc; c++) {
    cw.groupOf()[c] = -1;
    if (!cw.dim()[c]) continue;
    int s = ct.shape[cw.bodyA()[c]].reactive ? ct.shape[cw.bodyA()[c]].skel : ct.shape[cw.bodyB()[c]].skel;
    while (root[s] != s) s = root[s];
    int g = -1;
    for (int x = 0; x < ng; x++) if (groupRoot[x] == s) g = x;
    if (g < 0) { g = ng; groupRoot[ng++] = s; }
    cw.groupOf()[c] = g;
  }
  out.n_groups = ng;
  for (int i = 0; i < nb; i++) for (int k = 0; k < 6; k++) cw.imp(i)[k] = 0.0;
  }
  cw.mark(0, tmark);  // collision, constraints, groups
  int groupBase[DDB_MAX_BODIES];
  int mxSize = *n_cache;  // the solver's mX persists across groups and steps while its size is unchanged
  double* mx = cw.mx();   // committed to lcp_cache only when the whole stage has run
  if (!resuming) for (int k = 0; k < mxSize && k < ct.kmax; k++) mx[k] = lcp_cache[k];
  // returns the solver's bool, or -1 when the solve was deferred to the warp-cooperative kernel
  auto runPgs = [&](const LcpView& vv) -> int {
    if (!defer) return pgsSolve(vv, pgs) ? 1 : 0;
    const int k = defer->seen++;
    if (k >= defer->cap) return pgsSolve(vv, pgs) ? 1 : 0;
    if (k < defer->answered) {
      for (int i = 0; i < vv.n; i++) vv.x(i) = defer->ansX[(size_t)k * defer->kmax + i];
      return defer->ansOk[k];
    }
    defer->pendingN = vv.n;
    return -1;
  };
  int rowsTotal = 0;
  double* A = cw.A();
  if (resuming) {
    rowsTotal = sv[3]; mxSize = sv[4]; out.tier_mask = sv[5];
    for (int i = 0; i < sv[0]; i++) groupBase[i] = sv[8 + i];
    defer->seen = defer->answered;
  }
  // the loop scalars a later pass needs to continue behind a deferred solve of tier `kind` in group g
  auto saveState = [&](int g, int kind, int nReduced) {
    sv[0] = g; sv[1] = nc; sv[2] = ng; sv[3] = rowsTotal; sv[4] = mxSize; sv[5] = out.tier_mask; sv[6] = nReduced;
    for (int i = 0; i < g; i++) sv[8 + i] = groupBase[i];
    defer->resumeKind = kind;
  };
  for (int g = resuming ? sv[0] : 0; g < ng; g++) {
    cw.convoyGroup = g;
    const bool resumeHere = resuming && g == sv[0];
    // rows of this group
    int n = 0;
    for (int c = 0; c < nc; c++) if (cw.groupOf()[c] == g) { offs[c] = n; n += cw.dim()[c]; }
    if (rowsTotal + n > ct.kmax) {  // capacity: leave the remaining groups unsolved (reported through n_rows = -1)
      for (int c = 0; c < nc; c++) if (cw.groupOf()[c] >= g) cw.groupOf()[c] = -1;
      rowsTotal = -1;
      break;
    }
    double* b = cw.b(); double* lo = cw.lo(); double* hi = cw.hi(); double* x0 = cw.x0(); double* x = cw.x();
    int* findex = cw.findex();
    if (!resumeHere) {
    // ---- assembly by impulse tests ----
    for (int c = 0; c < nc; c++) {
      if (cw.groupOf()[c] != g) continue;
      const int o = offs[c], dim = cw.dim()[c];
      const int a = cw.bodyA()[c], bb = cw.bodyB()[c];
      const bool ra = ct.shape[a].reactive != 0, rb = ct.shape[bb].reactive != 0;
      const double mu = fmin(ct.shape[a].friction, ct.shape[bb].friction);
      const double rest = ct.shape[a].restitution * ct.shape[bb].restitution;
      for (int k = 0; k < dim; k++) {
        double rv = 0.0;
        rv -= dotSV6(cw.NA(c, k), cw.V(a));
        rv -= dotSV6(cw.NB(c, k), cw.V(bb));
        b[o + k] = rv;
        findex[o + k] = -1;
        if (keepI) keepI[ContactSnapLayout{nd, ct.cmax, ct.kmax}.oRowInfo() + rowsTotal + o + k] = c * 4 + k;
      }
      lo[o] = 0.0; hi[o] = INFINITY;
      if (dim == 3) { for (int t = 1; t < 3; t++) { lo[o + t] = -mu; hi[o + t] = mu; findex[o + t] = o; } }
      // penetration correction is off by default (World.cpp:85); restitution bounce (ContactConstraint.cpp:395-418)
      double bouncing = 0.0;
      bool didBounce = false;
      if (rest > 1e-3) {
        double restitutionVel = b[o] * rest;
        didBounce = restitutionVel > 1e-1;
        if (restitutionVel > 1e-1) { if (restitutionVel > bouncing) { bouncing = restitutionVel; if (bouncing > 1e+2) bouncing = 1e+2; } }
      }
      b[o] += bouncing;
      // ContactConstraint::getCoefficientOfRestitution (ContactConstraint.cpp:929-935) as registerConstraint records it per row
      // (ConstrainedGroupGradientMatrices.cpp:99-119): the coefficient on the normal row of a contact that bounced, 0 elsewhere
      if (keepD) for (int k = 0; k < dim; k++) keepD[ContactSnapLayout{nd, ct.cmax, ct.kmax}.oRest() + rowsTotal + o + k] = (k == 0 && didBounce) ? rest : 0.0;
      for (int k = 0; k < dim; k++) cw.keep()[o + k] = c * 4 + k;  // row -> (constraint, direction) for the tests below
    }
    // the impulse tests are independent: with one warp per world every lane runs one on its private pad
    coopSyncWarp(coopN);
    {
      const ImpulsePad pad = coopN > 1 ? cw.lanePad(coopLane) : cw.pad0();
      for (int base = 0; base < n; base += coopN) {
        const int t = base + coopLane;
        if (t >= n) continue;
        const int c = cw.keep()[t] >> 2, k = cw.keep()[t] & 3, o = offs[c];
        const int a = cw.bodyA()[c], bb = cw.bodyB()[c];
        const bool ra = ct.shape[a].reactive != 0, rb = ct.shape[bb].reactive != 0;
        for (int i = 0; i < nb; i++) for (int z = 0; z < 6; z++) pad.imp(i)[z] = 0.0;
        if (ra) for (int z = 0; z < 6; z++) pad.imp(a)[z] += cw.NA(c, k)[z];
        if (rb) for (int z = 0; z < 6; z++) pad.imp(bb)[z] += cw.NB(c, k)[z];
        for (int z = 0; z < nd; z++) pad.dqd()[z] = 0.0;
        impulsePass(sn, pad, nb);
        double* mcol = cw.massed(rowsTotal + o + k);
        for (int z = 0; z < nd; z++) mcol[z] = pad.dqd()[z];
        for (int m = c; m < nc; m++) {
          if (cw.groupOf()[m] != g) continue;
          const int am = cw.bodyA()[m], bm = cw.bodyB()[m];
          for (int l = 0; l < cw.dim()[m]; l++) {
            double v = 0.0;
            if (ct.shape[am].reactive) v += dotSV6(cw.NA(m, l), pad.delV(am));
            if (ct.shape[bm].reactive) v += dotSV6(cw.NB(m, l), pad.delV(bm));
            A[(size_t)(o + k) * n + offs[m] + l] = v;
          }
        }
      }
      coopSyncWarp(coopN);
      // the blocks below the diagonal mirror rows computed by other tests
      for (int base = 0; base < n; base += coopN) {
        const int t = base + coopLane;
        if (t >= n) continue;
        const int c = cw.keep()[t] >> 2, k = cw.keep()[t] & 3, o = offs[c];
        for (int m = 0; m < c; m++) {
          if (cw.groupOf()[m] != g) continue;
          for (int l = 0; l < cw.dim()[m]; l++) A[(size_t)(o + k) * n + offs[m] + l] = A[(size_t)(offs[m] + l) * n + o + k];
        }
      }
      coopSyncWarp(coopN);
    }
    // ---- squared column norms of A (aColNormGradientBackup) ----
    for (int j = 0; j < n; j++) { double sacc = 0.0; for (int r = 0; r < n; r++) sacc += A[(size_t)r * n + j] * A[(size_t)r * n + j]; cw.colnorm()[j] = sacc; }
    // ---- the solver's cached solution ----
    if (mxSize == n) { for (int k = 0; k < n; k++) x0[k] = mx[k]; }
    else { for (int k = 0; k < n; k++) x0[k] = 0.0; }
    for (int k = 0; k < n; k++) x[k] = x0[k];
    }
    cw.mark(1, tmark);  // assembly by impulse tests
    // ---- LCPUtils::reduce: merge duplicate columns (index map + column scale instead of explicit matrices) ----
    int* keep = cw.keep();
    int* mapTo = cw.mapTo();
    LcpView v;
    auto isValid = [&]() {
      return lcpValid(n, [&](int i, int j) { return A[(size_t)i * n + j]; }, [&](int i) { return x[i]; }, [&](int i) { return b[i]; },
                      [&](int i) { return hi[i]; }, [&](int i) { return lo[i]; }, [&](int i) { return findex[i]; }, false);
    };
    auto hasNaN = [&]() { for (int k = 0; k < n; k++) if (isnan(x[k])) return true; return false; };
    // column scales live in the (unused until the solve) L area of the LCP scratch
    auto buildReduced = [&](bool frictionless) -> int {
      int nr = 0;
      if (frictionless) {
        for (int k = 0; k < n; k++) { if (findex[k] == -1) { mapTo[k] = nr; keep[nr++] = k; } else mapTo[k] = -1; }
        v.init(lcpD, lcpI, nr, ct.kmax, lcpStride);
        coopSync(v);
        for (int a = 0; a < nr; a++) {
          for (int c2 = v.lane; c2 < nr; c2 += v.nl) v.Araw(a, c2) = A[(size_t)keep[a] * n + keep[c2]];
          v.b(a) = b[keep[a]]; v.lo(a) = lo[keep[a]]; v.hi(a) = hi[keep[a]]; v.findex(a) = -1; v.x(a) = x0[keep[a]];
          v.tmp(a) = 1.0;  // column scale (read by the warp-cooperative PGS)
        }
        coopSync(v);
        return nr;
      }
      nr = n;
      // scale[] and fidx[] of the reduced problem are kept in the view's d / C arrays until the copy below
      v.init(lcpD, lcpI, n, ct.kmax, lcpStride);
      for (int k = 0; k < n; k++) { keep[k] = k; mapTo[k] = k; v.d(k) = 1.0; v.C(k) = findex[k]; }
      for (;;) {
        bool found = false;
        for (int a = 0; a < nr - 1 && !found; a++) {
          // the first c2 > a whose column duplicates column a; with one warp per world the candidates are tested 32 at a time
          for (int base = a + 1; base < nr && !found; base += v.nl) {
            const int c2 = base + v.lane;
            bool dup = false;
            if (c2 < nr && v.C(a) == v.C(c2) && hi[keep[a]] == hi[keep[c2]] && lo[keep[a]] == lo[keep[c2]] &&
                fabs(b[keep[a]] - b[keep[c2]]) < 1e-4) {
              double sq = 0.0;
              for (int k = 0; k < nr; k++) {
                double dl = A[(size_t)keep[k] * n + keep[a]] * v.d(a) - A[(size_t)keep[k] * n + keep[c2]] * v.d(c2);
                sq += dl * dl;
              }
              dup = sq < 1e-4;
            }
            const int first = coopFirstTrue(v, dup, base);
            if (first < 0) continue;
            // mergeLCPColumns(a, first)
            v.d(a) = v.d(a) * 2.0;
            for (int k = 0; k < nr; k++) {
              int f = v.C(k);
              if (f == first) v.C(k) = a; else if (f > first) v.C(k) = f - 1;
            }
            for (int k = first; k < nr - 1; k++) { keep[k] = keep[k + 1]; v.d(k) = v.d(k + 1); v.C(k) = v.C(k + 1); }
            for (int k = 0; k < n; k++) { if (mapTo[k] == first) mapTo[k] = a; else if (mapTo[k] > first) mapTo[k]--; }
            nr--;
            found = true;
          }
        }
        if (!found) break;
      }
      // materialise the reduced problem in the view (scale/fidx are read before their slots are overwritten)
      v.n = nr; v.nskip = dpad(nr);
      for (int a = 0; a < nr; a++) { v.tmp(a) = v.d(a); v.findex(a) = v.C(a); }
      coopSync(v);
      for (int a = 0; a < nr; a++) {
        for (int c2 = v.lane; c2 < nr; c2 += v.nl) v.Araw(a, c2) = A[(size_t)keep[a] * n + keep[c2]] * v.tmp(c2);
        v.b(a) = b[keep[a]]; v.lo(a) = lo[keep[a]]; v.hi(a) = hi[keep[a]]; v.x(a) = x0[keep[a]];
      }
      coopSync(v);
      return nr;
    };
    auto mapOut = [&]() { for (int k = 0; k < n; k++) x[k] = mapTo[k] >= 0 ? v.x(mapTo[k]) : 0.0; };
    // ---- gradient-enabled pieces (ConstrainedGroupGradientMatrices.cpp:206-324,467-697; LCPUtils.cpp:87-141) ----
    int* mapping = cw.mapping();
    int* clampIdx = cw.clampIdx();
    int* ubIdx = cw.ubIdx();
    auto validWith = [&](const double* xv, bool ignoreFriction) {
      return lcpValid(n, [&](int i, int j) { return A[(size_t)i * n + j]; }, [&](int i) { return xv[i]; }, [&](int i) { return b[i]; },
                      [&](int i) { return hi[i]; }, [&](int i) { return lo[i]; }, [&](int i) { return findex[i]; }, ignoreFriction);
    };
    auto classify = [&](const double* xv, bool ignoreFriction, int& numClamping) {
      numClamping = 0;
      int numUB = 0;
      for (int j = 0; j < n; j++) { mapping[j] = findex[j]; clampIdx[j] = -1; ubIdx[j] = -1; }
      for (int j = 0; j < n; j++) {
        if (cw.colnorm()[j] < 1e-9) { mapping[j] = -2; continue; }
        const double force = xv[j];
        double upper = hi[j], lower = lo[j];
        const int fp = findex[j];
        if (fp != -1) { upper *= xv[fp]; lower *= xv[fp]; }
        if (fabs(force) < 1e-6) {
          if (fp != -1) {
            if (fabs(xv[fp]) < 1e-6) mapping[j] = -2;
            else if (ignoreFriction) mapping[j] = -2;
            else { mapping[j] = -1; clampIdx[j] = numClamping++; }
          } else mapping[j] = -2;
          continue;
        }
        const double tie = 1e-5;
        if ((xv[j] > lower + tie && xv[j] < upper - tie) || (lower - xv[j] > 1e-2 || xv[j] - upper > 1e-2)) {
          mapping[j] = -1; clampIdx[j] = numClamping++;
        } else if (lower - xv[j] > 1e-2 || xv[j] - upper > 1e-2) {
          mapping[j] = -3;
        } else if (fp != -1 && fabs(xv[fp]) > 1e-9) {
          if (cw.colnorm()[fp] > 1e-9 && (fp > j || mapping[fp] == -1)) { mapping[j] = fp; ubIdx[j] = numUB++; }
          else mapping[j] = -2;
        } else mapping[j] = -2;
      }
    };
    // constructMatrices + opportunisticallyStandardizeResults: xv is replaced by the standardised solution when valid
    auto constructAndStandardize = [&](double* xv, bool ignoreFriction) -> bool {
      double* xn = cw.xnew();
      for (int guard = 0; guard < 8; guard++) {
        int nc2 = 0;
        classify(xv, ignoreFriction, nc2);
        if (nc2 == 0) {
          for (int k = 0; k < n; k++) xn[k] = 0.0;
          if (validWith(xn, ignoreFriction)) { for (int k = 0; k < n; k++) xv[k] = 0.0; return true; }
          return false;
        }
        v.init(lcpD, lcpI, nc2, ct.kmax, lcpStride);
        auto Qm = [&](int i, int j) -> double& { return v.db[(size_t)(v.oL + i * nc2 + j) * v.st]; };
        int* rowOfClamp = keep;
        for (int j = 0; j < n; j++) if (clampIdx[j] >= 0) rowOfClamp[clampIdx[j]] = j;
        for (int a = 0; a < nc2; a++) {
          v.dw(a) = b[rowOfClamp[a]];
          v.w(a) = xv[rowOfClamp[a]];  // original clamping impulses
          for (int c2 = 0; c2 < nc2; c2++) Qm(a, c2) = A[(size_t)rowOfClamp[a] * n + rowOfClamp[c2]];
        }
        for (int j = 0; j < n; j++) {
          if (mapping[j] < 0) continue;
          const int fp = mapping[j];
          const double upper = xv[fp] * hi[j], lower = xv[fp] * lo[j];
          const double e = (fabs(xv[j] - upper) < fabs(xv[j] - lower)) ? hi[j] : lo[j];
          const int col = clampIdx[fp];
          for (int a = 0; a < nc2; a++) Qm(a, col) += A[(size_t)rowOfClamp[a] * n + j] * e;
        }
        codSolveDev(v, nc2);  // f_c in v.dx
        bool anyNewlyNotClamping = false;
        for (int i = 0; i < n; i++) {
          xn[i] = 0.0;
          if (clampIdx[i] != -1) {
            const double f = v.dx(clampIdx[i]);
            xn[i] = f;
            if (fabs(f) < 1e-6 && fabs(xv[i]) > 1e-6 && findex[i] == -1) anyNewlyNotClamping = true;
          }
          if (ubIdx[i] != -1) {
            const int fi = findex[i];
            const double originalMultiple = v.w(clampIdx[fi]) / xv[i];
            const double clean = (fabs(originalMultiple - hi[i]) < fabs(originalMultiple - lo[i])) ? hi[i] : lo[i];
            xn[i] = v.dx(clampIdx[fi]) * clean;
          }
        }
        if (!validWith(xn, ignoreFriction)) return false;
        for (int k = 0; k < n; k++) xv[k] = xn[k];
        if (!anyNewlyNotClamping) return true;
      }
      return true;
    };
    const bool gradients = ct.gradients != 0;
    if (!resumeHere && gradients && mxSize != n) {
      // LCPUtils::guessSolution: least squares on {normal rows with b > 0} + {all friction rows}
      int ncl = 0;
      for (int i = 0; i < n; i++) { if (findex[i] == -1) { if (b[i] > 0) keep[ncl++] = i; } else keep[ncl++] = i; }
      for (int k = 0; k < n; k++) x0[k] = 0.0;
      if (ncl > 0) {
        v.init(lcpD, lcpI, ncl, ct.kmax, lcpStride);
        for (int a = 0; a < ncl; a++) {
          v.dw(a) = b[keep[a]];
          for (int c2 = 0; c2 < ncl; c2++) v.db[(size_t)(v.oL + a * ncl + c2) * v.st] = A[(size_t)keep[a] * n + keep[c2]];
        }
        codSolveDev(v, ncl);
        for (int a = 0; a < ncl; a++) x0[keep[a]] = v.dx(a);
      }
      for (int k = 0; k < n; k++) x[k] = x0[k];
    }
    cw.mark(2, tmark);  // guessSolution
    bool success = false;
    int tier = 1;
    bool ignoreFrictionFinal = false;
    const int stage = resumeHere ? defer->resumeKind : 0;  // 2 / 3: continue behind the answered PGS solve of that tier
    // the answer of the solve this pass was waiting for, put where mapOut() expects it
    auto loadAnswer = [&]() -> int {
      const int k = defer->answered - 1;
      v.init(lcpD, lcpI, sv[6], ct.kmax, lcpStride);
      for (int i = 0; i < v.n; i++) v.x(i) = defer->ansX[(size_t)k * defer->kmax + i];
      return defer->ansOk[k];
    };
    if (stage == 0 && gradients) {
      success = constructAndStandardize(x, false);
      if (success) tier = 0;
      else for (int k = 0; k < n; k++) x[k] = x0[k];
    }
    cw.mark(3, tmark);  // short-circuit through the cached solution (classification + least squares)
    if (!success) {
      if (stage == 0) {
        tier = 1;
        buildReduced(false);
        success = dantzigSolve(v, 0, true);
        if (success) { mapOut(); if (!isValid()) success = false; }
        if (hasNaN()) { success = false; for (int k = 0; k < n; k++) x[k] = 0.0; }
        cw.mark(4, tmark);  // reduce + Dantzig + validity
      }
      if (stage == 2 || (stage == 0 && !success)) {
        tier = 2;
        int pr;
        if (stage == 2) pr = loadAnswer();
        else {
          buildReduced(false);
          pr = runPgs(v);
          cw.mark(5, tmark);  // reduce for PGS
          if (pr < 0) { saveState(g, 2, v.n); out.pending = 1; return out; }
        }
        success = pr != 0;
        if (success) { mapOut(); if (!isValid()) success = false; }
      }
      if (stage == 3 || !success) {
        tier = 3;
        ignoreFrictionFinal = true;
        if (stage == 3) loadAnswer();
        else {
          buildReduced(true);
          if (runPgs(v) < 0) { saveState(g, 3, v.n); out.pending = 1; return out; }
        }
        mapOut();
      }
      if (hasNaN()) { for (int k = 0; k < n; k++) x[k] = 0.0; }
      if (gradients) {
        double* xs = cw.xs();
        for (int k = 0; k < n; k++) xs[k] = x[k];
        if (constructAndStandardize(xs, ignoreFrictionFinal)) { for (int k = 0; k < n; k++) x[k] = xs[k]; }
      }
      cw.mark(6, tmark);  // after the PGS answers: validity, classification + least squares
    }
    if (gradients) {
      for (int k = 0; k < n; k++) cw.activeAll()[rowsTotal + k] = mapping[k] >= 0 ? mapping[k] + rowsTotal : mapping[k];
    }
    out.tier_mask |= 1 << tier;
    // the solution becomes the solver's mX
    mxSize = n;
    for (int k = 0; k < n; k++) mx[k] = x[k];
    // remember the multipliers of this group's rows for the final impulse accumulation
    groupBase[g] = rowsTotal;
    for (int k = 0; k < n; k++) cw.xall()[rowsTotal + k] = x[k];
    rowsTotal += n;
  }
  out.n_rows = rowsTotal;
  if (rowsTotal < 0) out.tier_mask |= 1 << 8;
  *n_cache = mxSize;
  for (int k = 0; k < mxSize && k < ct.kmax; k++) lcp_cache[k] = mx[k];
  if (keepD) {
    const ContactSnapLayout KL{nd, ct.cmax, ct.kmax};
    const int rows = rowsTotal < 0 ? 0 : rowsTotal;
    for (int c = 0; c < nc; c++) {
      for (int z = 0; z < 26; z++) keepD[KL.oContacts() + (size_t)c * 26 + z] = cw.contact(c)[z];
      keepI[KL.oBodyA() + c] = cw.bodyA()[c]; keepI[KL.oBodyB() + c] = cw.bodyB()[c];
      keepI[KL.oDim() + c] = cw.groupOf()[c] >= 0 ? cw.dim()[c] : 0;
    }
    for (int r = 0; r < rows; r++) {
      for (int z = 0; z < nd; z++) keepD[KL.oMassed() + (size_t)r * nd + z] = cw.massed(r)[z];
      keepD[KL.oX() + r] = cw.xall()[r];
      keepI[KL.oMapping() + r] = ct.gradients ? cw.activeAll()[r] : -2;
    }
    for (int z = 0; z < nd; z++) keepD[KL.oVf() + z] = qd_pre_in[vidx(layout, N, nd, w, z)];
    keepI[KL.oCounts()] = nc; keepI[KL.oCounts() + 1] = rows;
  }
  // ---- apply the impulses (ContactConstraint::applyImpulse order: normal, t1, t2 per constraint) ----
  for (int i = 0; i < nb; i++) for (int z = 0; z < 6; z++) cw.imp(i)[z] = 0.0;
  bool any = false;
  for (int c = 0; c < nc; c++) {
    const int dim = cw.dim()[c];
    if (!dim || cw.groupOf()[c] < 0) continue;
    any = true;
    const int a = cw.bodyA()[c], bb = cw.bodyB()[c];
    const double* xr = cw.xall() + groupBase[cw.groupOf()[c]] + offs[c];
    for (int k = 0; k < dim; k++) {
      const double lam = xr[k];
      if (ct.shape[a].reactive) for (int z = 0; z < 6; z++) cw.imp(a)[z] += cw.NA(c, k)[z] * lam;
      if (ct.shape[bb].reactive) for (int z = 0; z < 6; z++) cw.imp(bb)[z] += cw.NB(c, k)[z] * lam;
    }
  }
  for (int z = 0; z < nd; z++) cw.dqd()[z] = 0.0;
  if (any) impulsePass(sn, cw.pad0(), nb);
  for (int z = 0; z < nd; z++) {
    const size_t ix = vidx(layout, N, nd, w, z);
    qd_out[ix] = qd_pre_in[ix] + cw.dqd()[z];
  }
  cw.mark(7, tmark);  // commit, per-slot record, final impulse pass
  return out;
}

}  // namespace ddb
