// contact_bwd.cuh — BackpropSnapshot::backprop through a timestep WITH contacts, as one vector-Jacobian product per world
// (one world per thread, FP64). What it replaces in the reference:
//   BackpropSnapshot::getVelJacobianWrt              neural/BackpropSnapshot.cpp:1002-1076
//   getJacobianOfConstraintForce                     :3853-3894   (dF_c = Q^+ dB + d(Q^+ b))
//   getJacobianOfLCPConstraintMatrixClampingSubset   :3898-4055   (incl. the pseudo-inverse derivative for rank-deficient Q)
//   getJacobianOfLCPOffsetClampingSubset             :4252-4299
//   getJacobianOfClampingConstraints(Transpose) / UpperBound   :5005-5091
//   DifferentiableContactConstraint::getConstraintForcesJacobian and the contact geometry gradients behind it
//                                                    neural/DifferentiableContactConstraint.cpp:113-267,307-1107,1205-1293,1433-1576
//   world screw axes                                 dynamics/Joint.cpp:389-410, dynamics/FreeJoint.cpp:703-733
// With A = A_c + A_ub E, Q = A_c^T M^-1 A, P = Q^+, a = A^T M^-1 lambda and y1 = P^T a, every reference Jacobian applied
// to lambda = dL/dqd' collapses to
//   lambda~ = lambda - A_c y1,  dL/dtau = dt M^-1 lambda~,  dL/dqd = lambda~ - dt (dC/dqd)^T M^-1 lambda~ (+ velPos^T dL/dq'),
//   dL/dq   = posPos^T dL/dq' - dt d/dq[mu~^T ID(q, qd, qdd_f)] - sum_pairs mu_a^T (dM/dq) a_b + sum_rows (dA_row/dq)^T V_row,
// where the first q-term is the reverse sweep of the contact-free backward with the post-contact acceleration, the pairs
// are the extra mass-matrix terms of the pseudo-inverse derivative, and V_row collects, per clamping / upper-bound row,
// every vector that multiplies that row's constraint-force Jacobian. The reference evaluates the mass-matrix and Coriolis
// derivatives by finite differences (BackpropSnapshot.cpp:5990-6027,6121-6157); here they are closed-form sweeps.
// Restitution: with B = diag(1 + e_j) over the clamping rows that bounced (getBounceDiagonals, ConstrainedGroupGradientMatrices.cpp:753)
// the LCP offset is b_c = -B A_c^T v_f and its Jacobians carry the same factor (BackpropSnapshot.cpp:4252-4299), so wherever y1 meets
// dB it is yB = B y1: lambda~ = lambda - A_c yB, the row vectors use -yB_j v_f, M^-1 lambda = mu~ + M^-1 A_c yB, and the bilinear
// mass-matrix terms gain the pair (M^-1 A_c (yB - y1), M^-1 A f_c). The terms of d(Q^+) b keep the plain y1.
#pragma once
#include "contact.cuh"
#include "kernels.cuh"

namespace ddb {

// transient per-world scratch of the contact backward (world-major)
struct ContactBwdLayout {
  int nb, nd, kmax;
  __host__ __device__ size_t oW() const { return 0; }                                  // [nb][12] world transforms
  __host__ __device__ size_t oSF() const { return oW() + (size_t)nb * 12; }            // [nd][6] world screws (force / velocity)
  __host__ __device__ size_t oSP() const { return oSF() + (size_t)nd * 6; }            // [nd][6] world screws (position)
  __host__ __device__ size_t oF() const { return oSP() + (size_t)nd * 6; }             // [kmax][6] world force of each classified row
  __host__ __device__ size_t oAc() const { return oF() + (size_t)kmax * 6; }           // [kmax][nd] constraint-force columns
  __host__ __device__ size_t oMA() const { return oAc() + (size_t)kmax * nd; }         // [kmax][nd] M^-1 (A_c + A_ub E) columns
  __host__ __device__ size_t oCV() const { return oMA() + (size_t)kmax * nd; }         // 11 x [kmax]: a b f y1 r1 y2 r2 y3 r3 E yB
  __host__ __device__ size_t oNV() const { return oCV() + (size_t)11 * kmax; }         // 13 x [nd]
  __host__ __device__ size_t oPriv() const { return oNV() + (size_t)13 * nd; }         // 2 x [32][nd]: per-lane V_row and dL/dq contributions of one round of rows
  __host__ __device__ size_t doubles() const { return oPriv() + (size_t)2 * 32 * nd; }
  __host__ __device__ size_t ints() const { return (size_t)4 * kmax; }                 // clampRow ubRow ubNormal clampIdxOfRow
};

DDB_DI SV adT(const Xf& T, SV V) {  // Geometry.cpp AdT: [R w; R v + p x R w]
  V3 rw = mul(T.R, V.w);
  return sv(rw, mul(T.R, V.v) + cross(T.p, rw));
}
DDB_DI SV ldSVp(const double* p) { return sv(mk(p[0], p[1], p[2]), mk(p[3], p[4], p[5])); }
DDB_DI void stSVp(double* p, SV a) { p[0] = a.w.x; p[1] = a.w.y; p[2] = a.w.z; p[3] = a.v.x; p[4] = a.v.y; p[5] = a.v.z; }
DDB_DI V3 ldV3(const double* p) { return mk(p[0], p[1], p[2]); }

enum { CR_NONE = 0, CR_VERTEX, CR_FACE, CR_EDGE_A, CR_EDGE_B, CR_SELF, CR_SPHERE_A, CR_SPHERE_B, CR_SPHERE_TO_BOX, CR_BOX_TO_SPHERE };

// DifferentiableContactConstraint::getDofContactType for a dof that moves body A (sideA) or body B
DDB_DI int contactRole(int type, bool sideA) {
  switch (type) {
    case 1: return sideA ? CR_VERTEX : CR_FACE;
    case 2: return sideA ? CR_FACE : CR_VERTEX;
    case 3: return sideA ? CR_EDGE_A : CR_EDGE_B;
    case 4: return sideA ? CR_SPHERE_TO_BOX : CR_BOX_TO_SPHERE;
    case 5: return sideA ? CR_BOX_TO_SPHERE : CR_SPHERE_TO_BOX;
    case 6: return sideA ? CR_SPHERE_A : CR_SPHERE_B;
    default: return CR_NONE;
  }
}
DDB_DI V3 thetaGrad(SV s, V3 p) {  // math::gradientWrtTheta(s, p, 0)
  if (sqrt(dot(s.w, s.w)) > 1e-6) {
    const V3 wwv = cross(s.w, cross(s.w, s.v));
    return (cross(s.w, p) - wwv) + s.v + wwv;
  }
  return s.v;
}
DDB_DI V3 edgePointGradient(V3 pa, V3 dpa, V3 ua, V3 dua, V3 pb, V3 dpb, V3 ub, V3 dub) {  // math::getContactPointGradient, radii 1
  const V3 p = pb - pa, dp = dpb - dpa;
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
// contact record: point 0-2, normal 3-5, edgeAClosest 6-8, edgeAFixed 9-11, edgeADir 12-14, edgeBClosest 15-17, edgeBFixed 18-20,
// edgeBDir 21-23, depth 24, type 25 (sphere cases reuse the edge slots, see collide.cuh)
DDB_DI V3 projectLockedFaces(const double* ct, V3 g) {
  const int locks = (int)ct[13];
  for (int k = 0; k < 3; k++)
    if (locks & (1 << k)) { const V3 fn = ldV3(ct + 15 + 3 * k); g = g - dot(fn, g) * fn; }
  return g;
}
DDB_DI V3 contactPositionGradient(const double* ct, int role, SV sp) {
  const V3 point = ldV3(ct), eAF = ldV3(ct + 9), eAD = ldV3(ct + 12), eBF = ldV3(ct + 18), eBD = ldV3(ct + 21);
  switch (role) {
    case CR_VERTEX: case CR_SELF: return thetaGrad(sp, point);
    case CR_EDGE_A: return edgePointGradient(eAF, thetaGrad(sp, eAF), eAD, cross(sp.w, eAD), eBF, mk(0, 0, 0), eBD, mk(0, 0, 0));
    case CR_EDGE_B: return edgePointGradient(eAF, mk(0, 0, 0), eAD, mk(0, 0, 0), eBF, thetaGrad(sp, eBF), eBD, cross(sp.w, eBD));
    case CR_SPHERE_A: { const double rA = ct[12], rB = ct[21]; return (rB / (rA + rB)) * thetaGrad(sp, eAF); }
    case CR_SPHERE_B: { const double rA = ct[12], rB = ct[21]; return (rA / (rA + rB)) * thetaGrad(sp, eBF); }
    case CR_SPHERE_TO_BOX: return projectLockedFaces(ct, thetaGrad(sp, eAF));
    case CR_BOX_TO_SPHERE: return thetaGrad(sp, point) + projectLockedFaces(ct, -thetaGrad(sp, eAF));
    default: return mk(0, 0, 0);
  }
}
DDB_DI V3 contactNormalGradient(const double* ct, int role, SV sp) {
  const V3 point = ldV3(ct), n = ldV3(ct + 3), eAF = ldV3(ct + 9), eAD = ldV3(ct + 12), eBF = ldV3(ct + 18), eBD = ldV3(ct + 21);
  const int type = (int)ct[25];
  switch (role) {
    case CR_FACE: case CR_SELF: return cross(sp.w, n);
    case CR_EDGE_A: { const double sign = dot(cross(eBD, eAD), n) < 0 ? -1.0 : 1.0; return sign * cross(eBD, cross(sp.w, eAD)); }
    case CR_EDGE_B: { const double sign = dot(cross(eBD, eAD), n) < 0 ? -1.0 : 1.0; return sign * cross(cross(sp.w, eBD), eAD); }
    case CR_SPHERE_A: case CR_SPHERE_B: {
      const V3 dd = eAF - eBF;
      const double nn = sqrt(dot(dd, dd));
      V3 g = (1.0 / nn) * thetaGrad(sp, role == CR_SPHERE_A ? eAF : eBF);
      g = g - dot(n, g) * n;
      return role == CR_SPHERE_A ? g : -g;
    }
    case CR_SPHERE_TO_BOX: case CR_BOX_TO_SPHERE: {
      const V3 dd = eAF - point;
      const double nn = sqrt(dot(dd, dd));
      V3 cp = contactPositionGradient(ct, role, sp), sg = role == CR_SPHERE_TO_BOX ? thetaGrad(sp, eAF) : mk(0, 0, 0);
      if (nn > 1e-5) { cp = (1.0 / nn) * cp; sg = (1.0 / nn) * sg; }
      V3 t = type == 5 ? cp - sg : sg - cp;
      return t - dot(t, n) * n;
    }
    default: return mk(0, 0, 0);
  }
}
// ContactConstraint::getTangentBasisMatrixODEGradient (constraint/ContactConstraint.cpp:783-858), column col
DDB_DI V3 tangentBasisGradient(V3 n, V3 g, int col) {
  V3 cr = mk(0, 0, 1), t = cross(cr, n);
  if (dot(t, t) < 1e-12) {
    cr = mk(1, 0, 0); t = cross(cr, n);
    if (dot(t, t) < 1e-12) {
      cr = mk(0, 1, 0); t = cross(cr, n);
      if (dot(t, t) < 1e-12) { cr = mk(0, 0, 1); t = cross(cr, n); }
    }
  }
  const double tn = sqrt(dot(t, t));
  t = (1.0 / tn) * t;
  V3 gc = (1.0 / tn) * cross(cr, g), gt;
  if (fabs(tn - 1.0) > 1e-6) gt = gc - dot(gc, t) * t; else gt = gc;
  if (col == 0) return gt;
  return cross(g, t) + cross(n, gt);
}
// getContactWorldForceGradient for direction index k of the contact
DDB_DI SV worldForceGradient(const double* ct, int k, V3 dir, int role, SV sp) {
  if (role == CR_NONE) return svzero();
  V3 fg = mk(0, 0, 0);
  if (role != CR_VERTEX) {
    const V3 ng = contactNormalGradient(ct, role, sp);
    if (k == 0 || dot(ng, ng) <= 1e-12) fg = ng;
    else fg = tangentBasisGradient(ldV3(ct + 3), ng, k - 1);
  }
  const V3 pg = contactPositionGradient(ct, role, sp);
  return sv(cross(ldV3(ct), fg) + cross(pg, dir), fg);
}

struct ContactBwdWork {
  double* d;
  int* i;
  ContactBwdLayout L;
  DDB_DI double* W(int body) const { return d + L.oW() + (size_t)body * 12; }
  DDB_DI double* sF(int dof) const { return d + L.oSF() + (size_t)dof * 6; }
  DDB_DI double* sP(int dof) const { return d + L.oSP() + (size_t)dof * 6; }
  DDB_DI double* F(int slot) const { return d + L.oF() + (size_t)slot * 6; }
  DDB_DI double* Ac(int slot) const { return d + L.oAc() + (size_t)slot * L.nd; }
  DDB_DI double* MA(int j) const { return d + L.oMA() + (size_t)j * L.nd; }
  DDB_DI double* cv(int which) const { return d + L.oCV() + (size_t)which * L.kmax; }
  DDB_DI double* nv(int which) const { return d + L.oNV() + (size_t)which * L.nd; }
  DDB_DI double* privV(int lane) const { return d + L.oPriv() + (size_t)lane * L.nd; }
  DDB_DI double* privG(int lane) const { return d + L.oPriv() + (size_t)(32 + lane) * L.nd; }
  DDB_DI int* clampRow() const { return i; }
  DDB_DI int* ubRow() const { return i + L.kmax; }
  DDB_DI int* ubNormal() const { return i + 2 * L.kmax; }
  DDB_DI int* clampIdxOfRow() const { return i + 3 * L.kmax; }
};
enum { CV_A = 0, CV_B, CV_F, CV_Y1, CV_R1, CV_Y2, CV_R2, CV_Y3, CV_R3, CV_E, CV_YB };
enum { NV_LAMT = 0, NV_MUT, NV_M1, NV_M2, NV_M3, NV_M4, NV_M5, NV_M6, NV_M7, NV_V, NV_DA, NV_VFM1, NV_M8 };

// sum over the joints between `body` and its root: g_q[k] += sign * ad(sP_k, Y) . F + Wtot . dF_k, with Y the running sum of
// V_i sF_i over the dofs at or below the joint of k (getConstraintForcesJacobian's row sum folded over V)
DDB_DI void chainGeometryVjp(const ContactBwdWork& bw, const double* ct, int k, V3 dir, SV F, SV Wtot, const double* V, int body, bool sideA,
                             double sign, double* gq) {
  SV Y = svzero();
  const int role = contactRole((int)ct[25], sideA);
  for (int j = body; j >= 0; j = c_topo.b[j].parent) {
    const DevBody& b = c_topo.b[j];
    for (int z = 0; z < b.ndof; z++) Y = Y + V[b.dof + z] * ldSVp(bw.sF(b.dof + z));
    for (int z = 0; z < b.ndof; z++) {
      const SV sp = ldSVp(bw.sP(b.dof + z));
      const SV dF = worldForceGradient(ct, k, dir, role, sp);
      gq[b.dof + z] += sign * dot(ad(sp, Y), F) + dot(Wtot, dF);
    }
  }
}

template <int NB>
struct ContactBwdLocals {  // one object: see the note on BwdScratch (stack colouring)
  BwdScratch<NB> ws;
  SV dA[NB];
};

// Meeting points of the warps of a CTA (k_contact_backward with 8 warps per CTA; see ContactWork::mark in contact.cuh): the kernel is
// bound by instruction fetches and warps at the same place of the code share them. A per-warp progress word in shared memory and a
// bounded spin: a scheduling hint, nothing depends on it; a warp that has left reads as "ahead".
#ifndef DDB_CONVOY_SLEEP
#define DDB_CONVOY_SLEEP 100
#endif
constexpr int kConvoyMinRows = 8;
struct BwdConvoy {
  volatile int* reached = nullptr;
  int warps = 0;
  DDB_DI void leave() {  // this warp does not take part (any more): it reads as "ahead" of every meeting point
#ifndef DDB_HOST_EMULATION
    if (reached) reached[threadIdx.x >> 5] = 0x7fffffff;
    reached = nullptr;
#endif
  }
  DDB_DI void meet(int id) const {
#ifndef DDB_HOST_EMULATION
    if (!reached) return;
    reached[threadIdx.x >> 5] = id;
    __syncwarp();
    for (int spins = 0; spins < 4000; spins++) {
      int behind = 0;
      for (int k = 0; k < warps; k++) behind |= reached[k] < id;
      if (!behind) break;
      __nanosleep(DDB_CONVOY_SLEEP);
    }
    __syncwarp();
#endif
  }
};

// status bits written per world: 1 = contact rows present, 2 = pseudo-inverse branch (rank-deficient Q).
// One world per warp (lcpStride == 1): the tree sweeps run redundantly on the 32 lanes; the loops over independent rows, matrix
// entries and DOFs (constraint-force columns, Q, the n-vectors, the contact-geometry term of every classified row) are split across
// the lanes WITHOUT changing any summation order: a lane owns whole entries, and the per-row contributions to dL/dq are added in row
// order from per-lane buffers, so the result is the one-thread result bit for bit.
template <int NB>
DDB_DI int contactBackwardWorld(const Snap& sn, const ContactSnapLayout& KL, const double* keepD, const int* keepI, const ContactBwdWork& bw,
                                double* lcpD, int* lcpI, const double* gqn, const double* gqdn, double* gq, double* gqd, double* gtau,
                                int layout, int N, int w, int lcpStride = 32, BwdConvoy convoy = BwdConvoy()) {
  const int nb = c_topo.nb, nd = c_topo.nd;
  const double dt = c_topo.dt;
  const ContactTopo& ctp = c_ctopo;
  int coopLane = 0, coopN = 1;
#ifndef DDB_HOST_EMULATION
  if (lcpStride == 1) { coopLane = threadIdx.x & 31; coopN = 32; }
#endif
  ContactBwdLocals<NB> loc;
  BwdScratch<NB>& ws = loc.ws;
  SV* dA = loc.dA;
  int status = 0;
  const int rows = keepI[KL.oCounts() + 1];
  const int* mapping = keepI + KL.oMapping();
  const int* rowInfo = keepI + KL.oRowInfo();
  const double* xall = keepD + KL.oX();
  const double* vf = keepD + KL.oVf();
  const double* restOfRow = keepD + KL.oRest();
  auto massed = [&](int r) { return keepD + KL.oMassed() + (size_t)r * nd; };
  auto contactOf = [&](int r) { return keepD + KL.oContacts() + (size_t)(rowInfo[r] >> 2) * 26; };
  // ---- classification: clamping rows in row order, then upper-bound rows ----
  int c = 0, u = 0;
  for (int r = 0; r < rows; r++) { bw.clampIdxOfRow()[r] = -1; if (mapping[r] == -1) { bw.clampIdxOfRow()[r] = c; bw.clampRow()[c++] = r; } }
  for (int r = 0; r < rows; r++)
    if (mapping[r] >= 0) {
      const int fp = mapping[r], cons = rowInfo[r] >> 2;
      const int a = keepI[KL.oBodyA() + cons], bb = keepI[KL.oBodyB() + cons];
      const double mu = fmin(ctp.shape[a].friction, ctp.shape[bb].friction);
      const double upper = xall[fp] * mu, lower = -xall[fp] * mu;
      bw.ubRow()[u] = r; bw.ubNormal()[u] = bw.clampIdxOfRow()[fp];
      bw.cv(CV_E)[u] = (fabs(xall[r] - upper) < fabs(xall[r] - lower)) ? mu : -mu;
      u++;
    }
  if (c + u < kConvoyMinRows) convoy.leave();  // (a world of one or two contacts is through before a meeting pays off)
  double* lamT = bw.nv(NV_LAMT);
  for (int d = 0; d < nd; d++) lamT[d] = gqdn[vidx(layout, N, nd, w, d)];
  bool pinv = false, fcDiffers = false, bounced = false;
  if (c > 0) {
    status |= 1;
    // ---- world transforms and screws ----
    for (int i = 0; i < nb; i++) {
      const DevBody& b = c_topo.b[i];
      const Xf t = sn.ldT(i);
      double* Wi = bw.W(i);
      if (b.parent >= 0) {
        const double* Wp = bw.W(b.parent);
        for (int r = 0; r < 3; r++) {
          for (int cc = 0; cc < 3; cc++) Wi[3 * r + cc] = Wp[3 * r] * t.R.m[cc] + Wp[3 * r + 1] * t.R.m[3 + cc] + Wp[3 * r + 2] * t.R.m[6 + cc];
          Wi[9 + r] = Wp[3 * r] * t.p.x + Wp[3 * r + 1] * t.p.y + Wp[3 * r + 2] * t.p.z + Wp[9 + r];
        }
      } else {
        for (int k = 0; k < 9; k++) Wi[k] = t.R.m[k];
        Wi[9] = t.p.x; Wi[10] = t.p.y; Wi[11] = t.p.z;
      }
      Xf Wx;
      for (int k = 0; k < 9; k++) Wx.R.m[k] = Wi[k];
      Wx.p = mk(Wi[9], Wi[10], Wi[11]);
      if (b.type == DDB_JOINT_FREE) {
        Xf Wp, Tpj, tr;
        if (b.parent >= 0) { const double* P = bw.W(b.parent); for (int k = 0; k < 9; k++) Wp.R.m[k] = P[k]; Wp.p = mk(P[9], P[10], P[11]); }
        else { for (int k = 0; k < 9; k++) Wp.R.m[k] = (k % 4 == 0) ? 1.0 : 0.0; Wp.p = mk(0, 0, 0); }
        for (int k = 0; k < 9; k++) { Tpj.R.m[k] = b.N0[k]; tr.R.m[k] = (k % 4 == 0) ? 1.0 : 0.0; }
        Tpj.p = mk(b.P0[0], b.P0[1], b.P0[2]);
        double q6[6];
        ld6(sn, i, 56, q6);
        tr.p = mk(q6[3], q6[4], q6[5]);
        const M3 Jr = expRightJac(mk(q6[0], q6[1], q6[2]));  // expMapJac(q)^T: column k of expMapJac = row k of Jr
        for (int k = 0; k < 6; k++) {
          double col[6];
          for (int a = 0; a < 6; a++) col[a] = b.S6[a * 6 + k];
          stSVp(bw.sF(b.dof + k), adT(Wx, sv6(col)));
          SV g = svzero();
          if (k < 3) g = adT(tr, sv(mk(Jr.m[3 * k], Jr.m[3 * k + 1], Jr.m[3 * k + 2]), mk(0, 0, 0)));
          else g.v = mk(k == 3 ? 1.0 : 0.0, k == 4 ? 1.0 : 0.0, k == 5 ? 1.0 : 0.0);
          stSVp(bw.sP(b.dof + k), adT(Wp, adT(Tpj, g)));
        }
      } else if (b.type != DDB_JOINT_WELD) {
        const SV s = adT(Wx, b.S);
        stSVp(bw.sF(b.dof), s);
        stSVp(bw.sP(b.dof), s);
      }
    }
    convoy.meet(10);
    // ---- world force and constraint-force column of every classified row (clamping slots 0..c-1, upper-bound slots c..c+u-1) ----
    coopSyncWarp(coopN);
    for (int sbase = 0; sbase < c + u; sbase += coopN) {
      if (sbase) convoy.meet(10 + (sbase / coopN < 9 ? sbase / coopN : 9));
      const int slot = sbase + coopLane;
      if (slot >= c + u) continue;
      const int r = slot < c ? bw.clampRow()[slot] : bw.ubRow()[slot - c];
      const double* ct = contactOf(r);
      const int k = rowInfo[r] & 3, cons = rowInfo[r] >> 2;
      double dirs[3][3];
      for (int x = 0; x < 3; x++) dirs[0][x] = ct[3 + x];
      if (k > 0) tangentBasis(dirs[0], dirs[1], dirs[2]);
      const V3 dir = ldV3(dirs[k]);
      const SV F = sv(cross(ldV3(ct), dir), dir);
      stSVp(bw.F(slot), F);
      double* col = bw.Ac(slot);
      for (int d = 0; d < nd; d++) col[d] = 0.0;
      for (int side = 0; side < 2; side++) {
        const int body = keepI[(side == 0 ? KL.oBodyA() : KL.oBodyB()) + cons];
        for (int j = body; j >= 0; j = c_topo.b[j].parent) {
          const DevBody& b = c_topo.b[j];
          for (int z = 0; z < b.ndof; z++) col[b.dof + z] += (side == 0 ? 1.0 : -1.0) * dot(ldSVp(bw.sF(b.dof + z)), F);
        }
      }
    }
    convoy.meet(20);
    // ---- M^-1 A columns, b_c, f_c, a = A^T M^-1 lambda ----
    double *va = bw.cv(CV_A), *vb = bw.cv(CV_B), *vfc = bw.cv(CV_F), *y1 = bw.cv(CV_Y1), *r1 = bw.cv(CV_R1);
    double *y2 = bw.cv(CV_Y2), *r2 = bw.cv(CV_R2), *y3 = bw.cv(CV_Y3), *r3 = bw.cv(CV_R3);
    coopSyncWarp(coopN);
    for (int dbase = 0; dbase < nd; dbase += coopN) {  // a lane owns DOF d of every column (the upper-bound rows add in their order)
      const int d = dbase + coopLane;
      if (d >= nd) continue;
      for (int j = 0; j < c; j++) bw.MA(j)[d] = massed(bw.clampRow()[j])[d];
      for (int k = 0; k < u; k++) bw.MA(bw.ubNormal()[k])[d] += bw.cv(CV_E)[k] * massed(bw.ubRow()[k])[d];
    }
    coopSyncWarp(coopN);
    for (int j = 0; j < c; j++) if (restOfRow[bw.clampRow()[j]] != 0.0) bounced = true;
    for (int jbase = 0; jbase < c; jbase += coopN) {
      const int j = jbase + coopLane;
      if (j >= c) continue;
      double sb = 0.0, sa = 0.0;
      const double* col = bw.Ac(j);
      const double* ma = bw.MA(j);
      for (int d = 0; d < nd; d++) { sb += col[d] * vf[d]; sa += ma[d] * lamT[d]; }
      const double Bj = 1.0 + restOfRow[bw.clampRow()[j]];
      vb[j] = -(Bj * sb); va[j] = sa; vfc[j] = xall[bw.clampRow()[j]];
    }
    convoy.meet(30);
    // ---- Q = A_c^T M^-1 A and the pseudo-inverse products ----
    LcpView v;
    v.init(lcpD, lcpI, c, ctp.kmax, lcpStride);
    coopSyncWarp(coopN);
    for (int ebase = 0; ebase < c * c; ebase += coopN) {  // a lane owns whole entries of Q
      const int e = ebase + coopLane;
      if (e >= c * c) continue;
      const int i = e / c, j = e % c;
      const double* col = bw.Ac(i);
      const double* ma = bw.MA(j);
      double s = 0.0;
      for (int d = 0; d < nd; d++) s += col[d] * ma[d];
      v.Araw(i, j) = s;
      // one decomposition of Q serves P b, P^T a, P y1 and P^T r1 (BackpropSnapshot.cpp:3875-3876,3925-3932)
      v.db[(size_t)(v.oL + i * c + j) * v.st] = s;
    }
    coopSyncWarp(coopN);
    convoy.meet(31);
    const int rank = codFactorDev(v, c);
    convoy.meet(32);
    codApplyDev(v, c, rank, true, [&](int i) { return va[i]; }, [&](int i, double x) { y1[i] = x; });
    codApplyDev(v, c, rank, false, [&](int i) { return vb[i]; }, [&](int i, double x) { r1[i] = x; });
    pinv = rank < c;
    if (pinv) {
      status |= 2;
      coopSyncWarp(coopN);
      for (int ibase = 0; ibase < c; ibase += coopN) {
        const int i = ibase + coopLane;
        if (i >= c) continue;
        double s1 = 0.0, s2 = 0.0;
        for (int j = 0; j < c; j++) { s1 += v.Araw(i, j) * r1[j]; s2 += v.Araw(j, i) * y1[j]; }
        r2[i] = vb[i] - s1; y3[i] = va[i] - s2;
      }
      coopSyncWarp(coopN);
      codApplyDev(v, c, rank, false, [&](int i) { return y1[i]; }, [&](int i, double x) { y2[i] = x; });
      codApplyDev(v, c, rank, true, [&](int i) { return r1[i]; }, [&](int i, double x) { r3[i] = x; });
    } else {
      for (int i = 0; i < c; i++) { r2[i] = y2[i] = r3[i] = y3[i] = 0.0; }
    }
    double fcScale = 0.0, fcDiff = 0.0;
    for (int j = 0; j < c; j++) { fcScale = fmax(fcScale, fabs(vfc[j])); fcDiff = fmax(fcDiff, fabs(vfc[j] - r1[j])); }
    fcDiffers = fcDiff > 1e-13 * fmax(1.0, fcScale);
    convoy.meet(40);
    // ---- lambda~ and the n-vectors m1..m7, delta-acceleration ----
    double *m1 = bw.nv(NV_M1), *m2 = bw.nv(NV_M2), *m3 = bw.nv(NV_M3), *m4 = bw.nv(NV_M4), *m5 = bw.nv(NV_M5), *m6 = bw.nv(NV_M6), *m7 = bw.nv(NV_M7);
    double* da = bw.nv(NV_DA);
    double* m8 = bw.nv(NV_M8);   // M^-1 A_c (yB - y1): zero without restitution
    double* yB = bw.cv(CV_YB);
    for (int j = 0; j < c; j++) yB[j] = (1.0 + restOfRow[bw.clampRow()[j]]) * y1[j];  // (every lane: the stores coincide)
    coopSyncWarp(coopN);
    for (int dbase = 0; dbase < nd; dbase += coopN) {  // a lane owns DOF d of every n-vector; the rows add in their order
      const int d = dbase + coopLane;
      if (d >= nd) continue;
      double lt = lamT[d], a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, a6 = 0.0, a7 = 0.0, a8 = 0.0, ad = 0.0;
      for (int j = 0; j < c; j++) {
        const double col = bw.Ac(j)[d], ma = bw.MA(j)[d], mc = massed(bw.clampRow()[j])[d];
        const double e = restOfRow[bw.clampRow()[j]];
        lt -= yB[j] * col;
        a8 += (e * y1[j]) * mc;
        a1 += r1[j] * ma; a2 += y1[j] * mc; a3 += r2[j] * mc; a4 += y2[j] * ma;
        a5 += r3[j] * mc; a6 += y3[j] * ma; a7 += (vfc[j] - r1[j]) * ma;
        ad += vfc[j] * ma;
      }
      lamT[d] = lt; m1[d] = a1; m2[d] = a2; m3[d] = a3; m4[d] = a4; m5[d] = a5; m6[d] = a6; m7[d] = a7; m8[d] = a8;
      da[d] = ad * (1.0 / dt);
    }
    coopSyncWarp(coopN);
  }
  convoy.meet(50);
  // ---- mu~ = M^-1 lambda~, dL/dtau ----
  double* muT = bw.nv(NV_MUT);
  solveInvMass<NB, false>(sn, nb, [&](int d) { return lamT[d]; }, ws, [&](int d, double m) { muT[d] = m; });
  for (int i = 0; i < nb; i++) if (c_topo.b[i].ndof == 1) muT[c_topo.b[i].dof] = ws.mu[i];
  for (int d = 0; d < nd; d++) gtau[vidx(layout, N, nd, w, d)] = dt * muT[d];
  convoy.meet(60);
  // ---- reverse sweep over the inverse dynamics with the post-contact acceleration ----
  if (c > 0) twistsFromRates<NB>(sn, nb, [&](int d) { return bw.nv(NV_DA)[d]; }, dA, nullptr);
  else for (int i = 0; i < nb; i++) dA[i] = svzero();
  idAdjointEx<NB, false, 1>(sn, nb, ws, dA, [&](int d, double dq, double dqd) {
    const size_t ix = vidx(layout, N, nd, w, d);
    const double a = gqn[ix];
    gq[ix] = fma(-dt, dq, a);
    gqd[ix] = fma(dt, a, fma(-dt, dqd, lamT[d]));
  }, [&](int body, const double* dq6, const double* dqd6) {
    const DevBody& b = c_topo.b[body];
    double q6[6], qd6[6], pp[36], vp[36], a6[6];
    ld6(sn, body, 56, q6); ld6(sn, body, 48, qd6);
    freeJointPosJacobians(q6, qd6, dt, pp, vp);
    for (int k = 0; k < 6; k++) a6[k] = gqn[vidx(layout, N, nd, w, b.dof + k)];
    for (int k = 0; k < 6; k++) {
      double sp = 0.0, sv2 = 0.0;
      for (int r = 0; r < 6; r++) { sp += pp[r * 6 + k] * a6[r]; sv2 += vp[r * 6 + k] * a6[r]; }
      const size_t ix = vidx(layout, N, nd, w, b.dof + k);
      gq[ix] = sp - dt * dq6[k];
      gqd[ix] = sv2 + lamT[b.dof + k] - dt * dqd6[k];
    }
  });
  if (c == 0) return status;
  convoy.meet(70);
  // ---- extra mass-matrix pairs: dL/dq -= mu_a^T (dM/dq) a_b ----
  if (bounced) for (int d = 0; d < nd; d++) bw.nv(NV_DA)[d] *= dt;  // M^-1 A f_c (the reverse sweep above is done with it)
  for (int pair = 0; pair < 4; pair++) {
    if (pair) convoy.meet(70 + pair);
    if (pair == 0 && !fcDiffers) continue;
    if ((pair == 1 || pair == 2) && !pinv) continue;
    if (pair == 3 && !bounced) continue;
    const double* ma = pair == 0 ? bw.nv(NV_M2) : (pair == 1 ? bw.nv(NV_M3) : (pair == 2 ? bw.nv(NV_M5) : bw.nv(NV_M8)));
    const double* mb = pair == 0 ? bw.nv(NV_M7) : (pair == 1 ? bw.nv(NV_M4) : (pair == 2 ? bw.nv(NV_M6) : bw.nv(NV_DA)));
    twistsFromRates<NB>(sn, nb, [&](int d) { return ma[d]; }, ws.U, ws.mu);
    twistsFromRates<NB>(sn, nb, [&](int d) { return mb[d]; }, dA, nullptr);
    idAdjointEx<NB, false, 2>(sn, nb, ws, dA, [&](int d, double dq, double) { gq[vidx(layout, N, nd, w, d)] -= dq; },
                              [&](int body, const double* dq6, const double*) {
                                const int d0 = c_topo.b[body].dof;
                                for (int k = 0; k < 6; k++) gq[vidx(layout, N, nd, w, d0 + k)] -= dq6[k];
                              });
  }
  convoy.meet(80);
  // ---- contact geometry: dL/dq += sum_rows (dA_row/dq)^T V_row ----
  {
    const double *y1 = bw.cv(CV_Y1), *r1 = bw.cv(CV_R1), *y2 = bw.cv(CV_Y2), *r2 = bw.cv(CV_R2), *y3 = bw.cv(CV_Y3), *r3 = bw.cv(CV_R3), *vfc = bw.cv(CV_F);
    const double *m1 = bw.nv(NV_M1), *m2 = bw.nv(NV_M2), *m3 = bw.nv(NV_M3), *m4 = bw.nv(NV_M4), *m5 = bw.nv(NV_M5), *m6 = bw.nv(NV_M6), *m8 = bw.nv(NV_M8);
    const double* yB = bw.cv(CV_YB);
    double* acc = bw.nv(NV_VFM1);  // dL/dq contributions of the geometry, accumulated locally
    for (int d = 0; d < nd; d++) acc[d] = 0.0;  // (every lane: the stores coincide)
    coopSyncWarp(coopN);
    // rows in rounds of coopN: lane L works out the contribution of row base + L into its private buffer, then the lanes add the
    // buffers DOF by DOF in row order (a row touches a DOF at most once: its two bodies belong to different skeletons)
    for (int sbase = 0; sbase < c + u; sbase += coopN) {
      if (sbase) convoy.meet(80 + (sbase / coopN < 9 ? sbase / coopN : 9));
      const int slot = sbase + coopLane;
      double* V = bw.privV(coopLane);
      double* G = bw.privG(coopLane);
      for (int d = 0; d < nd; d++) G[d] = 0.0;
      if (slot < c + u) {
        const bool isUb = slot >= c;
        const int r = isUb ? bw.ubRow()[slot - c] : bw.clampRow()[slot];
        const int j = isUb ? bw.ubNormal()[slot - c] : slot;
        const double e = isUb ? bw.cv(CV_E)[slot - c] : 1.0;
        for (int d = 0; d < nd; d++) {
          const double mu = muT[d] + m2[d] + m8[d];  // M^-1 lambda = mu~ + M^-1 A_c yB
          double val = vfc[j] * mu - r1[j] * m2[d] + y2[j] * m3[d] + y3[j] * m5[d];
          if (!isUb) val += -yB[j] * vf[d] - y1[j] * m1[d] + r2[j] * m4[d] + r3[j] * m6[d];
          V[d] = e * val;
        }
        const double* ct = contactOf(r);
        const int k = rowInfo[r] & 3, cons = rowInfo[r] >> 2;
        const int bodyA = keepI[KL.oBodyA() + cons], bodyB = keepI[KL.oBodyB() + cons];
        const SV F = ldSVp(bw.F(slot));
        const V3 dir = F.v;
        SV Wtot = svzero();
        for (int side = 0; side < 2; side++)
          for (int jb = side == 0 ? bodyA : bodyB; jb >= 0; jb = c_topo.b[jb].parent) {
            const DevBody& b = c_topo.b[jb];
            for (int z = 0; z < b.ndof; z++) Wtot = Wtot + ((side == 0 ? 1.0 : -1.0) * V[b.dof + z]) * ldSVp(bw.sF(b.dof + z));
          }
        chainGeometryVjp(bw, ct, k, dir, F, Wtot, V, bodyA, true, 1.0, G);
        chainGeometryVjp(bw, ct, k, dir, F, Wtot, V, bodyB, false, -1.0, G);
      }
      coopSyncWarp(coopN);
      const int active = (c + u - sbase) < coopN ? (c + u - sbase) : coopN;
      for (int dbase = 0; dbase < nd; dbase += coopN) {
        const int d = dbase + coopLane;
        if (d >= nd) continue;
        double a = acc[d];
        for (int L = 0; L < active; L++) a += bw.privG(L)[d];
        acc[d] = a;
      }
      coopSyncWarp(coopN);
    }
    for (int d = 0; d < nd; d++) gq[vidx(layout, N, nd, w, d)] += acc[d];
  }
  return status;
}

}  // namespace ddb
