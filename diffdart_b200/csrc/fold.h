// fold.h — host-side folding of a ddb_topology into the device descriptors (DevTopo for the tree recursions,
// ContactTopo for the contact stage). Pure C++: shared by ddb_api.cu and the host-emulation tests.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace ddb {

struct HM3 { double m[9]; };
inline HM3 mmul(const HM3& A, const HM3& B) {
  HM3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return r;
}
inline void mvec(const HM3& A, const double* x, double* y) {
  for (int i = 0; i < 3; i++) y[i] = A.m[3 * i] * x[0] + A.m[3 * i + 1] * x[1] + A.m[3 * i + 2] * x[2];
}
inline void hcross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

// Fold T_parentToJoint * J(q) * T_childToJoint^-1 into N0 + c N1 + s N2, P0 + c P1 + s P2.
inline int foldTopology(const ddb_topology* t, const ddb_params* p, DevTopo& out, bool& chain, std::string& err) {
  if (!t || !p || !t->bodies) { err = "null topology/params"; return DDB_ERR_INVALID; }
  if (t->n_bodies < 1 || t->n_bodies > DDB_MAX_BODIES || t->n_dofs < 0 || t->n_dofs > DDB_MAX_DOFS) {
    err = "n_bodies/n_dofs out of range"; return DDB_ERR_INVALID;
  }
  std::memset(&out, 0, sizeof(out));
  out.nb = t->n_bodies; out.nd = t->n_dofs; out.layout = p->layout;
  out.dt = p->time_step;
  for (int k = 0; k < 3; k++) out.g[k] = p->gravity[k];
  chain = true;
  int dofCount = 0;
  for (int i = 0; i < t->n_bodies; i++) {
    const ddb_body& s = t->bodies[i];
    DevBody& b = out.b[i];
    if (s.parent >= i || s.parent < -1) { err = "bodies must be listed parent-before-child"; return DDB_ERR_INVALID; }
    if (s.joint_type != DDB_JOINT_WELD && s.joint_type != DDB_JOINT_REVOLUTE && s.joint_type != DDB_JOINT_PRISMATIC && s.joint_type != DDB_JOINT_FREE) {
      err = "unknown joint type"; return DDB_ERR_INVALID;
    }
    int nd = s.joint_type == DDB_JOINT_WELD ? 0 : (s.joint_type == DDB_JOINT_FREE ? 6 : 1);
    b.ndof = nd;
    b.snapOff = out.snapPairs;
    out.snapPairs += s.joint_type == DDB_JOINT_FREE ? kSnapPairsFree : kSnapPairs;
    if (s.joint_type == DDB_JOINT_FREE) out.has_free = 1;
    if (nd && (s.dof_offset < 0 || s.dof_offset + nd > t->n_dofs)) { err = "dof_offset out of range"; return DDB_ERR_INVALID; }
    dofCount += nd;
    b.parent = s.parent; b.type = s.joint_type; b.dof = nd ? s.dof_offset : -1; b.gravity = s.gravity_mode != 0;
    if (!(nd == 1 && s.dof_offset == i && s.parent == i - 1)) chain = false;
    HM3 Ra, Rc, Rb;
    std::memcpy(Ra.m, s.parent_to_joint.r, sizeof(Ra.m));
    std::memcpy(Rc.m, s.child_to_joint.r, sizeof(Rc.m));
    const double* pa = s.parent_to_joint.p;
    const double* pc = s.child_to_joint.p;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rb.m[3 * r + c] = Rc.m[3 * c + r];  // Rc^T
    double pb[3];
    mvec(Rb, pc, pb);
    for (int k = 0; k < 3; k++) pb[k] = -pb[k];
    double a[3] = {s.axis[0], s.axis[1], s.axis[2]};
    double an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (nd == 1) {
      if (!(an > 0)) { err = "zero joint axis"; return DDB_ERR_INVALID; }
      for (int k = 0; k < 3; k++) a[k] /= an;  // Joint::setAxis normalizes
    }
    if (s.joint_type == DDB_JOINT_FREE) {
      // T = T_p2j * [expMapRot(q0..2) | q3..5] * T_c2j^-1 (FreeJoint.cpp:673-683); S = getAdTMatrix(T_c2j) (:686-693)
      std::memcpy(b.N0, Ra.m, sizeof(Ra.m));
      std::memcpy(b.N1, Rb.m, sizeof(Rb.m));
      for (int k = 0; k < 3; k++) { b.P0[k] = pa[k]; b.P1[k] = pb[k]; }
      HM3 P = {{0, -pc[2], pc[1], pc[2], 0, -pc[0], -pc[1], pc[0], 0}};
      HM3 PR = mmul(P, Rc);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          b.S6[r * 6 + c] = Rc.m[3 * r + c];
          b.S6[r * 6 + 3 + c] = 0.0;
          b.S6[(3 + r) * 6 + c] = PR.m[3 * r + c];
          b.S6[(3 + r) * 6 + 3 + c] = Rc.m[3 * r + c];
        }
    }
    HM3 M0 = mmul(Ra, Rb);
    double Rapb[3];
    mvec(Ra, pb, Rapb);
    if (s.joint_type == DDB_JOINT_FREE) {
    } else if (s.joint_type == DDB_JOINT_REVOLUTE) {
      double Raa[3], Rbta[3];
      mvec(Ra, a, Raa);
      for (int c = 0; c < 3; c++) Rbta[c] = Rb.m[c] * a[0] + Rb.m[3 + c] * a[1] + Rb.m[6 + c] * a[2];  // (a^T Rb)_c
      HM3 K = {{0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0}};
      HM3 M2 = mmul(mmul(Ra, K), Rb);
      double adotpb = a[0] * pb[0] + a[1] * pb[1] + a[2] * pb[2];
      double axpb[3], RaAxpb[3];
      hcross(a, pb, axpb);
      mvec(Ra, axpb, RaAxpb);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          double m1 = Raa[r] * Rbta[c];
          b.N0[3 * r + c] = m1;
          b.N1[3 * r + c] = M0.m[3 * r + c] - m1;
          b.N2[3 * r + c] = M2.m[3 * r + c];
        }
      for (int k = 0; k < 3; k++) {
        b.P0[k] = pa[k] + Raa[k] * adotpb;
        b.P1[k] = Rapb[k] - Raa[k] * adotpb;
        b.P2[k] = RaAxpb[k];
      }
      // S = AdTAngular(T_c2j, axis) (RevoluteJoint.cpp:140-150)
      double w[3], v[3];
      mvec(Rc, a, w);
      hcross(pc, w, v);
      b.S.w = V3{w[0], w[1], w[2]};
      b.S.v = V3{v[0], v[1], v[2]};
    } else {
      std::memcpy(b.N0, M0.m, sizeof(M0.m));
      for (int k = 0; k < 3; k++) b.P0[k] = pa[k] + Rapb[k];
      if (s.joint_type == DDB_JOINT_PRISMATIC) {
        double Raa[3], v[3];
        mvec(Ra, a, Raa);
        for (int k = 0; k < 3; k++) b.P1[k] = Raa[k];
        // S = AdTLinear(T_c2j, axis) (PrismaticJoint.cpp:139-150)
        mvec(Rc, a, v);
        b.S.w = V3{0, 0, 0};
        b.S.v = V3{v[0], v[1], v[2]};
      }
    }
    // Inertia::computeSpatialTensor (Inertia.cpp:394-409)
    const double m = s.mass, cx = s.com[0], cy = s.com[1], cz = s.com[2];
    b.I.m = m;
    b.I.h = V3{m * cx, m * cy, m * cz};
    b.I.J[0] = s.moment[0] + m * (cy * cy + cz * cz);
    b.I.J[1] = s.moment[1] + m * (cx * cx + cz * cz);
    b.I.J[2] = s.moment[2] + m * (cx * cx + cy * cy);
    b.I.J[3] = s.moment[3] - m * cx * cy;
    b.I.J[4] = s.moment[4] - m * cx * cz;
    b.I.J[5] = s.moment[5] - m * cy * cz;
  }
  if (dofCount != t->n_dofs) { err = "n_dofs does not match the joints"; return DDB_ERR_INVALID; }
  auto finite = [&](const double* a, double def) {
    if (!a) return false;
    for (int i = 0; i < t->n_dofs; i++) if (a[i] != def) return true;
    return false;
  };
  out.has_limits = finite(t->position_lower, -INFINITY) || finite(t->position_upper, INFINITY) || finite(t->velocity_lower, -INFINITY) ||
                   finite(t->velocity_upper, INFINITY) || finite(t->force_lower, -INFINITY) || finite(t->force_upper, INFINITY);
  return DDB_OK;
}


// Serial chains of 1-DOF joints are run in canonical joint frames (chain.cuh): body frame i' sits at joint i with the joint axis
// as its z axis. Rewrites the DevBody entries of a topology foldTopology has accepted as a chain:
//   N0, P0 <- F_i = X_p^-1 T_parentBodyToJoint,i (A_i, 0)   with X_i = T_childBodyToJoint,i (A_i, 0), A_i e_z = axis_i
//   I      <- the rigid-body inertia about the new origin, in the new axes
//   S      <- (e_z; 0) / (0; e_z)  (kept for reference; the chain kernels pick components instead of multiplying by S)
struct HXf { HM3 R; double p[3]; };
inline HXf hxmul(const HXf& A, const HXf& B) {
  HXf r; r.R = mmul(A.R, B.R); mvec(A.R, B.p, r.p);
  for (int k = 0; k < 3; k++) r.p[k] += A.p[k];
  return r;
}
inline HXf hxinv(const HXf& A) {
  HXf r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.R.m[3 * i + j] = A.R.m[3 * j + i];
  mvec(r.R, A.p, r.p);
  for (int k = 0; k < 3; k++) r.p[k] = -r.p[k];
  return r;
}
// shortest rotation taking e_z to the unit vector a (exact 0 / +-1 entries for axis-aligned a)
inline HM3 rotZTo(const double* a) {
  HM3 R = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};
  const double c = a[2];
  if (c > 1.0 - 1e-15) return R;
  if (c < -1.0 + 1e-15) { R.m[4] = -1; R.m[8] = -1; return R; }  // half a turn about x
  const double v[3] = {-a[1], a[0], 0.0};  // e_z x a
  const HM3 K = {{0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0}};
  const HM3 K2 = mmul(K, K);
  for (int k = 0; k < 9; k++) R.m[k] += K.m[k] + K2.m[k] / (1.0 + c);
  return R;
}
inline void canonicalizeChain(const ddb_topology* t, DevTopo& out) {
  std::vector<HXf> X(t->n_bodies);
  out.snapPairs = 0;
  for (int i = 0; i < t->n_bodies; i++) {
    const ddb_body& s = t->bodies[i];
    DevBody& b = out.b[i];
    double a[3] = {s.axis[0], s.axis[1], s.axis[2]};
    const double an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    for (int k = 0; k < 3; k++) a[k] /= an;
    HXf A, Tp, Tc;
    A.R = rotZTo(a); A.p[0] = A.p[1] = A.p[2] = 0.0;
    std::memcpy(Tp.R.m, s.parent_to_joint.r, sizeof(Tp.R.m)); std::memcpy(Tp.p, s.parent_to_joint.p, sizeof(Tp.p));
    std::memcpy(Tc.R.m, s.child_to_joint.r, sizeof(Tc.R.m)); std::memcpy(Tc.p, s.child_to_joint.p, sizeof(Tc.p));
    X[i] = hxmul(Tc, A);
    HXf F = hxmul(Tp, A);
    if (i > 0) F = hxmul(hxinv(X[i - 1]), F);
    std::memset(b.N0, 0, sizeof(b.N0)); std::memset(b.N1, 0, sizeof(b.N1)); std::memset(b.N2, 0, sizeof(b.N2));
    std::memset(b.P1, 0, sizeof(b.P1)); std::memset(b.P2, 0, sizeof(b.P2));
    std::memcpy(b.N0, F.R.m, sizeof(b.N0));
    std::memcpy(b.P0, F.p, sizeof(b.P0));
    const bool rev = s.joint_type == DDB_JOINT_REVOLUTE;
    b.S.w = V3{0, 0, rev ? 1.0 : 0.0};
    b.S.v = V3{0, 0, rev ? 0.0 : 1.0};
    // inertia: centre of mass and moment (about the centre of mass) expressed in the new frame, then Inertia::computeSpatialTensor
    const HXf Xi = hxinv(X[i]);
    double c[3];
    mvec(Xi.R, s.com, c);
    for (int k = 0; k < 3; k++) c[k] += Xi.p[k];
    const HM3 Ic = {{s.moment[0], s.moment[3], s.moment[4], s.moment[3], s.moment[1], s.moment[5], s.moment[4], s.moment[5], s.moment[2]}};
    const HM3 In = mmul(mmul(Xi.R, Ic), hxinv(Xi).R);
    const double m = s.mass;
    b.I.m = m;
    b.I.h = V3{m * c[0], m * c[1], m * c[2]};
    b.I.J[0] = In.m[0] + m * (c[1] * c[1] + c[2] * c[2]);
    b.I.J[1] = In.m[4] + m * (c[0] * c[0] + c[2] * c[2]);
    b.I.J[2] = In.m[8] + m * (c[0] * c[0] + c[1] * c[1]);
    b.I.J[3] = In.m[1] - m * c[0] * c[1];
    b.I.J[4] = In.m[2] - m * c[0] * c[2];
    b.I.J[5] = In.m[5] - m * c[1] * c[2];
    b.snapOff = out.snapPairs;
    out.snapPairs += kChainPairs;
  }
}

// shapes, reactive flags, collidable pairs in collision-object order, capacities.
// BodyNode::isReactive (BodyNode.cpp:1999-2018) is per body: the skeleton is mobile and the body depends on at least one
// generalized coordinate, i.e. the body itself or one of its ancestors hangs on a joint with DOFs. A welded root link of a
// fixed-base robot (world -> fixed -> base) is not reactive although its skeleton moves.
inline int foldContactTopology(const ddb_topology* topo, const ddb_params* params, ContactTopo& ct, std::string& err) {
  const int nb = topo->n_bodies;
    std::memset(&ct, 0, sizeof(ct));
    ct.clip_depth = params->contact_clipping_depth;
    ct.gradients = params->reserved[1] == 1 ? 0 : 1;  // reserved[1] = 1: plain World::step (constraint solver gradients disabled)
    int nskel = 0;
    for (int i = 0; i < nb; i++) nskel = std::max(nskel, topo->bodies[i].skeleton + 1);
    if (nskel > DDB_MAX_BODIES) { err = "too many skeletons"; return DDB_ERR_INVALID; }
    ct.nskel = nskel;
    std::vector<char> dependsOnDof(nb, 0);  // bodies are listed parent-before-child (checked by foldTopology)
    for (int i = 0; i < nb; i++) {
      const int p = topo->bodies[i].parent;
      dependsOnDof[i] = topo->bodies[i].joint_type != DDB_JOINT_WELD || (p >= 0 && p < i && dependsOnDof[p]);
    }
    for (int i = 0; i < nb; i++) {
      const ddb_body& sb = topo->bodies[i];
      ShapeDesc& sd = ct.shape[i];
      sd.shape = sb.shape_type; sd.skel = sb.skeleton; sd.reactive = dependsOnDof[i];
      sd.friction = sb.friction; sd.restitution = sb.restitution;
      for (int k = 0; k < 3; k++) { sd.size[k] = sb.shape_size[k]; sd.offP[k] = sb.shape_offset.p[k]; }
      for (int k = 0; k < 9; k++) sd.offR[k] = sb.shape_offset.r[k];
    }
    int np = 0;
    for (int i = 0; i < nb; i++)
      for (int j = i + 1; j < nb; j++) {
        if (ct.shape[i].shape == DDB_SHAPE_NONE || ct.shape[j].shape == DDB_SHAPE_NONE) continue;
        if (ct.shape[i].skel == ct.shape[j].skel) continue;              // self-collision checking is off by default
        if (!ct.shape[i].reactive && !ct.shape[j].reactive) continue;    // such contacts never become active constraints
        if (np >= DDB_MAX_BODIES * 4) { err = "too many collidable pairs"; return DDB_ERR_INVALID; }
        ct.pairA[np] = i; ct.pairB[np] = j; np++;
        if (ct.shape[i].restitution * ct.shape[j].restitution > 1e-3) ct.bouncing = 1;
      }
    ct.n_pairs = np;
    int cap = params->reserved[0] > 0 ? params->reserved[0] : 64;  // reserved[0]: max contacts per world
    ct.cmax = std::min(8 * np, cap);
    ct.kmax = 3 * ct.cmax;
  return DDB_OK;
}

}  // namespace ddb
