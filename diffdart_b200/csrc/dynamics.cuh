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
  int snapOff, ndof;  // first snapshot pair of this body; joint DOFs (0, 1 or 6)
  // FreeJoint only (FreeJoint.cpp:673-693): N0/P0 hold T_parentToJoint, N1/P1 hold T_childToJoint^-1 and
  // S6 = getAdTMatrix(T_childToJoint), row-major 6x6
  double S6[36];
};

struct DevTopo {
  int nb, nd, layout, has_limits;
  int snapPairs, has_free;  // total snapshot pairs per world; any FreeJoint in the tree
  double g[3];
  double dt;
  DevBody b[DDB_MAX_BODIES];
};

// the topology of the handle that launched last lives here (one copy per translation unit).
// DDB_SPEC_TOPO names a file holding one topology as an aggregate initializer: kernels of that translation unit are then
// specialised for it (the compiler folds the constants and everything that follows from their zeros and ones).
#ifdef DDB_SPEC_TOPO
__device__ const DevTopo c_topo =
#include DDB_SPEC_TOPO
;
#else
__constant__ DevTopo c_topo;
#endif

// batched vector element (world w, dof d)
DDB_DI size_t vidx(int layout, int N, int n, int w, int d) {
  return layout == DDB_LAYOUT_WORLD_MAJOR ? (size_t)w * n + d : (size_t)d * N + w;
}

// ---------------------------------------------------------------------------------------------
// Snapshot of one step, written by the forward kernel and read by the backward / Jacobian
// kernels. 18 double2 per body; every access of a warp is one contiguous 512-byte segment
// (layout of the pairs: struct Snap below):
//   pairs 0-5  relative transform T_i (R row-major, p)      pairs 12-14 AIS_i = IA_i S_i
//   pairs 6-8  spatial velocity V_i                         pair  15    (psi_i, qdd_i)           written by sweep 3
//   pairs 9-11 spatial acceleration A_i                     pairs 16-17 (G_i = R_world^T g, qd_i) written by sweep 1
// Keeping these instead of recomputing them moves work from the FP64 pipe (the binding unit)
// to L2/HBM, which the path barely uses otherwise.
// ---------------------------------------------------------------------------------------------
// A FreeJoint body needs 6x6 blocks and uses 56 pairs instead: 0-5 T, 6-8 V, 9-11 A, 12-29 AIS (row-major 6x6),
// 30-47 psi (6x6), 48-50 qd, 51-53 qdd (u between the sweeps), 54-55 (G, pad), 56-58 q.
constexpr int kSnapPairs = 18;
constexpr int kSnapPairsFree = 59;  // + pairs 56-58: the joint coordinates q (6)

#ifndef DDB_SNAP_TILED
#define DDB_SNAP_TILED 1
#endif
struct Snap {
  double2* base;
  int N, w;
  // chains of 1-DOF joints: body i owns pairs [18 i, 18 i + 18), known at compile time, which (with the tile-major layout) turns
  // every snapshot address into one base register + an immediate; 0 = look the offset up in the topology
  int chainPairs = 0;
  DDB_DI int offOf(int body) const { return chainPairs ? body * chainPairs : c_topo.b[body].snapOff; }
#if DDB_SNAP_TILED
  // tile-major: [world / 32][pair][world % 32]. The pairs of 32 consecutive worlds form one contiguous block (46 KB for the
  // 5-link chain), so a warp streams through one DRAM-page-friendly region instead of touching one 512-byte piece in each
  // of 90 rows that lie N*16 bytes apart
  DDB_DI size_t rowStride() const { return 32; }
  DDB_DI size_t origin() const { return ((size_t)(w >> 5) * c_topo.snapPairs) * 32 + (w & 31); }
#else
  // pair-major: [pair][world]
  DDB_DI size_t rowStride() const { return (size_t)N; }
  DDB_DI size_t origin() const { return (size_t)w; }
#endif
  DDB_DI size_t at(int body, int k) const { return origin() + (size_t)(offOf(body) + k) * rowStride(); }
  DDB_DI void st(int body, int k, double a, double b) const { base[at(body, k)] = make_double2(a, b); }
  DDB_DI double2 ld(int body, int k) const { return __ldg(base + at(body, k)); }
  // plain (coherent) load: for values written earlier by the same thread in the same kernel
  DDB_DI double2 ldc(int body, int k) const { return base[at(body, k)]; }
  DDB_DI void beginUp(int) const {}
  DDB_DI void endDown(int) const {}
  DDB_DI void beginAdj(int) const {}
  DDB_DI void endAdj(int) const {}
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

// ---------------------------------------------------------------------------------------------
// Staged snapshot for the chain backward kernel. One world per thread at 255 registers leaves two warps per scheduler,
// and every sweep of the backward is a chain of dependent snapshot reads: issued as plain loads they stall the warp for
// an HBM round trip per body. Instead each thread copies its own rows global -> shared with cp.async (LDGSTS: no
// registers, no scoreboard), far ahead of their use, and the sweeps read shared memory.
//   phase A (M^-1 solve, up + down): rows T (0-5), AIS (12-14), (psi, qd) (15) of every body, 10 slots per body, all
//     requested at kernel entry, one commit group per body in the order the up sweep consumes them.
//   phase B (ID adjoint, leaves -> root): rows 0-11, 15-17 of one body, 15 slots, two buffers; the request for body j-2
//     is issued when body j is done. The buffers reuse the slots of phase A bodies the down sweep has finished with
//     (chains of >= 3 bodies), so a thread owns max(10 NB, 30) slots of 16 bytes: 800 B for the 5-link chain,
//     8 CTAs of 32 worlds per SM.
// A thread only reads slots it filled itself: cp.async.wait_group orders them, no CTA barrier is needed.
// ---------------------------------------------------------------------------------------------
#ifdef DDB_HOST_EMULATION  // the g++ build of these headers used by the CPU tests never takes the staged path
template <int N> DDB_DI void cpAsyncWait() {}
DDB_DI void cpAsyncCommit() {}
DDB_DI void cpAsync16(unsigned, const void*) {}
DDB_DI unsigned sharedAddr(const void*) { return 0; }
#else
template <int N> DDB_DI void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
DDB_DI void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DDB_DI void cpAsync16(unsigned dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
DDB_DI unsigned sharedAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
#endif
DDB_DI void cpAsyncWaitDyn(int n) {  // n is a compile-time constant after unrolling
  switch (n) {
    case 0: cpAsyncWait<0>(); break;
    case 1: cpAsyncWait<1>(); break;
    case 2: cpAsyncWait<2>(); break;
    case 3: cpAsyncWait<3>(); break;
    case 4: cpAsyncWait<4>(); break;
    case 5: cpAsyncWait<5>(); break;
    case 6: cpAsyncWait<6>(); break;
    default: cpAsyncWait<7>(); break;
  }
}

#ifndef DDB_STAGE_AHEAD
#define DDB_STAGE_AHEAD 3
#endif
#ifndef DDB_STAGE_NBUF
#define DDB_STAGE_NBUF 2
#endif
template <int NB, int BLOCK>
struct StageGeom {
  static constexpr int kRowsA = 10, kRowsB = 15, kBufs = DDB_STAGE_NBUF;
  // the phase B buffers reuse phase A slots when the chain is long enough for the down sweep to free them in time
  static constexpr bool kAlias = kRowsA * NB >= kBufs * kRowsB && NB >= 3;
  static constexpr int kBufBase = kAlias ? 0 : kRowsA * NB;
  static constexpr int kSlots = (kRowsA * NB > kBufBase + kBufs * kRowsB) ? kRowsA * NB : kBufBase + kBufs * kRowsB;
  static constexpr size_t kBytes = (size_t)kSlots * BLOCK * sizeof(double2);
};

template <class D>
struct SnapReaders {  // T / spatial-vector reads on top of D::ld(body, pair)
  DDB_DI const D& self() const { return *static_cast<const D*>(this); }
  DDB_DI SV ldSV(int body, int k) const {
    double2 a = self().ld(body, k), b = self().ld(body, k + 1), c = self().ld(body, k + 2);
    return sv(mk(a.x, a.y, b.x), mk(b.y, c.x, c.y));
  }
  DDB_DI Xf ldT(int body) const {
    double2 a = self().ld(body, 0), b = self().ld(body, 1), c = self().ld(body, 2), d = self().ld(body, 3), e = self().ld(body, 4),
            f = self().ld(body, 5);
    Xf t;
    t.R.m[0] = a.x; t.R.m[1] = a.y; t.R.m[2] = b.x; t.R.m[3] = b.y; t.R.m[4] = c.x; t.R.m[5] = c.y;
    t.R.m[6] = d.x; t.R.m[7] = d.y; t.R.m[8] = e.x; t.p = mk(e.y, f.x, f.y);
    return t;
  }
};

template <int NB, int BLOCK>
struct StageBase {
  typedef StageGeom<NB, BLOCK> G;
  const double2* g;   // this world's column of the snapshot in global memory
  size_t strideN;     // pairs between consecutive rows
  double2* sm;        // this thread's column of the CTA's staging area
  unsigned smAddr;
  int chainPairs;
  DDB_DI void init(const Snap& sn, double2* smemBase) {
    chainPairs = sn.chainPairs;
    g = sn.base + sn.origin();
    strideN = sn.rowStride();
    sm = smemBase + threadIdx.x;
    smAddr = sharedAddr(sm);
  }
  DDB_DI void copyRow(int slot, int body, int k) const {
    const double2* src = g + (size_t)(chainPairs * body + k) * strideN;  // (staging is used by chains only)
    cpAsync16(smAddr + (unsigned)(slot * BLOCK * (int)sizeof(double2)), src);
  }
  DDB_DI static int slotA(int body, int k) { return G::kRowsA * body + (k < 6 ? k : k - 6); }
  DDB_DI static int bufOf(int body) { return G::kBufBase + ((NB - 1 - body) % G::kBufs) * G::kRowsB; }
  DDB_DI static int slotB(int body, int k) { return bufOf(body) + (k < 12 ? k : k - 3); }
  DDB_DI void requestA(int body) const {
#pragma unroll
    for (int k = 0; k < 6; k++) copyRow(slotA(body, k), body, k);
#pragma unroll
    for (int k = 12; k < 16; k++) copyRow(slotA(body, k), body, k);
    cpAsyncCommit();
  }
  DDB_DI void requestB(int body) const {  // body < 0: an empty group keeps the group count uniform
    if (body >= 0) {
#pragma unroll
      for (int k = 0; k < 12; k++) copyRow(slotB(body, k), body, k);
#pragma unroll
      for (int k = 15; k < 18; k++) copyRow(slotB(body, k), body, k);
    }
    cpAsyncCommit();
  }
};

// phase A accessor: handed to solveInvMass
template <int NB, int BLOCK>
struct StageA : StageBase<NB, BLOCK>, SnapReaders<StageA<NB, BLOCK>> {
  typedef StageBase<NB, BLOCK> B;
  typedef StageGeom<NB, BLOCK> G;
  // Phase A rows are requested kAhead bodies ahead of the up sweep, not all at entry: with every warp of the first wave asking
  // for all of its rows at once (30 MB across the chip) the rows of the first body queue behind everybody's later rows.
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
  // when body i is needed, exactly kAhead - 1 younger groups (bodies i-1 ..., possibly empty) have been committed
  DDB_DI void beginUp(int i) const {
    cpAsyncWait<kAhead - 1>();
    requestAOrEmpty(i - kAhead);
  }
  // the down sweep is done with bodies 0..i: buffer b (slots [15 b, 15 b + 15)) is free once 10 (i + 1) >= 15 (b + 1)
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
// phase B accessor: handed to idAdjoint
template <int NB, int BLOCK>
struct StageB : StageBase<NB, BLOCK>, SnapReaders<StageB<NB, BLOCK>> {
  typedef StageBase<NB, BLOCK> B;
  DDB_DI double2 ld(int body, int k) const { return B::sm[B::slotB(body, k) * BLOCK]; }
  DDB_DI void beginUp(int) const {}
  DDB_DI void endDown(int) const {}
  // kBufs - 1 younger groups (bodies i-1 ..., possibly empty) may still be in flight
  DDB_DI void beginAdj(int) const { cpAsyncWait<StageGeom<NB, BLOCK>::kBufs - 1>(); }
  DDB_DI void endAdj(int i) const { B::requestB(i - StageGeom<NB, BLOCK>::kBufs); }
};

template <int NB, bool CHAIN>
struct Tree {
  DDB_DI static int parentOf(int i) { return CHAIN ? i - 1 : c_topo.b[i].parent; }
  DDB_DI static int dofOf(int i) { return CHAIN ? i : c_topo.b[i].dof; }
  DDB_DI static bool hasDof(int i) { return CHAIN ? true : c_topo.b[i].type != DDB_JOINT_WELD; }
};

DDB_DI void jointCS(const DevBody& b, double qi, double& c, double& s) {
  c = 0.0; s = 0.0;
  if (b.type == DDB_JOINT_REVOLUTE) sincos(qi, &s, &c);
  else if (b.type == DDB_JOINT_PRISMATIC) c = qi;
}
DDB_DI Xf jointTransformCS(const DevBody& b, double c, double s) {
  Xf t;
#pragma unroll
  for (int k = 0; k < 9; k++) t.R.m[k] = fma(s, b.N2[k], fma(c, b.N1[k], b.N0[k]));
  t.p.x = fma(s, b.P2[0], fma(c, b.P1[0], b.P0[0]));
  t.p.y = fma(s, b.P2[1], fma(c, b.P1[1], b.P0[1]));
  t.p.z = fma(s, b.P2[2], fma(c, b.P1[2], b.P0[2]));
  return t;
}
DDB_DI Xf jointTransform(const DevBody& b, double qi) {
  double c, s;
  jointCS(b, qi, c, s);
  return jointTransformCS(b, c, s);
}

// ---- FreeJoint pieces (generic instantiation only) ----
// math::expMapRot (Geometry.cpp:471-487): second-order Taylor below |q| = 1e-3
DDB_DI M3 expMapRot(V3 q) {
  const double theta = sqrt(dot(q, q));
  M3 K; K.m[0] = 0; K.m[1] = -q.z; K.m[2] = q.y; K.m[3] = q.z; K.m[4] = 0; K.m[5] = -q.x; K.m[6] = -q.y; K.m[7] = q.x; K.m[8] = 0;
  M3 K2 = mul(K, K);
  double a, b;
  if (theta < 1.0e-3) { a = 1.0; b = 0.5; }
  else { a = sin(theta) / theta; b = (1.0 - cos(theta)) / (theta * theta); }
  M3 R;
#pragma unroll
  for (int i = 0; i < 9; i++) R.m[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * K.m[i] + b * K2.m[i];
  return R;
}
// math::logMap(R) = AngleAxis(R).angle * axis, through Eigen's Quaternion(Matrix3) (Geometry.cpp:610-650)
DDB_DI V3 logMapRot(const M3& M) {
  double q[4];
  double t = M.m[0] + M.m[4] + M.m[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (M.m[7] - M.m[5]) * t; q[1] = (M.m[2] - M.m[6]) * t; q[2] = (M.m[3] - M.m[1]) * t;
  } else {
    int i = 0;
    if (M.m[4] > M.m[0]) i = 1;
    if (M.m[8] > M.m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(M.m[4 * i] - M.m[4 * j] - M.m[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (M.m[3 * k + j] - M.m[3 * j + k]) * t;
    q[j] = (M.m[3 * j + i] + M.m[3 * i + j]) * t;
    q[k] = (M.m[3 * k + i] + M.m[3 * i + k]) * t;
  }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  if (n != 0.0) {
    const double angle = 2.0 * atan2(n, fabs(q[3]));
    if (q[3] < 0.0) n = -n;
    return mk(angle * (q[0] / n), angle * (q[1] / n), angle * (q[2] / n));
  }
  return mk(0, 0, 0);
}
DDB_DI Xf mulXf(const Xf& A, const Xf& B) { Xf r; r.R = mul(A.R, B.R); r.p = mul(A.R, B.p) + A.p; return r; }
DDB_DI Xf freeJointTransform(const DevBody& b, const double* q6) {
  Xf A, J, Bt;
#pragma unroll
  for (int k = 0; k < 9; k++) { A.R.m[k] = b.N0[k]; Bt.R.m[k] = b.N1[k]; }
  A.p = mk(b.P0[0], b.P0[1], b.P0[2]);
  Bt.p = mk(b.P1[0], b.P1[1], b.P1[2]);
  J.R = expMapRot(mk(q6[0], q6[1], q6[2]));
  J.p = mk(q6[3], q6[4], q6[5]);
  return mulXf(mulXf(A, J), Bt);
}
DDB_DI SV mulS6(const DevBody& b, const double* x6) {  // S6 * x
  double r[6];
#pragma unroll
  for (int a = 0; a < 6; a++) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) acc += b.S6[a * 6 + k] * x6[k];
    r[a] = acc;
  }
  return sv(mk(r[0], r[1], r[2]), mk(r[3], r[4], r[5]));
}
DDB_DI void mulS6T(const DevBody& b, SV f, double* out6) {  // S6^T * f
  const double fv[6] = {f.w.x, f.w.y, f.w.z, f.v.x, f.v.y, f.v.z};
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 6; a++) acc += b.S6[a * 6 + k] * fv[a];
    out6[k] = acc;
  }
}
// inverse of a symmetric positive definite 6x6 (stands in for math::inverse<SE3Space>: ldlt().solve(I))
DDB_DI void spdInverse6(const double* Am, double* out) {
  double L[36], D[6];
  for (int j = 0; j < 6; j++) {
    double d = Am[j * 6 + j];
    for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
    D[j] = d;
    for (int i = j + 1; i < 6; i++) {
      double sacc = Am[i * 6 + j];
      for (int k = 0; k < j; k++) sacc -= L[i * 6 + k] * L[j * 6 + k] * D[k];
      L[i * 6 + j] = sacc / d;
    }
  }
  for (int c = 0; c < 6; c++) {
    double y[6];
    for (int i = 0; i < 6; i++) { double sacc = (i == c) ? 1.0 : 0.0; for (int k = 0; k < i; k++) sacc -= L[i * 6 + k] * y[k]; y[i] = sacc; }
    for (int i = 0; i < 6; i++) y[i] /= D[i];
    for (int i = 5; i >= 0; i--) { double sacc = y[i]; for (int k = i + 1; k < 6; k++) sacc -= L[k * 6 + i] * y[k]; y[i] = sacc; }
    for (int i = 0; i < 6; i++) out[i * 6 + c] = y[i];
  }
}
// 6-vectors in the snapshot: three pairs starting at pair k
DDB_DI void st6(const Snap& sn, int body, int k, const double* x) { sn.st(body, k, x[0], x[1]); sn.st(body, k + 1, x[2], x[3]); sn.st(body, k + 2, x[4], x[5]); }
DDB_DI void ld6c(const Snap& sn, int body, int k, double* x) {
  double2 a = sn.ldc(body, k), b = sn.ldc(body, k + 1), c = sn.ldc(body, k + 2);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y;
}
template <class SN>
DDB_DI void ld6(const SN& sn, int body, int k, double* x) {
  double2 a = sn.ld(body, k), b = sn.ld(body, k + 1), c = sn.ld(body, k + 2);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y;
}
DDB_DI void st36(const Snap& sn, int body, int k, const double* m) { for (int i = 0; i < 18; i++) sn.st(body, k + i, m[2 * i], m[2 * i + 1]); }
DDB_DI void ld36c(const Snap& sn, int body, int k, double* m) { for (int i = 0; i < 18; i++) { double2 a = sn.ldc(body, k + i); m[2 * i] = a.x; m[2 * i + 1] = a.y; } }
template <class SN>
DDB_DI void ld36(const SN& sn, int body, int k, double* m) { for (int i = 0; i < 18; i++) { double2 a = sn.ld(body, k + i); m[2 * i] = a.x; m[2 * i + 1] = a.y; } }
DDB_DI SV sv6(const double* x) { return sv(mk(x[0], x[1], x[2]), mk(x[3], x[4], x[5])); }

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
    SV V[NB], AIS[NB], A[NB], bias[NB], eta[NB];
    V3 G[NB];
    ABI IA[NB];
    double psi[NB], u[NB], qdv[NB];
    double cth[NB], sth[NB];
  } ws;
  Xf* T = ws.T;
  SV *V = ws.V, *AIS = ws.AIS, *A = ws.A, *bias = ws.bias, *eta = ws.eta;
  V3* G = ws.G;
  ABI* IA = ws.IA;
  double *psi = ws.psi, *u = ws.u, *qdv = ws.qdv;
#ifndef DDB_FWD_KEEP_T
  // chains: the 12 numbers of T_i are rebuilt from (cos q_i, sin q_i) in sweeps 2 and 3 (24 DFMA with constant-bank
  // operands each time) instead of occupying 24 registers per body across the whole kernel
  double *cth = ws.cth, *sth = ws.sth;
  // (the empty asm makes the value opaque: otherwise the compiler merges the rebuild with the first evaluation and keeps T live)
#ifdef DDB_HOST_EMULATION
  auto opaque = [](double x) { return x; };
#else
  auto opaque = [](double x) { asm volatile("" : "+d"(x)); return x; };
#endif
  auto Tof = [&](int i) -> Xf { return CHAIN ? jointTransformCS(c_topo.b[i], opaque(cth[i]), opaque(sth[i])) : T[i]; };
#else
  auto Tof = [&](int i) -> Xf { return T[i]; };
#endif
  auto down1 = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    const bool d = TS::hasDof(i);
    const bool isFree = !CHAIN && b.type == DDB_JOINT_FREE;
    Xf t;
    SV Sq;
    if (isFree) {
      double q6[6], qd6[6];
      for (int k = 0; k < 6; k++) { q6[k] = qOf(b.dof + k); qd6[k] = qdOf(b.dof + k); }
      t = freeJointTransform(b, q6);
      Sq = mulS6(b, qd6);
      qdv[i] = 0.0;
      st6(sn, i, 48, qd6);
      st6(sn, i, 56, q6);
    } else {
      const double qi = d ? qOf(TS::dofOf(i)) : 0.0, qdi = d ? qdOf(TS::dofOf(i)) : 0.0;
#ifndef DDB_FWD_KEEP_T
      if (CHAIN) { jointCS(b, qi, cth[i], sth[i]); t = jointTransformCS(b, cth[i], sth[i]); }
      else t = jointTransform(b, qi);
#else
      t = jointTransform(b, qi);
#endif
      qdv[i] = qdi;
      Sq = qdi * b.S;
    }
#ifndef DDB_FWD_KEEP_T
    if (!CHAIN) T[i] = t;
#else
    T[i] = t;
#endif
    V[i] = (p >= 0) ? AdInvT(t, V[p]) + Sq : Sq;
    V3 gp = (p >= 0) ? G[p] : mk(c_topo.g[0], c_topo.g[1], c_topo.g[2]);
    G[i] = mulT(t.R, gp);
    sn.stT(i, t);
    sn.stSV(i, 6, V[i]);
    // G_i and qd_i leave the register file here (sweeps 2 and 3 do not need them)
    if (isFree) { sn.st(i, 54, G[i].x, G[i].y); sn.st(i, 55, G[i].z, 0.0); }
    else { sn.st(i, 16, G[i].x, G[i].y); sn.st(i, 17, G[i].z, qdv[i]); }
  };
  // rigid-body terms of sweep 2, set right before it (they stay off the register file during sweep 1)
  auto rigid = [&](int i) {
    const DevBody& b = c_topo.b[i];
    IA[i] = toABI(b.I);
    // bias = -dad(V, I V) - I [0; R_w^T g]   (BodyNode.cpp:1692-1702)
    SV IV = mul(b.I, V[i]);
    SV fg = b.gravity ? mul(b.I, sv(mk(0, 0, 0), G[i])) : svzero();
    bias[i] = -dad(V[i], IV) - fg;
    // partial acceleration ad(V, S qd) (GenericJoint.hpp:1740-1750, dS = 0): all that sweeps 2 and 3 need of V_i
    eta[i] = ad(V[i], qdv[i] * b.S);
  };
  auto up = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    if (!CHAIN && b.type == DDB_JOINT_FREE) {
      // 6-DOF joint: AIS = IA S (6x6), psi = (S^T IA S)^-1, u = tau - S^T (IA eta + bias); a free joint transmits
      // only beta = bias + IA (eta + S psi u) to its parent and no articulated inertia (IA - AIS psi AIS^T = 0)
      double qd6[6], AISm[36], M6[36], psi6[36], u6[6], pu[6];
      ld6c(sn, i, 48, qd6);
      SV Sq = mulS6(b, qd6);
      SV eta = ad(V[i], Sq);
      for (int k = 0; k < 6; k++) {
        double col[6];
        for (int a = 0; a < 6; a++) col[a] = b.S6[a * 6 + k];
        SV c = mul(IA[i], sv6(col));
        AISm[0 * 6 + k] = c.w.x; AISm[1 * 6 + k] = c.w.y; AISm[2 * 6 + k] = c.w.z; AISm[3 * 6 + k] = c.v.x; AISm[4 * 6 + k] = c.v.y; AISm[5 * 6 + k] = c.v.z;
      }
      for (int k = 0; k < 6; k++)
        for (int l = 0; l < 6; l++) { double acc = 0.0; for (int a = 0; a < 6; a++) acc += b.S6[a * 6 + k] * AISm[a * 6 + l]; M6[k * 6 + l] = acc; }
      spdInverse6(M6, psi6);
      SV bf = mul(IA[i], eta) + bias[i];
      double sb[6];
      mulS6T(b, bf, sb);
      for (int k = 0; k < 6; k++) u6[k] = tauOf(b.dof + k) - sb[k];
      st36(sn, i, 12, AISm);
      st36(sn, i, 30, psi6);
      st6(sn, i, 51, u6);
      AIS[i] = svzero(); psi[i] = 0.0; u[i] = 0.0;
      if (p >= 0) {
        for (int k = 0; k < 6; k++) { double acc = 0.0; for (int l = 0; l < 6; l++) acc += psi6[k * 6 + l] * u6[l]; pu[k] = acc; }
        double av[6];
        for (int a = 0; a < 6; a++) { double acc = 0.0; for (int k = 0; k < 6; k++) acc += AISm[a * 6 + k] * pu[k]; av[a] = acc; }
        SV beta = bias[i] + mul(IA[i], eta) + sv6(av);
        bias[p] = bias[p] + dAdInvT(Tof(i), beta);
      }
      return;
    }
    const SV etai = eta[i];
    SV ais = svzero();
    double ps = 0.0, ui = 0.0;
    if (TS::hasDof(i)) {
      ais = mul(IA[i], b.S);
      ps = 1.0 / dot(b.S, ais);
      // total force u = tau - S^T (IA eta + bias)   (GenericJoint.hpp:2208-2229 with d = k = 0)
      ui = tauOf(TS::dofOf(i)) - dot(b.S, mul(IA[i], etai) + bias[i]);
    }
    AIS[i] = ais;
    psi[i] = ps;
    u[i] = ui;
    sn.stSV(i, 12, ais);
    if (p >= 0) {
      // beta = bias + IA (eta + S psi u)   (GenericJoint.hpp:2049-2074)
      SV beta = bias[i] + mul(IA[i], etai) + (ps * ui) * ais;
      const Xf Ti = Tof(i);
      bias[p] = bias[p] + dAdInvT(Ti, beta);
      ABI Pi = IA[i];
      if (TS::hasDof(i)) rank1Downdate(Pi, ais, ps);
      addTransformedInertia(IA[p], Ti, Pi);
    }
  };
  auto down2 = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    const Xf Ti = Tof(i);
    SV Ap = (p >= 0) ? AdInvT(Ti, A[p]) : svzero();
    if (!CHAIN && b.type == DDB_JOINT_FREE) {
      double AISm[36], psi6[36], u6[6], qd6[6], r6[6], acc6[6];
      ld36c(sn, i, 12, AISm); ld36c(sn, i, 30, psi6); ld6c(sn, i, 51, u6); ld6c(sn, i, 48, qd6);
      const double ap[6] = {Ap.w.x, Ap.w.y, Ap.w.z, Ap.v.x, Ap.v.y, Ap.v.z};
      for (int k = 0; k < 6; k++) { double acc = 0.0; for (int a = 0; a < 6; a++) acc += AISm[a * 6 + k] * ap[a]; r6[k] = u6[k] - acc; }
      for (int k = 0; k < 6; k++) { double acc = 0.0; for (int l = 0; l < 6; l++) acc += psi6[k * 6 + l] * r6[l]; acc6[k] = acc; store(b.dof + k, acc); }
      A[i] = Ap + mulS6(b, acc6) + ad(V[i], mulS6(b, qd6));
      sn.stSV(i, 9, A[i]);
      st6(sn, i, 51, acc6);
      return;
    }
    double acc = 0.0;
    if (TS::hasDof(i)) {
      // qdd = psi (u - (IA S)^T AdInvT(T, A_parent))   (GenericJoint.hpp:2316-2327)
      acc = psi[i] * (u[i] - dot(AIS[i], Ap));
      store(TS::dofOf(i), acc);
    }
    A[i] = Ap + acc * b.S + eta[i];
    sn.stSV(i, 9, A[i]);
    sn.st(i, 15, psi[i], acc);
  };
  if (CHAIN) {
#pragma unroll
#ifdef DDB_FWD_RIGID_LATE
    for (int i = 0; i < NB; i++) down1(i);
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) {
      if (i == NB - 1) rigid(i);
      if (i > 0) rigid(i - 1);  // the parent's own terms, set just before its child adds to them
      up(i);
    }
#else
    for (int i = 0; i < NB; i++) { down1(i); rigid(i); }
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) up(i);
#endif
#pragma unroll
    for (int i = 0; i < NB; i++) down2(i);
  } else {
#pragma unroll 1
    for (int i = 0; i < nb; i++) { down1(i); rigid(i); }
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

template <int NB, bool CHAIN, class SN, class LamF, class MuF>
DDB_DI void solveInvMass(const SN& sn, int nb, LamF lamOf, BwdScratch<NB>& ws, MuF muFree) {
  typedef Tree<NB, CHAIN> TS;
  SV* c = ws.c;
  double* a = ws.a;
  double* mu = ws.mu;
  SV* U = ws.U;
  auto init = [&](int i) { c[i] = svzero(); };
  auto up = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    sn.beginUp(i);
    if (!CHAIN && b.type == DDB_JOINT_FREE) {
      // a (6) = lambda - S^T c; the parent receives X^T (c + AIS psi a); a is kept in c[i] (no longer needed as bias)
      double a6[6], sc[6];
      mulS6T(b, c[i], sc);
      for (int k = 0; k < 6; k++) a6[k] = lamOf(b.dof + k) - sc[k];
      if (p >= 0) {
        double AISm[36], psi6[36], pa[6], av[6];
        ld36(sn, i, 12, AISm); ld36(sn, i, 30, psi6);
        for (int k = 0; k < 6; k++) { double acc = 0.0; for (int l = 0; l < 6; l++) acc += psi6[k * 6 + l] * a6[l]; pa[k] = acc; }
        for (int r = 0; r < 6; r++) { double acc = 0.0; for (int k = 0; k < 6; k++) acc += AISm[r * 6 + k] * pa[k]; av[r] = acc; }
        c[p] = c[p] + dAdInvT(sn.ldT(i), c[i] + sv6(av));
      }
      c[i] = sv6(a6);
      a[i] = 0.0;
      return;
    }
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
    if (!CHAIN && b.type == DDB_JOINT_FREE) {
      double AISm[36], psi6[36], r6[6], mu6[6];
      ld36(sn, i, 12, AISm); ld36(sn, i, 30, psi6);
      const double up6[6] = {Up.w.x, Up.w.y, Up.w.z, Up.v.x, Up.v.y, Up.v.z};
      const double a6[6] = {c[i].w.x, c[i].w.y, c[i].w.z, c[i].v.x, c[i].v.y, c[i].v.z};
      for (int k = 0; k < 6; k++) { double acc = 0.0; for (int r = 0; r < 6; r++) acc += AISm[r * 6 + k] * up6[r]; r6[k] = a6[k] - acc; }
      for (int k = 0; k < 6; k++) { double acc = 0.0; for (int l = 0; l < 6; l++) acc += psi6[k * 6 + l] * r6[l]; mu6[k] = acc; muFree(b.dof + k, acc); }
      mu[i] = 0.0;
      U[i] = Up + mulS6(b, mu6);
      return;
    }
    if (TS::hasDof(i)) {
      SV ais = sn.ldSV(i, 12);
      double ps = sn.ld(i, 15).x;
      m = ps * (a[i] - dot(ais, Up));
    }
    mu[i] = m;
    U[i] = Up + m * b.S;
    sn.endDown(i);
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
// expMapJac(q)^T: the right Jacobian of the exponential map (math::expMapJac, Geometry.cpp:489-505, is the left one),
// with the series used below |q| = 1e-3 so that it stays smooth
DDB_DI M3 expRightJac(V3 q) {
  const double t2 = dot(q, q), theta = sqrt(t2);
  double a, b;
  if (theta < 1.0e-3) { a = 0.5 - t2 / 24.0; b = 1.0 / 6.0 - t2 / 120.0; }
  else { a = (1.0 - cos(theta)) / t2; b = (theta - sin(theta)) / (t2 * theta); }
  M3 K; K.m[0] = 0; K.m[1] = -q.z; K.m[2] = q.y; K.m[3] = q.z; K.m[4] = 0; K.m[5] = -q.x; K.m[6] = -q.y; K.m[7] = q.x; K.m[8] = 0;
  M3 K2 = mul(K, K);
  M3 J;
#pragma unroll
  for (int i = 0; i < 9; i++) J.m[i] = ((i % 4 == 0) ? 1.0 : 0.0) - a * K.m[i] + b * K2.m[i];  // transpose flips the sign of K
  return J;
}
DDB_DI M3 inv3(const M3& A) {
  const double c00 = A.m[4] * A.m[8] - A.m[5] * A.m[7], c01 = A.m[5] * A.m[6] - A.m[3] * A.m[8], c02 = A.m[3] * A.m[7] - A.m[4] * A.m[6];
  const double det = A.m[0] * c00 + A.m[1] * c01 + A.m[2] * c02, id = 1.0 / det;
  M3 r;
  r.m[0] = c00 * id; r.m[1] = (A.m[2] * A.m[7] - A.m[1] * A.m[8]) * id; r.m[2] = (A.m[1] * A.m[5] - A.m[2] * A.m[4]) * id;
  r.m[3] = c01 * id; r.m[4] = (A.m[0] * A.m[8] - A.m[2] * A.m[6]) * id; r.m[5] = (A.m[2] * A.m[3] - A.m[0] * A.m[5]) * id;
  r.m[6] = c02 * id; r.m[7] = (A.m[1] * A.m[6] - A.m[0] * A.m[7]) * id; r.m[8] = (A.m[0] * A.m[4] - A.m[1] * A.m[3]) * id;
  return r;
}
// 6x6 blocks of d q+ / d q (pp) and d q+ / d qd (vp) for one FreeJoint (row-major), closed form of
// FreeJoint::integratePositionsExplicit (FreeJoint.cpp:596-601): R+ = R(q_r) exp(dt w), p+ = p + dt R(q_r) v.
// The reference differentiates it numerically (FreeJoint.cpp:604-653, eps 1e-6 / 1e-7).
DDB_DI void freeJointPosJacobians(const double* q6, const double* qd6, double dt, double* pp, double* vp) {
  const V3 qr = mk(q6[0], q6[1], q6[2]), w = mk(qd6[0] * dt, qd6[1] * dt, qd6[2] * dt), v = mk(qd6[3], qd6[4], qd6[5]);
  const M3 R = expMapRot(qr), E = expMapRot(w);
  const V3 qn = logMapRot(mul(R, E));
  const M3 JrQ = expRightJac(qr), JrW = expRightJac(w), JrNinv = inv3(expRightJac(qn));
  M3 Et; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Et.m[3 * i + j] = E.m[3 * j + i];
  const M3 Prr = mul(JrNinv, mul(Et, JrQ));   // d q+_r / d q_r
  const M3 Vrr = mul(JrNinv, JrW);            // d q+_r / d (dt w)
  M3 Vx; Vx.m[0] = 0; Vx.m[1] = -v.z; Vx.m[2] = v.y; Vx.m[3] = v.z; Vx.m[4] = 0; Vx.m[5] = -v.x; Vx.m[6] = -v.y; Vx.m[7] = v.x; Vx.m[8] = 0;
  const M3 Ppr = mul(mul(R, Vx), JrQ);        // d (R v) / d q_r = -R [v]x Jr(q_r)
  for (int i = 0; i < 36; i++) { pp[i] = 0.0; vp[i] = 0.0; }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      pp[i * 6 + j] = Prr.m[3 * i + j];
      pp[(3 + i) * 6 + j] = -dt * Ppr.m[3 * i + j];
      pp[(3 + i) * 6 + 3 + j] = (i == j) ? 1.0 : 0.0;
      vp[i * 6 + j] = dt * Vrr.m[3 * i + j];
      vp[(3 + i) * 6 + 3 + j] = dt * R.m[3 * i + j];
    }
}

// MODE 0: velocities / accelerations / gravity of the step as retained in the snapshot.
// MODE 1: the same with dA[i] added to every body acceleration (the joint accelerations of the contact impulses).
// MODE 2: V = 0, G = 0, A = dA: Phi = mu^T M(q) a, the bilinear mass-matrix form used for d(M^-1 w)/dq terms.
template <int NB, bool CHAIN, int MODE, class SN, class Out, class OutFree>
DDB_DI void idAdjointEx(const SN& sn, int nb, BwdScratch<NB>& ws, const SV* dA, Out out, OutFree outFree) {
  typedef Tree<NB, CHAIN> TS;
  auto accOf = [&](int i) { return MODE == 2 ? dA[i] : (MODE == 1 ? sn.ldSV(i, 9) + dA[i] : sn.ldSV(i, 9)); };
  auto velOf = [&](int i) { return MODE == 2 ? svzero() : sn.ldSV(i, 6); };
  const double* mu = ws.mu;
  const SV* U = ws.U;
  SV *pU = ws.pU, *pA = ws.pA, *pV = ws.pV;
  V3* pG = ws.pG;
  auto init = [&](int i) { pU[i] = svzero(); pA[i] = svzero(); pV[i] = svzero(); pG[i] = mk(0, 0, 0); };
  auto up = [&](int i) {
    const DevBody& b = c_topo.b[i];
    const int p = TS::parentOf(i);
    sn.beginAdj(i);
    const SV Vi = velOf(i), Ai = accOf(i);
    if (!CHAIN && b.type == DDB_JOINT_FREE) {
      double qd6[6] = {0, 0, 0, 0, 0, 0};
      if (MODE != 2) ld6(sn, i, 48, qd6);
      const double2 g0 = sn.ld(i, 54), g1 = sn.ld(i, 55);
      const V3 Gi = MODE == 2 ? mk(0, 0, 0) : mk(g0.x, g0.y, g1.x);
      const SV Ui = U[i];
      SV IV = mul(b.I, Vi);
      SV Aeff = Ai;
      if (b.gravity) Aeff.v = Aeff.v - Gi;
      SV P = mul(b.I, Ui);
      SV Sq = mulS6(b, qd6);
      SV Ub = mul(b.I, Aeff) - dad(Vi, IV) + pU[i];
      SV Ab = P + pA[i];
      SV Vb = dad(Ui, IV) + mul(b.I, ad(Ui, Vi)) - dad(Sq, Ab) + pV[i];
      V3 Gb = b.gravity ? pG[i] - P.v : pG[i];
      Xf t = sn.ldT(i);
      // what the parent contributes, expressed in this body's frame
      SV XU = svzero(), XV = svzero(), XA = svzero();
      if (p >= 0) { XU = AdInvT(t, U[p]); XV = AdInvT(t, velOf(p)); XA = AdInvT(t, accOf(p)); }
      // joint coordinates of the FreeJoint: the child frame moves with twist S6 * jhat_k(q) per unit q_k,
      // jhat_k = [Jr(q_r) e_k; 0] (k < 3), [0; R(q_r)^T e_k] (k >= 3)
      double q6[6], dq6[6], dqd6[6];
      ld6(sn, i, 56, q6);
      const V3 qr = mk(q6[0], q6[1], q6[2]);
      const M3 Jr = expRightJac(qr), Rq = expMapRot(qr);
      for (int k = 0; k < 6; k++) {
        double jh[6] = {0, 0, 0, 0, 0, 0};
        if (k < 3) { jh[0] = Jr.m[k]; jh[1] = Jr.m[3 + k]; jh[2] = Jr.m[6 + k]; }
        else { jh[3] = Rq.m[3 * (k - 3)]; jh[4] = Rq.m[3 * (k - 3) + 1]; jh[5] = Rq.m[3 * (k - 3) + 2]; }
        const SV Sk = mulS6(b, jh);
        dq6[k] = -(dot(dad(Sk, Ub), XU) + dot(dad(Sk, Ab), XA) + dot(dad(Sk, Vb), XV)) - dot(Gb, cross(Sk.w, Gi));
      }
      SV w2 = Vb + dad(Vi, Ab);
      mulS6T(b, w2, dqd6);
      outFree(i, dq6, dqd6);
      if (p >= 0) {
        pU[p] = pU[p] + dAdInvT(t, Ub);
        pA[p] = pA[p] + dAdInvT(t, Ab);
        pV[p] = pV[p] + dAdInvT(t, Vb);
        pG[p] = pG[p] + mul(t.R, Gb);
      }
      return;
    }
    const double2 g0 = sn.ld(i, 16), g1 = sn.ld(i, 17);
    const double qdi = MODE == 2 ? 0.0 : g1.y;
    const V3 Gi = MODE == 2 ? mk(0, 0, 0) : mk(g0.x, g0.y, g1.x);
    const double qddi = sn.ld(i, 15).y;
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
      if (MODE != 0) XA = (p >= 0) ? AdInvT(sn.ldT(i), accOf(p)) : svzero();
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
    sn.endAdj(i);
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

template <int NB, bool CHAIN, class SN, class Out, class OutFree>
DDB_DI void idAdjoint(const SN& sn, int nb, BwdScratch<NB>& ws, Out out, OutFree outFree) {
  idAdjointEx<NB, CHAIN, 0, SN>(sn, nb, ws, nullptr, out, outFree);
}

// U_i = J_i mu: body twists under joint rates mu (fills ws.U / ws.mu exactly as solveInvMass leaves them)
template <int NB, class RateF>
DDB_DI void twistsFromRates(const Snap& sn, int nb, RateF rateOf, SV* U, double* mu) {
#pragma unroll 1
  for (int i = 0; i < nb; i++) {
    const DevBody& b = c_topo.b[i];
    SV Up = svzero();
    if (b.parent >= 0) Up = AdInvT(sn.ldT(i), U[b.parent]);
    if (b.type == DDB_JOINT_FREE) {
      double r6[6];
      for (int k = 0; k < 6; k++) r6[k] = rateOf(b.dof + k);
      if (mu) mu[i] = 0.0;
      U[i] = Up + mulS6(b, r6);
    } else {
      const double m = b.type != DDB_JOINT_WELD ? rateOf(b.dof) : 0.0;
      if (mu) mu[i] = m;
      U[i] = Up + m * b.S;
    }
  }
}

}  // namespace ddb
