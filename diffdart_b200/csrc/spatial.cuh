// spatial.cuh — FP64 SE(3) spatial algebra for sm_100a device code.
// Conventions follow the reference (dart/math/Geometry.cpp): spatial vectors are
// [angular; linear] in the body frame; a transform T=(R,p) maps child to parent.
// Everything is __forceinline__ so that per-thread temporaries live in registers.
#pragma once
#include <cuda_runtime.h>

namespace ddb {

struct V3 {
  double x, y, z;
};
struct M3 {  // row-major
  double m[9];
};
struct SV {  // spatial vector (twist or wrench): angular part w, linear part v
  V3 w, v;
};
struct Xf {  // rigid transform
  M3 R;
  V3 p;
};
// Rigid-body spatial inertia [[J, [h]x],[-[h]x, m 1]] (Inertia.cpp:394-409): h = m*com,
// J = moment + m [c]x[c]x^T, stored Jxx Jyy Jzz Jxy Jxz Jyz.
struct RBI {
  double m;
  V3 h;
  double J[6];
};
// Articulated-body inertia [[A, B],[B^T, D]], A and D symmetric (xx yy zz xy xz yz), B full row-major.
struct ABI {
  double A[6], B[9], D[6];
};

#define DDB_DI __device__ __forceinline__

DDB_DI V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
DDB_DI V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
DDB_DI V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
DDB_DI V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
DDB_DI V3 operator*(double s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
DDB_DI double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DDB_DI V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DDB_DI V3 mul(const M3& R, V3 a) {
  return mk(R.m[0] * a.x + R.m[1] * a.y + R.m[2] * a.z, R.m[3] * a.x + R.m[4] * a.y + R.m[5] * a.z,
            R.m[6] * a.x + R.m[7] * a.y + R.m[8] * a.z);
}
DDB_DI V3 mulT(const M3& R, V3 a) {
  return mk(R.m[0] * a.x + R.m[3] * a.y + R.m[6] * a.z, R.m[1] * a.x + R.m[4] * a.y + R.m[7] * a.z,
            R.m[2] * a.x + R.m[5] * a.y + R.m[8] * a.z);
}
DDB_DI M3 mul(const M3& A, const M3& B) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return r;
}
DDB_DI M3 mulABt(const M3& A, const M3& B) {  // A * B^T
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = A.m[3 * i] * B.m[3 * j] + A.m[3 * i + 1] * B.m[3 * j + 1] + A.m[3 * i + 2] * B.m[3 * j + 2];
  return r;
}

DDB_DI SV sv(V3 w, V3 v) { SV r; r.w = w; r.v = v; return r; }
DDB_DI SV svzero() { return sv(mk(0, 0, 0), mk(0, 0, 0)); }
DDB_DI SV operator+(SV a, SV b) { return sv(a.w + b.w, a.v + b.v); }
DDB_DI SV operator-(SV a, SV b) { return sv(a.w - b.w, a.v - b.v); }
DDB_DI SV operator-(SV a) { return sv(-a.w, -a.v); }
DDB_DI SV operator*(double s, SV a) { return sv(s * a.w, s * a.v); }
DDB_DI double dot(SV a, SV b) { return dot(a.w, b.w) + dot(a.v, b.v); }

// Geometry.cpp:1121-1129  twist parent->child: [R^T w ; R^T (v + w x p)]
DDB_DI SV AdInvT(const Xf& T, SV V) { return sv(mulT(T.R, V.w), mulT(T.R, V.v + cross(V.w, T.p))); }
// Geometry.cpp:1195-1202  wrench child->parent: [R t + p x (R f) ; R f]
DDB_DI SV dAdInvT(const Xf& T, SV F) {
  V3 f = mul(T.R, F.v);
  return sv(mul(T.R, F.w) + cross(T.p, f), f);
}
// Geometry.cpp:1153-1167
DDB_DI SV ad(SV X, SV Y) { return sv(cross(X.w, Y.w), cross(X.w, Y.v) + cross(X.v, Y.w)); }
// Geometry.cpp:1760-1767  dad(s,t) = ad_s^T t   (s twist, t wrench)
DDB_DI SV dad(SV s, SV t) { return sv(cross(t.w, s.w) + cross(t.v, s.v), cross(t.v, s.w)); }

DDB_DI V3 symmul(const double* S, V3 a) {  // S = xx yy zz xy xz yz
  return mk(S[0] * a.x + S[3] * a.y + S[4] * a.z, S[3] * a.x + S[1] * a.y + S[5] * a.z, S[4] * a.x + S[5] * a.y + S[2] * a.z);
}
// I * V for a rigid body: [J w + h x v ; m v - h x w]
DDB_DI SV mul(const RBI& I, SV V) { return sv(symmul(I.J, V.w) + cross(I.h, V.v), I.m * V.v - cross(I.h, V.w)); }
// IA * V for an articulated inertia: [A w + B v ; B^T w + D v]
DDB_DI SV mul(const ABI& I, SV V) {
  V3 Bv = mk(I.B[0] * V.v.x + I.B[1] * V.v.y + I.B[2] * V.v.z, I.B[3] * V.v.x + I.B[4] * V.v.y + I.B[5] * V.v.z,
             I.B[6] * V.v.x + I.B[7] * V.v.y + I.B[8] * V.v.z);
  V3 Btw = mk(I.B[0] * V.w.x + I.B[3] * V.w.y + I.B[6] * V.w.z, I.B[1] * V.w.x + I.B[4] * V.w.y + I.B[7] * V.w.z,
              I.B[2] * V.w.x + I.B[5] * V.w.y + I.B[8] * V.w.z);
  return sv(symmul(I.A, V.w) + Bv, Btw + symmul(I.D, V.v));
}
DDB_DI ABI toABI(const RBI& I) {
  ABI r;
#pragma unroll
  for (int i = 0; i < 6; i++) r.A[i] = I.J[i];
  r.B[0] = 0; r.B[1] = -I.h.z; r.B[2] = I.h.y;
  r.B[3] = I.h.z; r.B[4] = 0; r.B[5] = -I.h.x;
  r.B[6] = -I.h.y; r.B[7] = I.h.x; r.B[8] = 0;
  r.D[0] = r.D[1] = r.D[2] = I.m;
  r.D[3] = r.D[4] = r.D[5] = 0;
  return r;
}
// IA -= psi * a a^T   (a = IA*S), GenericJoint.hpp:1825-1839
DDB_DI void rank1Downdate(ABI& I, SV a, double psi) {
  V3 pw = psi * a.w, pv = psi * a.v;
  I.A[0] -= pw.x * a.w.x; I.A[1] -= pw.y * a.w.y; I.A[2] -= pw.z * a.w.z;
  I.A[3] -= pw.x * a.w.y; I.A[4] -= pw.x * a.w.z; I.A[5] -= pw.y * a.w.z;
  I.D[0] -= pv.x * a.v.x; I.D[1] -= pv.y * a.v.y; I.D[2] -= pv.z * a.v.z;
  I.D[3] -= pv.x * a.v.y; I.D[4] -= pv.x * a.v.z; I.D[5] -= pv.y * a.v.z;
  I.B[0] -= pw.x * a.v.x; I.B[1] -= pw.x * a.v.y; I.B[2] -= pw.x * a.v.z;
  I.B[3] -= pw.y * a.v.x; I.B[4] -= pw.y * a.v.y; I.B[5] -= pw.y * a.v.z;
  I.B[6] -= pw.z * a.v.x; I.B[7] -= pw.z * a.v.y; I.B[8] -= pw.z * a.v.z;
}
DDB_DI M3 symToM3(const double* S) {
  M3 r;
  r.m[0] = S[0]; r.m[4] = S[1]; r.m[8] = S[2];
  r.m[1] = r.m[3] = S[3]; r.m[2] = r.m[6] = S[4]; r.m[5] = r.m[7] = S[5];
  return r;
}
// R S R^T for symmetric S -> symmetric (xx yy zz xy xz yz)
DDB_DI void rotSym(const M3& R, const double* S, double* out) {
  M3 RS = mul(R, symToM3(S));
  out[0] = RS.m[0] * R.m[0] + RS.m[1] * R.m[1] + RS.m[2] * R.m[2];
  out[1] = RS.m[3] * R.m[3] + RS.m[4] * R.m[4] + RS.m[5] * R.m[5];
  out[2] = RS.m[6] * R.m[6] + RS.m[7] * R.m[7] + RS.m[8] * R.m[8];
  out[3] = RS.m[0] * R.m[3] + RS.m[1] * R.m[4] + RS.m[2] * R.m[5];
  out[4] = RS.m[0] * R.m[6] + RS.m[1] * R.m[7] + RS.m[2] * R.m[8];
  out[5] = RS.m[3] * R.m[6] + RS.m[4] * R.m[7] + RS.m[5] * R.m[8];
}
// parent += X^T Pi X, X = AdInvT(T): math::transformInertia(T^-1, Pi) (Geometry.cpp:1769-1851;
// BodyNode.cpp:1659-1686). Rotate the blocks into the parent orientation, then shift by p:
//   D' = R D R^T, B' = R B R^T, A' = R A R^T
//   Y_D = D', Y_B = B' + P D', Y_A = A' + P B'^T + Y_B P^T,  P = [p]x
DDB_DI void addTransformedInertia(ABI& parent, const Xf& T, const ABI& Pi) {
  double Dp[6], Ap[6];
  rotSym(T.R, Pi.D, Dp);
  rotSym(T.R, Pi.A, Ap);
  M3 Bm;
#pragma unroll
  for (int i = 0; i < 9; i++) Bm.m[i] = Pi.B[i];
  M3 Bp = mulABt(mul(T.R, Bm), T.R);
  M3 Dm = symToM3(Dp);
  const V3 p = T.p;
  // Y_B = B' + P D' : column j of (P D') = p x (column j of D')
  M3 YB;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    V3 c = cross(p, mk(Dm.m[j], Dm.m[3 + j], Dm.m[6 + j]));
    YB.m[j] = Bp.m[j] + c.x; YB.m[3 + j] = Bp.m[3 + j] + c.y; YB.m[6 + j] = Bp.m[6 + j] + c.z;
  }
  // P B'^T : column j = p x (row j of B');  Y_B P^T : row i = -(row i of Y_B) x p ... (Y_B P^T)_i. = p x? no:
  // (Y_B P^T) row i = (P (Y_B row i)^T)^T = (p x rowi)^T
  V3 pb0 = cross(p, mk(Bp.m[0], Bp.m[1], Bp.m[2]));  // column 0 of P B'^T
  V3 pb1 = cross(p, mk(Bp.m[3], Bp.m[4], Bp.m[5]));
  V3 pb2 = cross(p, mk(Bp.m[6], Bp.m[7], Bp.m[8]));
  V3 yb0 = cross(p, mk(YB.m[0], YB.m[1], YB.m[2]));  // row 0 of Y_B P^T
  V3 yb1 = cross(p, mk(YB.m[3], YB.m[4], YB.m[5]));
  V3 yb2 = cross(p, mk(YB.m[6], YB.m[7], YB.m[8]));
  // (P B'^T)(r,c) = pb_c.[r];  (Y_B P^T)(r,c) = yb_r.[c]
  parent.A[0] += Ap[0] + pb0.x + yb0.x;
  parent.A[1] += Ap[1] + pb1.y + yb1.y;
  parent.A[2] += Ap[2] + pb2.z + yb2.z;
  parent.A[3] += Ap[3] + pb1.x + yb0.y;
  parent.A[4] += Ap[4] + pb2.x + yb0.z;
  parent.A[5] += Ap[5] + pb2.y + yb1.z;
#pragma unroll
  for (int i = 0; i < 9; i++) parent.B[i] += YB.m[i];
#pragma unroll
  for (int i = 0; i < 6; i++) parent.D[i] += Dp[i];
}

}  // namespace ddb
