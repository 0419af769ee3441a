// TEST INFRASTRUCTURE ONLY — CPU oracle for the batched differentiable timestep.
// Nothing under oracle/ is part of the product; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, load or call it.
//
// dd_spatial.hpp — SE(3) spatial algebra, restating (not copying) the reference's
// dart/math/Geometry.cpp. Templated on the scalar so the same arithmetic can be
// evaluated in double (reference precision) and long double (derivative checks).
// Spatial vectors are [angular; linear], body-frame, as in DART.
#pragma once
#include <cmath>
#include <cstring>

namespace ddo {

template <class S> struct Vec3 {
  S v[3];
  S& operator[](int i) { return v[i]; }
  const S& operator[](int i) const { return v[i]; }
};
template <class S> struct Mat3 {
  S m[3][3];
};
template <class S> struct Vec6 {
  S v[6];
  S& operator[](int i) { return v[i]; }
  const S& operator[](int i) const { return v[i]; }
};
template <class S> struct Mat6 {
  S m[6][6];
};
// Rigid transform (Eigen::Isometry3d): x_parent = R x_child + p
template <class S> struct Iso {
  Mat3<S> R;
  Vec3<S> p;
};

template <class S> inline Vec3<S> v3(S a, S b, S c) { return Vec3<S>{{a, b, c}}; }
template <class S> inline Vec3<S> zero3() { return Vec3<S>{{S(0), S(0), S(0)}}; }
template <class S> inline Vec6<S> zero6() { Vec6<S> r; for (int i = 0; i < 6; i++) r[i] = S(0); return r; }
template <class S> inline Mat3<S> eye3() {
  Mat3<S> r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = S(i == j);
  return r;
}
template <class S> inline Mat6<S> zero66() {
  Mat6<S> r;
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) r.m[i][j] = S(0);
  return r;
}
template <class S> inline Iso<S> isoIdentity() { return Iso<S>{eye3<S>(), zero3<S>()}; }

template <class S> inline Vec3<S> operator+(const Vec3<S>& a, const Vec3<S>& b) { return v3<S>(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
template <class S> inline Vec3<S> operator-(const Vec3<S>& a, const Vec3<S>& b) { return v3<S>(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
template <class S> inline Vec3<S> operator-(const Vec3<S>& a) { return v3<S>(-a[0], -a[1], -a[2]); }
template <class S> inline Vec3<S> operator*(S s, const Vec3<S>& a) { return v3<S>(s * a[0], s * a[1], s * a[2]); }
template <class S> inline S dot(const Vec3<S>& a, const Vec3<S>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class S> inline Vec3<S> cross(const Vec3<S>& a, const Vec3<S>& b) {
  return v3<S>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
template <class S> inline S norm(const Vec3<S>& a) { return std::sqrt(dot(a, a)); }

template <class S> inline Vec3<S> mul(const Mat3<S>& A, const Vec3<S>& x) {
  Vec3<S> r;
  for (int i = 0; i < 3; i++) r[i] = A.m[i][0] * x[0] + A.m[i][1] * x[1] + A.m[i][2] * x[2];
  return r;
}
template <class S> inline Vec3<S> mulT(const Mat3<S>& A, const Vec3<S>& x) {
  Vec3<S> r;
  for (int i = 0; i < 3; i++) r[i] = A.m[0][i] * x[0] + A.m[1][i] * x[1] + A.m[2][i] * x[2];
  return r;
}
template <class S> inline Mat3<S> mul(const Mat3<S>& A, const Mat3<S>& B) {
  Mat3<S> r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
  return r;
}
template <class S> inline Mat3<S> transpose(const Mat3<S>& A) {
  Mat3<S> r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = A.m[j][i];
  return r;
}
template <class S> inline Mat3<S> skew(const Vec3<S>& a) {
  Mat3<S> r;
  r.m[0][0] = 0; r.m[0][1] = -a[2]; r.m[0][2] = a[1];
  r.m[1][0] = a[2]; r.m[1][1] = 0; r.m[1][2] = -a[0];
  r.m[2][0] = -a[1]; r.m[2][1] = a[0]; r.m[2][2] = 0;
  return r;
}
template <class S> inline Iso<S> mul(const Iso<S>& A, const Iso<S>& B) { return Iso<S>{mul(A.R, B.R), mul(A.R, B.p) + A.p}; }
template <class S> inline Iso<S> inverse(const Iso<S>& A) {
  Iso<S> r;
  r.R = transpose(A.R);
  r.p = -mul(r.R, A.p);
  return r;
}
template <class S> inline Vec3<S> apply(const Iso<S>& T, const Vec3<S>& x) { return mul(T.R, x) + T.p; }

template <class S> inline Vec6<S> operator+(const Vec6<S>& a, const Vec6<S>& b) { Vec6<S> r; for (int i = 0; i < 6; i++) r[i] = a[i] + b[i]; return r; }
template <class S> inline Vec6<S> operator-(const Vec6<S>& a, const Vec6<S>& b) { Vec6<S> r; for (int i = 0; i < 6; i++) r[i] = a[i] - b[i]; return r; }
template <class S> inline Vec6<S> operator-(const Vec6<S>& a) { Vec6<S> r; for (int i = 0; i < 6; i++) r[i] = -a[i]; return r; }
template <class S> inline Vec6<S> operator*(S s, const Vec6<S>& a) { Vec6<S> r; for (int i = 0; i < 6; i++) r[i] = s * a[i]; return r; }
template <class S> inline S dot(const Vec6<S>& a, const Vec6<S>& b) { S s = 0; for (int i = 0; i < 6; i++) s += a[i] * b[i]; return s; }
template <class S> inline Vec3<S> ang(const Vec6<S>& a) { return v3<S>(a[0], a[1], a[2]); }
template <class S> inline Vec3<S> lin(const Vec6<S>& a) { return v3<S>(a[3], a[4], a[5]); }
template <class S> inline Vec6<S> join(const Vec3<S>& w, const Vec3<S>& v) { return Vec6<S>{{w[0], w[1], w[2], v[0], v[1], v[2]}}; }
template <class S> inline Vec6<S> mul(const Mat6<S>& A, const Vec6<S>& x) {
  Vec6<S> r;
  for (int i = 0; i < 6; i++) { S s = 0; for (int j = 0; j < 6; j++) s += A.m[i][j] * x[j]; r[i] = s; }
  return r;
}
template <class S> inline Mat6<S> operator+(const Mat6<S>& A, const Mat6<S>& B) {
  Mat6<S> r;
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) r.m[i][j] = A.m[i][j] + B.m[i][j];
  return r;
}

// Geometry.cpp:1027-1038  AdT(T,V): w' = R w, v' = p x (R w) + R v
template <class S> inline Vec6<S> AdT(const Iso<S>& T, const Vec6<S>& V) {
  Vec3<S> w = mul(T.R, ang(V));
  return join(w, mul(T.R, lin(V)) + cross(T.p, w));
}
// Geometry.cpp:1121-1129  AdInvT(T,V): [R^T w ; R^T (v + w x p)]
template <class S> inline Vec6<S> AdInvT(const Iso<S>& T, const Vec6<S>& V) {
  return join(mulT(T.R, ang(V)), mulT(T.R, lin(V) + cross(ang(V), T.p)));
}
// Geometry.cpp:1145-1151
template <class S> inline Vec6<S> AdInvRLinear(const Iso<S>& T, const Vec3<S>& v) { return join(zero3<S>(), mulT(T.R, v)); }
// Geometry.cpp:1153-1167  ad(X,Y) = [w1 x w2 ; w1 x v2 + v1 x w2]
template <class S> inline Vec6<S> ad(const Vec6<S>& X, const Vec6<S>& Y) {
  return join(cross(ang(X), ang(Y)), cross(ang(X), lin(Y)) + cross(lin(X), ang(Y)));
}
// Geometry.cpp:1760-1767  dad(s,t) = [t.w x s.w + t.v x s.v ; t.v x s.w] — (s twist, t wrench)
template <class S> inline Vec6<S> dad(const Vec6<S>& s, const Vec6<S>& t) {
  return join(cross(ang(t), ang(s)) + cross(lin(t), lin(s)), cross(lin(t), ang(s)));
}
// Geometry.cpp:1195-1202  dAdInvT(T,F) = [R tau + p x (R f) ; R f]
template <class S> inline Vec6<S> dAdInvT(const Iso<S>& T, const Vec6<S>& F) {
  Vec3<S> f = mul(T.R, lin(F));
  return join(mul(T.R, ang(F)) + cross(T.p, f), f);
}
// Geometry.cpp:1040-1051 getAdTMatrix
template <class S> inline Mat6<S> AdTMatrix(const Iso<S>& T) {
  Mat6<S> r = zero66<S>();
  Mat3<S> pR = mul(skew(T.p), T.R);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      r.m[i][j] = T.R.m[i][j];
      r.m[3 + i][j] = pR.m[i][j];
      r.m[3 + i][3 + j] = T.R.m[i][j];
    }
  return r;
}
// Geometry.cpp:1769-1851 transformInertia(T, I) = AdT(T)^T I AdT(T), evaluated
// blockwise (the reference expands the same product into scalars).
template <class S> inline Mat6<S> transformInertia(const Iso<S>& T, const Mat6<S>& I) {
  Mat6<S> X = AdTMatrix(T);
  Mat6<S> IX = zero66<S>(), r = zero66<S>();
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) { S s = 0; for (int k = 0; k < 6; k++) s += I.m[i][k] * X.m[k][j]; IX.m[i][j] = s; }
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++) { S s = 0; for (int k = 0; k < 6; k++) s += X.m[k][i] * IX.m[k][j]; r.m[i][j] = s; r.m[j][i] = s; }
  return r;
}

#define DDO_DART_EPSILON 1e-6          /* Geometry.cpp:45 */
#define DDO_EPSILON_EXPMAP_THETA 1e-3  /* Geometry.cpp:469 */

// Geometry.cpp:1668-1702 expAngular
template <class S> inline Mat3<S> expAngular(const Vec3<S>& s) {
  S s2[3] = {s[0] * s[0], s[1] * s[1], s[2] * s[2]};
  S s3[3] = {s[0] * s[1], s[1] * s[2], s[2] * s[0]};
  S theta = std::sqrt(s2[0] + s2[1] + s2[2]);
  S cos_t = std::cos(theta), alpha, beta;
  if (theta > S(DDO_DART_EPSILON)) {
    alpha = std::sin(theta) / theta;
    beta = (S(1) - cos_t) / theta / theta;
  } else {
    alpha = S(1) - theta * theta / S(6);
    beta = S(0.5) - theta * theta / S(24);
  }
  Mat3<S> r;
  r.m[0][0] = beta * s2[0] + cos_t;
  r.m[1][0] = beta * s3[0] + alpha * s[2];
  r.m[2][0] = beta * s3[2] - alpha * s[1];
  r.m[0][1] = beta * s3[0] - alpha * s[2];
  r.m[1][1] = beta * s2[1] + cos_t;
  r.m[2][1] = beta * s3[1] + alpha * s[0];
  r.m[0][2] = beta * s3[2] + alpha * s[1];
  r.m[1][2] = beta * s3[1] - alpha * s[0];
  r.m[2][2] = beta * s2[2] + cos_t;
  return r;
}
// Geometry.cpp:471-487 expMapRot
template <class S> inline Mat3<S> expMapRot(const Vec3<S>& q) {
  S theta = norm(q);
  Mat3<S> qss = skew(q), qss2 = mul(qss, qss), R;
  S a, b;
  if (theta < S(DDO_EPSILON_EXPMAP_THETA)) { a = S(1); b = S(0.5); }
  else { a = std::sin(theta) / theta; b = (S(1) - std::cos(theta)) / (theta * theta); }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R.m[i][j] = S(i == j) + a * qss.m[i][j] + b * qss2.m[i][j];
  return R;
}
// Geometry.cpp:489-505 expMapJac
template <class S> inline Mat3<S> expMapJac(const Vec3<S>& q) {
  S theta = norm(q);
  Mat3<S> qss = skew(q), qss2 = mul(qss, qss), J;
  S a, b;
  if (theta < S(DDO_EPSILON_EXPMAP_THETA)) { a = S(0.5); b = S(1) / S(6); }
  else { a = (S(1) - std::cos(theta)) / (theta * theta); b = (theta - std::sin(theta)) / (theta * theta * theta); }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) J.m[i][j] = S(i == j) + a * qss.m[i][j] + b * qss2.m[i][j];
  return J;
}
// Geometry.cpp:610-650 logMap(R) = AngleAxis(R).angle * axis. Eigen (3.3.7, not
// vendored) builds the AngleAxis through a quaternion: Quaternion(Matrix3)
// (Shoemake's branch on the trace) then angle = 2 atan2(|vec|, |w|).
template <class S> inline Vec3<S> logMapRot(const Mat3<S>& M) {
  S q[4];  // x y z w
  S t = M.m[0][0] + M.m[1][1] + M.m[2][2];
  if (t > S(0)) {
    t = std::sqrt(t + S(1));
    q[3] = S(0.5) * t;
    t = S(0.5) / t;
    q[0] = (M.m[2][1] - M.m[1][2]) * t;
    q[1] = (M.m[0][2] - M.m[2][0]) * t;
    q[2] = (M.m[1][0] - M.m[0][1]) * t;
  } else {
    int i = 0;
    if (M.m[1][1] > M.m[0][0]) i = 1;
    if (M.m[2][2] > M.m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(M.m[i][i] - M.m[j][j] - M.m[k][k] + S(1));
    q[i] = S(0.5) * t;
    t = S(0.5) / t;
    q[3] = (M.m[k][j] - M.m[j][k]) * t;
    q[j] = (M.m[j][i] + M.m[i][j]) * t;
    q[k] = (M.m[k][i] + M.m[i][k]) * t;
  }
  S n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  if (n != S(0)) {
    S angle = S(2) * std::atan2(n, std::fabs(q[3]));
    if (q[3] < S(0)) n = -n;
    return v3<S>(angle * (q[0] / n), angle * (q[1] / n), angle * (q[2] / n));
  }
  return zero3<S>();
}

}  // namespace ddo
