// TEST INFRASTRUCTURE ONLY.
// Shim that lets the REFERENCE's own box-box narrow phase (dart/collision/dart/DARTCollide.cpp:49-1427:
// cullPoints ... dBoxBox ... collideBoxBox) be compiled where it lies, without Eigen and without the rest of
// DART: oracle/Makefile pipes that line range of the reference file between this header and
// ref_boxbox_shim_tail.h straight into g++ (nothing from the reference is copied into the repo; the output is
// oracle/_ref/libboxbox_ref.so). The few Eigen / dart::collision names the fragment touches are stood in for
// below with plain structs of the same arithmetic (no expression templates, left-to-right evaluation).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
using std::abs;

namespace Eigen {
struct Vector3d;
struct CommaInit3 {
  Vector3d* v;
  int k;
  CommaInit3& operator,(double x);
};
struct Vector3d {
  double d[3];
  Vector3d() { d[0] = d[1] = d[2] = 0; }
  Vector3d(double a, double b, double c) { d[0] = a; d[1] = b; d[2] = c; }
  double& operator[](int i) { return d[i]; }
  const double& operator[](int i) const { return d[i]; }
  double& operator()(int i) { return d[i]; }
  CommaInit3 operator<<(double x) { d[0] = x; return CommaInit3{this, 1}; }
  double dot(const Vector3d& o) const { return d[0] * o.d[0] + d[1] * o.d[1] + d[2] * o.d[2]; }
  double squaredNorm() const { return dot(*this); }
  double norm() const { return std::sqrt(squaredNorm()); }
  // Eigen: normalize() divides by sqrt(squaredNorm()) when squaredNorm() > 0
  void normalize() { double z = squaredNorm(); if (z > 0) { double n = std::sqrt(z); d[0] /= n; d[1] /= n; d[2] /= n; } }
  Vector3d normalized() const { Vector3d r = *this; r.normalize(); return r; }
  Vector3d cross(const Vector3d& o) const { return Vector3d(d[1] * o.d[2] - d[2] * o.d[1], d[2] * o.d[0] - d[0] * o.d[2], d[0] * o.d[1] - d[1] * o.d[0]); }
  Vector3d& operator*=(double s) { d[0] *= s; d[1] *= s; d[2] *= s; return *this; }
  Vector3d& operator+=(const Vector3d& o) { d[0] += o.d[0]; d[1] += o.d[1]; d[2] += o.d[2]; return *this; }
  Vector3d& operator-=(const Vector3d& o) { d[0] -= o.d[0]; d[1] -= o.d[1]; d[2] -= o.d[2]; return *this; }
};
inline CommaInit3& CommaInit3::operator,(double x) { v->d[k++] = x; return *this; }
inline Vector3d operator+(const Vector3d& a, const Vector3d& b) { return Vector3d(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vector3d operator-(const Vector3d& a, const Vector3d& b) { return Vector3d(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Vector3d operator*(const Vector3d& a, double s) { return Vector3d(a[0] * s, a[1] * s, a[2] * s); }
inline Vector3d operator*(double s, const Vector3d& a) { return Vector3d(s * a[0], s * a[1], s * a[2]); }
struct Isometry3d {
  double m[3][4];
  double operator()(int r, int c) const { return m[r][c]; }
  Vector3d translation() const { return Vector3d(m[0][3], m[1][3], m[2][3]); }
};
}  // namespace Eigen

namespace dart {
namespace math {
struct constantsd { static double pi() { return 3.14159265358979323846; } };
}
namespace collision {
class CollisionObject;
enum ContactType { UNSUPPORTED = 0, VERTEX_FACE = 1, FACE_VERTEX = 2, EDGE_EDGE = 3, SPHERE_BOX = 4, BOX_SPHERE = 5, SPHERE_SPHERE = 6 };
struct Contact {
  CollisionObject* collisionObject1 = nullptr;
  CollisionObject* collisionObject2 = nullptr;
  Eigen::Vector3d point, normal;
  double penetrationDepth = 0;
  int type = 0;
  Eigen::Vector3d edgeAClosestPoint, edgeAFixedPoint, edgeADir, edgeBClosestPoint, edgeBFixedPoint, edgeBDir;
};
struct CollisionResult {
  std::vector<Contact> contacts;
  void addContact(const Contact& c) { contacts.push_back(c); }
};
}  // namespace collision
}  // namespace dart
// ---- the reference fragment (opens namespace dart { namespace collision {) follows ----
