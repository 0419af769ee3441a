// collide.cuh — narrow phase of the in-tree DART collision detector for the shapes on the hot path,
// one pair per call, FP64. Restates (arithmetic order kept; this header is compiled with -fmad=false):
//   boxBox()        collideBoxBox -> dBoxBox          dart/collision/dart/DARTCollide.cpp:713-1400,1402-1427
//                   (separating-axis search over 15 axes with the 1.05 edge fudge factor, edge-edge closest
//                   approach :225-250, incident-face clipping intersectRectQuad :483-541, up to 8 contacts per
//                   pair — DiffDART removed ODE's culling to 4 — and the VERTEX_FACE / FACE_VERTEX / EDGE_EDGE
//                   typing with edge metadata the backward needs :1207-1330)
//   boxSphere()     collideBoxSphere                  DARTCollide.cpp:1429-1600
//   sphereSphere()  collideSphereSphere               DARTCollide.cpp:1735-1800
// The reference's own dBoxBox is compiled as an oracle from where it lies (oracle/_ref/libboxbox_ref.so).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ddb {

#define DDB_COL_DI __device__ __forceinline__

// dart/collision/Contact.hpp:43-80
enum { CT_UNSUPPORTED = 0, CT_VERTEX_FACE = 1, CT_FACE_VERTEX = 2, CT_EDGE_EDGE = 3, CT_SPHERE_BOX = 4, CT_BOX_SPHERE = 5, CT_SPHERE_SPHERE = 6 };

struct ContactPoint {  // collision::Contact (Contact.hpp:83-229), the fields this path fills
  double point[3], normal[3], depth;
  int type;
  double edgeAClosest[3], edgeAFixed[3], edgeADir[3], edgeBClosest[3], edgeBFixed[3], edgeBDir[3];
};

struct Pose {  // world transform of a shape: R row-major, p
  double R[9], p[3];
};

DDB_COL_DI double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// column c of R dotted with a (Inner41), column with column (Inner44)
DDB_COL_DI double colDot(const double* R, int c, const double* a) { return R[c] * a[0] + R[3 + c] * a[1] + R[6 + c] * a[2]; }
DDB_COL_DI double colCol(const double* Ra, int ca, const double* Rb, int cb) { return Ra[ca] * Rb[cb] + Ra[3 + ca] * Rb[3 + cb] + Ra[6 + ca] * Rb[6 + cb]; }
DDB_COL_DI void normalize3(double* v) {  // Eigen normalize(): divide by sqrt(squaredNorm) when > 0
  double z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (z > 0) { double n = sqrt(z); v[0] /= n; v[1] /= n; v[2] /= n; }
}

// DARTCollide.cpp:483-541 intersectRectQuad: clip quad p (4 points) against the rectangle |x|<h0, |y|<h1
DDB_COL_DI int intersectRectQuad(const double* h, const double* p, double* ret) {
  int nq = 4, nr = 0;
  double bufA[16], bufB[16];
  for (int i = 0; i < 8; i++) bufA[i] = p[i];
  double* q = bufA;
  double* r = bufB;
  bool full = false;
  for (int dir = 0; dir <= 1 && !full; dir++) {
    for (int sign = -1; sign <= 1 && !full; sign += 2) {
      nr = 0;
      for (int i = 0; i < nq; i++) {
        const double* pq = q + 2 * i;
        if (sign * pq[dir] < h[dir]) {
          r[2 * nr] = pq[0]; r[2 * nr + 1] = pq[1];
          nr++;
          if (nr & 8) { full = true; break; }
        }
        const double* nextq = (i < nq - 1) ? pq + 2 : q;
        if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {
          r[2 * nr + (1 - dir)] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
          r[2 * nr + dir] = sign * h[dir];
          nr++;
          if (nr & 8) { full = true; break; }
        }
      }
      double* t = q; q = r; r = t;
      if (!full) nq = nr;
    }
  }
  for (int i = 0; i < nr * 2; i++) ret[i] = q[i];
  return nr;
}

// collideBoxBox(size0, T0, size1, T1): contacts point from object 2 toward object 1 (normal = -n), depth >= 0.
// Returns the number of contacts written (<= 8).
DDB_COL_DI int boxBox(const double* size0, const Pose& T0, const double* size1, const Pose& T1, ContactPoint* out) {
  const double fudge_factor = 1.05;
  const double* R1 = T0.R; const double* R2 = T1.R; const double* p1 = T0.p; const double* p2 = T1.p;
  double A[3] = {0.5 * size0[0], 0.5 * size0[1], 0.5 * size0[2]};
  double B[3] = {0.5 * size1[0], 0.5 * size1[1], 0.5 * size1[2]};
  double p[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double pp[3] = {colDot(R1, 0, p), colDot(R1, 1, p), colDot(R1, 2, p)};
  double R[3][3], Q[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { R[i][j] = colCol(R1, i, R2, j); Q[i][j] = fabs(R[i][j]); }
  double s = -1E12, s2, l;
  int invert_normal = 0, code = 0;
  int normalRbox = 0, normalRcol = 0;  // normalR = column normalRcol of R1 (box 1) or R2 (box 2)
  bool haveNormalR = false;
  double normalC[3] = {0, 0, 0};
#define DDB_TST1(expr1, expr2, boxi, col, cc) { double e1 = (expr1); s2 = fabs(e1) - (expr2); if (s2 > s) { s = s2; haveNormalR = true; normalRbox = boxi; normalRcol = col; invert_normal = (e1 < 0); code = (cc); } }
  DDB_TST1(pp[0], (A[0] + B[0] * Q[0][0] + B[1] * Q[0][1] + B[2] * Q[0][2]), 1, 0, 1);
  DDB_TST1(pp[1], (A[1] + B[0] * Q[1][0] + B[1] * Q[1][1] + B[2] * Q[1][2]), 1, 1, 2);
  DDB_TST1(pp[2], (A[2] + B[0] * Q[2][0] + B[1] * Q[2][1] + B[2] * Q[2][2]), 1, 2, 3);
  DDB_TST1(colDot(R2, 0, p), (A[0] * Q[0][0] + A[1] * Q[1][0] + A[2] * Q[2][0] + B[0]), 2, 0, 4);
  DDB_TST1(colDot(R2, 1, p), (A[0] * Q[0][1] + A[1] * Q[1][1] + A[2] * Q[2][1] + B[1]), 2, 1, 5);
  DDB_TST1(colDot(R2, 2, p), (A[0] * Q[0][2] + A[1] * Q[1][2] + A[2] * Q[2][2] + B[2]), 2, 2, 6);
#undef DDB_TST1
#define DDB_TST2(expr1, expr2, n1, n2, n3, cc) { double e1 = (expr1); s2 = fabs(e1) - (expr2); l = sqrt((n1) * (n1) + (n2) * (n2) + (n3) * (n3)); \
    if (l > 0) { s2 /= l; if (s2 * fudge_factor > s) { s = s2; haveNormalR = false; normalC[0] = (n1) / l; normalC[1] = (n2) / l; normalC[2] = (n3) / l; invert_normal = (e1 < 0); code = (cc); } } }
  DDB_TST2(pp[2] * R[1][0] - pp[1] * R[2][0], (A[1] * Q[2][0] + A[2] * Q[1][0] + B[1] * Q[0][2] + B[2] * Q[0][1]), 0, -R[2][0], R[1][0], 7);
  DDB_TST2(pp[2] * R[1][1] - pp[1] * R[2][1], (A[1] * Q[2][1] + A[2] * Q[1][1] + B[0] * Q[0][2] + B[2] * Q[0][0]), 0, -R[2][1], R[1][1], 8);
  DDB_TST2(pp[2] * R[1][2] - pp[1] * R[2][2], (A[1] * Q[2][2] + A[2] * Q[1][2] + B[0] * Q[0][1] + B[1] * Q[0][0]), 0, -R[2][2], R[1][2], 9);
  DDB_TST2(pp[0] * R[2][0] - pp[2] * R[0][0], (A[0] * Q[2][0] + A[2] * Q[0][0] + B[1] * Q[1][2] + B[2] * Q[1][1]), R[2][0], 0, -R[0][0], 10);
  DDB_TST2(pp[0] * R[2][1] - pp[2] * R[0][1], (A[0] * Q[2][1] + A[2] * Q[0][1] + B[0] * Q[1][2] + B[2] * Q[1][0]), R[2][1], 0, -R[0][1], 11);
  DDB_TST2(pp[0] * R[2][2] - pp[2] * R[0][2], (A[0] * Q[2][2] + A[2] * Q[0][2] + B[0] * Q[1][1] + B[1] * Q[1][0]), R[2][2], 0, -R[0][2], 12);
  DDB_TST2(pp[1] * R[0][0] - pp[0] * R[1][0], (A[0] * Q[1][0] + A[1] * Q[0][0] + B[1] * Q[2][2] + B[2] * Q[2][1]), -R[1][0], R[0][0], 0, 13);
  DDB_TST2(pp[1] * R[0][1] - pp[0] * R[1][1], (A[0] * Q[1][1] + A[1] * Q[0][1] + B[0] * Q[2][2] + B[2] * Q[2][0]), -R[1][1], R[0][1], 0, 14);
  DDB_TST2(pp[1] * R[0][2] - pp[0] * R[1][2], (A[0] * Q[1][2] + A[1] * Q[0][2] + B[0] * Q[2][1] + B[1] * Q[2][0]), -R[1][2], R[0][2], 0, 15);
#undef DDB_TST2
  if (!code) return 0;
  if (s > 0.0) return 0;

  double normal[3];
  if (haveNormalR) {
    const double* Rn = normalRbox == 1 ? R1 : R2;
    normal[0] = Rn[normalRcol]; normal[1] = Rn[3 + normalRcol]; normal[2] = Rn[6 + normalRcol];
  } else {
    normal[0] = dot3(R1, normalC); normal[1] = dot3(R1 + 3, normalC); normal[2] = dot3(R1 + 6, normalC);
    normalize3(normal);
  }
  if (invert_normal) { normal[0] *= -1.0; normal[1] *= -1.0; normal[2] *= -1.0; }

  if (code > 6) {
    // edge-edge: closest points of the two edges (DARTCollide.cpp:918-1000)
    double pa[3] = {p1[0], p1[1], p1[2]};
    for (int j = 0; j < 3; j++) {
      double val = normal[0] * R1[j] + normal[1] * R1[3 + j] + normal[2] * R1[6 + j];
      double sign = (val > -1e-10) ? 1.0 : -1.0;
      for (int i = 0; i < 3; i++) pa[i] += sign * A[j] * R1[i * 3 + j];
    }
    double pb[3] = {p2[0], p2[1], p2[2]};
    for (int j = 0; j < 3; j++) {
      double val = normal[0] * R2[j] + normal[1] * R2[3 + j] + normal[2] * R2[6 + j];
      double sign = (val > -1e-3) ? -1.0 : 1.0;
      for (int i = 0; i < 3; i++) pb[i] += sign * B[j] * R2[i * 3 + j];
    }
    double ua[3], ub[3];
    for (int i = 0; i < 3; i++) { ua[i] = R1[(code - 7) / 3 + i * 3]; ub[i] = R2[(code - 7) % 3 + i * 3]; }
    // dLineClosestApproach (:225-250)
    double alpha, beta;
    {
      double pd[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
      double uaub = dot3(ua, ub), q1 = dot3(ua, pd), q2 = -dot3(ub, pd);
      double d = 1 - uaub * uaub;
      if (d <= 0) { alpha = 0; beta = 0; }
      else { d = 1.0 / d; alpha = (q1 + uaub * q2) * d; beta = (uaub * q1 + q2) * d; }
    }
    ContactPoint& c = out[0];
    for (int i = 0; i < 3; i++) { c.edgeAFixed[i] = pa[i]; c.edgeBFixed[i] = pb[i]; }
    for (int i = 0; i < 3; i++) pa[i] += ua[i] * alpha;
    for (int i = 0; i < 3; i++) pb[i] += ub[i] * beta;
    for (int i = 0; i < 3; i++) {
      c.point[i] = 0.5 * (pa[i] + pb[i]);
      c.normal[i] = normal[i] * -1;
      c.edgeAClosest[i] = pa[i]; c.edgeBClosest[i] = pb[i];
      c.edgeADir[i] = ua[i]; c.edgeBDir[i] = ub[i];
    }
    normalize3(c.edgeADir);
    normalize3(c.edgeBDir);
    c.depth = -s;
    c.type = CT_EDGE_EDGE;
    return 1;
  }

  // face-something: clip the incident face of the other box against the reference face (:1002-1190)
  const double *Ra, *Rb, *pa, *pb, *Sa, *Sb;
  bool flip;
  if (code <= 3) { Ra = R1; Rb = R2; pa = p1; pb = p2; Sa = A; Sb = B; flip = false; }
  else { Ra = R2; Rb = R1; pa = p2; pb = p1; Sa = B; Sb = A; flip = true; }
  double normal2[3], nr[3], anr[3];
  for (int i = 0; i < 3; i++) normal2[i] = (code <= 3) ? normal[i] : -normal[i];
  for (int i = 0; i < 3; i++) { nr[i] = colDot(Rb, i, normal2); anr[i] = fabs(nr[i]); }
  int lanr, a1, a2;
  if (anr[1] > anr[0]) {
    if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  } else {
    if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  }
  double center[3];
  if (nr[lanr] < 0) { for (int i = 0; i < 3; i++) center[i] = pb[i] - pa[i] + Sb[lanr] * Rb[i * 3 + lanr]; }
  else { for (int i = 0; i < 3; i++) center[i] = pb[i] - pa[i] - Sb[lanr] * Rb[i * 3 + lanr]; }
  int codeN = (code <= 3) ? code - 1 : code - 4, code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  double quad[8];
  double c1 = center[0] * Ra[code1] + center[1] * Ra[3 + code1] + center[2] * Ra[6 + code1];
  double c2 = center[0] * Ra[code2] + center[1] * Ra[3 + code2] + center[2] * Ra[6 + code2];
  double m11 = colCol(Ra, code1, Rb, a1), m12 = colCol(Ra, code1, Rb, a2), m21 = colCol(Ra, code2, Rb, a1), m22 = colCol(Ra, code2, Rb, a2);
  {
    double k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  double rect[2] = {Sa[code1], Sa[code2]};
  double ret[16];
  int n = intersectRectQuad(rect, quad, ret);
  if (n < 1) return 0;
  double point[24], dep[8];
  double det1 = 1.0 / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnum = 0;
  for (int j = 0; j < n; j++) {
    double k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    double k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    for (int i = 0; i < 3; i++) point[cnum * 3 + i] = center[i] + k1 * Rb[i * 3 + a1] + k2 * Rb[i * 3 + a2];
    dep[cnum] = Sa[codeN] - dot3(normal2, point + cnum * 3);
    if (dep[cnum] >= 0) { ret[cnum * 2] = ret[j * 2]; ret[cnum * 2 + 1] = ret[j * 2 + 1]; cnum++; }
  }
  if (cnum < 1) return 0;
  double otherNormal[3] = {Rb[lanr], Rb[3 + lanr], Rb[6 + lanr]};
  if (dot3(otherNormal, normal) < 0) { otherNormal[0] *= -1; otherNormal[1] *= -1; otherNormal[2] *= -1; }
  double ortho1[3] = {Rb[a1], Rb[3 + a1], Rb[6 + a1]}, ortho2[3] = {Rb[a2], Rb[3 + a2], Rb[6 + a2]};
  double centerB[3] = {pb[0], pb[1], pb[2]};
  double faceCenter[3];
  for (int i = 0; i < 3; i++) faceCenter[i] = centerB[i] - Sb[lanr] * otherNormal[i];
  for (int j = 0; j < cnum; j++) {
    ContactPoint& c = out[j];
    const double x = ret[j * 2], y = ret[j * 2 + 1];
    const double rectWidthX = rect[0], rectWidthY = rect[1];
    for (int i = 0; i < 3; i++) { c.point[i] = point[j * 3 + i] + pa[i]; c.normal[i] = normal[i] * -1; }
    c.depth = dep[j];
    for (int i = 0; i < 3; i++) { c.edgeAClosest[i] = c.edgeAFixed[i] = c.edgeADir[i] = c.edgeBClosest[i] = c.edgeBFixed[i] = c.edgeBDir[i] = 0; }
    const bool onEdgeX = fabs(x) == rectWidthX, onEdgeY = fabs(y) == rectWidthY;
    if (onEdgeX && onEdgeY) {
      if (flip) { c.type = CT_FACE_VERTEX; for (int i = 0; i < 3; i++) c.point[i] += c.normal[i] * c.depth; }
      else { c.type = CT_VERTEX_FACE; for (int i = 0; i < 3; i++) c.point[i] -= c.normal[i] * c.depth; }
    } else if (!onEdgeX && !onEdgeY) {
      c.type = flip ? CT_VERTEX_FACE : CT_FACE_VERTEX;
    } else {
      c.type = CT_EDGE_EDGE;
      const double faceX = x > 0 ? rectWidthX : -rectWidthX, faceY = y > 0 ? rectWidthY : -rectWidthY;
      double eAF[3], eAD[3];
      for (int i = 0; i < 3; i++) {
        double faceCenterA = pa[i] + Sa[codeN] * normal[i];
        eAF[i] = faceCenterA + faceX * Ra[3 * i + code1] + faceY * Ra[3 * i + code2];
        eAD[i] = c.point[i] - eAF[i];
      }
      normalize3(eAD);
      const double incX = dot3(ortho1, c.point) - dot3(ortho1, centerB), incY = dot3(ortho2, c.point) - dot3(ortho2, centerB);
      const double signX = incX == 0 ? 1.0 : (incX / fabs(incX)), signY = incY == 0 ? 1.0 : (incY / fabs(incY));
      const double distX = fabs(fabs(incX) - Sb[a1]), distY = fabs(fabs(incY) - Sb[a2]);
      double nearest[3], other[3], eBD[3];
      for (int i = 0; i < 3; i++) {
        nearest[i] = (signX * Sb[a1]) * ortho1[i] + (signY * Sb[a2]) * ortho2[i] + faceCenter[i];
        if (distX < distY) other[i] = (signX * Sb[a1]) * ortho1[i] + (-1 * signY * Sb[a2]) * ortho2[i] + faceCenter[i];
        else other[i] = (-1 * signX * Sb[a1]) * ortho1[i] + (signY * Sb[a2]) * ortho2[i] + faceCenter[i];
        eBD[i] = nearest[i] - other[i];
      }
      normalize3(eBD);
      for (int i = 0; i < 3; i++) {
        c.edgeAClosest[i] = c.point[i];
        c.edgeAFixed[i] = flip ? nearest[i] : eAF[i];
        c.edgeADir[i] = flip ? eBD[i] : eAD[i];
        c.edgeBFixed[i] = flip ? eAF[i] : nearest[i];
        c.edgeBDir[i] = flip ? eAD[i] : eBD[i];
      }
    }
  }
  return cnum;
}

// collideBoxSphere (DARTCollide.cpp:1429-1581): box is object 1, sphere object 2.
// sphereIsFirst = true gives collideSphereBox (:1583-1733): same clamp, opposite normal, SPHERE_BOX / VERTEX_FACE typing.
DDB_COL_DI int boxSphereImpl(const double* boxSize, const Pose& Tb, double r, const Pose& Ts, bool sphereIsFirst, ContactPoint* out) {
  const double half[3] = {0.5 * boxSize[0], 0.5 * boxSize[1], 0.5 * boxSize[2]};
  const double* c0 = Ts.p;
  // p = Tb^-1 * c0
  double d[3] = {c0[0] - Tb.p[0], c0[1] - Tb.p[1], c0[2] - Tb.p[2]};
  double p[3] = {colDot(Tb.R, 0, d), colDot(Tb.R, 1, d), colDot(Tb.R, 2, d)};
  bool inside = true;
  int locks = 0;  // face{1,2,3}Locked (DARTCollide.cpp:1452-1493)
  for (int k = 0; k < 3; k++) {
    if (p[k] < -half[k]) { p[k] = -half[k]; inside = false; locks |= 1 << k; }
    if (p[k] > half[k]) { p[k] = half[k]; inside = false; locks |= 1 << k; }
  }
  ContactPoint& c = out[0];
  for (int i = 0; i < 3; i++) c.edgeAClosest[i] = c.edgeAFixed[i] = c.edgeADir[i] = c.edgeBClosest[i] = c.edgeBFixed[i] = c.edgeBDir[i] = 0;
  for (int i = 0; i < 3; i++) c.edgeAFixed[i] = c0[i];  // sphereCenter travels in edgeAFixed
  c.edgeADir[0] = r;                                      // sphereRadius
  c.edgeADir[1] = (double)locks;                          // face lock mask; face normals = box axes:
  for (int i = 0; i < 3; i++) { c.edgeBClosest[i] = Tb.R[3 * i]; c.edgeBFixed[i] = Tb.R[3 * i + 1]; c.edgeBDir[i] = Tb.R[3 * i + 2]; }
  c.type = sphereIsFirst ? CT_SPHERE_BOX : CT_BOX_SPHERE;
  auto nearestSide = [&](int& idx) {
    double mn = half[0] - fabs(p[0]);
    double t = half[1] - fabs(p[1]);
    idx = 0;
    if (t < mn) { mn = t; idx = 1; }
    t = half[2] - fabs(p[2]);
    if (t < mn) { mn = t; idx = 2; }
    return mn;
  };
  auto axisNormal = [&](int idx, double* n) {
    double e[3] = {0, 0, 0};
    // box-sphere: (p>0 ? -1 : 1); sphere-box: (p>0 ? 1 : -1)
    e[idx] = sphereIsFirst ? (p[idx] > 0.0 ? 1.0 : -1.0) : (p[idx] > 0.0 ? -1.0 : 1.0);
    for (int i = 0; i < 3; i++) n[i] = Tb.R[3 * i] * e[0] + Tb.R[3 * i + 1] * e[1] + Tb.R[3 * i + 2] * e[2];
  };
  if (inside) {
    int idx;
    double mn = nearestSide(idx);
    axisNormal(idx, c.normal);
    c.depth = mn + r;
    c.type = sphereIsFirst ? CT_VERTEX_FACE : CT_FACE_VERTEX;
    for (int i = 0; i < 3; i++) c.point[i] = c0[i];
    return 1;
  }
  double cp[3];
  for (int i = 0; i < 3; i++) cp[i] = Tb.R[3 * i] * p[0] + Tb.R[3 * i + 1] * p[1] + Tb.R[3 * i + 2] * p[2] + Tb.p[i];
  double n[3];
  for (int i = 0; i < 3; i++) n[i] = sphereIsFirst ? c0[i] - cp[i] : cp[i] - c0[i];
  const double mag = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  const double pen = r - mag;
  if (pen < 0.0) return 0;
  if (mag > 1E-6) {
    const double inv = 1.0 / mag;
    for (int i = 0; i < 3; i++) c.normal[i] = n[i] * inv;
  } else {
    int idx;
    nearestSide(idx);
    axisNormal(idx, c.normal);
  }
  for (int i = 0; i < 3; i++) c.point[i] = cp[i];
  c.depth = pen;
  return 1;
}

// collideSphereSphere (DARTCollide.cpp:1735-1800)
DDB_COL_DI int sphereSphere(double r0_, const Pose& c0, double r1_, const Pose& c1, ContactPoint* out) {
  double r0 = r0_, r1 = r1_;
  const double rsum = r0 + r1;
  double n[3] = {c0.p[0] - c1.p[0], c0.p[1] - c1.p[1], c0.p[2] - c1.p[2]};
  double nsq = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  if (nsq > rsum * rsum) return 0;
  r0 /= rsum;
  r1 /= rsum;
  ContactPoint& c = out[0];
  for (int i = 0; i < 3; i++) {
    c.point[i] = r1 * c0.p[i] + r0 * c1.p[i];
    c.edgeAClosest[i] = c.edgeADir[i] = c.edgeBClosest[i] = c.edgeBDir[i] = 0;
    c.edgeAFixed[i] = c0.p[i];  // centerA
    c.edgeBFixed[i] = c1.p[i];  // centerB
  }
  c.edgeADir[0] = r0 * rsum;  // radiusA
  c.edgeBDir[0] = r1 * rsum;  // radiusB
  c.type = CT_SPHERE_SPHERE;
  if (nsq < 1E-6) {
    c.normal[0] = c.normal[1] = c.normal[2] = 0;
    c.depth = rsum;
    return 1;
  }
  nsq = sqrt(nsq);
  const double inv = 1.0 / nsq;
  for (int i = 0; i < 3; i++) c.normal[i] = n[i] * inv;
  c.depth = rsum - nsq;
  return 1;
}

// collideCapsuleBox / collideBoxCapsule (DARTCollide.cpp:4217-4361): when the contact lies beyond an end plane of the cylinder part the
// reference collides that end's sphere with the box (collideSphereBox / collideBoxSphere on T * translate(0, 0, +-height / 2)); it takes
// the contact position from libccd's MPR, which is not part of the reference tree. Here the end is chosen in closed form: with
// f(t) = squared distance of the axis point e0 + t (e1 - e0) to the box (convex in t), end 0 is the closest feature iff f'(0) > 0,
// end 1 iff f'(1) < 0; an end point inside the box counts as that end. Everything else (closest feature on the cylinder part: the
// reference's pipe contacts, createCapsuleMeshContact :2887-3300) is not carried: returns -1.
DDB_COL_DI int capsuleBox(const double* cap, const Pose& Tc, const double* boxSize, const Pose& Tb, bool capsuleIsFirst, ContactPoint* out) {
  const double r = cap[0], hh = 0.5 * cap[1];
  const double half[3] = {0.5 * boxSize[0], 0.5 * boxSize[1], 0.5 * boxSize[2]};
  double e[2][3], pl[2][3];  // end centres in the world and in the box frame
  for (int s = 0; s < 2; s++) {
    const double z = s == 0 ? -hh : hh;
    for (int i = 0; i < 3; i++) e[s][i] = Tc.R[3 * i + 2] * z + Tc.p[i];
    const double d[3] = {e[s][0] - Tb.p[0], e[s][1] - Tb.p[1], e[s][2] - Tb.p[2]};
    for (int k = 0; k < 3; k++) pl[s][k] = colDot(Tb.R, k, d);
  }
  // derivative of the squared distance along the axis at both ends, and the distances themselves
  double slope[2] = {0.0, 0.0}, dist2[2] = {0.0, 0.0};
  for (int s = 0; s < 2; s++)
    for (int k = 0; k < 3; k++) {
      double g = 0.0;
      if (pl[s][k] > half[k]) g = pl[s][k] - half[k];
      else if (pl[s][k] < -half[k]) g = pl[s][k] + half[k];
      slope[s] += g * (pl[1][k] - pl[0][k]);
      dist2[s] += g * g;
    }
  int end = -1;
  if (dist2[0] == 0.0 && dist2[1] == 0.0) return -1;
  if (dist2[0] == 0.0) end = 0;
  else if (dist2[1] == 0.0) end = 1;
  else if (slope[0] > 0.0) end = 0;
  else if (slope[1] < 0.0) end = 1;
  if (end < 0) {
    // closest feature on the cylinder part: only a miss can be reported (no point of the axis within the radius of the box)
    // f is convex: its minimum over [0, 1] is at least min(f(0) + slope terms ...) - decided by bisection on f'
    double lo = 0.0, hi = 1.0;
    for (int it = 0; it < 60; it++) {
      const double t = 0.5 * (lo + hi);
      double sl = 0.0;
      for (int k = 0; k < 3; k++) {
        const double x = pl[0][k] + t * (pl[1][k] - pl[0][k]);
        double g = 0.0;
        if (x > half[k]) g = x - half[k];
        else if (x < -half[k]) g = x + half[k];
        sl += g * (pl[1][k] - pl[0][k]);
      }
      if (sl > 0.0) hi = t; else lo = t;
    }
    const double t = 0.5 * (lo + hi);
    double f = 0.0;
    for (int k = 0; k < 3; k++) {
      const double x = pl[0][k] + t * (pl[1][k] - pl[0][k]);
      double g = 0.0;
      if (x > half[k]) g = x - half[k];
      else if (x < -half[k]) g = x + half[k];
      f += g * g;
    }
    return f > r * r ? 0 : -1;
  }
  Pose Ts = Tc;
  for (int i = 0; i < 3; i++) Ts.p[i] = e[end][i];
  return boxSphereImpl(boxSize, Tb, r, Ts, capsuleIsFirst, out);
}

// collide(o1, o2) dispatcher (DARTCollide.cpp:4675-5183) for SHAPE_BOX / SHAPE_SPHERE / SHAPE_CAPSULE; -1 = a pair this path does not carry
DDB_COL_DI int collidePair(int shape1, const double* size1, const Pose& T1, int shape2, const double* size2, const Pose& T2, ContactPoint* out) {
  if (shape1 == 2) {                                                     // sphere
    if (shape2 == 2) return sphereSphere(size1[0], T1, size2[0], T2, out);
    if (shape2 == 1) return boxSphereImpl(size2, T2, size1[0], T1, true, out);   // collideSphereBox
    if (shape2 == 3) return -1;
  } else if (shape1 == 1) {                                              // box
    if (shape2 == 2) return boxSphereImpl(size1, T1, size2[0], T2, false, out);  // collideBoxSphere
    if (shape2 == 1) return boxBox(size1, T1, size2, T2, out);
    if (shape2 == 3) return capsuleBox(size2, T2, size1, T1, false, out);        // collideBoxCapsule
  } else if (shape1 == 3) {                                              // capsule
    if (shape2 == 1) return capsuleBox(size1, T1, size2, T2, true, out);         // collideCapsuleBox
    return -1;
  }
  return 0;
}

}  // namespace ddb
