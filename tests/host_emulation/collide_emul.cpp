// TEST INFRASTRUCTURE ONLY — see cuda_shim.h. Host build of diffdart_b200/csrc/collide.cuh.
#include "cuda_shim.h"
#include "collide_nocuda.cuh"
using namespace ddb;
extern "C" int emul_box_box(const double* size0, const double* T0, const double* size1, const double* T1, double* out) {
  Pose a, b;
  for (int i = 0; i < 9; i++) { a.R[i] = T0[i]; b.R[i] = T1[i]; }
  for (int i = 0; i < 3; i++) { a.p[i] = T0[9 + i]; b.p[i] = T1[9 + i]; }
  ContactPoint c[8];
  int n = boxBox(size0, a, size1, b, c);
  for (int k = 0; k < n; k++) {
    double* o = out + 26 * k;
    const double* vs[8] = {c[k].point, c[k].normal, c[k].edgeAClosest, c[k].edgeAFixed, c[k].edgeADir, c[k].edgeBClosest, c[k].edgeBFixed, c[k].edgeBDir};
    for (int v = 0; v < 8; v++) for (int i = 0; i < 3; i++) o[3 * v + i] = vs[v][i];
    o[24] = c[k].depth; o[25] = c[k].type;
  }
  return n;
}
extern "C" int emul_collide_pair(int shape1, const double* size1, const double* T1, int shape2, const double* size2, const double* T2, double* out) {
  Pose a, b;
  for (int i = 0; i < 9; i++) { a.R[i] = T1[i]; b.R[i] = T2[i]; }
  for (int i = 0; i < 3; i++) { a.p[i] = T1[9 + i]; b.p[i] = T2[9 + i]; }
  ContactPoint c[8];
  int n = collidePair(shape1, size1, a, shape2, size2, b, c);
  for (int k = 0; k < n; k++) {
    double* o = out + 26 * k;
    const double* vs[8] = {c[k].point, c[k].normal, c[k].edgeAClosest, c[k].edgeAFixed, c[k].edgeADir, c[k].edgeBClosest, c[k].edgeBFixed, c[k].edgeBDir};
    for (int v = 0; v < 8; v++) for (int i = 0; i < 3; i++) o[3 * v + i] = vs[v][i];
    o[24] = c[k].depth; o[25] = c[k].type;
  }
  return n;
}
