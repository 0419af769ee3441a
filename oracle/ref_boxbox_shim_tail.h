// ---- end of the reference fragment ----
}  // namespace collision
}  // namespace dart

// C entry: collideBoxBox(size0, T0, size1, T1) exactly as DARTCollisionDetector calls it.
// T: 12 doubles (R row-major, p). out: per contact 24 doubles:
// point(3) normal(3) depth(1) type(1) edgeAClosest(3) edgeAFixed(3) edgeADir(3) edgeBClosest(3)... see below.
extern "C" int ddref_collide_box_box(const double* size0, const double* T0, const double* size1, const double* T1, double* out, int max_contacts) {
  Eigen::Isometry3d a, b;
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) { a.m[r][c] = T0[3 * r + c]; b.m[r][c] = T1[3 * r + c]; } a.m[r][3] = T0[9 + r]; b.m[r][3] = T1[9 + r]; }
  dart::collision::CollisionResult res;
  dart::collision::collideBoxBox(nullptr, nullptr, Eigen::Vector3d(size0[0], size0[1], size0[2]), a, Eigen::Vector3d(size1[0], size1[1], size1[2]), b, res);
  int n = (int)res.contacts.size();
  for (int k = 0; k < n && k < max_contacts; k++) {
    const dart::collision::Contact& c = res.contacts[k];
    double* o = out + 26 * k;
    const Eigen::Vector3d* vs[8] = {&c.point, &c.normal, &c.edgeAClosestPoint, &c.edgeAFixedPoint, &c.edgeADir, &c.edgeBClosestPoint, &c.edgeBFixedPoint, &c.edgeBDir};
    for (int v = 0; v < 8; v++) for (int i = 0; i < 3; i++) o[3 * v + i] = (*vs[v])[i];
    o[24] = c.penetrationDepth;
    o[25] = (double)c.type;
  }
  return n;
}
