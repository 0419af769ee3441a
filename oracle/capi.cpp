// TEST INFRASTRUCTURE ONLY — see dd_spatial.hpp.
// C entry points of the CPU oracle for ctypes (tests/, smoke(), bench.py cpu_baseline).
// Batched arrays are world-major: x[world * n + dof]; matrices [world][col][row].
#include <cstring>
#include <thread>
#include <vector>
#include <functional>
#include "dd_step.hpp"
#include "dd_lcp.hpp"
#include "dd_contact.hpp"
#include "dd_contact_grad.hpp"
#include <dlfcn.h>

using namespace ddo;

struct OracleWorld {
  std::vector<ddb_body> bodies;
  Model M;
};

static void parallelFor(int N, int threads, const std::function<void(int, int)>& fn) {
  if (threads <= 1 || N < 2) { fn(0, N); return; }
  std::vector<std::thread> pool;
  int per = (N + threads - 1) / threads;
  for (int t = 0; t < threads; t++) {
    int lo = t * per, hi = std::min(N, lo + per);
    if (lo >= hi) break;
    pool.emplace_back(fn, lo, hi);
  }
  for (auto& th : pool) th.join();
}

extern "C" {

void* ddo_create(const ddb_topology* t, const ddb_params* p) {
  OracleWorld* w = new OracleWorld();
  w->M = buildModel(*t, *p);
  return w;
}
void ddo_destroy(void* h) { delete (OracleWorld*)h; }
int ddo_num_dofs(void* h) { return ((OracleWorld*)h)->M.nd; }

void ddo_step(void* h, int N, int threads, const double* q, const double* qd, const double* tau, double* qn, double* qdn) {
  const Model& M = ((OracleWorld*)h)->M;
  int n = M.nd;
  parallelFor(N, threads, [&](int lo, int hi) {
    for (int w = lo; w < hi; w++) stepNoContact<double>(M, q + w * n, qd + w * n, tau + w * n, qn + w * n, qdn + w * n);
  });
}

// mode 0: the reference's recipe (its own finite-difference eps); mode 1: exact (long double, Richardson)
void ddo_jacobians(void* h, int N, int threads, int mode, const double* q, const double* qd, const double* tau, double* posPos,
                   double* posVel, double* velPos, double* velVel, double* forceVel) {
  const Model& M = ((OracleWorld*)h)->M;
  int n = M.nd, nn = n * n;
  parallelFor(N, threads, [&](int lo, int hi) {
    Jacobians J;
    for (int w = lo; w < hi; w++) {
      if (mode == 0) jacobiansReferenceNoContact(M, q + w * n, qd + w * n, tau + w * n, J);
      else jacobiansExactNoContact(M, q + w * n, qd + w * n, tau + w * n, J);
      if (posPos) std::memcpy(posPos + (size_t)w * nn, J.posPos.data(), nn * sizeof(double));
      if (posVel) std::memcpy(posVel + (size_t)w * nn, J.posVel.data(), nn * sizeof(double));
      if (velPos) std::memcpy(velPos + (size_t)w * nn, J.velPos.data(), nn * sizeof(double));
      if (velVel) std::memcpy(velVel + (size_t)w * nn, J.velVel.data(), nn * sizeof(double));
      if (forceVel) std::memcpy(forceVel + (size_t)w * nn, J.forceVel.data(), nn * sizeof(double));
    }
  });
}

// forward + backprop of one step per world: the CPU "world-step" of the baseline.
void ddo_step_backprop(void* h, int N, int threads, int mode, const double* q, const double* qd, const double* tau, const double* gqn,
                       const double* gqdn, double* qn, double* qdn, double* gq, double* gqd, double* gtau) {
  const Model& M = ((OracleWorld*)h)->M;
  int n = M.nd;
  parallelFor(N, threads, [&](int lo, int hi) {
    Jacobians J;
    for (int w = lo; w < hi; w++) {
      const double *qw = q + w * n, *qdw = qd + w * n, *tw = tau + w * n;
      if (qn) stepNoContact<double>(M, qw, qdw, tw, qn + w * n, qdn + w * n);
      if (mode == 0) jacobiansReferenceNoContact(M, qw, qdw, tw, J);
      else jacobiansExactNoContact(M, qw, qdw, tw, J);
      backprop(M, J, qw, qdw, tw, gqn + w * n, gqdn + w * n, gq + w * n, gqd + w * n, gtau + w * n);
    }
  });
}

void ddo_inv_mass_matrix(void* h, int N, const double* q, double* Minv) {
  const Model& M = ((OracleWorld*)h)->M;
  int n = M.nd;
  std::vector<double> zero(n, 0.0);
  for (int w = 0; w < N; w++) {
    Dyn<double> d(M);
    d.setState(q + w * n, zero.data());
    d.invMassMatrix(Minv + (size_t)w * n * n);
  }
}
void ddo_coriolis_gravity(void* h, int N, const double* q, const double* qd, double* C) {
  const Model& M = ((OracleWorld*)h)->M;
  int n = M.nd;
  for (int w = 0; w < N; w++) {
    Dyn<double> d(M);
    d.setState(q + w * n, qd + w * n);
    d.coriolisGravity(C + w * n);
  }
}
void ddo_forward_dynamics(void* h, int N, const double* q, const double* qd, const double* tau, double* qdd) {
  const Model& M = ((OracleWorld*)h)->M;
  int n = M.nd;
  for (int w = 0; w < N; w++) {
    Dyn<double> d(M);
    d.setState(q + w * n, qd + w * n);
    d.forwardDynamics(tau + w * n, qdd + w * n);
  }
}
// world transforms of every body: out[world][body][12] (R row-major, p)
void ddo_body_transforms(void* h, int N, const double* q, double* out) {
  const Model& M = ((OracleWorld*)h)->M;
  int n = M.nd;
  std::vector<double> zero(n, 0.0);
  for (int w = 0; w < N; w++) {
    Dyn<double> d(M);
    d.setState(q + w * n, zero.data());
    for (int i = 0; i < M.nb; i++) {
      double* o = out + ((size_t)w * M.nb + i) * 12;
      for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) o[3 * r + c] = d.b[i].W.R.m[r][c]; o[9 + r] = d.b[i].W.p[r]; }
    }
  }
}
// spatial inertia tensors: out[body][36] row-major
void ddo_spatial_inertias(void* h, double* out) {
  const Model& M = ((OracleWorld*)h)->M;
  for (int i = 0; i < M.nb; i++) for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) out[i * 36 + r * 6 + c] = M.b[i].I.m[r][c];
}
// math primitives exposed for the golden-vector tests
void ddo_transform_inertia(const double* T12, const double* I36, double* out36) {
  Iso<double> T;
  Mat6<double> I;
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T.R.m[r][c] = T12[3 * r + c]; T.p[r] = T12[9 + r]; }
  for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) I.m[r][c] = I36[6 * r + c];
  Mat6<double> o = transformInertia(T, I);
  for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) out36[6 * r + c] = o.m[r][c];
}
void ddo_exp_angular(const double* s, double* R9) {
  Mat3<double> R = expAngular(v3<double>(s[0], s[1], s[2]));
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R9[3 * r + c] = R.m[r][c];
}
void ddo_exp_map_rot(const double* s, double* R9) {
  Mat3<double> R = expMapRot(v3<double>(s[0], s[1], s[2]));
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R9[3 * r + c] = R.m[r][c];
}
void ddo_log_map_rot(const double* R9, double* w) {
  Mat3<double> R;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R.m[r][c] = R9[3 * r + c];
  Vec3<double> o = logMapRot(R);
  for (int r = 0; r < 3; r++) w[r] = o[r];
}

// minimum-norm least squares (the oracle's stand-in for Eigen's completeOrthogonalDecomposition().solve)
int ddo_cod_solve(int m, int n, const double* A, const double* b, double* x) {
  std::vector<double> Av(A, A + (size_t)m * n), bv(b, b + m);
  int rank = 0;
  std::vector<double> xv = codSolve(m, n, Av, bv, &rank);
  for (int i = 0; i < n; i++) x[i] = xv[i];
  return rank;
}
// PGS on one problem: A n x n row-major (copied into the padded layout), x warm start in/out
int ddo_pgs_solve(int n, const double* A, double* x, const double* b, const double* lo, const double* hi, const int* findex) {
  int ns = dPad(n);
  std::vector<double> Ap((size_t)n * ns, 0.0), bb(b, b + n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ap[(size_t)i * ns + j] = A[i * n + j];
  return pgs(n, Ap.data(), x, bb.data(), lo, hi, findex) ? 1 : 0;
}
int ddo_lcp_valid(int n, const double* A, const double* x, const double* b, const double* hi, const double* lo, const int* findex, int ignoreFriction) {
  return lcpValid(n, A, x, b, hi, lo, findex, ignoreFriction != 0) ? 1 : 0;
}

// ---- contact path ----
// load the reference's own solvers (oracle/_ref/*.so); returns a bit mask of what was found
int ddo_load_ref_libs(const char* dantzig_so, const char* boxbox_so) {
  int mask = 0;
  if (dantzig_so) { void* h = dlopen(dantzig_so, RTLD_NOW | RTLD_LOCAL); if (h) { refLibs().dantzig = (RefDantzigFn)dlsym(h, "ddref_dantzig_solve"); if (refLibs().dantzig) mask |= 1; } }
  if (boxbox_so) { void* h = dlopen(boxbox_so, RTLD_NOW | RTLD_LOCAL); if (h) { refLibs().boxbox = (RefBoxBoxFn)dlsym(h, "ddref_collide_box_box"); if (refLibs().boxbox) mask |= 2; } }
  return mask;
}
// One world, World::step with contacts (gradients disabled). kmax: capacity of the per-row outputs.
// contacts_out [cmax][10]: bodyA bodyB point(3) normal(3) depth type; returns through counts[0..3]: n contacts, n rows, n groups, tier mask
void ddo_step_contacts(void* h, const double* q, const double* qd, const double* tau, double* lcp_cache, int* n_cache, int kmax, double* qn,
                       double* qdn, double* pre_vel, int cmax, double* contacts_out, double* A_out, double* b_out, double* lo_out, double* hi_out,
                       int* findex_out, double* x_out, int* counts, int gradients, int* mapping_out) {
  const Model& M = ((OracleWorld*)h)->M;
  std::vector<double> cache(lcp_cache, lcp_cache + *n_cache);
  ContactStepResult R;
  stepWithContacts(M, q, qd, tau, cache, R, gradients != 0);
  for (int i = 0; i < M.nd; i++) { qn[i] = R.qNext[i]; qdn[i] = R.qdNext[i]; if (pre_vel) pre_vel[i] = R.preConstraintVel[i]; }
  *n_cache = (int)std::min<size_t>(cache.size(), kmax);
  for (int i = 0; i < *n_cache; i++) lcp_cache[i] = cache[i];
  counts[0] = (int)R.contacts.size();
  for (int k = 0; k < (int)R.contacts.size() && k < cmax && contacts_out; k++) {
    double* o = contacts_out + 10 * k;
    const Contact& c = R.contacts[k];
    o[0] = c.bodyA; o[1] = c.bodyB;
    for (int x = 0; x < 3; x++) { o[2 + x] = c.point[x]; o[5 + x] = c.normal[x]; }
    o[8] = c.depth; o[9] = c.type;
  }
  int rows = 0, tiers = 0;
  // groups are concatenated in order; A_out is block diagonal in a kmax x kmax row-major matrix
  for (const GroupLCP& g : R.groups) {
    for (int i = 0; i < g.n && rows + i < kmax; i++) {
      if (A_out) for (int j = 0; j < g.n && rows + j < kmax; j++) A_out[(size_t)(rows + i) * kmax + rows + j] = g.A[(size_t)i * g.n + j];
      if (b_out) b_out[rows + i] = g.b[i];
      if (lo_out) lo_out[rows + i] = g.lo[i];
      if (hi_out) hi_out[rows + i] = g.hi[i];
      if (findex_out) findex_out[rows + i] = g.findex[i] >= 0 ? g.findex[i] + rows : -1;
      if (x_out) x_out[rows + i] = g.x[i];
      if (mapping_out && !g.mapping.empty()) mapping_out[rows + i] = g.mapping[i] >= 0 ? g.mapping[i] + rows : g.mapping[i];
    }
    rows += g.n;
    tiers |= 1 << g.tier;
  }
  if (R.unsupportedPair) tiers |= 1 << 9;
  counts[1] = rows; counts[2] = (int)R.groups.size(); counts[3] = tiers;
}

// One world: forwardPass with contacts (gradient-enabled solve) + the BackpropSnapshot Jacobians through the contact
// solution (reference recipe). Matrices n x n column-major; info[0..2] = numClamping, numUpperBound, pseudo-inverse branch.
// x_override (optional, all rows of the world in group order): multipliers to evaluate the Jacobians at.
// mode 0: analytic (reference recipe); mode 1: central differences (eps) of the frozen-active-set velocity map with the
// collision detector re-run at every perturbed state (posVel / velVel / forceVel only; info[3] = 1 when every
// perturbed state kept the contact structure).
void ddo_contact_jacobians(void* h, int mode, double eps, const double* q, const double* qd, const double* tau, const double* lcp_cache,
                           int n_cache, double* posPos, double* posVel, double* velPos, double* velVel, double* forceVel, int* info,
                           double* qn, double* qdn, const double* x_override, int n_override, double* fc_vs_ls, const int* mapping_override, double q_noise) {
  qNoise() = q_noise;
  const Model& M = ((OracleWorld*)h)->M;
  const int n = M.nd;
  std::vector<double> cache(lcp_cache, lcp_cache + n_cache);
  ContactStepResult R;
  stepWithContacts(M, q, qd, tau, cache, R, true);
  if (qn) for (int i = 0; i < n; i++) { qn[i] = R.qNext[i]; qdn[i] = R.qdNext[i]; }
  if (x_override) {
    // The LCP multipliers of a rank-deficient group are determined only up to the null space of A (the next state does not
    // depend on it, the reference's Jacobian formulas do): evaluate the Jacobians at the caller's multipliers.
    int row = 0;
    for (GroupLCP& g : R.groups) {
      const int base = row;
      for (int i = 0; i < g.n && row < n_override; i++, row++) {
        g.x[i] = x_override[row];
        if (mapping_override) g.mapping[i] = mapping_override[row] >= 0 ? mapping_override[row] - base : mapping_override[row];  // world rows -> group rows
      }
    }
  }
  Jacobians J;
  ContactGradInfo gi;
  if (mode == 0) {
    jacobiansReferenceContact(M, q, qd, tau, R, J, &gi);
    info[0] = gi.numClamping; info[1] = gi.numUpperBound; info[2] = gi.pseudoInverseBranch ? 1 : 0; info[3] = 1;
    if (fc_vs_ls) *fc_vs_ls = gi.fcVsLeastSquares;
  } else {
    J.resize(n);
    bool ok = true;
    std::vector<double> x(3 * n), xp, a, b;
    for (int i = 0; i < n; i++) { x[i] = q[i]; x[n + i] = qd[i]; x[2 * n + i] = tau[i]; }
    for (int i = 0; i < 3 * n; i++) {
      xp = x; xp[i] += eps; ok = frozenSetVelocity(M, &xp[0], &xp[n], &xp[2 * n], R, a) && ok;
      xp = x; xp[i] -= eps; ok = frozenSetVelocity(M, &xp[0], &xp[n], &xp[2 * n], R, b) && ok;
      if (a.size() != (size_t)n || b.size() != (size_t)n) { ok = false; continue; }
      std::vector<double>& out = i < n ? J.posVel : (i < 2 * n ? J.velVel : J.forceVel);
      for (int r = 0; r < n; r++) out[(size_t)(i % n) * n + r] = (a[r] - b[r]) / (2 * eps);
    }
    info[0] = info[1] = info[2] = 0; info[3] = ok ? 1 : 0;
  }
  const size_t nn = (size_t)n * n * sizeof(double);
  if (posPos) std::memcpy(posPos, J.posPos.data(), nn);
  if (posVel) std::memcpy(posVel, J.posVel.data(), nn);
  if (velPos) std::memcpy(velPos, J.velPos.data(), nn);
  if (velVel) std::memcpy(velVel, J.velVel.data(), nn);
  if (forceVel) std::memcpy(forceVel, J.forceVel.data(), nn);
}

// contact geometry gradients of ONE world at (q, qd, tau): out [constraint][dof][6] = d point / d q_dof, d normal / d q_dof
// (getContactPositionGradient / getContactNormalGradient); returns the number of constraints (<= cmax written).
int ddo_contact_geometry_gradients(void* h, const double* q, const double* qd, const double* tau, int cmax, double* out, double* points) {
  const Model& M = ((OracleWorld*)h)->M;
  const int n = M.nd;
  std::vector<double> cache;
  ContactStepResult R;
  stepWithContacts(M, q, qd, tau, cache, R, true);
  Dyn<double> d(M);
  d.setState(q, qd);
  const Screws sc = worldScrews(M, d);
  int k = 0;
  for (const Constraint& c : R.constraints) {
    if (k >= cmax) break;
    for (int w = 0; w < n; w++) {
      const DofRole role = dofRole(M, c, w);
      const Vec3<double> pg = contactPositionGradient(c, role, sc.pos[w]), ng = contactNormalGradient(c, role, sc.pos[w]);
      for (int x = 0; x < 3; x++) { out[((size_t)k * n + w) * 6 + x] = pg[x]; out[((size_t)k * n + w) * 6 + 3 + x] = ng[x]; }
    }
    for (int x = 0; x < 3; x++) { points[k * 8 + x] = c.contact.point[x]; points[k * 8 + 3 + x] = c.contact.normal[x]; }
    points[k * 8 + 6] = c.contact.type; points[k * 8 + 7] = c.contact.depth;
    k++;
  }
  return (int)R.constraints.size();
}

}  // extern "C"
