// TEST INFRASTRUCTURE ONLY — see cuda_shim.h. Host build of the forward kernel + contact stage (generic tree
// instantiation) so that the whole contact step can be checked against the oracle on the CPU.
#include "cuda_shim.h"
#include <string>
#include <vector>
struct dim3_ { int x; };
static dim3_ blockIdx, threadIdx, blockDim;
inline void sincos(double x, double* s, double* c);
#include "kernels_nocuda.cuh"
#include "contact_nocuda.cuh"
#include "contact_bwd_nocuda.cuh"
#include "fold.h"
using namespace ddb;

extern "C" int emul_contact_step(const ddb_topology* topo, const ddb_params* params, const double* q, const double* qd, const double* tau,
                                 double* lcp_cache, int* n_cache, double* qn, double* qdn, int* counts, double* contacts_out, int* bodies_out,
                                 int* dims_out, double* A_out, double* b_out, double* x_out, int* caps, int* active_out) {
  std::string err;
  bool chain;
  if (foldTopology(topo, params, c_topo, chain, err) != DDB_OK) return 1;
  if (foldContactTopology(topo, params, c_ctopo, err) != DDB_OK) return 2;
  const int N = 1, nb = c_topo.nb, nd = c_topo.nd, cmax = c_ctopo.cmax, kmax = c_ctopo.kmax;
  caps[0] = cmax; caps[1] = kmax;
  std::vector<double> snap(snapDoublesPerWorld(c_topo.snapPairs, nd, c_topo.has_limits != 0) * snapWorldsPadded(N) + 8, 0.0);
  blockIdx.x = 0; threadIdx.x = 0; blockDim.x = 64;
  k_forward<DDB_MAX_BODIES, false, 64, 1>(N, 0, N, q, qd, tau, qn, qdn, snap.data());
  if (c_ctopo.n_pairs == 0) { counts[0] = counts[1] = counts[2] = counts[3] = 0; return 0; }
  ContactLayout L{nb, nd, cmax, kmax};
  std::vector<double> wd(L.doubles(), 0.0), ld(lcpScratchDoubles(kmax) * 32, 0.0);
  std::vector<int> wi(L.ints(), 0), li(lcpScratchInts(kmax) * 32, 0);
  Snap sn{reinterpret_cast<double2*>(snap.data()), N, 0};
  ContactWork cw{wd.data(), wi.data(), L};
  PgsOptions pgs{10000, 1e-10, 1e-8, 1e-8};
  ContactStepOut o = contactStage(sn, cw, ld.data() + 3, li.data() + 3, qdn, qdn, c_topo.layout, N, 0, lcp_cache, n_cache, pgs);
  counts[0] = o.n_contacts; counts[1] = o.n_rows; counts[2] = o.n_groups; counts[3] = o.tier_mask;
  for (int c = 0; c < cmax; c++) {
    for (int k = 0; k < 26; k++) contacts_out[c * 26 + k] = cw.contact(c)[k];
    bodies_out[2 * c] = cw.bodyA()[c]; bodies_out[2 * c + 1] = cw.bodyB()[c]; dims_out[c] = cw.dim()[c];
  }
  for (int k = 0; k < kmax * kmax; k++) A_out[k] = cw.A()[k];
  for (int k = 0; k < kmax; k++) { b_out[k] = cw.b()[k]; x_out[k] = cw.xall()[k]; if (active_out) active_out[k] = cw.activeAll()[k]; }
  return 0;
}

// forward + backward + Jacobians of one world through the generic tree kernels (no contact stage)
extern "C" int emul_step_backward(const ddb_topology* topo, const ddb_params* params, const double* q, const double* qd, const double* tau,
                                  const double* gqn, const double* gqdn, double* qn, double* qdn, double* gq, double* gqd, double* gtau,
                                  double* posPos, double* posVel, double* velPos, double* velVel, double* forceVel, double* minv) {
  std::string err;
  bool chain;
  if (foldTopology(topo, params, c_topo, chain, err) != DDB_OK) return 1;
  const int N = 1, nd = c_topo.nd;
  std::vector<double> snap(snapDoublesPerWorld(c_topo.snapPairs, nd, c_topo.has_limits != 0) * snapWorldsPadded(N) + 8, 0.0);
  blockIdx.x = 0; threadIdx.x = 0; blockDim.x = 64;
  k_forward<DDB_MAX_BODIES, false, 64, 1>(N, 0, N, q, qd, tau, qn, qdn, snap.data());
  Limits lim{};
  k_backward<DDB_MAX_BODIES, false, 64, 1>(N, 0, N, snap.data(), gqn, gqdn, nullptr, nullptr, gq, gqd, gtau, lim);
  k_jacobians<DDB_MAX_BODIES, false, 64, 1>(N, snap.data(), posPos, posVel, velPos, velVel, forceVel);
  k_inv_mass<DDB_MAX_BODIES, false, 64, 1>(N, snap.data(), minv);
  return 0;
}

// forwardPass with contacts (gradient-enabled solve) followed by the backward through the contact solution, one world
extern "C" int emul_contact_step_backward(const ddb_topology* topo, const ddb_params* params, const double* q, const double* qd, const double* tau,
                                          double* lcp_cache, int* n_cache, const double* gqn, const double* gqdn, double* qn, double* qdn,
                                          double* gq, double* gqd, double* gtau, int* status_out, double* x_out, int* rows_out) {
  std::string err;
  bool chain;
  if (foldTopology(topo, params, c_topo, chain, err) != DDB_OK) return 1;
  if (foldContactTopology(topo, params, c_ctopo, err) != DDB_OK) return 2;
  const int N = 1, nb = c_topo.nb, nd = c_topo.nd, cmax = c_ctopo.cmax, kmax = c_ctopo.kmax;
  std::vector<double> snap(snapDoublesPerWorld(c_topo.snapPairs, nd, c_topo.has_limits != 0) * snapWorldsPadded(N) + 8, 0.0);
  blockIdx.x = 0; threadIdx.x = 0; blockDim.x = 64;
  k_forward<DDB_MAX_BODIES, false, 64, 1>(N, 0, N, q, qd, tau, qn, qdn, snap.data());
  ContactLayout L{nb, nd, cmax, kmax};
  ContactSnapLayout KL{nd, cmax, kmax};
  ContactBwdLayout BL{nb, nd, kmax};
  std::vector<double> wd(std::max(L.doubles(), BL.doubles()), 0.0), ld(lcpScratchDoubles(kmax) * 32, 0.0), kd(KL.doubles(), 0.0);
  std::vector<int> wi(std::max(L.ints(), BL.ints()), 0), li(lcpScratchInts(kmax) * 32, 0), ki(KL.ints(), 0);
  Snap sn{reinterpret_cast<double2*>(snap.data()), N, 0};
  ContactWork cw{wd.data(), wi.data(), L};
  PgsOptions pgs{10000, 1e-10, 1e-8, 1e-8};
  if (c_ctopo.n_pairs > 0) contactStage(sn, cw, ld.data() + 3, li.data() + 3, qdn, qdn, c_topo.layout, N, 0, lcp_cache, n_cache, pgs, kd.data(), ki.data());
  ContactBwdWork bw{wd.data(), wi.data(), BL};
  const int st = contactBackwardWorld<DDB_MAX_BODIES>(sn, KL, kd.data(), ki.data(), bw, ld.data() + 3, li.data() + 3, gqn, gqdn, gq, gqd, gtau,
                                                      c_topo.layout, N, 0);
  if (status_out) *status_out = st;
  const int rows = ki[KL.oCounts() + 1];
  if (rows_out) *rows_out = rows;
  if (x_out) for (int r = 0; r < rows; r++) x_out[r] = kd[KL.oX() + r];
  return 0;
}

// d/dq [mu_a^T M(q) a_b] through the MODE-2 reverse sweep (the bilinear mass-matrix terms of the contact backward)
extern "C" int emul_mass_bilinear_grad(const ddb_topology* topo, const ddb_params* params, const double* q, const double* mu_a, const double* a_b,
                                       double* dq_out) {
  std::string err;
  bool chain;
  if (foldTopology(topo, params, c_topo, chain, err) != DDB_OK) return 1;
  const int N = 1, nb = c_topo.nb, nd = c_topo.nd;
  std::vector<double> snap(snapDoublesPerWorld(c_topo.snapPairs, nd, c_topo.has_limits != 0) * snapWorldsPadded(N) + 8, 0.0), zero(nd, 0.0), qn(nd), qdn(nd);
  blockIdx.x = 0; threadIdx.x = 0; blockDim.x = 64;
  k_forward<DDB_MAX_BODIES, false, 64, 1>(N, 0, N, q, zero.data(), zero.data(), qn.data(), qdn.data(), snap.data());
  Snap sn{reinterpret_cast<double2*>(snap.data()), N, 0};
  ContactBwdLocals<DDB_MAX_BODIES> loc;
  twistsFromRates<DDB_MAX_BODIES>(sn, nb, [&](int d) { return mu_a[d]; }, loc.ws.U, loc.ws.mu);
  twistsFromRates<DDB_MAX_BODIES>(sn, nb, [&](int d) { return a_b[d]; }, loc.dA, nullptr);
  idAdjointEx<DDB_MAX_BODIES, false, 2>(sn, nb, loc.ws, loc.dA, [&](int d, double dq, double) { dq_out[d] = dq; },
                                        [&](int body, const double* dq6, const double*) { for (int k = 0; k < 6; k++) dq_out[c_topo.b[body].dof + k] = dq6[k]; });
  return 0;
}

// P rhs / P^T rhs with P = Q^+ from one complete orthogonal decomposition (codFactorDev / codApplyDev); returns the rank
extern "C" int emul_cod_apply(int n, const double* Q, const double* rhs, int transposed, double* out) {
  std::vector<double> ld(lcpScratchDoubles(n) * 32, 0.0);
  std::vector<int> li(lcpScratchInts(n) * 32, 0);
  LcpView v;
  v.init(ld.data() + 5, li.data() + 5, n, n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) v.db[(size_t)(v.oL + i * n + j) * 32] = Q[i * n + j];
  const int rank = codFactorDev(v, n);
  codApplyDev(v, n, rank, transposed != 0, [&](int i) { return rhs[i]; }, [&](int i, double x) { out[i] = x; });
  return rank;
}

// The deferred-PGS control flow of the GPU (k_contact passes + PGS kernel) on the host: the stage is re-entered behind every
// answered solve; the solves themselves use the scalar pgsSolve, so the result must equal emul_contact_step bit for bit.
extern "C" int emul_contact_step_deferred(const ddb_topology* topo, const ddb_params* params, const double* q, const double* qd, const double* tau,
                                          double* lcp_cache, int* n_cache, double* qn, double* qdn, int* counts, double* x_out, int* active_out,
                                          int* passes_out) {
  std::string err;
  bool chain;
  if (foldTopology(topo, params, c_topo, chain, err) != DDB_OK) return 1;
  if (foldContactTopology(topo, params, c_ctopo, err) != DDB_OK) return 2;
  const int N = 1, nb = c_topo.nb, nd = c_topo.nd, cmax = c_ctopo.cmax, kmax = c_ctopo.kmax;
  std::vector<double> snap(snapDoublesPerWorld(c_topo.snapPairs, nd, c_topo.has_limits != 0) * snapWorldsPadded(N) + 8, 0.0);
  blockIdx.x = 0; threadIdx.x = 0; blockDim.x = 64;
  k_forward<DDB_MAX_BODIES, false, 64, 1>(N, 0, N, q, qd, tau, qn, qdn, snap.data());
  if (c_ctopo.n_pairs == 0) { counts[0] = counts[1] = counts[2] = counts[3] = 0; *passes_out = 0; return 0; }
  ContactLayout L{nb, nd, cmax, kmax};
  std::vector<double> wd(L.doubles(), 0.0), ld(lcpScratchDoubles(kmax) * 32, 0.0);
  std::vector<int> wi(L.ints(), 0), li(lcpScratchInts(kmax) * 32, 0);
  Snap sn{reinterpret_cast<double2*>(snap.data()), N, 0};
  ContactWork cw{wd.data(), wi.data(), L};
  PgsOptions pgs{10000, 1e-10, 1e-8, 1e-8};
  const int cap = 4;
  std::vector<double> ansX((size_t)cap * kmax, 0.0);
  std::vector<int> ansOk(cap, 0);
  int answered = 0, resumeKind = 0, passes = 0;
  for (;;) {
    PgsDefer df{answered, 0, kmax, 0, cap, resumeKind, ansX.data(), ansOk.data()};
    ContactStepOut o = contactStage(sn, cw, ld.data() + 3, li.data() + 3, qdn, qdn, c_topo.layout, N, 0, lcp_cache, n_cache, pgs, nullptr, nullptr, &df);
    passes++;
    if (!o.pending) {
      counts[0] = o.n_contacts; counts[1] = o.n_rows; counts[2] = o.n_groups; counts[3] = o.tier_mask;
      break;
    }
    resumeKind = df.resumeKind;
    LcpView v;
    v.init(ld.data() + 3, li.data() + 3, df.pendingN, kmax);
    const bool ok = pgsSolve(v, pgs);
    for (int i = 0; i < df.pendingN; i++) ansX[(size_t)answered * kmax + i] = v.x(i);
    ansOk[answered++] = ok ? 1 : 0;
  }
  *passes_out = passes;
  for (int k = 0; k < kmax; k++) { x_out[k] = cw.xall()[k]; if (active_out) active_out[k] = cw.activeAll()[k]; }
  return 0;
}
