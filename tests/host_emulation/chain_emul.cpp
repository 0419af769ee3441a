// TEST INFRASTRUCTURE ONLY — see cuda_shim.h. Host build of the chain kernels (canonical joint frames, chain.cuh).
#include "cuda_shim.h"
#include <string>
#include <vector>
struct dim3_ { int x; };
static dim3_ blockIdx, threadIdx, blockDim;
inline void sincos(double x, double* s, double* c);
#include "kernels_nocuda.cuh"
#include "contact_nocuda.cuh"
#include "fold.h"
using namespace ddb;

// The chain kernels (canonical joint frames, chain.cuh) for one world: forward step, backward VJP and M^-1 (nb = 1..8)
template <int NB, int KIND>
static void chainRun(const double* q, const double* qd, const double* tau, const double* gqn, const double* gqdn, double* qn, double* qdn,
                     double* gq, double* gqd, double* gtau, double* minv, double* jac5) {
  const int N = 1, nd = c_topo.nd;
  std::vector<double> snap(snapDoublesPerWorld(c_topo.snapPairs, nd, c_topo.has_limits != 0) * snapWorldsPadded(N) + 8, 0.0);
  blockIdx.x = 0; threadIdx.x = 0; blockDim.x = 64;
  k_forward<NB, KIND, 64, 1>(N, 0, N, q, qd, tau, qn, qdn, snap.data());
  Limits lim{};
  k_backward<NB, KIND, 32, 1>(N, 0, N, snap.data(), gqn, gqdn, nullptr, nullptr, gq, gqd, gtau, lim);
  k_inv_mass<NB, KIND, 64, 1>(N, snap.data(), minv);
  const size_t M = (size_t)nd * nd;
  k_jacobians<NB, KIND, 64, 1>(N, snap.data(), jac5, jac5 + M, jac5 + 2 * M, jac5 + 3 * M, jac5 + 4 * M);
}
extern "C" int emul_chain_step(const ddb_topology* topo, const ddb_params* params, const double* q, const double* qd, const double* tau,
                               const double* gqn, const double* gqdn, double* qn, double* qdn, double* gq, double* gqd, double* gtau, double* minv,
                               double* jac5, int kind) {
  std::string err;
  bool chain;
  if (foldTopology(topo, params, c_topo, chain, err) != DDB_OK) return 1;
  if (!chain || c_topo.has_limits) return 2;
  for (int i = 0; i < c_topo.nb; i++) if (kind == 2 && c_topo.b[i].type != DDB_JOINT_REVOLUTE) return 4;  // kind 2: revolute joints only
  canonicalizeChain(topo, c_topo);
#define RUN(NBV) (kind == 2 ? chainRun<NBV, 2>(q, qd, tau, gqn, gqdn, qn, qdn, gq, gqd, gtau, minv, jac5) : chainRun<NBV, 1>(q, qd, tau, gqn, gqdn, qn, qdn, gq, gqd, gtau, minv, jac5))
  switch (c_topo.nb) {
    case 2: RUN(2); break;
    case 3: RUN(3); break;
    case 5: RUN(5); break;
    default: return 3;
  }
#undef RUN
  return 0;
}
