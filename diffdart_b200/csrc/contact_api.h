// contact_api.h — host-side interface between the two translation units: ddb_api.cu (dynamics kernels, FMA allowed)
// and lcp_api.cu (LCP / narrow phase / contact stage, compiled with -fmad=false to keep the reference's operation order).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
namespace ddb {
struct DevTopo;
struct ContactTopo;
// upload the topology of the launching handle into the contact TU's constant memory (stream-ordered)
cudaError_t contactBindTopology(const DevTopo* topo, const void* ctopo_host, cudaStream_t s);
size_t contactTopoBytes();
// scratch sizes per handle
void contactScratchSizes(int N, int nb, int nd, int cmax, int kmax, size_t* work_doubles, size_t* work_ints, size_t* lcp_doubles, size_t* lcp_ints);
// the contact stage after k_forward: qd (in: unconstrained update, out: post-contact) ; lcp_cache [N][kmax], n_cache [N],
// counts_out [N][4] = n_contacts, n_rows, n_groups, tier mask
cudaError_t contactLaunch(cudaStream_t s, int N, int nb, int nd, int cmax, int kmax, int layout, double* snap, double* qd_inout, double* lcp_cache,
                          int* n_cache, int* counts_out, int* active_out, double* work_d, int* work_i, double* lcp_d, int* lcp_i,
                          double* keep_d, int* keep_i, int* pgs_ctrl, double* pgs_ans_x, int* pgs_ans_ok, int* launches);
// control words and answer buffers of the deferred PGS solves (ints, doubles, ints for N worlds)
void contactPgsSizes(int N, int kmax, size_t* ctrl_ints, size_t* ans_doubles, size_t* ans_ints);
// per-slot storage of what the backward pass needs of the contact stage (doubles / ints for N worlds)
// DDB_CONTACT_PROFILE=1: read (and reset) the clock64 ticks summed over worlds per section of the contact stage
int contactProfileRead(long long* out16);
void contactKeepSizes(int N, int nd, int cmax, int kmax, size_t* keep_doubles, size_t* keep_ints);
// BackpropSnapshot::backprop through the contact solution retained in (snap, keep_d, keep_i); lim6 = 6 x [nd] limits or NULL
cudaError_t contactBackwardLaunch(cudaStream_t s, int N, int nb, int nd, int cmax, int kmax, int layout, double* snap, const double* keep_d,
                                  const int* keep_i, const double* gqn, const double* gqdn, double* gq, double* gqd, double* gtau, double* work_d,
                                  int* work_i, double* lcp_d, int* lcp_i, const double* lim6, int* status);
cudaError_t contactUnitVector(cudaStream_t s, int N, int n, int layout, int j, double* out);
cudaError_t contactScatterRow(cudaStream_t s, int N, int n, int layout, int j, const double* src, double* dst);
}  // namespace ddb
