// ddb_api.cu — the C ABI declared in include/ddb.h: handle management, topology folding,
// kernel dispatch. No torch types, no CPU fallback.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "contact.cuh"
#include "contact_api.h"
#include "fold.h"

using namespace ddb;

namespace {

#ifndef DDB_BLOCK
#define DDB_BLOCK 64
#endif
#ifndef DDB_BLOCK_BWD
#define DDB_BLOCK_BWD 32  // measured: the forward wants two warps per CTA marching together (instruction cache), the backward CTAs of one warp
#endif
#ifndef DDB_MINB
#define DDB_MINB 1
#endif
#ifndef DDB_MINB_FWD
#define DDB_MINB_FWD DDB_MINB
#endif
#ifndef DDB_MINB_BWD
#define DDB_MINB_BWD DDB_MINB
#endif
constexpr int kBlock = DDB_BLOCK, kBlockB = DDB_BLOCK_BWD;
constexpr int kMinB = DDB_MINB, kMinBF = DDB_MINB_FWD, kMinBB = DDB_MINB_BWD;

struct KernelTable {
  // forward / backward over the worlds [w0, w1) of the batch
  void (*fwd)(cudaStream_t, int, int, int, const double*, const double*, const double*, double*, double*, double*);
  void (*bwd)(cudaStream_t, int, int, int, double*, const double*, const double*, const double*, const double*, double*, double*, double*, Limits);
  void (*jac)(cudaStream_t, int, double*, double*, double*, double*, double*, double*);
  void (*minv)(cudaStream_t, int, double*, double*);
  // chains only (else NULL): the same kernels with their vectors in pinned host memory (HOSTIO, kernels.cuh)
  void (*fwdHost)(cudaStream_t, int, int, int, const double*, const double*, const double*, double*, double*, double*);
  void (*bwdHost)(cudaStream_t, int, int, int, double*, const double*, const double*, double*, double*, double*, Limits);
};

template <int NB, int CHAIN>
struct Launch {
  static int grid(int N, int block = kBlock) { return (N + block - 1) / block; }
  static void fwd(cudaStream_t s, int N, int w0, int w1, const double* q, const double* qd, const double* tau, double* qn, double* qdn, double* snap) {
    k_forward<NB, CHAIN, kBlock, kMinBF><<<grid(w1 - w0), kBlock, 0, s>>>(N, w0, w1, q, qd, tau, qn, qdn, snap);
  }
  static void bwd(cudaStream_t s, int N, int w0, int w1, double* snap, const double* gqn, const double* gqdn, const double* gqnAdd,
                  const double* gqdnAdd, double* gq, double* gqd, double* gtau, Limits lim) {
    // chains stage their snapshot rows in shared memory (dynamics.cuh, StageGeom): 16 bytes x slots per world
    const size_t smem = (CHAIN && DDB_BWD_STAGE) ? ChainStageGeom<NB, kBlockB>::kBytes : 0;
    if (smem) {
      static bool once[64] = {false};  // per instantiation and device: the attribute belongs to the function in that context
      int dev = 0;
      cudaGetDevice(&dev);
      if (!once[dev & 63]) {
        cudaFuncSetAttribute(k_backward<NB, CHAIN, kBlockB, kMinBB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_backward<NB, CHAIN, kBlockB, kMinBB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        once[dev & 63] = true;
      }
    }
    k_backward<NB, CHAIN, kBlockB, kMinBB><<<grid(w1 - w0, kBlockB), kBlockB, smem, s>>>(N, w0, w1, snap, gqn, gqdn, gqnAdd, gqdnAdd, gq, gqd, gtau, lim);
  }
  static void jac(cudaStream_t s, int N, double* snap, double* pp, double* pv, double* vp, double* vv, double* fv) {
    // chains small enough keep their snapshot in shared memory for the n pull-backs (kernels.cuh, kJacResident)
    const size_t smem = kJacResident<NB, CHAIN, kBlock> ? ChainResident<NB, kBlock>::kBytes : 0;
    if (smem) {
      static bool once[64] = {false};  // per instantiation and device
      int dev = 0;
      cudaGetDevice(&dev);
      if (!once[dev & 63]) {
        cudaFuncSetAttribute(k_jacobians<NB, CHAIN, kBlock, kMinB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_jacobians<NB, CHAIN, kBlock, kMinB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        once[dev & 63] = true;
      }
    }
    k_jacobians<NB, CHAIN, kBlock, kMinB><<<grid(N), kBlock, smem, s>>>(N, snap, pp, pv, vp, vv, fv);
  }
  static void minv(cudaStream_t s, int N, double* snap, double* out) {
    k_inv_mass<NB, CHAIN, kBlock, kMinB><<<grid(N), kBlock, 0, s>>>(N, snap, out);
  }
  // per instantiation and device: the attribute belongs to the function in that context
  template <class K>
  static void allowShared(K kernel, size_t smem, bool* once) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && !once[dev & 63]) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      once[dev & 63] = true;
    }
  }
  static void fwdHost(cudaStream_t s, int N, int w0, int w1, const double* q, const double* qd, const double* tau, double* qn, double* qdn, double* snap) {
    const size_t smem = (size_t)3 * kBlock * NB * sizeof(double);
    k_forward<NB, CHAIN, kBlock, kMinBF, true><<<grid(w1 - w0), kBlock, smem, s>>>(N, w0, w1, q, qd, tau, qn, qdn, snap);
  }
  static void bwdHost(cudaStream_t s, int N, int w0, int w1, double* snap, const double* gqn, const double* gqdn, double* gq, double* gqd, double* gtau,
                      Limits lim) {
    const size_t smem = (DDB_BWD_STAGE ? ChainStageGeom<NB, kBlockB>::kBytes : 0) + (size_t)3 * kBlockB * NB * sizeof(double);
    static bool once[64] = {false};
    allowShared(k_backward<NB, CHAIN, kBlockB, kMinBB, true>, smem, once);
    k_backward<NB, CHAIN, kBlockB, kMinBB, true><<<grid(w1 - w0, kBlockB), kBlockB, smem, s>>>(N, w0, w1, snap, gqn, gqdn, nullptr, nullptr, gq, gqd, gtau, lim);
  }
  static KernelTable table() {
    KernelTable t{fwd, bwd, jac, minv, nullptr, nullptr};
    if (CHAIN) { t.fwdHost = fwdHost; t.bwdHost = bwdHost; }
    return t;
  }
};

// chain: 0 = tree, 1 = serial chain of 1-DOF joints, 2 = such a chain with revolute joints only (kernels.cuh)
KernelTable pickTable(int nb, int chain) {
  if (chain == 1) {
    switch (nb) {
      case 1: return Launch<1, 1>::table();
      case 2: return Launch<2, 1>::table();
      case 3: return Launch<3, 1>::table();
      case 4: return Launch<4, 1>::table();
      case 5: return Launch<5, 1>::table();
      case 6: return Launch<6, 1>::table();
      case 7: return Launch<7, 1>::table();
      case 8: return Launch<8, 1>::table();
      default: break;
    }
  }
  if (chain == 2) {
    switch (nb) {
      case 1: return Launch<1, 2>::table();
      case 2: return Launch<2, 2>::table();
      case 3: return Launch<3, 2>::table();
      case 4: return Launch<4, 2>::table();
      case 5: return Launch<5, 2>::table();
      case 6: return Launch<6, 2>::table();
      case 7: return Launch<7, 2>::table();
      case 8: return Launch<8, 2>::table();
      default: break;
    }
  }
  if (nb <= 8) return Launch<8, 0>::table();
  if (nb <= 16) return Launch<16, 0>::table();
  if (nb <= 32) return Launch<32, 0>::table();
  return Launch<DDB_MAX_BODIES, 0>::table();
}

// The __constant__ topology slot of a device has one owner (handle) at a time. Every entry point that launches kernels holds
// the device's mutex from the bind to its last launch (ApiScope), so host threads driving different handles on one GPU cannot
// swap the topology between another thread's bind and its kernels. Recursive: the rollout entry points call ddb_forward /
// ddb_backward.
std::recursive_mutex g_devMu[64];
const void* g_owner[64] = {nullptr};

}  // namespace

struct ddb_handle {
  int device = 0;
  int N = 0, n = 0, nb = 0;
  bool chain = false;
  DevTopo topo;
  KernelTable k;
  std::vector<double*> slots;  // device snapshot buffers
  int cur = 0;
  bool forwarded = false;
  std::vector<char> slotValid;
  double* lim_dev = nullptr;  // 6*n
  Limits lim{};
  std::string err;
  int64_t launches = 0;
  // contact stage
  bool hasContacts = false;
  ContactTopo ctopo;
  int cmax = 0, kmax = 0;
  double *cwork_d = nullptr, *clcp_d = nullptr, *lcp_cache = nullptr;
  int *cwork_i = nullptr, *clcp_i = nullptr, *n_cache = nullptr, *ccounts = nullptr;
  std::vector<double*> keepD;  // per snapshot slot: what the backward needs of the contact stage
  std::vector<int*> keepI;
  double* jacTmp = nullptr;    // 5 x [N][n] staging of ddb_jacobians with contacts
  int* bwdStatus = nullptr;
  int *pgsCtrl = nullptr, *pgsOk = nullptr;  // deferred PGS solves (contact_api.h)
  double* pgsX = nullptr;
  // lanes: internal streams on which chunks of worlds run side by side (rollouts, host-buffer path)
  std::vector<cudaStream_t> lanes;
  std::vector<cudaEvent_t> laneDone;
  cudaEvent_t forkEv = nullptr;
  int nLanes = 2;
  double* gradRun = nullptr;  // 4 x [N*n]: running loss gradient of a rollout's reverse sweep (ping-pong)
  // host-buffer path
  cudaStream_t hstream = nullptr;
  double* stage = nullptr;  // device staging: 3n in + 2n out (+ 2n + 3n for backward), see *_host
  size_t stageDoubles = 0;
  // CUDA graphs of the contact-free rollouts (RolloutGraphs below)
  struct RolloutGraph {
    int kind = 0, T = 0, K = 0;
    const void* args[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaGraphExec_t exec = nullptr;
    bool same(int kind_, int T_, int K_, const void* const* a) const {
      if (kind != kind_ || T != T_ || K != K_) return false;
      for (int i = 0; i < 6; i++) if (args[i] != a[i]) return false;
      return true;
    }
  };
  std::vector<RolloutGraph> graphs;  // instantiated graphs, most recently used last
  RolloutGraph lastCall;             // the arguments of the previous rollout call (a graph is built when a call repeats)
  cudaStream_t capStream = nullptr;  // origin stream of the captures (the caller's stream may be the legacy default stream, which cannot be captured)
  // the upload of this handle's topology into constant memory: the stream it was ordered on and an event behind it
  cudaEvent_t topoReady = nullptr;
  cudaStream_t topoStream = nullptr;
  bool topoDone = false;
};

namespace {

int fail(ddb_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
int cudaFail(ddb_handle* h, cudaError_t e, const char* what) {
  return fail(h, DDB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define DDB_CUDA(h, call)                                   \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return cudaFail(h, e__, #call); \
  } while (0)

struct ApiScope {
  std::unique_lock<std::recursive_mutex> lk;
  explicit ApiScope(const ddb_handle* h) : lk(g_devMu[h->device & 63]) {}
};

// make this handle's topology the one in constant memory; the caller holds the device's ApiScope. The upload is ordered on s and,
// unless s is being captured into a CUDA graph, waited for on the host: afterwards kernels of this handle may run on any stream
// without further ordering (binding happens once per handle, or once per switch between handles that share a device).
bool isCapturing(cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); return false; }
  return st != cudaStreamCaptureStatusNone;
}
int bindTopology(ddb_handle* h, cudaStream_t s) {
  if (g_owner[h->device] == h) {
    if (!h->topoDone && s != h->topoStream) {
      // the upload was captured on another stream and has not run yet
      if (isCapturing(s)) return fail(h, DDB_ERR_STATE, "first use of this handle happened inside a stream capture on another stream");
      DDB_CUDA(h, cudaEventSynchronize(h->topoReady));
      h->topoDone = true;
    }
    return DDB_OK;
  }
  const bool capturing = isCapturing(s);
  // another handle's kernels (on any stream or lane) may still be reading the slot: switching owners drains the device first
  if (g_owner[h->device] != nullptr) {
    if (capturing) return fail(h, DDB_ERR_STATE, "switching between handles of one device inside a stream capture");
    DDB_CUDA(h, cudaDeviceSynchronize());
  }
  g_owner[h->device] = nullptr;
  if (!h->topoReady) DDB_CUDA(h, cudaEventCreateWithFlags(&h->topoReady, cudaEventDisableTiming));
  // (the host copy lives in the handle, pageable: the runtime stages it before the call returns)
  DDB_CUDA(h, cudaMemcpyToSymbolAsync(c_topo, &h->topo, sizeof(DevTopo), 0, cudaMemcpyHostToDevice, s));
  if (h->hasContacts) DDB_CUDA(h, contactBindTopology(&h->topo, &h->ctopo, s));
  h->topoStream = s;
  h->topoDone = false;
  if (!capturing) {
    DDB_CUDA(h, cudaStreamSynchronize(s));
    h->topoDone = true;
  } else {
    DDB_CUDA(h, cudaEventRecord(h->topoReady, s));
  }
  g_owner[h->device] = h;
  return DDB_OK;
}

int enter(ddb_handle* h, cudaStream_t s) {
  if (!h) return DDB_ERR_INVALID;
  int dev = -1;
  DDB_CUDA(h, cudaGetDevice(&dev));
  if (dev != h->device) DDB_CUDA(h, cudaSetDevice(h->device));
  return bindTopology(h, s);
}

int allocKeep(ddb_handle* h, int slot) {
  if (!h->hasContacts) return DDB_OK;
  size_t kd, ki;
  contactKeepSizes(h->N, h->n, h->cmax, h->kmax, &kd, &ki);
  if ((int)h->keepD.size() <= slot) { h->keepD.resize(slot + 1, nullptr); h->keepI.resize(slot + 1, nullptr); }
  if (h->keepD[slot]) return DDB_OK;
  DDB_CUDA(h, cudaMalloc(&h->keepD[slot], kd * sizeof(double)));
  DDB_CUDA(h, cudaMalloc(&h->keepI[slot], ki * sizeof(int)));
  DDB_CUDA(h, cudaMemset(h->keepI[slot], 0, ki * sizeof(int)));
  return DDB_OK;
}

// ---- lanes ----
// Worlds are independent, so a batch can run as chunks on separate streams. Inside one kernel that changes nothing; across
// the kernels of a rollout (or the copies of the host-buffer path) it does: a 65536-world kernel lasts ~30 us, of which the
// ramp-up and the drain of its last, partly filled wave leave many SMs idle; with chunks on lanes the next kernel of one chunk
// fills the SMs the other chunk is draining, and host<->device copies of one chunk hide behind the kernels of another.
int ensureLanes(ddb_handle* h) {
  if ((int)h->lanes.size() >= h->nLanes && h->forkEv) return DDB_OK;
  if (!h->forkEv) DDB_CUDA(h, cudaEventCreateWithFlags(&h->forkEv, cudaEventDisableTiming));
  while ((int)h->lanes.size() < h->nLanes) {
    cudaStream_t st = nullptr;
    cudaEvent_t ev = nullptr;
    DDB_CUDA(h, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    DDB_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    h->lanes.push_back(st);
    h->laneDone.push_back(ev);
  }
  return DDB_OK;
}
// chunk c of K: [w0, w1), boundaries on multiples of 64 worlds (CTA size and snapshot tile)
void chunkRange(int N, int K, int c, int* w0, int* w1) {
  const int per = ((N + K - 1) / K + 63) / 64 * 64;
  *w0 = std::min(N, c * per);
  *w1 = std::min(N, (c + 1) * per);
}
int forkLanes(ddb_handle* h, cudaStream_t s, int K) {
  DDB_CUDA(h, cudaEventRecord(h->forkEv, s));
  for (int c = 0; c < K; c++) DDB_CUDA(h, cudaStreamWaitEvent(h->lanes[c], h->forkEv, 0));
  return DDB_OK;
}
int joinLanes(ddb_handle* h, cudaStream_t s, int K) {
  for (int c = 0; c < K; c++) {
    DDB_CUDA(h, cudaEventRecord(h->laneDone[c], h->lanes[c]));
    DDB_CUDA(h, cudaStreamWaitEvent(s, h->laneDone[c], 0));
  }
  return DDB_OK;
}

size_t snapDoubles(const ddb_handle* h) { return snapDoublesPerWorld(h->topo.snapPairs, h->n, h->topo.has_limits != 0) * snapWorldsPadded(h->N); }

}  // namespace

__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" {

int ddb_measure_fp64_peak(int32_t device, double* tflops_out) {
  if (!tflops_out) return DDB_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DDB_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return DDB_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DDB_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return DDB_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_fp64_peak<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return DDB_ERR_CUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops_out = best;
  return DDB_OK;
}

int ddb_abi_version(void) { return DDB_ABI_VERSION; }

void ddb_default_params(ddb_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->gravity[0] = 0; p->gravity[1] = 0; p->gravity[2] = -9.81;
  p->time_step = 1e-3;
  p->contact_clipping_depth = 0.03;
  p->layout = DDB_LAYOUT_WORLD_MAJOR;
  p->n_snapshot_slots = 1;
}

int ddb_create(const ddb_topology* topo, const ddb_params* params, int32_t n_worlds, int32_t device, ddb_handle** out) {
  if (!out) return DDB_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DDB_ERR_NO_DEVICE;
  if (device < 0 || device >= count || device >= 64 || n_worlds < 1) return DDB_ERR_INVALID;
  ddb_handle* h = new ddb_handle();
  *out = h;  // returned even on failure so that ddb_last_error can be read; caller destroys it
  h->device = device;
  h->N = n_worlds;
  int rc = foldTopology(topo, params, h->topo, h->chain, h->err);
  if (rc != DDB_OK) return rc;
  h->n = h->topo.nd; h->nb = h->topo.nb;
  if (h->n == 0) return fail(h, DDB_ERR_INVALID, "world has no degrees of freedom");
  {
    int rc2 = foldContactTopology(topo, params, h->ctopo, h->err);
    if (rc2 != DDB_OK) return rc2;
    h->hasContacts = h->ctopo.n_pairs > 0;
    h->cmax = h->ctopo.cmax; h->kmax = h->ctopo.kmax;
  }
  // DDB_FORCE_GENERIC=1 routes chains through the generic tree kernels (used by the tests to cover both)
  const char* fg = getenv("DDB_FORCE_GENERIC");
  if (fg && fg[0] == '1') h->chain = false;
  if (h->chain && h->nb > 8) h->chain = false;  // chain kernels exist for up to 8 links
  if (h->chain) canonicalizeChain(topo, h->topo);  // chain kernels work in canonical joint frames (chain.cuh)
  int chainKind = h->chain ? 2 : 0;
  for (int i = 0; h->chain && i < h->nb; i++) if (h->topo.b[i].type != DDB_JOINT_REVOLUTE) chainKind = 1;
  h->k = pickTable(h->nb, chainKind);
  DDB_CUDA(h, cudaSetDevice(device));
  int slots = params->n_snapshot_slots < 1 ? 1 : params->n_snapshot_slots;
  h->slots.assign(slots, nullptr);
  h->slotValid.assign(slots, 0);
  for (int i = 0; i < slots; i++) DDB_CUDA(h, cudaMalloc(&h->slots[i], snapDoubles(h) * sizeof(double)));
  if (h->topo.has_limits) {
    std::vector<double> lim(6 * h->n);
    const double* src[6] = {topo->position_lower, topo->position_upper, topo->velocity_lower, topo->velocity_upper, topo->force_lower, topo->force_upper};
    for (int a = 0; a < 6; a++)
      for (int i = 0; i < h->n; i++) lim[a * h->n + i] = src[a] ? src[a][i] : ((a & 1) ? INFINITY : -INFINITY);
    DDB_CUDA(h, cudaMalloc(&h->lim_dev, lim.size() * sizeof(double)));
    DDB_CUDA(h, cudaMemcpy(h->lim_dev, lim.data(), lim.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->lim = Limits{h->lim_dev, h->lim_dev + h->n, h->lim_dev + 2 * h->n, h->lim_dev + 3 * h->n, h->lim_dev + 4 * h->n, h->lim_dev + 5 * h->n};
  }
  DDB_CUDA(h, cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking));
  if (h->hasContacts) {
    size_t wd, wi, ld, li;
    contactScratchSizes(h->N, h->nb, h->n, h->cmax, h->kmax, &wd, &wi, &ld, &li);
    DDB_CUDA(h, cudaMalloc(&h->cwork_d, wd * sizeof(double)));
    DDB_CUDA(h, cudaMalloc(&h->cwork_i, wi * sizeof(int)));
    DDB_CUDA(h, cudaMalloc(&h->clcp_d, ld * sizeof(double)));
    DDB_CUDA(h, cudaMalloc(&h->clcp_i, li * sizeof(int)));
    DDB_CUDA(h, cudaMalloc(&h->lcp_cache, (size_t)h->N * h->kmax * sizeof(double)));
    DDB_CUDA(h, cudaMalloc(&h->n_cache, (size_t)h->N * sizeof(int)));
    DDB_CUDA(h, cudaMalloc(&h->ccounts, (size_t)h->N * 4 * sizeof(int)));
    DDB_CUDA(h, cudaMemset(h->lcp_cache, 0, (size_t)h->N * h->kmax * sizeof(double)));
    DDB_CUDA(h, cudaMemset(h->n_cache, 0, (size_t)h->N * sizeof(int)));
    DDB_CUDA(h, cudaMemset(h->ccounts, 0, (size_t)h->N * 4 * sizeof(int)));
    DDB_CUDA(h, cudaMemset(h->cwork_d, 0, wd * sizeof(double)));
    DDB_CUDA(h, cudaMemset(h->cwork_i, 0, wi * sizeof(int)));
    DDB_CUDA(h, cudaMalloc(&h->bwdStatus, (size_t)h->N * sizeof(int)));
    {
      size_t ci, ad, ai;
      contactPgsSizes(h->N, h->kmax, &ci, &ad, &ai);
      DDB_CUDA(h, cudaMalloc(&h->pgsCtrl, ci * sizeof(int)));
      DDB_CUDA(h, cudaMalloc(&h->pgsX, ad * sizeof(double)));
      DDB_CUDA(h, cudaMalloc(&h->pgsOk, ai * sizeof(int)));
    }
    for (int i = 0; i < slots; i++) { int rk = allocKeep(h, i); if (rk) return rk; }
  }
  return DDB_OK;
}

int ddb_destroy(ddb_handle* h) {
  if (!h) return DDB_OK;
  ApiScope scope(h);
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (g_owner[h->device] == h) g_owner[h->device] = nullptr;
  if (h->topoReady) cudaEventDestroy(h->topoReady);
  for (double* s : h->slots) if (s) cudaFree(s);
  if (h->lim_dev) cudaFree(h->lim_dev);
  if (h->stage) cudaFree(h->stage);
  for (void* p : {(void*)h->cwork_d, (void*)h->cwork_i, (void*)h->clcp_d, (void*)h->clcp_i, (void*)h->lcp_cache, (void*)h->n_cache, (void*)h->ccounts,
                  (void*)h->jacTmp, (void*)h->bwdStatus, (void*)h->pgsCtrl, (void*)h->pgsX, (void*)h->pgsOk})
    if (p) cudaFree(p);
  for (double* p : h->keepD) if (p) cudaFree(p);
  for (int* p : h->keepI) if (p) cudaFree(p);
  if (h->hstream) cudaStreamDestroy(h->hstream);
  for (cudaStream_t l : h->lanes) cudaStreamDestroy(l);
  for (cudaEvent_t e : h->laneDone) cudaEventDestroy(e);
  if (h->forkEv) cudaEventDestroy(h->forkEv);
  if (h->gradRun) cudaFree(h->gradRun);
  for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (h->capStream) cudaStreamDestroy(h->capStream);
  delete h;
  return DDB_OK;
}

const char* ddb_last_error(const ddb_handle* h) { return h ? h->err.c_str() : "null handle"; }
int32_t ddb_num_worlds(const ddb_handle* h) { return h ? h->N : 0; }
int32_t ddb_num_dofs(const ddb_handle* h) { return h ? h->n : 0; }
int64_t ddb_launch_count(const ddb_handle* h) { return h ? h->launches : 0; }

int ddb_snapshot_select(ddb_handle* h, int32_t slot) {
  if (!h) return DDB_ERR_INVALID;
  if (slot < 0 || slot >= (int)h->slots.size()) return fail(h, DDB_ERR_INVALID, "snapshot slot out of range");
  h->cur = slot;
  return DDB_OK;
}
int32_t ddb_snapshot_current(const ddb_handle* h) { return h ? h->cur : -1; }
int ddb_snapshot_reserve(ddb_handle* h, int32_t n_slots, int32_t* n_out) {
  if (!h) return DDB_ERR_INVALID;
  DDB_CUDA(h, cudaSetDevice(h->device));
  while ((int)h->slots.size() < n_slots) {
    double* p = nullptr;
    DDB_CUDA(h, cudaMalloc(&p, snapDoubles(h) * sizeof(double)));
    h->slots.push_back(p);
    h->slotValid.push_back(0);
    int rk = allocKeep(h, (int)h->slots.size() - 1);
    if (rk) return rk;
  }
  if (n_out) *n_out = (int32_t)h->slots.size();
  return DDB_OK;
}

int ddb_forward(ddb_handle* h, void* stream, const double* q, const double* qd, const double* tau, double* qn, double* qdn,
                double* lcp_cache_inout, int32_t* n_rows_inout, int32_t* active_set_out, int32_t* status_out) {
  if (!h || !q || !qd || !tau || !qn || !qdn) return fail(h, DDB_ERR_INVALID, "ddb_forward: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  ApiScope scope(h);
  int rc = enter(h, s);
  if (rc) return rc;
  h->k.fwd(s, h->N, 0, h->N, q, qd, tau, qn, qdn, h->slots[h->cur]);
  DDB_CUDA(h, cudaGetLastError());
  h->launches++;
  if (h->hasContacts) {
    // collision, contact rows, LCP assembly and solve, impulses: qdn holds the unconstrained update on entry
    double* cache = lcp_cache_inout ? lcp_cache_inout : h->lcp_cache;
    int* ncache = n_rows_inout ? n_rows_inout : h->n_cache;
    int nl = 0;
    DDB_CUDA(h, contactLaunch(s, h->N, h->nb, h->n, h->cmax, h->kmax, h->topo.layout, h->slots[h->cur], qdn, cache, ncache, h->ccounts,
                              active_set_out, h->cwork_d, h->cwork_i, h->clcp_d, h->clcp_i, h->keepD[h->cur], h->keepI[h->cur], h->pgsCtrl, h->pgsX,
                              h->pgsOk, &nl));
    h->launches += nl;
    if (status_out)
      DDB_CUDA(h, cudaMemcpy2DAsync(status_out, sizeof(int), h->ccounts + 3, 4 * sizeof(int), sizeof(int), h->N, cudaMemcpyDeviceToDevice, s));
  }
  h->slotValid[h->cur] = 1;
  return DDB_OK;
}

// World::getCachedLCPSolution / setCachedLCPSolution (BoxedLcpConstraintSolver.cpp:174-186) for every world: the solution the
// handle carries from step to step when ddb_forward is given no cache of the caller's. [N][max_rows] doubles and [N] row counts.
int ddb_lcp_cache_get(ddb_handle* h, void* stream, double* cache_out, int32_t* n_rows_out) {
  if (!h || !cache_out || !n_rows_out) return fail(h, DDB_ERR_INVALID, "ddb_lcp_cache_get: null pointer");
  if (!h->hasContacts) return fail(h, DDB_ERR_STATE, "no contact stage on this handle");
  cudaStream_t s = (cudaStream_t)stream;
  DDB_CUDA(h, cudaSetDevice(h->device));
  DDB_CUDA(h, cudaMemcpyAsync(cache_out, h->lcp_cache, (size_t)h->N * h->kmax * sizeof(double), cudaMemcpyDeviceToDevice, s));
  DDB_CUDA(h, cudaMemcpyAsync(n_rows_out, h->n_cache, (size_t)h->N * sizeof(int), cudaMemcpyDeviceToDevice, s));
  return DDB_OK;
}
int ddb_lcp_cache_set(ddb_handle* h, void* stream, const double* cache, const int32_t* n_rows) {
  if (!h || !cache || !n_rows) return fail(h, DDB_ERR_INVALID, "ddb_lcp_cache_set: null pointer");
  if (!h->hasContacts) return fail(h, DDB_ERR_STATE, "no contact stage on this handle");
  cudaStream_t s = (cudaStream_t)stream;
  DDB_CUDA(h, cudaSetDevice(h->device));
  DDB_CUDA(h, cudaMemcpyAsync(h->lcp_cache, cache, (size_t)h->N * h->kmax * sizeof(double), cudaMemcpyDeviceToDevice, s));
  DDB_CUDA(h, cudaMemcpyAsync(h->n_cache, n_rows, (size_t)h->N * sizeof(int), cudaMemcpyDeviceToDevice, s));
  return DDB_OK;
}

int32_t ddb_contact_capacity(const ddb_handle* h, int32_t* max_contacts, int32_t* max_rows) {
  if (!h) return DDB_ERR_INVALID;
  if (max_contacts) *max_contacts = h->cmax;
  if (max_rows) *max_rows = h->kmax;
  return DDB_OK;
}
// Copies the contact-stage results of the last ddb_forward to the host (synchronises the device):
// counts [N][4] = n_contacts, n_rows, n_groups, solver tier mask (bit 1 Dantzig, 2 PGS, 3 friction dropped, 8 capacity);
// for one world: contacts [cmax][26] (layout of ddb_collide_pairs), bodies [cmax][2], dims [cmax] (0 = filtered out),
// the last group's boxed LCP A [n*n] packed, b/lo/hi/x [n] and findex [n]; any pointer may be NULL.
int ddb_contact_debug_host(ddb_handle* h, int32_t* counts, int32_t world, double* contacts, int32_t* bodies, int32_t* dims, double* A,
                           double* b, double* lo, double* hi, double* x, int32_t* findex) {
  if (!h || !h->hasContacts) return fail(h, DDB_ERR_STATE, "no contact stage on this handle");
  DDB_CUDA(h, cudaSetDevice(h->device));
  DDB_CUDA(h, cudaDeviceSynchronize());
  if (counts) DDB_CUDA(h, cudaMemcpy(counts, h->ccounts, (size_t)h->N * 4 * sizeof(int), cudaMemcpyDeviceToHost));
  if (world < 0 || world >= h->N) return DDB_OK;
  ContactLayout L{h->nb, h->n, h->cmax, h->kmax};
  const double* wd = h->cwork_d + (size_t)world * L.doubles();
  const int* wi = h->cwork_i + (size_t)world * L.ints();
  const size_t K = h->kmax, Cm = h->cmax;
  if (contacts) DDB_CUDA(h, cudaMemcpy(contacts, wd + L.oContacts(), Cm * 26 * sizeof(double), cudaMemcpyDeviceToHost));
  if (bodies) {
    DDB_CUDA(h, cudaMemcpy2D(bodies, 2 * sizeof(int), wi, sizeof(int), sizeof(int), Cm, cudaMemcpyDeviceToHost));
    DDB_CUDA(h, cudaMemcpy2D(bodies + 1, 2 * sizeof(int), wi + Cm, sizeof(int), sizeof(int), Cm, cudaMemcpyDeviceToHost));
  }
  if (dims) DDB_CUDA(h, cudaMemcpy(dims, wi + 2 * Cm, Cm * sizeof(int), cudaMemcpyDeviceToHost));
  if (A) DDB_CUDA(h, cudaMemcpy(A, wd + L.oA(), K * K * sizeof(double), cudaMemcpyDeviceToHost));
  if (b) DDB_CUDA(h, cudaMemcpy(b, wd + L.oB(), K * sizeof(double), cudaMemcpyDeviceToHost));
  if (lo) DDB_CUDA(h, cudaMemcpy(lo, wd + L.oB() + K, K * sizeof(double), cudaMemcpyDeviceToHost));
  if (hi) DDB_CUDA(h, cudaMemcpy(hi, wd + L.oB() + 2 * K, K * sizeof(double), cudaMemcpyDeviceToHost));
  if (x) DDB_CUDA(h, cudaMemcpy(x, wd + L.oB() + 5 * K, K * sizeof(double), cudaMemcpyDeviceToHost));
  if (findex) DDB_CUDA(h, cudaMemcpy(findex, wi + 5 * Cm, K * sizeof(int), cudaMemcpyDeviceToHost));
  return DDB_OK;
}

// DDB_CONTACT_PROFILE=1: clock ticks per section of the contact stage, summed over worlds since the last call (host array [16])
int ddb_contact_profile_host(int64_t* ticks_out) { return contactProfileRead(reinterpret_cast<long long*>(ticks_out)) ? DDB_OK : DDB_ERR_STATE; }

// Statistics of the last ddb_forward: Gauss-Seidel sweeps spent per world by the warp-cooperative PGS solver (host array [N]).
int ddb_contact_pgs_sweeps_host(ddb_handle* h, int32_t* sweeps_out) {
  if (!h || !h->hasContacts || !sweeps_out) return fail(h, DDB_ERR_STATE, "no contact stage on this handle");
  DDB_CUDA(h, cudaSetDevice(h->device));
  DDB_CUDA(h, cudaDeviceSynchronize());
  DDB_CUDA(h, cudaMemcpy2D(sweeps_out, sizeof(int), h->pgsCtrl + 3, 8 * sizeof(int), sizeof(int), h->N, cudaMemcpyDeviceToHost));
  return DDB_OK;
}

int ddb_contact_pgs_rows_host(ddb_handle* h, int32_t* rows_out) {
  if (!h || !h->hasContacts || !rows_out) return fail(h, DDB_ERR_STATE, "no contact stage on this handle");
  DDB_CUDA(h, cudaSetDevice(h->device));
  DDB_CUDA(h, cudaDeviceSynchronize());
  DDB_CUDA(h, cudaMemcpy2D(rows_out, sizeof(int), h->pgsCtrl + 5, 8 * sizeof(int), sizeof(int), h->N, cudaMemcpyDeviceToHost));
  return DDB_OK;
}

int ddb_backward(ddb_handle* h, void* stream, const double* gqn, const double* gqdn, double* gq, double* gqd, double* gtau) {
  if (!h || !gqn || !gqdn || !gq || !gqd || !gtau) return fail(h, DDB_ERR_INVALID, "ddb_backward: null pointer");
  if (!h->slotValid[h->cur]) return fail(h, DDB_ERR_STATE, "ddb_backward: no forward pass retained in the current snapshot slot");
  cudaStream_t s = (cudaStream_t)stream;
  ApiScope scope(h);
  int rc = enter(h, s);
  if (rc) return rc;
  if (h->hasContacts) {
    if (h->ctopo.gradients == 0) return fail(h, DDB_ERR_STATE, "ddb_backward: this handle steps with the solver's gradients disabled (reserved[1] = 1)");
    DDB_CUDA(h, contactBackwardLaunch(s, h->N, h->nb, h->n, h->cmax, h->kmax, h->topo.layout, h->slots[h->cur], h->keepD[h->cur], h->keepI[h->cur],
                                      gqn, gqdn, gq, gqd, gtau, h->cwork_d, h->cwork_i, h->clcp_d, h->clcp_i, h->lim_dev, h->bwdStatus));
    h->launches++;
    return DDB_OK;
  }
  h->k.bwd(s, h->N, 0, h->N, h->slots[h->cur], gqn, gqdn, nullptr, nullptr, gq, gqd, gtau, h->lim);
  DDB_CUDA(h, cudaGetLastError());
  h->launches++;
  return DDB_OK;
}

int ddb_jacobians(ddb_handle* h, void* stream, double* pp, double* pv, double* vp, double* vv, double* fv) {
  if (!h) return DDB_ERR_INVALID;
  if (!h->slotValid[h->cur]) return fail(h, DDB_ERR_STATE, "ddb_jacobians: no forward pass retained in the current snapshot slot");
  cudaStream_t s = (cudaStream_t)stream;
  ApiScope scope(h);
  int rc = enter(h, s);
  if (rc) return rc;
  if (h->hasContacts) {
    // rows of the five Jacobians by pulling unit loss gradients back through the step (2 n backward passes)
    if (h->ctopo.gradients == 0) return fail(h, DDB_ERR_STATE, "ddb_jacobians: this handle steps with the solver's gradients disabled (reserved[1] = 1)");
    const size_t V = (size_t)h->N * h->n;
    if (!h->jacTmp) DDB_CUDA(h, cudaMalloc(&h->jacTmp, 5 * V * sizeof(double)));
    double *unit = h->jacTmp, *zero = unit + V, *o1 = unit + 2 * V, *o2 = unit + 3 * V, *o3 = unit + 4 * V;
    DDB_CUDA(h, cudaMemsetAsync(zero, 0, V * sizeof(double), s));
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 0 && !pv && !vv && !fv) continue;
      if (pass == 1 && !pp && !vp) continue;
      for (int j = 0; j < h->n; j++) {
        DDB_CUDA(h, contactUnitVector(s, h->N, h->n, h->topo.layout, j, unit));
        DDB_CUDA(h, contactBackwardLaunch(s, h->N, h->nb, h->n, h->cmax, h->kmax, h->topo.layout, h->slots[h->cur], h->keepD[h->cur],
                                          h->keepI[h->cur], pass == 0 ? zero : unit, pass == 0 ? unit : zero, o1, o2, o3, h->cwork_d, h->cwork_i,
                                          h->clcp_d, h->clcp_i, nullptr, nullptr));
        h->launches++;
        double* dq = pass == 0 ? pv : pp;
        double* dqd = pass == 0 ? vv : vp;
        if (dq) DDB_CUDA(h, contactScatterRow(s, h->N, h->n, h->topo.layout, j, o1, dq));
        if (dqd) DDB_CUDA(h, contactScatterRow(s, h->N, h->n, h->topo.layout, j, o2, dqd));
        if (pass == 0 && fv) DDB_CUDA(h, contactScatterRow(s, h->N, h->n, h->topo.layout, j, o3, fv));
      }
    }
    return DDB_OK;
  }
  h->k.jac(s, h->N, h->slots[h->cur], pp, pv, vp, vv, fv);
  DDB_CUDA(h, cudaGetLastError());
  h->launches++;
  return DDB_OK;
}

int ddb_inv_mass_matrix(ddb_handle* h, void* stream, double* minv) {
  if (!h || !minv) return fail(h, DDB_ERR_INVALID, "ddb_inv_mass_matrix: null pointer");
  if (!h->slotValid[h->cur]) return fail(h, DDB_ERR_STATE, "ddb_inv_mass_matrix: no forward pass retained in the current snapshot slot");
  cudaStream_t s = (cudaStream_t)stream;
  ApiScope scope(h);
  int rc = enter(h, s);
  if (rc) return rc;
  h->k.minv(s, h->N, h->slots[h->cur], minv);
  DDB_CUDA(h, cudaGetLastError());
  h->launches++;
  return DDB_OK;
}

int ddb_set_lanes(ddb_handle* h, int32_t n_lanes) {
  if (!h || n_lanes < 1 || n_lanes > 16) return fail(h, DDB_ERR_INVALID, "ddb_set_lanes: 1..16 lanes");
  h->nLanes = n_lanes;
  return DDB_OK;
}

}  // extern "C"

namespace {
// Optional (DDB_ROLLOUT_GRAPHS=1): when a contact-free rollout call repeats with the same arguments (the same buffers, as in a training
// loop that reuses its tensors) the second call records its 2 x T x lanes launches into a CUDA graph on an internal stream, and this and
// every later call with those arguments replays it with one cudaGraphLaunch on the caller's stream. Results are the kernels' own: the
// graph holds exactly the launches the direct path makes (tests/test_gpu_rollout.py). Off by default: measured on B200 the host issues
// the launches of 12 us kernels fast enough, on one GPU (33.1 us per step either way) and with eight ranks on one host (33.3 us direct,
// 34.2 us replayed, 2000-step timed regions), and building the graph costs about a millisecond on the call that repeats.
bool rolloutGraphsEnabled() { const char* e = getenv("DDB_ROLLOUT_GRAPHS"); return e && e[0] == '1'; }
constexpr size_t kMaxRolloutGraphs = 8;

// returns the instantiated graph for these arguments, building it with `enqueue(origin stream)` when the call repeats; nullptr = launch directly
template <class Enqueue>
cudaGraphExec_t rolloutGraph(ddb_handle* h, cudaStream_t s, int kind, int T, int K, const void* const* args, Enqueue enqueue) {
  if (!rolloutGraphsEnabled() || isCapturing(s)) return nullptr;
  for (size_t i = 0; i < h->graphs.size(); i++)
    if (h->graphs[i].same(kind, T, K, args)) {
      ddb_handle::RolloutGraph g = h->graphs[i];
      h->graphs.erase(h->graphs.begin() + i);
      h->graphs.push_back(g);
      return g.exec;
    }
  const bool repeated = h->lastCall.same(kind, T, K, args);
  h->lastCall.kind = kind; h->lastCall.T = T; h->lastCall.K = K;
  for (int i = 0; i < 6; i++) h->lastCall.args[i] = args[i];
  if (!repeated) return nullptr;
  if (!h->capStream && cudaStreamCreateWithFlags(&h->capStream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (cudaStreamBeginCapture(h->capStream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  const int rc = enqueue(h->capStream);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(h->capStream, &graph);
  ddb_handle::RolloutGraph g = h->lastCall;
  if (rc != DDB_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    return nullptr;
  }
  cudaGraphDestroy(graph);
  if (h->graphs.size() >= kMaxRolloutGraphs) { cudaGraphExecDestroy(h->graphs.front().exec); h->graphs.erase(h->graphs.begin()); }
  h->graphs.push_back(g);
  return g.exec;
}
}  // namespace

extern "C" {

__global__ void k_add_inplace(size_t n, double* __restrict__ a, const double* __restrict__ b) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] += b[i];
}

// T steps from (q0, qd0) under tau[t], every step retained in snapshot slot t (SingleShot::getSnapshots, SingleShot.cpp:719-760).
int ddb_rollout_forward(ddb_handle* h, void* stream, int32_t T, const double* q0, const double* qd0, const double* tau, double* q_seq,
                        double* qd_seq) {
  if (!h || T < 1 || !q0 || !qd0 || !tau || !q_seq || !qd_seq) return fail(h, DDB_ERR_INVALID, "ddb_rollout_forward: bad argument");
  ApiScope scope(h);
  int rc = ddb_snapshot_reserve(h, T, nullptr);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t V = (size_t)h->N * h->n;
  if (h->hasContacts) {
    // the contact stage works on the whole batch (deferred PGS passes, shared scratch): steps run one after the other on s
    for (int t = 0; t < T; t++) {
      h->cur = t;
      rc = ddb_forward(h, stream, t ? q_seq + (t - 1) * V : q0, t ? qd_seq + (t - 1) * V : qd0, tau + t * V, q_seq + t * V, qd_seq + t * V, nullptr,
                       nullptr, nullptr, nullptr);
      if (rc) return rc;
    }
    return DDB_OK;
  }
  rc = enter(h, s);
  if (rc) return rc;
  rc = ensureLanes(h);
  if (rc) return rc;
  const int K = h->nLanes;
  auto enqueue = [&](cudaStream_t origin) -> int {
    int r = forkLanes(h, origin, K);
    if (r) return r;
    for (int t = 0; t < T; t++)
      for (int c = 0; c < K; c++) {
        int w0, w1;
        chunkRange(h->N, K, c, &w0, &w1);
        if (w0 >= w1) continue;
        h->k.fwd(h->lanes[c], h->N, w0, w1, t ? q_seq + (t - 1) * V : q0, t ? qd_seq + (t - 1) * V : qd0, tau + t * V, q_seq + t * V, qd_seq + t * V,
                 h->slots[t]);
      }
    DDB_CUDA(h, cudaGetLastError());
    return joinLanes(h, origin, K);
  };
  const void* gargs[6] = {q0, qd0, tau, q_seq, qd_seq, nullptr};
  if (cudaGraphExec_t exec = rolloutGraph(h, s, 0, T, K, gargs, enqueue)) DDB_CUDA(h, cudaGraphLaunch(exec, s));
  else if ((rc = enqueue(s)) != DDB_OK) return rc;
  for (int c = 0; c < K; c++) { int w0, w1; chunkRange(h->N, K, c, &w0, &w1); if (w1 > w0) h->launches += T; }
  for (int t = 0; t < T; t++) h->slotValid[t] = 1;
  h->cur = T - 1;
  return DDB_OK;
}

// Reverse sweep over the retained steps (SingleShot::backpropGradientWrt, SingleShot.cpp:650-700): g_q_seq[t], g_qd_seq[t] are the
// loss gradients w.r.t. the state after step t; what the later steps pulled back is added before step t is pulled back.
int ddb_rollout_backward(ddb_handle* h, void* stream, int32_t T, const double* g_q_seq, const double* g_qd_seq, double* g_q0, double* g_qd0,
                         double* g_tau) {
  if (!h || T < 1 || !g_q_seq || !g_qd_seq || !g_q0 || !g_qd0 || !g_tau) return fail(h, DDB_ERR_INVALID, "ddb_rollout_backward: bad argument");
  ApiScope scope(h);
  if ((int)h->slots.size() < T) return fail(h, DDB_ERR_STATE, "ddb_rollout_backward: fewer retained steps than T");
  for (int t = 0; t < T; t++)
    if (!h->slotValid[t]) return fail(h, DDB_ERR_STATE, "ddb_rollout_backward: a step of the rollout has no retained forward pass");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t V = (size_t)h->N * h->n;
  DDB_CUDA(h, cudaSetDevice(h->device));
  if (!h->gradRun) DDB_CUDA(h, cudaMalloc(&h->gradRun, 4 * V * sizeof(double)));
  double* run[2][2] = {{h->gradRun, h->gradRun + V}, {h->gradRun + 2 * V, h->gradRun + 3 * V}};
  if (h->hasContacts) {
    for (int t = T - 1; t >= 0; t--) {
      const int o = t & 1;  // step t writes run[o], reads run[1-o] (written by step t+1)
      double* oq = t ? run[o][0] : g_q0;
      double* oqd = t ? run[o][1] : g_qd0;
      const double *iq = g_q_seq + t * V, *iqd = g_qd_seq + t * V;
      if (t < T - 1) {
        // incoming = direct loss on state t + what steps > t pulled back
        k_add_inplace<<<296, 256, 0, s>>>(V, run[1 - o][0], iq);
        k_add_inplace<<<296, 256, 0, s>>>(V, run[1 - o][1], iqd);
        iq = run[1 - o][0]; iqd = run[1 - o][1];
      }
      h->cur = t;
      int rc = ddb_backward(h, stream, iq, iqd, oq, oqd, g_tau + t * V);
      if (rc) return rc;
    }
    return DDB_OK;
  }
  int rc = enter(h, s);
  if (rc) return rc;
  rc = ensureLanes(h);
  if (rc) return rc;
  const int K = h->nLanes;
  auto enqueue = [&](cudaStream_t origin) -> int {
    int r = forkLanes(h, origin, K);
    if (r) return r;
    for (int t = T - 1; t >= 0; t--)
      for (int c = 0; c < K; c++) {
        int w0, w1;
        chunkRange(h->N, K, c, &w0, &w1);
        if (w0 >= w1) continue;
        const int o = t & 1;
        h->k.bwd(h->lanes[c], h->N, w0, w1, h->slots[t], g_q_seq + t * V, g_qd_seq + t * V, t < T - 1 ? run[1 - o][0] : nullptr,
                 t < T - 1 ? run[1 - o][1] : nullptr, t ? run[o][0] : g_q0, t ? run[o][1] : g_qd0, g_tau + t * V, h->lim);
      }
    DDB_CUDA(h, cudaGetLastError());
    return joinLanes(h, origin, K);
  };
  const void* gargs[6] = {g_q_seq, g_qd_seq, g_q0, g_qd0, g_tau, nullptr};
  if (cudaGraphExec_t exec = rolloutGraph(h, s, 1, T, K, gargs, enqueue)) DDB_CUDA(h, cudaGraphLaunch(exec, s));
  else if ((rc = enqueue(s)) != DDB_OK) return rc;
  for (int c = 0; c < K; c++) { int w0, w1; chunkRange(h->N, K, c, &w0, &w1); if (w1 > w0) h->launches += T; }
  return DDB_OK;
}

// ---- host-buffer entry points (the e2e path): H2D, kernel, D2H on the handle's own stream ----
static int ensureStage(ddb_handle* h, size_t doubles) {
  if (h->stageDoubles >= doubles) return DDB_OK;
  if (h->stage) cudaFree(h->stage);
  h->stage = nullptr;
  h->stageDoubles = 0;
  DDB_CUDA(h, cudaMalloc(&h->stage, doubles * sizeof(double)));
  h->stageDoubles = doubles;
  return DDB_OK;
}

// Contact-free handles with world-major vectors run the call as chunks on the lanes: the copy-in of chunk c+1, the kernel of chunk c
// and the copy-out of chunk c-1 overlap (PCIe is full duplex). Measured on B200 / PCIe 5 (65536 five-link worlds, 13 MB each way per
// forward + backward): 1 chunk 0.62 ms, 2 chunks 0.53 ms, 4 chunks 0.57-0.63 ms, 8 chunks 0.75 ms - every extra copy costs ~5 us of DMA
// set-up, also when the pipeline is replayed from a CUDA graph (tried: no gain), and letting the kernels read / write the pinned
// host buffers themselves is 3x slower (40-byte strides over PCIe). Hence two chunks (DDB_HOST_CHUNKS overrides).
static bool hostChunked(const ddb_handle* h) { return !h->hasContacts && h->topo.layout == DDB_LAYOUT_WORLD_MAJOR && h->N >= 4096; }
static bool hostZeroCopy(const ddb_handle* h) { return !h->hasContacts && h->topo.layout == DDB_LAYOUT_WORLD_MAJOR && h->k.fwdHost != nullptr; }
static int envInt(const char* name, int dflt) { const char* e = getenv(name); return e ? std::max(1, atoi(e)) : dflt; }

// kind 0: forward (in: q, qd, tau; out: qn, qdn), kind 1: backward (in: gqn, gqdn; out: gq, gqd, gtau); ptr = the five host pointers in that order
static int enqueueHostChunks(ddb_handle* h, int kind, const void* const* ptr, int K) {
  const size_t V = (size_t)h->N * h->n;
  double* d = h->stage;
  const int L = std::min((int)h->lanes.size(), K), nin = kind == 0 ? 3 : 2;
  int rc = forkLanes(h, h->hstream, L);
  if (rc) return rc;
  for (int c = 0; c < K; c++) {
    int w0, w1;
    chunkRange(h->N, K, c, &w0, &w1);
    if (w0 >= w1) continue;
    cudaStream_t st = h->lanes[c % L];
    const size_t o = (size_t)w0 * h->n, b = (size_t)(w1 - w0) * h->n * sizeof(double);
    for (int a = 0; a < nin; a++) DDB_CUDA(h, cudaMemcpyAsync(d + a * V + o, (const double*)ptr[a] + o, b, cudaMemcpyHostToDevice, st));
    if (kind == 0) h->k.fwd(st, h->N, w0, w1, d, d + V, d + 2 * V, d + 3 * V, d + 4 * V, h->slots[h->cur]);
    else h->k.bwd(st, h->N, w0, w1, h->slots[h->cur], d, d + V, nullptr, nullptr, d + 2 * V, d + 3 * V, d + 4 * V, h->lim);
    for (int a = nin; a < 5; a++) DDB_CUDA(h, cudaMemcpyAsync((double*)ptr[a] + o, d + a * V + o, b, cudaMemcpyDeviceToHost, st));
  }
  DDB_CUDA(h, cudaGetLastError());
  return joinLanes(h, h->hstream, L);
}

// Pinned (page-locked, mapped) host memory can be addressed by the kernels themselves; returns the device-side aliases of the five
// vectors, or false when one of them is ordinary pageable memory (then the copy path below is taken).
static bool mappedHostPointers(const void* const* ptr, const void** dev) {
  for (int i = 0; i < 5; i++) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr[i]) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
    if ((reinterpret_cast<uintptr_t>(a.devicePointer) & 15) != 0) return false;  // 16-byte vector accesses
    dev[i] = a.devicePointer;
  }
  return true;
}
// DDB_HOST_ZEROCOPY=0 forces the cudaMemcpyAsync pipeline (A/B timing)
static bool zeroCopyEnabled() { const char* e = getenv("DDB_HOST_ZEROCOPY"); return !(e && e[0] == '0'); }

// One launch of the HOSTIO kernel over the whole batch: see kernels.cuh. Measured on B200 / PCIe 5 against the two-chunk copy pipeline
// in bench.py's e2e leg.
static int runHostZeroCopy(ddb_handle* h, int kind, const void* const* dev) {
  {
    ApiScope scope(h);
    int rc = enter(h, h->hstream);
    if (rc) return rc;
    if (kind == 0)
      h->k.fwdHost(h->hstream, h->N, 0, h->N, (const double*)dev[0], (const double*)dev[1], (const double*)dev[2], (double*)dev[3], (double*)dev[4],
                   h->slots[h->cur]);
    else
      h->k.bwdHost(h->hstream, h->N, 0, h->N, h->slots[h->cur], (const double*)dev[0], (const double*)dev[1], (double*)dev[2], (double*)dev[3],
                   (double*)dev[4], h->lim);
    DDB_CUDA(h, cudaGetLastError());
    h->launches++;
  }
  DDB_CUDA(h, cudaStreamSynchronize(h->hstream));
  return DDB_OK;
}

static int runHostChunked(ddb_handle* h, int kind, const void* const* ptr) {
  static const int K = envInt("DDB_HOST_CHUNKS", 2);
  {
    ApiScope scope(h);  // released before the host waits for the device
    int rc = enter(h, h->hstream);
    if (rc) return rc;
    rc = ensureLanes(h);
    if (rc) return rc;
    rc = enqueueHostChunks(h, kind, ptr, K);
    if (rc) return rc;
  }
  for (int c = 0; c < K; c++) { int w0, w1; chunkRange(h->N, K, c, &w0, &w1); h->launches += w1 > w0; }
  DDB_CUDA(h, cudaStreamSynchronize(h->hstream));
  return DDB_OK;
}

int ddb_forward_host(ddb_handle* h, const double* q, const double* qd, const double* tau, double* qn, double* qdn) {
  if (!h || !q || !qd || !tau || !qn || !qdn) return fail(h, DDB_ERR_INVALID, "ddb_forward_host: null pointer");
  cudaSetDevice(h->device);
  const size_t V = (size_t)h->N * h->n, B = V * sizeof(double);
  int rc = ensureStage(h, 8 * V);
  if (rc) return rc;
  double* d = h->stage;
  const void* ptr[5] = {q, qd, tau, qn, qdn};
  const void* dev[5];
  const bool zc = hostZeroCopy(h) && zeroCopyEnabled() && mappedHostPointers(ptr, dev);
  if (zc || hostChunked(h)) {
    rc = zc ? runHostZeroCopy(h, 0, dev) : runHostChunked(h, 0, ptr);
    if (rc) return rc;
    h->slotValid[h->cur] = 1;
    return DDB_OK;
  }
  DDB_CUDA(h, cudaMemcpyAsync(d, q, B, cudaMemcpyHostToDevice, h->hstream));
  DDB_CUDA(h, cudaMemcpyAsync(d + V, qd, B, cudaMemcpyHostToDevice, h->hstream));
  DDB_CUDA(h, cudaMemcpyAsync(d + 2 * V, tau, B, cudaMemcpyHostToDevice, h->hstream));
  rc = ddb_forward(h, h->hstream, d, d + V, d + 2 * V, d + 3 * V, d + 4 * V, nullptr, nullptr, nullptr, nullptr);
  if (rc) return rc;
  DDB_CUDA(h, cudaMemcpyAsync(qn, d + 3 * V, B, cudaMemcpyDeviceToHost, h->hstream));
  DDB_CUDA(h, cudaMemcpyAsync(qdn, d + 4 * V, B, cudaMemcpyDeviceToHost, h->hstream));
  DDB_CUDA(h, cudaStreamSynchronize(h->hstream));
  return DDB_OK;
}

int ddb_backward_host(ddb_handle* h, const double* gqn, const double* gqdn, double* gq, double* gqd, double* gtau) {
  if (!h || !gqn || !gqdn || !gq || !gqd || !gtau) return fail(h, DDB_ERR_INVALID, "ddb_backward_host: null pointer");
  cudaSetDevice(h->device);
  const size_t V = (size_t)h->N * h->n, B = V * sizeof(double);
  int rc = ensureStage(h, 8 * V);
  if (rc) return rc;
  double* d = h->stage;
  const void* ptr[5] = {gqn, gqdn, gq, gqd, gtau};
  const void* dev[5];
  const bool zc = hostZeroCopy(h) && zeroCopyEnabled() && mappedHostPointers(ptr, dev);
  if (zc || hostChunked(h)) {
    if (!h->slotValid[h->cur]) return fail(h, DDB_ERR_STATE, "ddb_backward_host: no forward pass retained in the current snapshot slot");
    return zc ? runHostZeroCopy(h, 1, dev) : runHostChunked(h, 1, ptr);
  }
  DDB_CUDA(h, cudaMemcpyAsync(d, gqn, B, cudaMemcpyHostToDevice, h->hstream));
  DDB_CUDA(h, cudaMemcpyAsync(d + V, gqdn, B, cudaMemcpyHostToDevice, h->hstream));
  rc = ddb_backward(h, h->hstream, d, d + V, d + 2 * V, d + 3 * V, d + 4 * V);
  if (rc) return rc;
  DDB_CUDA(h, cudaMemcpyAsync(gq, d + 2 * V, B, cudaMemcpyDeviceToHost, h->hstream));
  DDB_CUDA(h, cudaMemcpyAsync(gqd, d + 3 * V, B, cudaMemcpyDeviceToHost, h->hstream));
  DDB_CUDA(h, cudaMemcpyAsync(gtau, d + 4 * V, B, cudaMemcpyDeviceToHost, h->hstream));
  DDB_CUDA(h, cudaStreamSynchronize(h->hstream));
  return DDB_OK;
}

int ddb_jacobians_host(ddb_handle* h, double* pp, double* pv, double* vp, double* vv, double* fv) {
  if (!h) return DDB_ERR_INVALID;
  cudaSetDevice(h->device);
  const size_t M = (size_t)h->N * h->n * h->n, B = M * sizeof(double);
  double* host[5] = {pp, pv, vp, vv, fv};
  double* dev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int rc = ensureStage(h, 5 * M > 8 * (size_t)h->N * h->n ? 5 * M : 8 * (size_t)h->N * h->n);
  if (rc) return rc;
  for (int i = 0; i < 5; i++) if (host[i]) dev[i] = h->stage + i * M;
  rc = ddb_jacobians(h, h->hstream, dev[0], dev[1], dev[2], dev[3], dev[4]);
  if (rc) return rc;
  for (int i = 0; i < 5; i++) if (host[i]) DDB_CUDA(h, cudaMemcpyAsync(host[i], dev[i], B, cudaMemcpyDeviceToHost, h->hstream));
  DDB_CUDA(h, cudaStreamSynchronize(h->hstream));
  return DDB_OK;
}

}  // extern "C"
