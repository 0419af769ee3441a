// ddb_api.cu — the C ABI declared in include/ddb.h: handle management, topology folding,
// kernel dispatch. No torch types, no CPU fallback.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace ddb;

namespace {

#ifndef DDB_BLOCK
#define DDB_BLOCK 64
#endif
#ifndef DDB_MINB
#define DDB_MINB 1
#endif
constexpr int kBlock = DDB_BLOCK;
constexpr int kMinB = DDB_MINB;

struct KernelTable {
  void (*fwd)(cudaStream_t, int, const double*, const double*, const double*, double*, double*, double*);
  void (*bwd)(cudaStream_t, int, double*, const double*, const double*, double*, double*, double*, Limits);
  void (*jac)(cudaStream_t, int, double*, double*, double*, double*, double*, double*);
  void (*minv)(cudaStream_t, int, double*, double*);
};

template <int NB, bool CHAIN>
struct Launch {
  static int grid(int N) { return (N + kBlock - 1) / kBlock; }
  static void fwd(cudaStream_t s, int N, const double* q, const double* qd, const double* tau, double* qn, double* qdn, double* snap) {
    k_forward<NB, CHAIN, kBlock, kMinB><<<grid(N), kBlock, 0, s>>>(N, q, qd, tau, qn, qdn, snap);
  }
  static void bwd(cudaStream_t s, int N, double* snap, const double* gqn, const double* gqdn, double* gq, double* gqd, double* gtau,
                  Limits lim) {
    k_backward<NB, CHAIN, kBlock, kMinB><<<grid(N), kBlock, 0, s>>>(N, snap, gqn, gqdn, gq, gqd, gtau, lim);
  }
  static void jac(cudaStream_t s, int N, double* snap, double* pp, double* pv, double* vp, double* vv, double* fv) {
    k_jacobians<NB, CHAIN, kBlock, kMinB><<<grid(N), kBlock, 0, s>>>(N, snap, pp, pv, vp, vv, fv);
  }
  static void minv(cudaStream_t s, int N, double* snap, double* out) {
    k_inv_mass<NB, CHAIN, kBlock, kMinB><<<grid(N), kBlock, 0, s>>>(N, snap, out);
  }
  static KernelTable table() { return KernelTable{fwd, bwd, jac, minv}; }
};

KernelTable pickTable(int nb, bool chain) {
  if (chain) {
    switch (nb) {
      case 1: return Launch<1, true>::table();
      case 2: return Launch<2, true>::table();
      case 3: return Launch<3, true>::table();
      case 4: return Launch<4, true>::table();
      case 5: return Launch<5, true>::table();
      case 6: return Launch<6, true>::table();
      case 7: return Launch<7, true>::table();
      case 8: return Launch<8, true>::table();
      default: break;
    }
  }
  if (nb <= 8) return Launch<8, false>::table();
  if (nb <= 16) return Launch<16, false>::table();
  if (nb <= 32) return Launch<32, false>::table();
  return Launch<DDB_MAX_BODIES, false>::table();
}

// per-device owner of the __constant__ topology slot
std::mutex g_mu;
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
  // host-buffer path
  cudaStream_t hstream = nullptr;
  double* stage = nullptr;  // device staging: 3n in + 2n out (+ 2n + 3n for backward), see *_host
  size_t stageDoubles = 0;
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

struct HM3 { double m[9]; };
HM3 mmul(const HM3& A, const HM3& B) {
  HM3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return r;
}
void mvec(const HM3& A, const double* x, double* y) {
  for (int i = 0; i < 3; i++) y[i] = A.m[3 * i] * x[0] + A.m[3 * i + 1] * x[1] + A.m[3 * i + 2] * x[2];
}
void hcross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

// Fold T_parentToJoint * J(q) * T_childToJoint^-1 into N0 + c N1 + s N2, P0 + c P1 + s P2.
int foldTopology(const ddb_topology* t, const ddb_params* p, DevTopo& out, bool& chain, std::string& err) {
  if (!t || !p || !t->bodies) { err = "null topology/params"; return DDB_ERR_INVALID; }
  if (t->n_bodies < 1 || t->n_bodies > DDB_MAX_BODIES || t->n_dofs < 0 || t->n_dofs > DDB_MAX_DOFS) {
    err = "n_bodies/n_dofs out of range"; return DDB_ERR_INVALID;
  }
  std::memset(&out, 0, sizeof(out));
  out.nb = t->n_bodies; out.nd = t->n_dofs; out.layout = p->layout;
  out.dt = p->time_step;
  for (int k = 0; k < 3; k++) out.g[k] = p->gravity[k];
  chain = true;
  int dofCount = 0;
  for (int i = 0; i < t->n_bodies; i++) {
    const ddb_body& s = t->bodies[i];
    DevBody& b = out.b[i];
    if (s.parent >= i || s.parent < -1) { err = "bodies must be listed parent-before-child"; return DDB_ERR_INVALID; }
    if (s.joint_type != DDB_JOINT_WELD && s.joint_type != DDB_JOINT_REVOLUTE && s.joint_type != DDB_JOINT_PRISMATIC) {
      err = "joint type not supported by the contact-free kernels (weld/revolute/prismatic only)"; return DDB_ERR_INVALID;
    }
    int nd = s.joint_type == DDB_JOINT_WELD ? 0 : 1;
    if (nd && (s.dof_offset < 0 || s.dof_offset >= t->n_dofs)) { err = "dof_offset out of range"; return DDB_ERR_INVALID; }
    dofCount += nd;
    b.parent = s.parent; b.type = s.joint_type; b.dof = nd ? s.dof_offset : -1; b.gravity = s.gravity_mode != 0;
    if (!(nd == 1 && s.dof_offset == i && s.parent == i - 1)) chain = false;
    HM3 Ra, Rc, Rb;
    std::memcpy(Ra.m, s.parent_to_joint.r, sizeof(Ra.m));
    std::memcpy(Rc.m, s.child_to_joint.r, sizeof(Rc.m));
    const double* pa = s.parent_to_joint.p;
    const double* pc = s.child_to_joint.p;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rb.m[3 * r + c] = Rc.m[3 * c + r];  // Rc^T
    double pb[3];
    mvec(Rb, pc, pb);
    for (int k = 0; k < 3; k++) pb[k] = -pb[k];
    double a[3] = {s.axis[0], s.axis[1], s.axis[2]};
    double an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (nd) {
      if (!(an > 0)) { err = "zero joint axis"; return DDB_ERR_INVALID; }
      for (int k = 0; k < 3; k++) a[k] /= an;  // Joint::setAxis normalizes
    }
    HM3 M0 = mmul(Ra, Rb);
    double Rapb[3];
    mvec(Ra, pb, Rapb);
    if (s.joint_type == DDB_JOINT_REVOLUTE) {
      double Raa[3], Rbta[3];
      mvec(Ra, a, Raa);
      for (int c = 0; c < 3; c++) Rbta[c] = Rb.m[c] * a[0] + Rb.m[3 + c] * a[1] + Rb.m[6 + c] * a[2];  // (a^T Rb)_c
      HM3 K = {{0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0}};
      HM3 M2 = mmul(mmul(Ra, K), Rb);
      double adotpb = a[0] * pb[0] + a[1] * pb[1] + a[2] * pb[2];
      double axpb[3], RaAxpb[3];
      hcross(a, pb, axpb);
      mvec(Ra, axpb, RaAxpb);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          double m1 = Raa[r] * Rbta[c];
          b.N0[3 * r + c] = m1;
          b.N1[3 * r + c] = M0.m[3 * r + c] - m1;
          b.N2[3 * r + c] = M2.m[3 * r + c];
        }
      for (int k = 0; k < 3; k++) {
        b.P0[k] = pa[k] + Raa[k] * adotpb;
        b.P1[k] = Rapb[k] - Raa[k] * adotpb;
        b.P2[k] = RaAxpb[k];
      }
      // S = AdTAngular(T_c2j, axis) (RevoluteJoint.cpp:140-150)
      double w[3], v[3];
      mvec(Rc, a, w);
      hcross(pc, w, v);
      b.S.w = V3{w[0], w[1], w[2]};
      b.S.v = V3{v[0], v[1], v[2]};
    } else {
      std::memcpy(b.N0, M0.m, sizeof(M0.m));
      for (int k = 0; k < 3; k++) b.P0[k] = pa[k] + Rapb[k];
      if (s.joint_type == DDB_JOINT_PRISMATIC) {
        double Raa[3], v[3];
        mvec(Ra, a, Raa);
        for (int k = 0; k < 3; k++) b.P1[k] = Raa[k];
        // S = AdTLinear(T_c2j, axis) (PrismaticJoint.cpp:139-150)
        mvec(Rc, a, v);
        b.S.w = V3{0, 0, 0};
        b.S.v = V3{v[0], v[1], v[2]};
      }
    }
    // Inertia::computeSpatialTensor (Inertia.cpp:394-409)
    const double m = s.mass, cx = s.com[0], cy = s.com[1], cz = s.com[2];
    b.I.m = m;
    b.I.h = V3{m * cx, m * cy, m * cz};
    b.I.J[0] = s.moment[0] + m * (cy * cy + cz * cz);
    b.I.J[1] = s.moment[1] + m * (cx * cx + cz * cz);
    b.I.J[2] = s.moment[2] + m * (cx * cx + cy * cy);
    b.I.J[3] = s.moment[3] - m * cx * cy;
    b.I.J[4] = s.moment[4] - m * cx * cz;
    b.I.J[5] = s.moment[5] - m * cy * cz;
  }
  if (dofCount != t->n_dofs) { err = "n_dofs does not match the joints"; return DDB_ERR_INVALID; }
  auto finite = [&](const double* a, double def) {
    if (!a) return false;
    for (int i = 0; i < t->n_dofs; i++) if (a[i] != def) return true;
    return false;
  };
  out.has_limits = finite(t->position_lower, -INFINITY) || finite(t->position_upper, INFINITY) || finite(t->velocity_lower, -INFINITY) ||
                   finite(t->velocity_upper, INFINITY) || finite(t->force_lower, -INFINITY) || finite(t->force_upper, INFINITY);
  return DDB_OK;
}

// make this handle's topology the one in constant memory (stream-ordered)
int bindTopology(ddb_handle* h, cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_owner[h->device] == h) return DDB_OK;
  DDB_CUDA(h, cudaMemcpyToSymbolAsync(c_topo, &h->topo, sizeof(DevTopo), 0, cudaMemcpyHostToDevice, s));
  // the host copy lives in the handle (pageable): make sure the copy has been staged before anyone frees it
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

size_t snapDoubles(const ddb_handle* h) { return snapDoublesPerWorld(h->nb, h->n, h->topo.has_limits != 0) * (size_t)h->N; }

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
  // DDB_FORCE_GENERIC=1 routes chains through the generic tree kernels (used by the tests to cover both)
  const char* fg = getenv("DDB_FORCE_GENERIC");
  if (fg && fg[0] == '1') h->chain = false;
  h->k = pickTable(h->nb, h->chain);
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
  return DDB_OK;
}

int ddb_destroy(ddb_handle* h) {
  if (!h) return DDB_OK;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_owner[h->device] == h) g_owner[h->device] = nullptr;
  }
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (double* s : h->slots) if (s) cudaFree(s);
  if (h->lim_dev) cudaFree(h->lim_dev);
  if (h->stage) cudaFree(h->stage);
  if (h->hstream) cudaStreamDestroy(h->hstream);
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
  }
  if (n_out) *n_out = (int32_t)h->slots.size();
  return DDB_OK;
}

int ddb_forward(ddb_handle* h, void* stream, const double* q, const double* qd, const double* tau, double* qn, double* qdn,
                double* /*lcp_cache_inout*/, int32_t* /*n_rows_inout*/, int32_t* /*active_set_out*/, int32_t* /*status_out*/) {
  if (!h || !q || !qd || !tau || !qn || !qdn) return fail(h, DDB_ERR_INVALID, "ddb_forward: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = enter(h, s);
  if (rc) return rc;
  h->k.fwd(s, h->N, q, qd, tau, qn, qdn, h->slots[h->cur]);
  DDB_CUDA(h, cudaGetLastError());
  h->launches++;
  h->slotValid[h->cur] = 1;
  return DDB_OK;
}

int ddb_backward(ddb_handle* h, void* stream, const double* gqn, const double* gqdn, double* gq, double* gqd, double* gtau) {
  if (!h || !gqn || !gqdn || !gq || !gqd || !gtau) return fail(h, DDB_ERR_INVALID, "ddb_backward: null pointer");
  if (!h->slotValid[h->cur]) return fail(h, DDB_ERR_STATE, "ddb_backward: no forward pass retained in the current snapshot slot");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = enter(h, s);
  if (rc) return rc;
  h->k.bwd(s, h->N, h->slots[h->cur], gqn, gqdn, gq, gqd, gtau, h->lim);
  DDB_CUDA(h, cudaGetLastError());
  h->launches++;
  return DDB_OK;
}

int ddb_jacobians(ddb_handle* h, void* stream, double* pp, double* pv, double* vp, double* vv, double* fv) {
  if (!h) return DDB_ERR_INVALID;
  if (!h->slotValid[h->cur]) return fail(h, DDB_ERR_STATE, "ddb_jacobians: no forward pass retained in the current snapshot slot");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = enter(h, s);
  if (rc) return rc;
  h->k.jac(s, h->N, h->slots[h->cur], pp, pv, vp, vv, fv);
  DDB_CUDA(h, cudaGetLastError());
  h->launches++;
  return DDB_OK;
}

int ddb_inv_mass_matrix(ddb_handle* h, void* stream, double* minv) {
  if (!h || !minv) return fail(h, DDB_ERR_INVALID, "ddb_inv_mass_matrix: null pointer");
  if (!h->slotValid[h->cur]) return fail(h, DDB_ERR_STATE, "ddb_inv_mass_matrix: no forward pass retained in the current snapshot slot");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = enter(h, s);
  if (rc) return rc;
  h->k.minv(s, h->N, h->slots[h->cur], minv);
  DDB_CUDA(h, cudaGetLastError());
  h->launches++;
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

int ddb_forward_host(ddb_handle* h, const double* q, const double* qd, const double* tau, double* qn, double* qdn) {
  if (!h || !q || !qd || !tau || !qn || !qdn) return fail(h, DDB_ERR_INVALID, "ddb_forward_host: null pointer");
  cudaSetDevice(h->device);
  const size_t V = (size_t)h->N * h->n, B = V * sizeof(double);
  int rc = ensureStage(h, 8 * V);
  if (rc) return rc;
  double* d = h->stage;
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
