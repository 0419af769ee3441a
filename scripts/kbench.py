#!/usr/bin/env python
"""Kernel-level A/B timing of the contact-free path (config 2: 65536 five-link pendulums): forward and backward
launched alone and back to back, cold inputs (rotating sets larger than L2), CUDA events on the launching stream.
  DDB_LIB=diffdart_b200/lib/variants/libddb_X.so python scripts/kbench.py"""
import argparse, ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffdart_b200 import _abi, skeletons

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=65536)
ap.add_argument("--links", type=int, default=5)
ap.add_argument("--reps", type=int, default=200)
ap.add_argument("--sets", type=int, default=8)
args = ap.parse_args()
lib = _abi.load_library()
N, n, SETS = args.n, args.links, args.sets
desc = skeletons.create_multiarm_robot(n)
topo, params, keep = desc.to_ctypes(n_snapshot_slots=SETS)
h = _abi._vp()
assert lib.ddb_create(C.byref(topo), C.byref(params), N, 0, C.byref(h)) == 0
dev = torch.device("cuda", 0)
rng = np.random.default_rng(1)
mk = lambda: torch.tensor(rng.uniform(-1, 1, (N, n)), device=dev)
sets = [([mk() for _ in range(5)], [torch.empty(N, n, dtype=torch.float64, device=dev) for _ in range(5)]) for _ in range(SETS)]
st = torch.cuda.current_stream(dev)
sp = C.c_void_p(st.cuda_stream)
ptr = [([t.data_ptr() for t in a], [t.data_ptr() for t in b]) for a, b in sets]

def fwd(s):
    i, o = ptr[s % SETS]
    lib.ddb_snapshot_select(h, s % SETS)
    assert lib.ddb_forward(h, sp, i[0], i[1], i[2], o[0], o[1], None, None, None, None) == 0

def bwd(s):
    i, o = ptr[s % SETS]
    lib.ddb_snapshot_select(h, s % SETS)
    assert lib.ddb_backward(h, sp, i[3], i[4], o[2], o[3], o[4]) == 0

def timeit(f, reps):
    for s in range(SETS):
        f(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for s in range(reps):
        f(s)
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for s in range(SETS):
    fwd(s)
out = {"lib": os.path.basename(_abi.LIB_PATH), "N": N, "links": n}
out["fwd_us"] = timeit(fwd, args.reps)
out["bwd_us"] = timeit(bwd, args.reps)
out["step_us"] = timeit(lambda s: (fwd(s), bwd(s)), args.reps)
# the same step replayed from a CUDA graph (no host launch cost between the kernels)
g = torch.cuda.CUDAGraph()
cs = torch.cuda.Stream(dev)
cs.wait_stream(st)
with torch.cuda.stream(cs):
    csp = C.c_void_p(cs.cuda_stream)
    def step_on(s, spx):
        i, o = ptr[s % SETS]
        lib.ddb_snapshot_select(h, s % SETS)
        lib.ddb_forward(h, spx, i[0], i[1], i[2], o[0], o[1], None, None, None, None)
        lib.ddb_backward(h, spx, i[3], i[4], o[2], o[3], o[4])
    step_on(0, csp)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=cs):
        for s in range(SETS):
            step_on(s, csp)
torch.cuda.synchronize()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
R = max(1, args.reps // SETS)
e0.record()
for _ in range(R):
    g.replay()
e1.record()
torch.cuda.synchronize()
out["graph_step_us"] = e0.elapsed_time(e1) / (R * SETS) * 1e3
out["ws_per_s_graph"] = N / (out["graph_step_us"] * 1e-6)
# forward + the five materialised Jacobians (BM_Cartpole_Jacobians semantics, unittests/benchmarks/bench_Jacobians.cpp:131-139)
J = [torch.empty(N, n, n, dtype=torch.float64, device=dev) for _ in range(5)]
def jac(s):
    lib.ddb_snapshot_select(h, s % SETS)
    assert lib.ddb_jacobians(h, sp, *[t.data_ptr() for t in J]) == 0
out["jacobians_us"] = timeit(jac, max(8, args.reps // 4))
Minv = torch.empty(N, n, n, dtype=torch.float64, device=dev)
def minv(s):
    lib.ddb_snapshot_select(h, s % SETS)
    assert lib.ddb_inv_mass_matrix(h, sp, Minv.data_ptr()) == 0
out["inv_mass_us"] = timeit(minv, max(8, args.reps // 4))
# T-step rollout with backprop through the rollout entry points, chunks of worlds on 1..4 lanes
T = SETS
taus = torch.tensor(rng.uniform(-1, 1, (T, N, n)), device=dev)
gp, gv = torch.randn(T, N, n, dtype=torch.float64, device=dev), torch.randn(T, N, n, dtype=torch.float64, device=dev)
poses, vels, gtau = torch.empty_like(taus), torch.empty_like(taus), torch.empty_like(taus)
q0, qd0 = sets[0][0][0], sets[0][0][1]
gq0, gqd0 = torch.empty_like(q0), torch.empty_like(q0)
def roll():
    assert lib.ddb_rollout_forward(h, sp, T, q0.data_ptr(), qd0.data_ptr(), taus.data_ptr(), poses.data_ptr(), vels.data_ptr()) == 0
    assert lib.ddb_rollout_backward(h, sp, T, gp.data_ptr(), gv.data_ptr(), gq0.data_ptr(), gqd0.data_ptr(), gtau.data_ptr()) == 0
for lanes in (1, 2, 3, 4, 6):
    lib.ddb_set_lanes(h, lanes)
    for _ in range(3):
        roll()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R = max(1, args.reps // T)
    e0.record(st)
    for _ in range(R):
        roll()
    e1.record(st)
    torch.cuda.synchronize()
    out["rollout_step_us_lanes%d" % lanes] = e0.elapsed_time(e1) / (R * T) * 1e3
print(json.dumps(out))
lib.ddb_destroy(h)
