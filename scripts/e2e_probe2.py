"""e2e path A/B on one GPU: HOSTIO kernels (zero-copy through shared memory) vs the cudaMemcpyAsync chunk pipeline, next to the raw
PCIe copy rates that bound both. Usage: python scripts/e2e_probe2.py"""
import ctypes as C, os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from diffdart_b200 import _abi, skeletons
lib = _abi.load_library(); N, n = 65536, 5
topo, params, keep = skeletons.create_multiarm_robot(n).to_ctypes()
h = _abi._vp(); assert lib.ddb_create(C.byref(topo), C.byref(params), N, 0, C.byref(h)) == 0
rng = np.random.default_rng(0)
hin = [torch.tensor(rng.uniform(-1, 1, (N, n))).pin_memory() for _ in range(5)]
hout = [torch.empty(N, n, dtype=torch.float64).pin_memory() for _ in range(5)]
def step():
    assert lib.ddb_forward_host(h, *[t.data_ptr() for t in hin[:3]], hout[0].data_ptr(), hout[1].data_ptr()) == 0
    assert lib.ddb_backward_host(h, hin[3].data_ptr(), hin[4].data_ptr(), *[t.data_ptr() for t in hout[2:]]) == 0
def fwd_only():
    assert lib.ddb_forward_host(h, *[t.data_ptr() for t in hin[:3]], hout[0].data_ptr(), hout[1].data_ptr()) == 0
def bwd_only():
    assert lib.ddb_backward_host(h, hin[3].data_ptr(), hin[4].data_ptr(), *[t.data_ptr() for t in hout[2:]]) == 0
def timeit(f, reps=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
res = {}
for zc in ("0", "1"):
    os.environ["DDB_HOST_ZEROCOPY"] = zc
    for t in hout: t.zero_()
    dt = timeit(step); res[zc] = [t.clone() for t in hout]
    print("zero_copy=%s  ms/step %.3f  ws/s %.3e   fwd %.3f ms  bwd %.3f ms" % (zc, dt * 1e3, N / dt, timeit(fwd_only) * 1e3, timeit(bwd_only) * 1e3))
print("bit-equal:", all(torch.equal(a, b) for a, b in zip(res["0"], res["1"])))
# raw PCIe rates: one 7.86 MB copy each way, alone and both directions at once
d = torch.empty(3 * N * n, dtype=torch.float64, device="cuda"); hp = torch.empty(3 * N * n, dtype=torch.float64).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d); hp2 = torch.empty_like(hp).pin_memory()
def h2d(): d.copy_(hp, non_blocking=True)
def d2h(): hp2.copy_(d2, non_blocking=True)
def both():
    with torch.cuda.stream(s1): d.copy_(hp, non_blocking=True)
    with torch.cuda.stream(s2): hp2.copy_(d2, non_blocking=True)
for name, f in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    dt = timeit(f, 100); print("%s 7.86 MB: %.1f us = %.1f GB/s per direction" % (name, dt * 1e6, d.numel() * 8 / dt / 1e9))
