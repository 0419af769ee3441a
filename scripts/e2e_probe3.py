"""e2e scaling probe: time of ddb_forward_host / ddb_backward_host vs batch size for both host paths, raw copies of the same byte counts."""
import ctypes as C, os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from diffdart_b200 import _abi, skeletons
lib = _abi.load_library(); n = 5
def timeit(f, reps=40):
    for _ in range(5): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return np.median(ts), np.min(ts)
for N in (16384, 65536, 262144):
    topo, params, keep = skeletons.create_multiarm_robot(n).to_ctypes()
    h = _abi._vp(); assert lib.ddb_create(C.byref(topo), C.byref(params), N, 0, C.byref(h)) == 0
    rng = np.random.default_rng(0)
    hin = [torch.tensor(rng.uniform(-1, 1, (N, n))).pin_memory() for _ in range(5)]
    hout = [torch.empty(N, n, dtype=torch.float64).pin_memory() for _ in range(5)]
    fwd = lambda: lib.ddb_forward_host(h, *[t.data_ptr() for t in hin[:3]], hout[0].data_ptr(), hout[1].data_ptr())
    bwd = lambda: lib.ddb_backward_host(h, hin[3].data_ptr(), hin[4].data_ptr(), *[t.data_ptr() for t in hout[2:]])
    for zc in ("0", "1"):
        os.environ["DDB_HOST_ZEROCOPY"] = zc
        (fm, fb), (bm, bb) = timeit(fwd), timeit(bwd)
        print("N %6d zero_copy=%s  fwd med %.1f us (min %.1f)  bwd med %.1f us (min %.1f)   in/out MB fwd %.2f/%.2f" % (N, zc, fm * 1e6, fb * 1e6, bm * 1e6, bb * 1e6, 3 * N * n * 8 / 1e6, 2 * N * n * 8 / 1e6))
    d = torch.empty(3 * N * n, dtype=torch.float64, device="cuda"); hp = torch.empty(3 * N * n, dtype=torch.float64).pin_memory()
    (a, b), (c, e) = timeit(lambda: d.copy_(hp, non_blocking=True)), timeit(lambda: hp.copy_(d, non_blocking=True))
    print("   raw copy of %.2f MB: h2d med %.1f us (min %.1f) = %.1f GB/s   d2h med %.1f us (min %.1f) = %.1f GB/s" % (d.numel() * 8 / 1e6, a * 1e6, b * 1e6, d.numel() * 8 / a / 1e9, c * 1e6, e * 1e6, d.numel() * 8 / c / 1e9))
    lib.ddb_destroy(h)
