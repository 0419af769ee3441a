"""H2D / D2H copy rates from pinned buffers allocated under different CPU affinities (NUMA placement of the pinned pages)."""
import os, time, subprocess, torch
print(subprocess.run("nvidia-smi topo -m | head -20; lscpu | grep -i 'numa\\|socket\\|model name\\|^CPU(s)'; cat /sys/devices/system/node/node*/cpulist", shell=True, capture_output=True, text=True).stdout)
import pynvml
pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
try:
    print("pci bus", pynvml.nvmlDeviceGetPciInfo(hnd).busId, "numa_node file:", open("/sys/bus/pci/devices/%s/numa_node" % pynvml.nvmlDeviceGetPciInfo(hnd).busId.lower()[4:]).read().strip())
except Exception as e:
    print("numa_node lookup failed", e)
def timeit(f, reps=30):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def probe(tag):
    for mb in (8, 256):
        n = mb * 1024 * 1024 // 8
        d = torch.empty(n, dtype=torch.float64, device="cuda"); hp = torch.empty(n, dtype=torch.float64).pin_memory(); hp.zero_()
        t1 = timeit(lambda: d.copy_(hp, non_blocking=True)); t2 = timeit(lambda: hp.copy_(d, non_blocking=True))
        print("%-28s %4d MB  h2d %.1f GB/s  d2h %.1f GB/s" % (tag, mb, n * 8 / t1 / 1e9, n * 8 / t2 / 1e9))
        del d, hp
torch.cuda.init(); torch.zeros(1, device="cuda")
print("affinity now:", sorted(os.sched_getaffinity(0)))
probe("default affinity")
ncpu = os.cpu_count()
nodes = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
for nd in nodes:
    cl = open("/sys/devices/system/node/%s/cpulist" % nd).read().strip()
    cpus = set()
    for part in cl.split(","):
        if "-" in part: a, b = part.split("-"); cpus |= set(range(int(a), int(b) + 1))
        elif part: cpus.add(int(part))
    cpus &= os.sched_getaffinity(0) | cpus
    try:
        os.sched_setaffinity(0, cpus); probe("bound to " + nd)
    except Exception as e:
        print(nd, "bind failed", e)
try:
    pynvml.nvmlDeviceSetCpuAffinity(hnd); print("nvml affinity:", sorted(os.sched_getaffinity(0))[:8], "..."); probe("nvmlDeviceSetCpuAffinity")
except Exception as e:
    print("nvml affinity failed", e)
