"""Forward-only throughput of the contact path on BASELINE config 3 (box stacking, 5 FreeJoint boxes on a floor)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
from contact_scenes import box_stack, box_stack_states
from diffdart_b200.world import BatchedWorld, forwardPass

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
desc = box_stack()
q, qd, tau = box_stack_states(N, 7)
world = BatchedWorld(desc, N)
cu = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
world.setPositions(cu(q)); world.setVelocities(cu(qd)); world.setForces(cu(tau))
for _ in range(3):  # warm-up steps also populate the LCP cache
    world.setForces(cu(tau)); forwardPass(world)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    world.setForces(cu(tau)); forwardPass(world)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
# backward through the contact solution of the last step
from diffdart_b200.world import LossGradient
snap = forwardPass(world, idempotent=True)
nxt, this = LossGradient(), LossGradient()
g = torch.randn(N, q.shape[1], dtype=torch.float64, device="cuda")
nxt.lossWrtPosition, nxt.lossWrtVelocity = g, g.clone()
snap.backprop(world, this, nxt)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    snap.backprop(world, this, nxt)
e1.record(); torch.cuda.synchronize()
ms_bwd = e0.elapsed_time(e1) / steps
counts = np.zeros((N, 4), dtype=np.int32)
world._lib.ddb_contact_debug_host(world._h, counts.ctypes.data_as(C.c_void_p), -1, None, None, None, None, None, None, None, None, None)
prof = np.zeros(16, dtype=np.int64)
if world._lib.ddb_contact_profile_host(prof.ctypes.data_as(C.c_void_p)) == 0:
    names = ["collide+constraints", "assembly", "guess", "short_circuit", "dantzig", "reduce_for_pgs", "post_pgs", "commit"]
    tot = max(int(prof[:8].sum()), 1)
    print("contact stage profile (share of thread time):", {k: round(float(v) / tot, 3) for k, v in zip(names, prof[:8])}, file=sys.stderr)
sw = np.zeros(N, dtype=np.int32)
world._lib.ddb_contact_pgs_sweeps_host(world._h, sw.ctypes.data_as(C.c_void_p))
rows = counts[:, 1]
stats = {"sweeps_mean": float(sw.mean()), "sweeps_p50": float(np.percentile(sw, 50)), "sweeps_p90": float(np.percentile(sw, 90)), "sweeps_max": int(sw.max()),
         "frac_10000": float((sw >= 10000).mean()), "rows_hist": {str(b): int(((rows > lo) & (rows <= b)).sum()) for lo, b in ((0, 64), (64, 92), (92, 116), (116, 160), (160, 999))}}
pr = np.zeros(N, dtype=np.int32)
world._lib.ddb_contact_pgs_rows_host(world._h, pr.ctypes.data_as(C.c_void_p))
stats["pgs_rows_exact"] = {int(k): int(v) for k, v in zip(*np.unique(pr, return_counts=True))}
tiers = {int(k): int(v) for k, v in zip(*np.unique(counts[:, 3], return_counts=True))}
print(json.dumps({"workload": "box_stack forward, backward", "N": N, "ms_per_step": ms, "ms_backward": ms_bwd, "world_steps_per_s": N / ((ms + ms_bwd) * 1e-3),
                  "contacts_mean": float(counts[:, 0].mean()), "rows_mean": float(counts[:, 1].mean()), "tier_mask_hist": tiers, "pgs": stats}))
