import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from contact_scenes import walker_on_ground, walker_on_ground_states, half_cheetah_on_ground, half_cheetah_on_ground_states, half_cheetah_available
from diffdart_b200.world import BatchedWorld, forwardPass
def prof(name, make, states, N=32768):
    desc = make(); q, qd, tau = states(N, 11)
    world = BatchedWorld(desc, N)
    cu = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
    world.setPositions(cu(q)); world.setVelocities(cu(qd)); world.setForces(cu(tau))
    for _ in range(3): world.setForces(cu(tau)); forwardPass(world)
    p = np.zeros(16, dtype=np.int64); world._lib.ddb_contact_profile_host(p.ctypes.data_as(C.c_void_p))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): world.setForces(cu(tau)); forwardPass(world)
    e1.record(); torch.cuda.synchronize()
    world._lib.ddb_contact_profile_host(p.ctypes.data_as(C.c_void_p))
    from diffdart_b200.world import LossGradient
    snap = forwardPass(world, idempotent=True)
    nxt, this = LossGradient(), LossGradient()
    g = torch.randn(N, q.shape[1], dtype=torch.float64, device="cuda")
    nxt.lossWrtPosition, nxt.lossWrtVelocity = g, g.clone()
    snap.backprop(world, this, nxt)
    torch.cuda.synchronize()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for _ in range(5): snap.backprop(world, this, nxt)
    b1.record(); torch.cuda.synchronize()
    bwd_ms = b0.elapsed_time(b1) / 5
    names = ["collide+constraints", "assembly", "guess", "short_circuit", "dantzig", "reduce_for_pgs", "post_pgs", "commit"]
    tot = max(int(p[:8].sum()), 1)
    counts = np.zeros((N, 4), dtype=np.int32)
    world._lib.ddb_contact_debug_host(world._h, counts.ctypes.data_as(C.c_void_p), -1, None, None, None, None, None, None, None, None, None)
    tiers = {int(k): int(v) for k, v in zip(*np.unique(counts[:, 3], return_counts=True))}
    print(name, "ms/step %.2f bwd %.2f" % (e0.elapsed_time(e1) / 5, bwd_ms), {k: round(float(v) / tot, 3) for k, v in zip(names, p[:8])}, "contacts", counts[:, 0].mean(), "rows", counts[:, 1].mean(), "tiers", tiers)
prof("walker", walker_on_ground, walker_on_ground_states)
if half_cheetah_available(): prof("cheetah", half_cheetah_on_ground, half_cheetah_on_ground_states)
