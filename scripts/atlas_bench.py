"""BASELINE configs[4]: Atlas (27 revolute joints + FreeJoint root, box feet) standing on the ground box, contact-rich,
N worlds per GPU: forwardPass with the contact solve + backward through the contact solution, world-steps/s, and a check of a few worlds
against the CPU oracle. The URDFs are the reference's own data files (data/sdf/atlas/atlas_v3_box_colliders.urdf, ground.urdf): they are
not part of this repo; copy them to baseline/_ref/atlas/ (git-ignored, travels to the GPU box) before running:
    mkdir -p baseline/_ref/atlas && cp /root/reference/data/sdf/atlas/{atlas_v3_box_colliders,ground}.urdf baseline/_ref/atlas/
    python scripts/atlas_bench.py [N] [steps]"""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from diffdart_b200.skeletons import WorldDescription
from diffdart_b200.urdf_loader import load_urdf
from diffdart_b200.world import BatchedWorld, LossGradient, forwardPass

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
d = os.path.join(ROOT, "baseline", "_ref", "atlas")
atlas, rep = load_urdf(os.path.join(d, "atlas_v3_box_colliders.urdf"))
ground, _ = load_urdf(os.path.join(d, "ground.urdf"))
desc = WorldDescription().addSkeleton(atlas).addSkeleton(ground).setGravity([0.0, -9.81, 0.0])
n = desc.getNumDofs()
rng = np.random.default_rng(13)
q = np.zeros((N, n)); q[:, 0] = -0.5 * math.pi; q[:, 4] = -0.01   # test_AtlasGradients.cpp:222-223
q += rng.uniform(-0.01, 0.01, (N, n)); qd = rng.uniform(-0.1, 0.1, (N, n)); tau = np.zeros((N, n))
cu = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
world = BatchedWorld(desc, N)
world.setPositions(cu(q)); world.setVelocities(cu(qd))
g = torch.randn(N, n, dtype=torch.float64, device="cuda")
nxt, this = LossGradient(), LossGradient()
nxt.lossWrtPosition, nxt.lossWrtVelocity = g, g.clone()
def step():
    world.setForces(cu(tau))
    snap = forwardPass(world)
    snap.backprop(world, this, nxt)
    return snap
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
status = world.last_solver_status.cpu().numpy()
active = world.last_active_set.cpu().numpy()
rows = (active != -4).sum(axis=1)
out = {"workload": "atlas (33 DOFs, 35 bodies, box feet on ground box), forwardPass + backward", "N": N, "ms_per_step": ms, "world_steps_per_s": N / (ms * 1e-3),
       "rows_mean": float(rows.mean()), "rows_max": int(rows.max()),
       "tier_mask_hist": {str(int(k)): int(v) for k, v in zip(*np.unique(status, return_counts=True))},
       "finite": bool(torch.isfinite(world.getPositions()).all() and torch.isfinite(this.lossWrtPosition).all())}
# parity of the first step on a few worlds against the oracle (fresh worlds: no cached LCP solution on either side)
try:
    from oracle.oracle import OracleWorld
    K = 4
    w2 = BatchedWorld(desc, K)
    w2.setPositions(cu(q[:K])); w2.setVelocities(cu(qd[:K])); w2.setForces(cu(tau[:K]))
    forwardPass(w2)
    o = OracleWorld(desc)
    errs, same_set, tiers = [], [], []
    act2 = w2.last_active_set.cpu().numpy()
    for k in range(K):
        r = o.step_contacts(q[k], qd[k], tau[k], gradients=True)
        errs.append(float(np.max(np.abs(w2.getVelocities()[k].cpu().numpy() - r["qd_next"]))))
        kk = len(r["active_set"])
        same_set.append(bool(np.array_equal(act2[k][:kk], r["active_set"]) and np.all(act2[k][kk:] == -4)))
        tiers.append((int(w2.last_solver_status[k]), r["tier_mask"]))
    out["oracle_active_set_equal"] = same_set
    out["oracle_tiers_gpu_vs_cpu"] = tiers
    out["oracle_qd_next_max_abs_err"] = errs
except Exception as exc:
    out["oracle_check"] = "skipped: %s" % str(exc)[:120]
print(json.dumps(out))
