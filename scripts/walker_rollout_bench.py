"""BASELINE configs[3]: half-cheetah-style URDF walker with ground contact, T-step trajectory rollout with backprop
(loss = |q_T|^2 + |qd_T|^2), through dart_rollout (ddb_rollout_forward / ddb_rollout_backward; every step keeps its snapshot slot
and its contact record). The walker is tests/fixtures/walker.urdf (box feet; written for this repo).
    python scripts/walker_rollout_bench.py [N=32768] [T=100]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from contact_scenes import walker_on_ground, walker_on_ground_states
from diffdart_b200.rollout import dart_rollout
from diffdart_b200.world import BatchedWorld

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
desc = walker_on_ground()
q, qd, tau = walker_on_ground_states(N, 11)
cu = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
free0 = torch.cuda.mem_get_info()[0]
world = BatchedWorld(desc, N)
rng = np.random.default_rng(5)
taus = cu(np.broadcast_to(tau, (T, N, 9)) + rng.uniform(-1, 1, (T, 1, 9)) * (np.arange(9) >= 3))
q0, qd0 = cu(q).requires_grad_(), cu(qd).requires_grad_()
taus.requires_grad_()
def run():
    for t in (q0, qd0, taus):
        t.grad = None
    poses, vels = dart_rollout(world, q0, qd0, taus)
    loss = (poses[-1] ** 2).sum() + (vels[-1] ** 2).sum()
    loss.backward()
    return poses, vels
run()  # warm-up: grows the snapshot pool to T slots
torch.cuda.synchronize()
used = (free0 - torch.cuda.mem_get_info()[0]) / 2 ** 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
poses, vels = run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
status = world.last_solver_status
out = {"workload": "walker (9 DOFs, box feet) on the ground: %d-step rollout with backprop" % T, "N": N, "T": T, "ms_per_rollout": ms,
       "world_steps_per_s": N * T / (ms * 1e-3), "device_memory_GiB": used,
       "finite": bool(torch.isfinite(poses).all() and torch.isfinite(q0.grad).all() and torch.isfinite(taus.grad).all()),
       "grad_norms": [float(q0.grad.norm()), float(qd0.grad.norm()), float(taus.grad.norm())],
       "torso_height_end_mean": float(poses[-1][:, 1].mean())}
print(json.dumps(out))
