"""Parity of the contact step against the CPU oracle on more seeded worlds than the tests take (same comparison as
tests/test_gpu_contact.py::test_contact_step_matches_oracle): python scripts/big_parity.py"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_gpu_contact as T
from test_gpu_contact import *
def run(scene, N, seed):
    make, states = SCENES[scene]
    desc = make()
    q, qd, tau = states(N, seed)
    world = BatchedWorld(PlainStepDescription(desc), N)
    world.setPositions(cuda(q)); world.setVelocities(cuda(qd)); world.setForces(cuda(tau))
    forwardPass(world)
    qn, qdn = world.getPositions().cpu().numpy(), world.getVelocities().cpu().numpy()
    counts, _, _ = contact_debug(world, -1)
    o = OracleWorld(desc)
    bad = 0; worst = 0.0; tiers = {}
    t0 = time.time()
    for w in range(N):
        r = o.step_contacts(q[w], qd[w], tau[w])
        tiers[r["tier_mask"]] = tiers.get(r["tier_mask"], 0) + 1
        ok = list(counts[w]) == [len(r["contacts"]), len(r["x"]), r["n_groups"], r["tier_mask"]]
        tol = (1e-6, 1e-8) if r["tier_mask"] & 12 else (1e-9, 1e-11)
        err = np.abs(qdn[w] - r["qd_next"]).max() / max(1.0, np.abs(r["qd_next"]).max())
        worst = max(worst, err)
        if not ok or not np.allclose(qdn[w], r["qd_next"], rtol=tol[0], atol=tol[1]): bad += 1
    print(scene, "N", N, "mismatches", bad, "worst rel err %.2e" % worst, "tiers", tiers, "oracle s %.0f" % (time.time() - t0), flush=True)
run("box_stack", 4000, 21)
run("walker_on_ground", 20000, 22)
run("two_box_stack", 10000, 23)
if "half_cheetah_on_ground" in SCENES: run("half_cheetah_on_ground", 20000, 24)
if "atlas_on_ground" in SCENES: run("atlas_on_ground", 1000, 25)


def run_grad(scene, N, seed, T=2):
    """neural::forwardPass (gradient-enabled solve, cached solution carried over T steps): counts, tier, active set, next velocity."""
    make, states = SCENES[scene]
    desc = make()
    q, qd, tau = states(N, seed)
    world = BatchedWorld(desc, N)
    world.setPositions(cuda(q)); world.setVelocities(cuda(qd))
    o = OracleWorld(desc)
    ref = [(q[w], qd[w], None) for w in range(N)]
    dropped = set()
    bad_counts = bad_active = bad_state = 0
    worst = 0.0
    for t in range(T):
        world.setForces(cuda(tau))
        forwardPass(world)
        qdn = world.getVelocities().cpu().numpy()
        active = world.last_active_set.cpu().numpy()
        counts, _, _ = contact_debug(world, -1)
        for w in range(N):
            if w in dropped:
                continue
            r = o.step_contacts(ref[w][0], ref[w][1], tau[w], lcp_cache=ref[w][2], gradients=True)
            ref[w] = (r["q_next"], r["qd_next"], r["lcp_cache"])
            k = len(r["x"])
            if list(counts[w]) != [len(r["contacts"]), k, r["n_groups"], r["tier_mask"]]:
                bad_counts += 1; dropped.add(w); continue
            if not (np.array_equal(active[w, :k], r["active_set"]) and np.all(active[w, k:] == -4)):
                bad_active += 1; dropped.add(w); continue
            tol = 1e-6 if r["tier_mask"] & 12 else 1e-9
            err = np.abs(qdn[w] - r["qd_next"]).max() / max(1.0, np.abs(r["qd_next"]).max())
            worst = max(worst, err)
            if not np.allclose(qdn[w], r["qd_next"], rtol=tol, atol=tol):
                bad_state += 1; dropped.add(w)
    print(scene, "forwardPass N", N, "steps", T, "count/tier mismatches", bad_counts, "active-set mismatches", bad_active, "state mismatches", bad_state,
          "worst rel err %.2e" % worst, flush=True)


run_grad("box_stack", 1500, 31)
run_grad("two_box_stack", 5000, 32)
run_grad("walker_on_ground", 5000, 33)
if "half_cheetah_on_ground" in SCENES: run_grad("half_cheetah_on_ground", 5000, 34)
if "atlas_on_ground" in SCENES: run_grad("atlas_on_ground", 500, 35)
