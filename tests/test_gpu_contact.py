"""GPU: World::step with contacts (collision -> contact rows -> LCP by impulse tests -> Dantzig/PGS chain -> impulses)
through ddb_forward, against the oracle (restatement + the reference's own dBoxBox and Dantzig) world by world."""
import ctypes as C

import numpy as np
import pytest
import torch

from contact_scenes import SCENES
from diffdart_b200 import _abi
from diffdart_b200.world import BatchedWorld, forwardPass
from oracle.oracle import OracleWorld

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


@pytest.fixture(autouse=True)
def _quiet_ode(capfd):
    yield
    capfd.readouterr()


def contact_debug(world, w):
    lib = world._lib
    cm, km = C.c_int32(0), C.c_int32(0)
    lib.ddb_contact_capacity(world._h, C.byref(cm), C.byref(km))
    cm, km = cm.value, km.value
    counts = np.zeros((world.N, 4), dtype=np.int32)
    contacts, x = np.zeros((cm, 26)), np.zeros(km)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.ddb_contact_debug_host(world._h, p(counts), w, p(contacts), None, None, None, None, None, None, p(x), None)
    assert rc == 0
    return counts, contacts, x


class PlainStepDescription:
    """Wraps a WorldDescription so that the handle runs World::step with the solver's gradients disabled (reserved[1] = 1)."""

    def __init__(self, desc):
        self._d = desc

    def to_ctypes(self, *a, **k):
        topo, params, keep = self._d.to_ctypes(*a, **k)
        params.reserved[1] = 1
        return topo, params, keep


@pytest.mark.parametrize("scene", sorted(SCENES))
def test_contact_step_matches_oracle(scene):
    """World::step with the constraint solver's gradients disabled."""
    make, states = SCENES[scene]
    desc = make()
    N = 64 if scene == "box_stack" else 200
    q, qd, tau = states(N, 5)
    world = BatchedWorld(PlainStepDescription(desc), N)
    world.setPositions(cuda(q)); world.setVelocities(cuda(qd)); world.setForces(cuda(tau))
    forwardPass(world)
    qn, qdn = world.getPositions().cpu().numpy(), world.getVelocities().cpu().numpy()
    counts, _, _ = contact_debug(world, -1)
    o = OracleWorld(desc)
    for w in range(N):
        r = o.step_contacts(q[w], qd[w], tau[w])
        assert list(counts[w]) == [len(r["contacts"]), len(r["x"]), r["n_groups"], r["tier_mask"]], w
        assert np.allclose(qdn[w], r["qd_next"], rtol=1e-9, atol=1e-11), w
        assert np.allclose(qn[w], r["q_next"], rtol=0, atol=1e-14), w
        if w % 37 == 0 and len(r["x"]):
            _, contacts, x = contact_debug(world, w)
            assert np.allclose(contacts[:len(r["contacts"]), 0:3], r["contacts"][:, 2:5], atol=1e-13)
            assert np.allclose(x[:len(r["x"])], r["x"], rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("scene", sorted(SCENES))
def test_forward_pass_rollout_active_set_bit_exact(scene):
    """neural::forwardPass (gradient-enabled solve) over a rollout that carries the solver's cached solution: the active
    set (ConstraintMapping per LCP row), contact / row / group counts and solver tier must equal the oracle's exactly;
    next states within 1e-6 (north-star tolerance for contact-rich solves) for PGS worlds, 1e-9 otherwise."""
    make, states = SCENES[scene]
    desc = make()
    N, T = (24, 3) if scene == "box_stack" else (64, 4)
    q, qd, tau = states(N, 11)
    world = BatchedWorld(desc, N)
    world.setPositions(cuda(q)); world.setVelocities(cuda(qd))
    o = OracleWorld(desc)
    ref = [(q[w], qd[w], None) for w in range(N)]
    tiers = set()
    for t in range(T):
        world.setForces(cuda(tau))
        forwardPass(world)
        qn, qdn = world.getPositions().cpu().numpy(), world.getVelocities().cpu().numpy()
        active = world.last_active_set.cpu().numpy()
        status = world.last_solver_status.cpu().numpy()
        counts, _, _ = contact_debug(world, -1)
        for w in range(N):
            r = o.step_contacts(ref[w][0], ref[w][1], tau[w], lcp_cache=ref[w][2], gradients=True)
            ref[w] = (r["q_next"], r["qd_next"], r["lcp_cache"])
            k = len(r["x"])
            tiers.add(r["tier_mask"])
            assert list(counts[w]) == [len(r["contacts"]), k, r["n_groups"], r["tier_mask"]], (t, w)
            assert status[w] == r["tier_mask"]
            assert np.array_equal(active[w, :k], r["active_set"]) and np.all(active[w, k:] == -4), (t, w)
            tol = 1e-6 if r["tier_mask"] & 12 else 1e-9
            assert np.allclose(qdn[w], r["qd_next"], rtol=tol, atol=tol), (t, w)
            assert np.allclose(qn[w], r["q_next"], rtol=0, atol=1e-9), (t, w)  # q' inherits dt * (qd differences of earlier steps)
    if scene in ("sliding_box", "two_box_stack"):
        assert 1 in tiers


def test_backward_through_contacts_is_refused_not_faked():
    make, states = SCENES["sliding_box"]
    desc = make()
    q, qd, tau = states(4, 1)
    world = BatchedWorld(desc, 4)
    world.setPositions(cuda(q)); world.setVelocities(cuda(qd)); world.setForces(cuda(tau))
    snap = forwardPass(world)
    with pytest.raises(RuntimeError, match="not implemented"):
        snap.getJacobians()
