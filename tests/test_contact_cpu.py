"""CPU: the device source of the whole contact step (forward kernel + contact stage, generic instantiation) compiled
for the host, against the oracle (restatement + the reference's own dBoxBox and Dantzig) on seeded contact scenes:
same contact count, group count, solver tier, LCP solution, next state."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from contact_scenes import SCENES
from oracle.oracle import OracleWorld

HERE = os.path.dirname(os.path.abspath(__file__))
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", os.path.join(HERE, "host_emulation"), "-s"], check=True)
    lib = C.CDLL(os.path.join(HERE, "host_emulation", "_build", "libemul.so"))

    def step(desc, q, qd, tau, cache=None, gradients=True):
        topo, params, keep = desc.to_ctypes()
        params.reserved[1] = 0 if gradients else 1
        n = topo.n_dofs
        q, qd, tau = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, qd, tau))
        cachebuf, nc = np.zeros(1024), C.c_int(0)
        if cache is not None:
            cachebuf[:len(cache)] = cache
            nc.value = len(cache)
        qn, qdn = np.zeros(n), np.zeros(n)
        counts, caps = np.zeros(4, dtype=np.int32), np.zeros(2, dtype=np.int32)
        contacts, bodies, dims = np.zeros(64 * 26), np.zeros(128, dtype=np.int32), np.zeros(64, dtype=np.int32)
        A, b, x = np.zeros(192 * 192), np.zeros(192), np.zeros(192)
        act = np.zeros(192, dtype=np.int32)
        rc = lib.emul_contact_step(C.byref(topo), C.byref(params), q.ctypes.data_as(dp), qd.ctypes.data_as(dp), tau.ctypes.data_as(dp),
                                   cachebuf.ctypes.data_as(dp), C.byref(nc), qn.ctypes.data_as(dp), qdn.ctypes.data_as(dp),
                                   counts.ctypes.data_as(ip), contacts.ctypes.data_as(dp), bodies.ctypes.data_as(ip), dims.ctypes.data_as(ip),
                                   A.ctypes.data_as(dp), b.ctypes.data_as(dp), x.ctypes.data_as(dp), caps.ctypes.data_as(ip), act.ctypes.data_as(ip))
        assert rc == 0
        step.lib = lib
        return {"q_next": qn, "qd_next": qdn, "counts": counts, "contacts": contacts.reshape(64, 26)[:counts[0]],
                "x": x[:max(counts[1], 0)], "lcp_cache": cachebuf[:nc.value].copy(), "active_set": act[:max(counts[1], 0)].copy()}

    return step


@pytest.fixture(autouse=True)
def _quiet_ode(capfd):
    yield
    capfd.readouterr()


@pytest.mark.parametrize("scene", sorted(SCENES))
def test_contact_step_matches_oracle(emul, scene):
    """World::step with the constraint solver's gradients disabled: reduce -> Dantzig -> PGS chain."""
    make, states = SCENES[scene]
    desc = make()
    o = OracleWorld(desc)
    q, qd, tau = states(6 if scene == "box_stack" else 40, 3)
    tiers = set()
    flips = 0
    for w in range(len(q)):
        r = o.step_contacts(q[w], qd[w], tau[w])
        e = emul(desc, q[w], qd[w], tau[w], gradients=False)
        if scene == "box_on_pedestal" and e["counts"][3] != r["tier_mask"]:
            # eight coplanar edge contacts: A is rank deficient and whether Dantzig's answer passes isLCPSolutionValid hangs
            # on the last bit of A (the two sides assemble it with different instruction orders); the velocities still agree
            assert list(e["counts"][:3]) == [len(r["contacts"]), len(r["x"]), r["n_groups"]], w
            assert np.allclose(e["qd_next"], r["qd_next"], rtol=1e-6, atol=1e-6)
            flips += 1
            continue
        assert list(e["counts"]) == [len(r["contacts"]), len(r["x"]), r["n_groups"], r["tier_mask"]], w
        tiers.add(r["tier_mask"])
        if len(r["contacts"]):
            assert np.allclose(e["contacts"][:, 0:3], r["contacts"][:, 2:5], rtol=0, atol=1e-13)
            assert np.allclose(e["x"], r["x"], rtol=1e-8, atol=1e-11)
        assert np.allclose(e["qd_next"], r["qd_next"], rtol=1e-10, atol=1e-12)
        assert np.allclose(e["q_next"], r["q_next"], rtol=0, atol=1e-15)
    assert flips <= 2
    if scene == "pendulum_on_floor":
        assert {2, 4, 8} <= tiers  # Dantzig, PGS and friction-dropped tiers all exercised


@pytest.mark.parametrize("scene", sorted(SCENES))
def test_forward_pass_with_gradients_matches_oracle(emul, scene):
    """The gradient-enabled solve of neural::forwardPass over a 3-step rollout that carries the solver's cached solution:
    guessSolution, short-circuit, classification, standardisation. The ACTIVE SET (ConstraintMapping per row) must be
    identical; states agree to 1e-6 (north-star tolerance for contact-rich solves; PGS worlds) / 1e-10 otherwise."""
    make, states = SCENES[scene]
    desc = make()
    o = OracleWorld(desc)
    nw = 3 if scene == "box_stack" else 16
    q, qd, tau = states(nw, 3)
    tiers = set()
    for w in range(nw):
        qo, qdo, co = q[w], qd[w], None
        qe, qde, ce = q[w], qd[w], None
        for t in range(3):
            r = o.step_contacts(qo, qdo, tau[w], lcp_cache=co, gradients=True)
            e = emul(desc, qe, qde, tau[w], cache=ce)
            tiers.add(r["tier_mask"])
            assert list(e["counts"]) == [len(r["contacts"]), len(r["x"]), r["n_groups"], r["tier_mask"]], (w, t)
            assert np.array_equal(e["active_set"], r["active_set"]), (w, t)
            tol = 1e-6 if r["tier_mask"] & 12 else 1e-10
            assert np.allclose(e["qd_next"], r["qd_next"], rtol=tol, atol=tol) and np.allclose(e["q_next"], r["q_next"], atol=1e-13)
            qo, qdo, co = r["q_next"], r["qd_next"], r["lcp_cache"]
            qe, qde, ce = e["q_next"], e["qd_next"], e["lcp_cache"]
    if scene in ("sliding_box", "two_box_stack"):
        assert 1 in tiers  # the short-circuit through the cached solution was taken


def test_cod_min_norm_properties(emul):
    """The oracle's stand-in for Eigen's completeOrthogonalDecomposition().solve: least-squares residual orthogonal to
    the range, solution orthogonal to the null space, on full-rank and rank-deficient matrices (parity with Eigen itself
    is unpinned: Eigen is not vendored)."""
    from oracle import oracle as orc
    rng = np.random.default_rng(0)
    for t in range(200):
        n = int(rng.integers(1, 40))
        r = int(rng.integers(1, n + 1))
        A = rng.normal(size=(n, r)) @ rng.normal(size=(r, n))
        if t % 3 == 0:
            A = A @ A.T
        b = rng.normal(size=n)
        x, rank = orc.cod_solve(A, b)
        assert rank == r
        assert np.allclose(A.T @ (A @ x - b), 0, atol=1e-8 * max(1.0, np.abs(A).max() ** 2))
        xp = np.linalg.pinv(A, rcond=1e-12) @ b
        assert np.allclose(x, xp, rtol=1e-8, atol=1e-9 * max(1.0, np.abs(xp).max()))


@pytest.mark.parametrize("scene", sorted(SCENES))
def test_deferred_pgs_control_flow_equals_inline(emul, scene):
    """The GPU hands its PGS solves to a warp-cooperative kernel and re-enters the stage behind each answered solve. The same
    control flow on the host (scalar PGS for the solves) must reproduce the single-pass stage bit for bit over a rollout."""
    make, states = SCENES[scene]
    desc = make()
    nw = 3 if scene == "box_stack" else 12
    q, qd, tau = states(nw, 9)
    n = q.shape[1]
    most_passes = 1
    for w in range(nw):
        qi, qdi, ci = q[w], qd[w], None
        qe, qde, cache, nc = q[w].copy(), qd[w].copy(), np.zeros(1024), C.c_int(0)
        for t in range(3):
            a = emul(desc, qi, qdi, tau[w], cache=ci)
            topo, params, keep = desc.to_ctypes()
            qn, qdn = np.zeros(n), np.zeros(n)
            counts, x, act, passes = np.zeros(4, dtype=np.int32), np.zeros(192), np.zeros(192, dtype=np.int32), C.c_int(0)
            rc = emul.lib.emul_contact_step_deferred(C.byref(topo), C.byref(params), qe.ctypes.data_as(dp), qde.ctypes.data_as(dp),
                                                     np.ascontiguousarray(tau[w]).ctypes.data_as(dp), cache.ctypes.data_as(dp), C.byref(nc),
                                                     qn.ctypes.data_as(dp), qdn.ctypes.data_as(dp), counts.ctypes.data_as(ip), x.ctypes.data_as(dp),
                                                     act.ctypes.data_as(ip), C.byref(passes))
            assert rc == 0
            most_passes = max(most_passes, passes.value)
            k = max(int(a["counts"][1]), 0)
            assert list(counts) == list(a["counts"]), (w, t)
            assert np.array_equal(x[:k], a["x"]) and np.array_equal(act[:k], a["active_set"]), (w, t)
            assert np.array_equal(qdn, a["qd_next"]) and np.array_equal(qn, a["q_next"]), (w, t)
            assert np.array_equal(cache[:nc.value], a["lcp_cache"]), (w, t)
            qi, qdi, ci = a["q_next"], a["qd_next"], a["lcp_cache"]
            qe, qde = qn, qdn
    if scene in ("box_stack", "pendulum_on_floor"):
        assert most_passes >= 2  # some world did hand a PGS solve over
