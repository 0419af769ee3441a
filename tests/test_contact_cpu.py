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

    def step(desc, q, qd, tau, cache=None):
        topo, params, keep = desc.to_ctypes()
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
        rc = lib.emul_contact_step(C.byref(topo), C.byref(params), q.ctypes.data_as(dp), qd.ctypes.data_as(dp), tau.ctypes.data_as(dp),
                                   cachebuf.ctypes.data_as(dp), C.byref(nc), qn.ctypes.data_as(dp), qdn.ctypes.data_as(dp),
                                   counts.ctypes.data_as(ip), contacts.ctypes.data_as(dp), bodies.ctypes.data_as(ip), dims.ctypes.data_as(ip),
                                   A.ctypes.data_as(dp), b.ctypes.data_as(dp), x.ctypes.data_as(dp), caps.ctypes.data_as(ip))
        assert rc == 0
        return {"q_next": qn, "qd_next": qdn, "counts": counts, "contacts": contacts.reshape(64, 26)[:counts[0]],
                "x": x[:max(counts[1], 0)], "lcp_cache": cachebuf[:nc.value].copy()}

    return step


@pytest.fixture(autouse=True)
def _quiet_ode(capfd):
    yield
    capfd.readouterr()


@pytest.mark.parametrize("scene", sorted(SCENES))
def test_contact_step_matches_oracle(emul, scene):
    make, states = SCENES[scene]
    desc = make()
    o = OracleWorld(desc)
    q, qd, tau = states(16 if scene == "box_stack" else 60, 3)
    tiers = set()
    for w in range(len(q)):
        r = o.step_contacts(q[w], qd[w], tau[w])
        e = emul(desc, q[w], qd[w], tau[w])
        assert list(e["counts"]) == [len(r["contacts"]), len(r["x"]), r["n_groups"], r["tier_mask"]], w
        tiers.add(r["tier_mask"])
        if len(r["contacts"]):
            assert np.allclose(e["contacts"][:, 0:3], r["contacts"][:, 2:5], rtol=0, atol=1e-13)
            assert np.allclose(e["x"], r["x"], rtol=1e-8, atol=1e-11)
        assert np.allclose(e["qd_next"], r["qd_next"], rtol=1e-10, atol=1e-12)
        assert np.allclose(e["q_next"], r["q_next"], rtol=0, atol=1e-15)
    if scene == "pendulum_on_floor":
        assert {2, 4, 8} <= tiers  # Dantzig, PGS and friction-dropped tiers all exercised


def test_rollout_with_lcp_cache(emul):
    """Several consecutive steps carrying the solver's cached solution (mX) from step to step."""
    make, states = SCENES["two_box_stack"]
    desc = make()
    o = OracleWorld(desc)
    q, qd, tau = states(4, 9)
    for w in range(4):
        qo, qdo, co = q[w], qd[w], None
        qe, qde, ce = q[w], qd[w], None
        for t in range(6):
            r = o.step_contacts(qo, qdo, tau[w], lcp_cache=co)
            e = emul(desc, qe, qde, tau[w], cache=ce)
            assert np.allclose(e["qd_next"], r["qd_next"], rtol=1e-9, atol=1e-11) and np.allclose(e["q_next"], r["q_next"], atol=1e-13)
            assert np.allclose(e["lcp_cache"], r["lcp_cache"], rtol=1e-8, atol=1e-11)
            qo, qdo, co = r["q_next"], r["qd_next"], r["lcp_cache"]
            qe, qde, ce = e["q_next"], e["qd_next"], e["lcp_cache"]
