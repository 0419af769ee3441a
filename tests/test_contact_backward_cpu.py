"""CPU: the backward pass through a timestep with contacts.

1. The oracle's restatement of the reference's analytic Jacobians (BackpropSnapshot::getVelJacobianWrt and the contact
   geometry gradients of DifferentiableContactConstraint) against central differences of the frozen-active-set velocity
   map that re-run the collision detector at every perturbed state (vertex / face / sphere contacts, where the reference's
   gradient conventions coincide with how the detector's contacts move), on the worlds whose clamping impulses are the
   standardised least-squares ones (f_c = Q^+ b), which is what the reference's formulas differentiate.
2. The device source of the contact backward (contact_bwd.cuh, a vector-Jacobian product in closed form) compiled for the
   host, against the oracle's Jacobians applied to the same loss gradients. The LCP multipliers of a rank-deficient group
   are fixed only up to the null space of A (the next state does not depend on it, the reference's Jacobian formulas do),
   so the oracle evaluates its Jacobians at the multipliers the device's own forward pass found."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from contact_scenes import SCENES
from oracle.oracle import OracleWorld

HERE = os.path.dirname(os.path.abspath(__file__))
dp = C.POINTER(C.c_double)


@pytest.fixture(autouse=True)
def _quiet_ode(capfd):
    yield
    capfd.readouterr()


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))  # the reference's `equals`: abs below 1, relative above


@pytest.mark.parametrize("scene", ["sliding_box", "pendulum_on_floor", "sphere_pendulum_on_floor", "free_box_on_floor", "two_box_stack", "separate_groups", "bouncing_boxes"])
def test_oracle_contact_jacobians_vs_detector_finite_differences(scene):
    make, states = SCENES[scene]
    o = OracleWorld(make())
    q, qd, tau = states(10, 3)
    checked = 0
    for w in range(len(q)):
        a = o.contact_jacobians(q[w], qd[w], tau[w], mode="reference")
        f = o.contact_jacobians(q[w], qd[w], tau[w], mode="frozen_fd", eps=1e-6)
        if not f["structure_kept"] or a["fc_vs_least_squares"] > 1e-9:
            continue  # the frozen map is the reference's formula only where f_c is the least-squares solution Q^+ b
        checked += 1
        for k in ("velVel", "forceVel"):
            assert rel(a[k], f[k]) < 1e-7, (scene, w, k)
        # pos-vel: the detector also moves a vertex-face contact point by O(depth) with the penetration, which the reference's
        # conventions (the point rides on the vertex body) leave out
        assert rel(a["posVel"], f["posVel"]) < 2e-6, (scene, w)
    assert checked >= 5


@pytest.fixture(scope="module")
def emul_backward():
    subprocess.run(["make", "-C", os.path.join(HERE, "host_emulation"), "-s"], check=True)
    lib = C.CDLL(os.path.join(HERE, "host_emulation", "_build", "libemul.so"))

    def run(desc, q, qd, tau, gqn, gqdn):
        topo, params, keep = desc.to_ctypes()
        n = topo.n_dofs
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (q, qd, tau, gqn, gqdn)]
        cachebuf, nc = np.zeros(1024), C.c_int(0)
        outs = [np.zeros(n) for _ in range(5)]
        st, rows, x = C.c_int(0), C.c_int(0), np.zeros(1024)
        rc = lib.emul_contact_step_backward(C.byref(topo), C.byref(params), *[v.ctypes.data_as(dp) for v in a[:3]], cachebuf.ctypes.data_as(dp),
                                            C.byref(nc), a[3].ctypes.data_as(dp), a[4].ctypes.data_as(dp), *[v.ctypes.data_as(dp) for v in outs],
                                            C.byref(st), x.ctypes.data_as(dp), C.byref(rows))
        assert rc == 0
        return outs, st.value, x[:rows.value].copy()

    return run


def vjp_check(o, q, qd, tau, x, mapping, gqn, gqdn, gq, gqd, gtau, tag):
    """Compare one world's backward with the oracle's Jacobians applied to the same loss gradients at 1e-8, widened by the
    measured amplification of a few-ulp perturbation of Q where the numerical rank of Q sits on the edge of the
    decomposition's threshold (there the reference's pseudo-inverse formulas are themselves only that well determined)."""
    J = o.contact_jacobians(q, qd, tau, mode="reference", x_override=x, mapping_override=mapping)
    Jn = o.contact_jacobians(q, qd, tau, mode="reference", x_override=x, mapping_override=mapping, q_noise=4e-16)
    for name, got, f in (("dL/dq", gq, lambda K: K["posPos"].T @ gqn + K["posVel"].T @ gqdn),
                         ("dL/dqd", gqd, lambda K: K["velPos"].T @ gqn + K["velVel"].T @ gqdn),
                         ("dL/dtau", gtau, lambda K: K["forceVel"].T @ gqdn)):
        ref, refn = f(J), f(Jn)
        tol = 1e-8 + 20.0 * rel(refn, ref)
        assert tol < 1e-2, (tag, name, "the oracle itself is not determined here", tol)
        assert rel(got, ref) < tol, (tag, name, rel(got, ref), tol)
    return J


@pytest.mark.parametrize("scene", sorted(SCENES))
def test_device_contact_backward_matches_oracle(emul_backward, scene):
    make, states = SCENES[scene]
    desc = make()
    o = OracleWorld(desc)
    n = o.n
    q, qd, tau = states(8 if scene == "box_stack" else 12, 21)
    rng = np.random.default_rng(5)
    seen = 0
    # box_stack (n = 30, up to 66 classified rows) is slow on the CPU: one well-conditioned world and world 7, whose Q has
    # its numerical rank on the edge of the decomposition's threshold (friction dropped, f_c != Q^+ b)
    for w in ((0, 7) if scene == "box_stack" else range(len(q))):
        gqn, gqdn = rng.normal(size=n), rng.normal(size=n)
        (qn, qdn, gq, gqd, gtau), status, x = emul_backward(desc, q[w], qd[w], tau[w], gqn, gqdn)
        J = vjp_check(o, q[w], qd[w], tau[w], x, None, gqn, gqdn, gq, gqd, gtau, (scene, w))
        seen |= status
        assert (status & 1) == (J["num_clamping"] > 0)
        assert np.allclose(qdn, J["qd_next"], rtol=1e-6, atol=1e-6)
    assert seen & 1  # some world had clamping contacts
