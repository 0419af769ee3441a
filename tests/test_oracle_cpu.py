"""CPU tests pinning the oracle (test infrastructure) before it is trusted as the checker.

Pins, in order of strength:
  * the reference's own scalar-expanded transformInertia / expAngular (tests/golden/geometry.json,
    transliterated from dart/math/Geometry.cpp by tests/golden/make_golden.py);
  * closed-form cartpole dynamics for the reference's cartpole fixture
    (unittests/comprehensive/test_Gradients.cpp:1254-1299);
  * first-principles checks that do not share code with the recursions: the mass matrix as
    sum_i J_i^T I_i J_i with body Jacobians from finite-differenced transforms, Lagrange's
    equations for C+g, and energy conservation of the step;
  * the reference-recipe Jacobians (the reference's finite differences) against exact ones
    under the reference's own tolerance rule.
"""
import json
import os

import numpy as np
import pytest

from diffdart_b200 import skeletons
from oracle import oracle as orc
from util import config_inputs, max_ref_err, random_inputs, ref_equals

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_geometry_golden():
    g = json.load(open(os.path.join(GOLDEN, "geometry.json")))
    for case in g["transformInertia"]:
        out = orc.transform_inertia(case["T"], case["I"])
        assert np.allclose(out, np.array(case["out"]), rtol=1e-13, atol=1e-13)
    for case in g["expAngular"]:
        assert np.allclose(orc.exp_angular(case["s"]), np.array(case["out"]), rtol=1e-14, atol=1e-15)


def test_exp_log_roundtrip():
    rng = np.random.default_rng(3)
    for k in range(50):
        w = rng.normal(size=3)
        w *= rng.uniform(1e-5, 3.1) / np.linalg.norm(w)
        R = orc.exp_map_rot(w)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6 if np.linalg.norm(w) < 1e-3 else 1e-12)
        if np.linalg.norm(w) > 1e-3:
            assert np.allclose(orc.log_map_rot(R), w, atol=1e-9)
    # the reference switches to the 2nd-order Taylor series below 1e-3 (Geometry.cpp:479-483)
    w = np.array([3e-4, -2e-4, 1e-4])
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    assert np.array_equal(orc.exp_map_rot(w), np.eye(3) + K + 0.5 * K @ K)


def test_cartpole_closed_form():
    desc = skeletons.create_cartpole()
    o = orc.OracleWorld(desc)
    q, qd, tau = config_inputs("cartpole", 64)
    th, thd, xd = q[:, 1], qd[:, 1], qd[:, 0]
    # cart m=1; pole m=1, I_zz=1 about its COM which sits 0.5 above the pivot at th=0; g=(0,-9.81,0)
    l, mp, mc, Izz, g = 0.5, 1.0, 1.0, 1.0, 9.81
    qdd = o.forward_dynamics(q, qd, tau)
    for i in range(len(th)):
        M = np.array([[mc + mp, -mp * l * np.cos(th[i])], [-mp * l * np.cos(th[i]), Izz + mp * l * l]])
        # x_com = x - l sin th, y_com = l cos th
        c = np.array([mp * l * np.sin(th[i]) * thd[i] ** 2, -mp * g * l * np.sin(th[i])])
        assert np.allclose(qdd[i], np.linalg.solve(M, tau[i] - c), rtol=1e-11, atol=1e-11)


def _numeric_mass_matrix(o, q, eps=1e-6):
    """M = sum_i J_i^T I_i J_i with body Jacobians from central differences of the body transforms."""
    n, nb = o.n, o.nb
    I = o.spatial_inertias()
    T0 = o.body_transforms(q[None])[0]
    J = np.zeros((nb, 6, n))
    for k in range(n):
        qp, qm = q.copy(), q.copy()
        qp[k] += eps
        qm[k] -= eps
        Tp, Tm = o.body_transforms(qp[None])[0], o.body_transforms(qm[None])[0]
        for i in range(nb):
            R = T0[i, :9].reshape(3, 3)
            dR = (Tp[i, :9] - Tm[i, :9]).reshape(3, 3) / (2 * eps)
            dp = (Tp[i, 9:] - Tm[i, 9:]) / (2 * eps)
            W = R.T @ dR  # body angular velocity (skew)
            J[i, :3, k] = [W[2, 1], W[0, 2], W[1, 0]]
            J[i, 3:, k] = R.T @ dp
    return sum(J[i].T @ I[i] @ J[i] for i in range(nb))


@pytest.mark.parametrize("seed,nb", [(1, 4), (2, 7), (3, 12)])
def test_mass_matrix_and_lagrange(seed, nb):
    desc = skeletons.create_random_tree(nb, seed)
    o = orc.OracleWorld(desc)
    n = o.n
    q, qd, tau = random_inputs(n, 1, seed + 100)
    q, qd, tau = q[0], qd[0], tau[0]
    M = _numeric_mass_matrix(o, q)
    Minv = o.inv_mass_matrix(q[None])[0]
    assert np.allclose(Minv @ M, np.eye(n), atol=2e-7)
    assert np.allclose(Minv, Minv.T, atol=1e-13)
    # Lagrange: C = dM/dt qd - 1/2 d(qd^T M qd)/dq + dU/dq, with U from body COM heights
    eps = 1e-5
    g = np.array(desc.gravity)
    I = o.spatial_inertias()

    def potential(qq):
        T = o.body_transforms(qq[None])[0]
        U = 0.0
        for i in range(o.nb):
            m = I[i, 3, 3]
            com = np.array([I[i, 2, 4], I[i, 0, 5], I[i, 1, 3]]) / m  # m*skew(c) block
            U -= m * g @ (T[i, :9].reshape(3, 3) @ com + T[i, 9:])
        return U

    dMdq = np.zeros((n, n, n))
    dU = np.zeros(n)
    for k in range(n):
        qp, qm = q.copy(), q.copy()
        qp[k] += eps
        qm[k] -= eps
        dMdq[k] = (_numeric_mass_matrix(o, qp) - _numeric_mass_matrix(o, qm)) / (2 * eps)
        dU[k] = (potential(qp) - potential(qm)) / (2 * eps)
    Mdot = np.einsum("kij,k->ij", dMdq, qd)
    C_lagrange = Mdot @ qd - 0.5 * np.einsum("kij,i,j->k", dMdq, qd, qd) + dU
    C = o.coriolis_gravity(q[None], qd[None])[0]
    assert np.allclose(C, C_lagrange, atol=5e-5)
    # forward dynamics consistency: M qdd + C = tau
    qdd = o.forward_dynamics(q[None], qd[None], tau[None])[0]
    assert np.allclose(M @ qdd + C, tau, atol=1e-5)


def test_energy_conservation_small_dt():
    desc = skeletons.create_multiarm_robot(3)
    desc.setTimeStep(1e-5)
    o = orc.OracleWorld(desc)
    q, qd, _ = random_inputs(3, 1, 5)
    tau = np.zeros((1, 3))
    I = o.spatial_inertias()

    def energy(q, qd):
        M = np.linalg.inv(o.inv_mass_matrix(q)[0])
        T = o.body_transforms(q)[0]
        U = sum(-I[i, 3, 3] * np.array(desc.gravity) @ T[i, 9:] for i in range(o.nb))
        return 0.5 * qd[0] @ M @ qd[0] + U

    e0 = energy(q, qd)
    for _ in range(2000):
        q, qd = o.step(q, qd, tau)
    assert abs(energy(q, qd) - e0) < 5e-3 * max(1.0, abs(e0))


@pytest.mark.parametrize("name,build", [("cartpole", skeletons.create_cartpole), ("pendulum5", lambda: skeletons.create_multiarm_robot(5))])
def test_reference_recipe_matches_exact_jacobians(name, build):
    """The reference's Jacobians contain central differences (SURVEY finding 1); they must agree
    with exact derivatives under the reference's own rule and tolerance (1e-8)."""
    o = orc.OracleWorld(build())
    q, qd, tau = config_inputs(name, 32)
    Jr = o.jacobians(q, qd, tau, "reference")
    Je = o.jacobians(q, qd, tau, "exact")
    for k in Jr:
        assert ref_equals(Jr[k], Je[k], 1e-8), (k, max_ref_err(Jr[k], Je[k]))
    # force-vel = dt * Minv exactly (BackpropSnapshot.cpp:464-467)
    assert np.allclose(Jr["forceVel"], o.dt * o.inv_mass_matrix(q), rtol=0, atol=1e-18)


def test_cartpole_kat():
    """Config 1: the reference's CARTPOLE_15_DEG world (test_Gradients.cpp:1254-1318): q=(0, 15 deg)
    after one unconstrained velocity integration from rest."""
    desc = skeletons.create_cartpole()
    o = orc.OracleWorld(desc)
    q = np.array([[0.0, 15.0 / 180.0 * 3.1415]])
    qdd = o.forward_dynamics(q, np.zeros((1, 2)), np.zeros((1, 2)))
    qd = qdd * o.dt
    qn, qdn = o.step(q, qd, np.zeros((1, 2)))
    assert np.allclose(qn, q + o.dt * qd, atol=0)
    Jr = o.jacobians(q, qd, np.zeros((1, 2)), "reference")
    Je = o.jacobians(q, qd, np.zeros((1, 2)), "exact")
    for k in Jr:
        assert ref_equals(Jr[k], Je[k], 1e-8)
    # backprop = J^T g
    g = np.array([[1.0, -2.0]]), np.array([[0.5, 3.0]])
    _, _, gq, gqd, gtau = o.step_backprop(q, qd, np.zeros((1, 2)), g[0], g[1])
    assert np.allclose(gq[0], Jr["posPos"][0].T @ g[0][0] + Jr["posVel"][0].T @ g[1][0])
    assert np.allclose(gqd[0], Jr["velPos"][0].T @ g[0][0] + Jr["velVel"][0].T @ g[1][0])
    assert np.allclose(gtau[0], Jr["forceVel"][0].T @ g[1][0])


def test_ref_dantzig_matches_golden():
    """The reference's own Dantzig solver (compiled from /root/reference into oracle/_ref) reproduces the
    answers recorded for the literal LCP instances of unittests/unit/test_LCPUtils.cpp."""
    cases = json.load(open(os.path.join(GOLDEN, "lcp_instances.json")))
    assert len(cases) >= 8
    for c in cases:
        for early in (0, 1):
            ok, x = orc.ref_dantzig(np.array(c["A"]), c["b"], c["lo"], c["hi"], c["findex"], bool(early))
            want = c["ref_dantzig_early%d" % early]
            assert ok == want["ok"], c["name"]
            if ok:
                assert np.allclose(x, want["x"], rtol=1e-12, atol=1e-14), c["name"]
