"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): next state and Jacobians within 1e-8 under the reference's own
comparison rule (absolute for |x|<=1, relative above; unittests/TestHelpers.hpp:75-106) against the
oracle in reference mode (which reproduces the reference's finite-difference recipes); the closed-form
GPU derivatives are additionally held to 1e-10 against the exact (long double) oracle mode.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from diffdart_b200 import _abi, skeletons
from diffdart_b200.dart_layer import BackpropSnapshotPointer, dart_layer
from diffdart_b200.world import BatchedWorld, LossGradient, forwardPass
from oracle import oracle as orc
from util import config_inputs, max_ref_err, random_inputs, ref_equals

pytestmark = pytest.mark.gpu

TOL_REFERENCE = 1e-8  # north_star tolerance vs the reference recipe
TOL_EXACT = 1e-10     # closed-form derivatives vs exact derivatives
TOL_STATE = 1e-12     # next state (no finite differences involved)


def cuda(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def run_gpu(desc, q, qd, tau, gq, gqd):
    N = q.shape[0]
    world = BatchedWorld(desc, N)
    world.setPositions(cuda(q))
    world.setVelocities(cuda(qd))
    world.setForces(cuda(tau))
    snap = forwardPass(world)
    qn, qdn = world.getPositions().cpu().numpy(), world.getVelocities().cpu().numpy()
    nxt, this = LossGradient(), LossGradient()
    nxt.lossWrtPosition, nxt.lossWrtVelocity = cuda(gq), cuda(gqd)
    snap.backprop(world, this, nxt)
    J = {k: v.cpu().numpy() for k, v in snap.getJacobians().items()}
    Minv = snap.getInvMassMatrix().cpu().numpy()
    return qn, qdn, this.lossWrtPosition.cpu().numpy(), this.lossWrtVelocity.cpu().numpy(), this.lossWrtTorque.cpu().numpy(), J, Minv


def check_against_oracle(desc, q, qd, tau, seed=0, clipped=False):
    rng = np.random.default_rng(seed)
    gq, gqd = rng.normal(size=q.shape), rng.normal(size=q.shape)
    qn, qdn, dq, dqd, dtau, J, Minv = run_gpu(desc, q, qd, tau, gq, gqd)
    o = orc.OracleWorld(desc)
    oqn, oqdn = o.step(q, qd, tau)
    assert ref_equals(qn, oqn, TOL_STATE), max_ref_err(qn, oqn)
    assert ref_equals(qdn, oqdn, TOL_STATE), max_ref_err(qdn, oqdn)
    assert ref_equals(Minv, o.inv_mass_matrix(q), 1e-11)
    Jr = o.jacobians(q, qd, tau, "reference", threads=4)
    Je = o.jacobians(q, qd, tau, "exact", threads=4)
    for k in ("posPos", "posVel", "velPos", "velVel", "forceVel"):
        assert ref_equals(J[k], Jr[k], TOL_REFERENCE), (k, "vs reference recipe", max_ref_err(J[k], Jr[k]))
        assert ref_equals(J[k], Je[k], TOL_EXACT), (k, "vs exact", max_ref_err(J[k], Je[k]))
    _, _, ogq, ogqd, ogtau = o.step_backprop(q, qd, tau, gq, gqd, mode="reference", threads=4, with_step=False)
    for name, a, b in (("dL/dq", dq, ogq), ("dL/dqd", dqd, ogqd), ("dL/dtau", dtau, ogtau)):
        assert ref_equals(a, b, TOL_REFERENCE), (name, max_ref_err(a, b))
    if clipped:
        return
    # backward == J^T g with the GPU's own Jacobians (VJP consistency)
    assert np.allclose(dq, np.einsum("wrc,wr->wc", J["posPos"], gq) + np.einsum("wrc,wr->wc", J["posVel"], gqd), rtol=1e-12, atol=1e-13)
    assert np.allclose(dqd, np.einsum("wrc,wr->wc", J["velPos"], gq) + np.einsum("wrc,wr->wc", J["velVel"], gqd), rtol=1e-12, atol=1e-13)
    assert np.allclose(dtau, np.einsum("wrc,wr->wc", J["forceVel"], gqd), rtol=1e-12, atol=1e-14)


def test_cartpole_kat_single_world():
    """Config 1 — the reference's CARTPOLE_15_DEG fixture (test_Gradients.cpp:1254-1318), one world, 1-D tensors."""
    desc = skeletons.create_cartpole()
    o = orc.OracleWorld(desc)
    q = np.array([[0.0, 15.0 / 180.0 * 3.1415]])
    qd = o.forward_dynamics(q, np.zeros((1, 2)), np.zeros((1, 2))) * o.dt
    tau = np.zeros((1, 2))
    check_against_oracle(desc, q, qd, tau)
    # N = 1 with 1-D tensors reproduces the reference call shape
    world = BatchedWorld(desc, 1)
    pos, vel, trq = (cuda(a[0]).requires_grad_() for a in (q, qd, tau))
    ptr = BackpropSnapshotPointer()
    npos, nvel = dart_layer(world, pos, vel, trq, ptr)
    assert npos.shape == (2,) and ptr.backprop_snapshot is not None
    loss = (npos ** 2).sum() + (nvel ** 2).sum()
    loss.backward()
    oqn, oqdn, gq, gqd, gtau = o.step_backprop(q, qd, tau, 2 * npos.detach().cpu().numpy()[None], 2 * nvel.detach().cpu().numpy()[None])
    assert ref_equals(pos.grad.cpu().numpy(), gq[0], TOL_REFERENCE)
    assert ref_equals(vel.grad.cpu().numpy(), gqd[0], TOL_REFERENCE)
    assert ref_equals(trq.grad.cpu().numpy(), gtau[0], TOL_REFERENCE)


@pytest.mark.parametrize("N", [1, 33, 257])
def test_cartpole_batch(N):
    check_against_oracle(skeletons.create_cartpole(), *config_inputs("cartpole", N))


@pytest.mark.parametrize("N", [1, 64, 1000])
def test_pendulum5_batch(N):
    """Config 2 topology (TestHelpers.hpp:217-266), seeded as SURVEY §8(d)."""
    check_against_oracle(skeletons.create_multiarm_robot(5), *config_inputs("pendulum5", N))


@pytest.mark.parametrize("links", [1, 2, 3, 4, 6, 7, 8, 9, 12])
def test_chain_lengths(links):
    """Every chain specialisation (1..8) and the generic tree path (>8)."""
    desc = skeletons.create_multiarm_robot(links)
    check_against_oracle(desc, *random_inputs(links, 40, links), seed=links)


@pytest.mark.parametrize("links,prismatic_every", [(2, 3), (3, 3), (4, 0), (5, 3), (5, 0), (6, 2), (7, 3), (8, 0), (8, 3)])
def test_skewed_chains(links, prismatic_every):
    """The chain kernels work in canonical joint frames (joint axis = local z, chain.cuh) and come in two instantiations per length:
    revolute joints only (prismatic_every = 0) and revolute / prismatic mixed. Chains with arbitrary axes, joint frames offset and rotated
    on both sides, off-centre masses and full inertia tensors: nothing in them is aligned with the frames the kernels use."""
    from util import skewed_arm

    desc = skewed_arm(links, 100 + links, prismatic_every)
    q, qd, tau = random_inputs(links, 40, links)
    check_against_oracle(desc, q, 0.5 * qd, tau, seed=links)


def test_largest_tree():
    """DDB_MAX_BODIES bodies / DOFs: the capacity of the topology descriptor."""
    desc = skeletons.create_random_tree(48, 99)
    n = desc.getNumDofs()
    assert n == 48
    q, qd, tau = random_inputs(n, 8, 99)
    check_against_oracle(desc, q, 0.3 * qd, tau, seed=99)


@pytest.mark.parametrize("seed,nb", [(11, 3), (12, 6), (13, 10), (14, 20), (15, 40)])
def test_random_trees(seed, nb):
    """Branching trees with revolute + prismatic joints, skewed axes/offsets/COMs, 3-D gravity."""
    desc = skeletons.create_random_tree(nb, seed)
    n = desc.getNumDofs()
    q, qd, tau = random_inputs(n, 24, seed)
    check_against_oracle(desc, q, 0.5 * qd, tau, seed=seed)


def test_weld_joint_in_tree():
    desc = skeletons.WorldDescription().setGravity([0.0, -9.81, 0.0])
    s = skeletons.Skeleton("s")
    a = s.createRevoluteJointAndBodyNodePair(None).setAxis([0, 0, 1])
    w = s.createWeldJointAndBodyNodePair(a).setTransformFromParentBodyNode(p=[0.2, 0.5, 0.0]).setMass(2.0)
    b = s.createRevoluteJointAndBodyNodePair(w).setAxis([1, 0, 0]).setTransformFromParentBodyNode(p=[0.0, 0.3, 0.1])
    b.setLocalCOM([0.0, 0.2, 0.0])
    desc.addSkeleton(s)
    check_against_oracle(desc, *random_inputs(2, 16, 77))


def test_two_skeletons_block_diagonal():
    desc = skeletons.create_multiarm_robot(2)
    other = skeletons.Skeleton("other")
    other.createPrismaticJointAndBodyNodePair(None).setAxis([0, 1, 0]).setMass(3.0)
    desc.addSkeleton(other)
    q, qd, tau = random_inputs(3, 8, 5)
    check_against_oracle(desc, q, qd, tau)
    _, _, _, _, _, J, Minv = run_gpu(desc, q, qd, tau, q, q)
    assert np.all(Minv[:, 2, :2] == 0) and np.all(Minv[:, :2, 2] == 0)


def test_clip_loss_gradients_to_bounds():
    """clipLossGradientsToBounds (BackpropSnapshot.cpp:374-428): exact-equality tests at the limits."""
    desc = skeletons.create_multiarm_robot(3)
    bodies = desc.skeletons[0].bodies
    bodies[0].position_lower[:] = -0.5
    bodies[1].velocity_upper[:] = 0.25
    bodies[2].force_lower[:] = -1.0
    q, qd, tau = random_inputs(3, 12, 9)
    q[:6, 0] = -0.5
    qd[3:9, 1] = 0.25
    tau[::2, 2] = -1.0
    check_against_oracle(desc, q, qd, tau, seed=3, clipped=True)
    rng = np.random.default_rng(3)
    gq, gqd = rng.normal(size=q.shape), rng.normal(size=q.shape)
    _, _, dq, dqd, dtau, _, _ = run_gpu(desc, q, qd, tau, gq, gqd)
    assert np.all(dq[:6, 0] <= 0) and np.all(dqd[3:9, 1] >= 0) and np.all(dtau[::2, 2] <= 0)


def test_dof_major_layout_and_host_entry_points():
    """The same step through the raw C ABI in the structure-of-arrays layout and through the
    host-buffer entry points."""
    lib = _abi.load_library()
    desc = skeletons.create_multiarm_robot(5)
    N, n = 100, 5
    q, qd, tau = config_inputs("pendulum5", N)
    o = orc.OracleWorld(desc)
    oqn, oqdn = o.step(q, qd, tau)
    rng = np.random.default_rng(1)
    gq, gqd = rng.normal(size=q.shape), rng.normal(size=q.shape)
    _, _, ogq, ogqd, ogtau = o.step_backprop(q, qd, tau, gq, gqd, mode="exact", with_step=False)
    for layout in (_abi.LAYOUT_DOF_MAJOR, _abi.LAYOUT_WORLD_MAJOR):
        topo, params, keep = desc.to_ctypes(layout=layout)
        h = _abi._vp()
        assert lib.ddb_create(C.byref(topo), C.byref(params), N, 0, C.byref(h)) == 0
        lay = (lambda a: np.ascontiguousarray(a.T)) if layout == _abi.LAYOUT_DOF_MAJOR else np.ascontiguousarray
        unlay = (lambda a: a.reshape(n, N).T) if layout == _abi.LAYOUT_DOF_MAJOR else (lambda a: a.reshape(N, n))
        hq, hqd, htau, hgq, hgqd = (lay(a) for a in (q, qd, tau, gq, gqd))
        outs = [np.empty(N * n) for _ in range(5)]
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        assert lib.ddb_forward_host(h, p(hq), p(hqd), p(htau), p(outs[0]), p(outs[1])) == 0
        assert lib.ddb_backward_host(h, p(hgq), p(hgqd), p(outs[2]), p(outs[3]), p(outs[4])) == 0
        assert ref_equals(unlay(outs[0]), oqn, TOL_STATE) and ref_equals(unlay(outs[1]), oqdn, TOL_STATE)
        assert ref_equals(unlay(outs[2]), ogq, TOL_EXACT) and ref_equals(unlay(outs[3]), ogqd, TOL_EXACT)
        assert ref_equals(unlay(outs[4]), ogtau, TOL_EXACT)
        vv = np.empty((N, n, n))
        assert lib.ddb_jacobians_host(h, None, None, None, p(vv), None) == 0
        assert ref_equals(vv.transpose(0, 2, 1), o.jacobians(q, qd, tau, "exact")["velVel"], TOL_EXACT)
        assert lib.ddb_launch_count(h) == 3
        assert lib.ddb_destroy(h) == 0


def test_error_paths():
    lib = _abi.load_library()
    desc = skeletons.create_cartpole()
    topo, params, keep = desc.to_ctypes()
    h = _abi._vp()
    assert lib.ddb_create(C.byref(topo), C.byref(params), 8, 0, C.byref(h)) == 0
    g = torch.zeros(8, 2, dtype=torch.float64, device="cuda")
    # backward before any forward: DDB_ERR_STATE, with a message
    assert lib.ddb_backward(h, None, g.data_ptr(), g.data_ptr(), g.data_ptr(), g.data_ptr(), g.data_ptr()) == 4
    assert b"forward" in lib.ddb_last_error(h)
    assert lib.ddb_snapshot_select(h, 3) == 1
    assert lib.ddb_forward(h, None, None, g.data_ptr(), g.data_ptr(), g.data_ptr(), g.data_ptr(), None, None, None, None) == 1
    lib.ddb_destroy(h)
    # a handle that steps with the constraint solver's gradients disabled (plain World::step, ddb_params.reserved[1] = 1) keeps no
    # classification of its contact rows: the backward through such a step is refused with a message, never approximated
    stack = skeletons.create_box_stack(2)
    topo, params, keep = stack.to_ctypes()
    params.reserved[1] = 1
    h = _abi._vp()
    assert lib.ddb_create(C.byref(topo), C.byref(params), 8, 0, C.byref(h)) == 0
    g6 = torch.zeros(8, 12, dtype=torch.float64, device="cuda")
    assert lib.ddb_forward(h, None, g6.data_ptr(), g6.data_ptr(), g6.data_ptr(), g6.clone().data_ptr(), g6.clone().data_ptr(), None, None, None, None) == 0
    assert lib.ddb_backward(h, None, g6.data_ptr(), g6.data_ptr(), g6.data_ptr(), g6.data_ptr(), g6.data_ptr()) == 4
    assert b"gradients disabled" in lib.ddb_last_error(h)
    assert lib.ddb_jacobians(h, None, g6.data_ptr(), None, None, None, None) == 4
    # rollouts: the reverse sweep needs the forward sweep
    assert lib.ddb_rollout_backward(h, None, 2, g6.data_ptr(), g6.data_ptr(), g6.data_ptr(), g6.data_ptr(), g6.data_ptr()) == 4
    lib.ddb_destroy(h)


def test_rollout_with_snapshot_slots():
    """T-step rollout through dart_layer with autograd (the loop shape of SingleShot::getSnapshots,
    dart/trajectory/SingleShot.cpp:650-760) against the oracle's chained backprop."""
    desc = skeletons.create_multiarm_robot(3)
    N, n, T = 16, 3, 12
    q0, qd0, _ = random_inputs(n, N, 21)
    rng = np.random.default_rng(22)
    taus = rng.uniform(-1, 1, (T, N, n))
    world = BatchedWorld(desc, N, n_snapshot_slots=2)  # grows on demand
    pos, vel = cuda(q0).requires_grad_(), cuda(qd0).requires_grad_()
    tq = [cuda(t).requires_grad_() for t in taus]
    p, v = pos, vel
    for t in range(T):
        p, v = dart_layer(world, p, v, tq[t])
    loss = (p ** 2).sum() + (v ** 2).sum()
    loss.backward()
    o = orc.OracleWorld(desc)
    qs, qds = [q0], [qd0]
    for t in range(T):
        a, b = o.step(qs[-1], qds[-1], taus[t])
        qs.append(a)
        qds.append(b)
    assert ref_equals(p.detach().cpu().numpy(), qs[-1], 1e-11)
    gq, gqd = 2 * qs[-1], 2 * qds[-1]
    gtaus = [None] * T
    for t in reversed(range(T)):
        _, _, gq, gqd, gtaus[t] = o.step_backprop(qs[t], qds[t], taus[t], gq, gqd, mode="exact", with_step=False)
    assert ref_equals(pos.grad.cpu().numpy(), gq, 1e-9)
    assert ref_equals(vel.grad.cpu().numpy(), gqd, 1e-9)
    for t in range(T):
        assert ref_equals(tq[t].grad.cpu().numpy(), gtaus[t], 1e-9)


def test_full_size_properties():
    """BASELINE config 2 at full size (65536 worlds): size-independent properties —
    (a) a strided sample agrees with the oracle; (b) backward == J^T g for every world;
    (c) force-vel is symmetric positive definite (dt * M^-1); (d) energy drift of the step is O(dt^2)."""
    desc = skeletons.create_multiarm_robot(5)
    N = 65536
    q, qd, tau = config_inputs("pendulum5", N)
    rng = np.random.default_rng(5)
    gq, gqd = rng.normal(size=q.shape), rng.normal(size=q.shape)
    qn, qdn, dq, dqd, dtau, J, Minv = run_gpu(desc, q, qd, tau, gq, gqd)
    o = orc.OracleWorld(desc)
    idx = np.arange(0, N, 257)
    oqn, oqdn, ogq, ogqd, ogtau = o.step_backprop(q[idx], qd[idx], tau[idx], gq[idx], gqd[idx], mode="reference", threads=4)
    assert ref_equals(qn[idx], oqn, TOL_STATE) and ref_equals(qdn[idx], oqdn, TOL_STATE)
    assert ref_equals(dq[idx], ogq, TOL_REFERENCE) and ref_equals(dqd[idx], ogqd, TOL_REFERENCE) and ref_equals(dtau[idx], ogtau, TOL_REFERENCE)
    assert np.allclose(dq, gq + np.einsum("wrc,wr->wc", J["posVel"], gqd), rtol=1e-12, atol=1e-13)
    assert np.allclose(dqd, o.dt * gq + np.einsum("wrc,wr->wc", J["velVel"], gqd), rtol=1e-12, atol=1e-13)
    fv = J["forceVel"]
    assert np.allclose(fv, fv.transpose(0, 2, 1), rtol=0, atol=1e-16)
    assert np.all(np.linalg.eigvalsh(fv) > 0)
    assert np.allclose(fv, o.dt * Minv, rtol=1e-15, atol=1e-18)
    assert np.all(np.isfinite(qn)) and np.all(np.isfinite(J["posVel"]))


@pytest.mark.parametrize("seed,nb", [(21, 1), (22, 4), (23, 9)])
def test_free_joint_forward(seed, nb):
    """FreeJoint root (FreeJoint.cpp:64-81,587-601,673-693) + revolute/prismatic children, no contact: next state vs the
    oracle, incl. the exp/log-map position integration at small and large rotation vectors."""
    desc = skeletons.create_random_tree(nb, seed, free_root=True)
    n = desc.getNumDofs()
    q, qd, tau = random_inputs(n, 48, seed)
    q[:8, 0:3] *= 1e-4   # below expMapRot's 1e-3 Taylor switch
    q[8:16, 0:3] *= 2.5  # large rotation vectors
    world = BatchedWorld(desc, 48)
    world.setPositions(cuda(q)); world.setVelocities(cuda(qd)); world.setForces(cuda(tau))
    forwardPass(world)
    o = orc.OracleWorld(desc)
    oqn, oqdn = o.step(q, qd, tau)
    assert ref_equals(world.getVelocities().cpu().numpy(), oqdn, 1e-11), max_ref_err(world.getVelocities().cpu().numpy(), oqdn)
    assert ref_equals(world.getPositions().cpu().numpy(), oqn, 1e-11), max_ref_err(world.getPositions().cpu().numpy(), oqn)


@pytest.mark.parametrize("seed,nb", [(21, 1), (22, 4), (23, 9)])
def test_free_joint_backward(seed, nb):
    """Closed-form backward through a FreeJoint root: pos-pos / vel-pos blocks of the exp/log-map integration (the
    reference finite-differences them, FreeJoint.cpp:604-653), 6-DOF blocks in the M^-1 solve and the ID adjoint.
    Bar: 1e-8 (north star) against both oracle modes; the residual ~1e-9 comes from the reference's second-order
    Taylor branch of expMapRot below 1e-3 (Geometry.cpp:479-483), which dt*omega always falls into."""
    desc = skeletons.create_random_tree(nb, seed, free_root=True)
    n = desc.getNumDofs()
    q, qd, tau = random_inputs(n, 40, seed)
    q[:8, 0:3] *= 1e-4
    q[8:16, 0:3] *= 2.5
    rng = np.random.default_rng(seed)
    gq, gqd = rng.normal(size=q.shape), rng.normal(size=q.shape)
    qn, qdn, dq, dqd, dtau, J, Minv = run_gpu(desc, q, qd, tau, gq, gqd)
    o = orc.OracleWorld(desc)
    assert ref_equals(Minv, o.inv_mass_matrix(q), 1e-11)
    Je = o.jacobians(q, qd, tau, "exact", threads=4)
    Jr = o.jacobians(q, qd, tau, "reference", threads=4)
    for k in Je:
        assert ref_equals(J[k], Je[k], 1e-8), (k, max_ref_err(J[k], Je[k]))
        assert ref_equals(J[k], Jr[k], 1e-8), (k, max_ref_err(J[k], Jr[k]))
    for k in ("posVel", "velVel", "forceVel"):  # no exp/log map involved: tight
        assert ref_equals(J[k], Je[k], 1e-10), (k, max_ref_err(J[k], Je[k]))
    _, _, ogq, ogqd, ogtau = o.step_backprop(q, qd, tau, gq, gqd, mode="exact", threads=4, with_step=False)
    assert ref_equals(dq, ogq, 1e-8) and ref_equals(dqd, ogqd, 1e-8) and ref_equals(dtau, ogtau, 1e-11)
    assert np.allclose(dq, np.einsum("wrc,wr->wc", J["posPos"], gq) + np.einsum("wrc,wr->wc", J["posVel"], gqd), rtol=1e-12, atol=1e-13)
    assert np.allclose(dqd, np.einsum("wrc,wr->wc", J["velPos"], gq) + np.einsum("wrc,wr->wc", J["velVel"], gqd), rtol=1e-12, atol=1e-13)
