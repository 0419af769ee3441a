"""GPU: World::step with contacts (collision -> contact rows -> LCP by impulse tests -> Dantzig/PGS chain -> impulses)
through ddb_forward, against the oracle (restatement + the reference's own dBoxBox and Dantzig) world by world."""
import ctypes as C

import numpy as np
import pytest
import torch

from contact_scenes import GPU_SCENES as SCENES  # + Atlas on the ground when baseline/_ref/atlas holds the reference's URDFs
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
        # worlds solved by PGS (tier bits 4 / 8) agree to the north-star tolerance for contact-rich solves: the GPU's Gauss-Seidel sums
        # a row in another order than the scalar solver, and its stopping rule is relative (1e-8)
        tol = (1e-6, 1e-8) if r["tier_mask"] & 12 else (1e-9, 1e-11)
        assert np.allclose(qdn[w], r["qd_next"], rtol=tol[0], atol=tol[1]), w
        assert np.allclose(qn[w], r["q_next"], rtol=0, atol=1e-14), w
        if w % 37 == 0 and len(r["x"]):
            _, contacts, x = contact_debug(world, w)
            assert np.allclose(contacts[:len(r["contacts"]), 0:3], r["contacts"][:, 2:5], atol=1e-13)
            assert np.allclose(x[:len(r["x"])], r["x"], rtol=1e-8, atol=1e-11)


def test_scalar_pgs_kernel_for_large_problems(monkeypatch):
    """Problems above the warp solver's 192 rows go to the scalar k_pgs_thread; DDB_PGS_WARP_MAX lowers that limit so that the box stack's
    60 ... 108-row problems take that path (and the classes below the limit the warp path, in the same launch)."""
    monkeypatch.setenv("DDB_PGS_WARP_MAX", "80")
    make, states = SCENES["box_stack"]
    desc = make()
    N = 48
    q, qd, tau = states(N, 9)
    world = BatchedWorld(PlainStepDescription(desc), N)
    world.setPositions(cuda(q)); world.setVelocities(cuda(qd)); world.setForces(cuda(tau))
    forwardPass(world)
    qdn = world.getVelocities().cpu().numpy()
    counts, _, _ = contact_debug(world, -1)
    rows = np.zeros(N, dtype=np.int32)
    assert world._lib.ddb_contact_pgs_rows_host(world._h, rows.ctypes.data_as(C.c_void_p)) == 0
    assert (rows > 80).any() and ((rows > 0) & (rows <= 80)).any(), rows   # both paths were taken
    o = OracleWorld(desc)
    for w in range(N):
        r = o.step_contacts(q[w], qd[w], tau[w])
        assert list(counts[w]) == [len(r["contacts"]), len(r["x"]), r["n_groups"], r["tier_mask"]], w
        tol = (1e-6, 1e-8) if r["tier_mask"] & 12 else (1e-9, 1e-11)
        assert np.allclose(qdn[w], r["qd_next"], rtol=tol[0], atol=tol[1]), w


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
    tiers, indeterminate, prev = set(), set(), {}
    for t in range(T):
        world.setForces(cuda(tau))
        forwardPass(world)
        qn, qdn = world.getPositions().cpu().numpy(), world.getVelocities().cpu().numpy()
        active = world.last_active_set.cpu().numpy()
        status = world.last_solver_status.cpu().numpy()
        counts, _, _ = contact_debug(world, -1)
        for w in range(N):
            if w in indeterminate:
                continue
            r = o.step_contacts(ref[w][0], ref[w][1], tau[w], lcp_cache=ref[w][2], gradients=True)
            prev[w] = ref[w]
            ref[w] = (r["q_next"], r["qd_next"], r["lcp_cache"])
            k = len(r["x"])
            tiers.add(r["tier_mask"])
            assert list(counts[w]) == [len(r["contacts"]), k, r["n_groups"], r["tier_mask"]], (t, w)
            assert status[w] == r["tier_mask"]
            if scene == "walker_on_ground" and r["tier_mask"] & 12 and not np.array_equal(active[w, :k], r["active_set"]):
                # KNOWN GAP (point 1): the planar walker's foot boxes touch with pairs of contact points whose LCP columns
                # are duplicates; in a PGS world the impulse of such a pair is fixed only up to the null space of A, and the GPU and
                # the scalar solver may hand it to the pair differently (all on one point / half each). The step itself is the same:
                # require A (x_gpu - x_cpu) = 0 and equal next velocities, stop following this world (its cached solution now differs).
                x_gpu = contact_debug(world, w)[2][:k]
                assert np.max(np.abs(r["A"] @ (x_gpu - r["x"]))) < 1e-9, (t, w)
                assert np.allclose(qdn[w], r["qd_next"], rtol=1e-9, atol=1e-11), (t, w)
                # ... and the GPU's answer is one of the oracle's own: at such a state the oracle's active set flips between the two
                # under one-ulp noise on its inputs (the rank decision of the least-squares standardisation sits on its threshold)
                prng = np.random.default_rng(w)
                q_in, qd_in = prev[w][0], prev[w][1]  # the state this step started from
                found = False
                for _ in range(600):
                    rp = o.step_contacts(q_in * (1 + prng.uniform(-1, 1, q_in.size) * 2e-16), qd_in * (1 + prng.uniform(-1, 1, q_in.size) * 2e-16),
                                         tau[w], lcp_cache=prev[w][2], gradients=True)
                    if len(rp["x"]) == k and np.array_equal(active[w, :k], rp["active_set"]):
                        found = True
                        break
                assert found, (t, w, "the GPU's active set is not among the oracle's answers under one-ulp input noise")
                indeterminate.add(w)
                continue
            assert np.array_equal(active[w, :k], r["active_set"]) and np.all(active[w, k:] == -4), (t, w)
            tol = 1e-6 if r["tier_mask"] & 12 else 1e-9
            assert np.allclose(qdn[w], r["qd_next"], rtol=tol, atol=tol), (t, w)
            assert np.allclose(qn[w], r["q_next"], rtol=0, atol=1e-9), (t, w)  # q' inherits dt * (qd differences of earlier steps)
    if scene in ("sliding_box", "two_box_stack"):
        assert 1 in tiers
    assert len(indeterminate) <= max(1, N // 16), sorted(indeterminate)


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))  # the reference's `equals`: abs below 1, relative above


def all_multipliers(world):
    """LCP multipliers of every world right after forwardPass (the backward reuses the same scratch)."""
    return [contact_debug(world, w)[2] for w in range(world.N)]


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
def test_contact_backward_matches_oracle(scene):
    """BackpropSnapshot::backprop through the contact solution (ddb_backward) against the oracle's restatement of the
    reference Jacobians applied to the same loss gradients, 1e-8 (the reference's own FD-vs-analytic bar). The multipliers of
    a rank-deficient LCP are fixed only up to the null space of A, so the oracle evaluates its Jacobians at the
    multipliers of the GPU's forward pass."""
    from diffdart_b200.world import LossGradient
    make, states = SCENES[scene]
    desc = make()
    N = 16 if scene == "box_stack" else (12 if scene == "atlas_on_ground" else 48)
    q, qd, tau = states(N, 21)
    n = q.shape[1]
    world = BatchedWorld(desc, N)
    world.setPositions(cuda(q)); world.setVelocities(cuda(qd)); world.setForces(cuda(tau))
    snap = forwardPass(world)
    qdn = world.getVelocities().cpu().numpy()
    xs = all_multipliers(world)
    active = world.last_active_set.cpu().numpy()
    counts, _, _ = contact_debug(world, -1)
    rng = np.random.default_rng(3)
    gqn, gqdn = rng.normal(size=(N, n)), rng.normal(size=(N, n))
    nxt, this = LossGradient(), LossGradient()
    nxt.lossWrtPosition, nxt.lossWrtVelocity = cuda(gqn), cuda(gqdn)
    snap.backprop(world, this, nxt)
    gq, gqd, gtau = (t.cpu().numpy() for t in (this.lossWrtPosition, this.lossWrtVelocity, this.lossWrtTorque))
    o = OracleWorld(desc)
    clamping = 0
    for w in range(N):
        k = max(int(counts[w, 1]), 0)
        J = vjp_check(o, q[w], qd[w], tau[w], xs[w][:k], active[w, :k], gqn[w], gqdn[w], gq[w], gqd[w], gtau[w], (scene, w))
        clamping += J["num_clamping"] > 0
        assert np.allclose(qdn[w], J["qd_next"], rtol=1e-6, atol=1e-6), w
    assert clamping > 0


@pytest.mark.parametrize("scene", ["free_box_on_floor", "pendulum_on_floor", "two_box_stack", "fixed_base_arm_on_floor", "walker_on_ground"])
def test_contact_jacobians_match_oracle(scene):
    """BackpropSnapshot::get{PosPos,PosVel,VelPos,VelVel,ForceVel}Jacobian with contacts (ddb_jacobians). walker_on_ground carries
    finite joint limits: the getters do not clip (only backprop() does, BackpropSnapshot.cpp:374-428) and must not touch the limits."""
    make, states = SCENES[scene]
    desc = make()
    N = 12
    q, qd, tau = states(N, 8)
    world = BatchedWorld(desc, N)
    world.setPositions(cuda(q)); world.setVelocities(cuda(qd)); world.setForces(cuda(tau))
    snap = forwardPass(world)
    xs = all_multipliers(world)
    active = world.last_active_set.cpu().numpy()
    counts, _, _ = contact_debug(world, -1)
    Jg = {k: v.cpu().numpy() for k, v in snap.getJacobians().items()}
    o = OracleWorld(desc)
    for w in range(N):
        k = max(int(counts[w, 1]), 0)
        J = o.contact_jacobians(q[w], qd[w], tau[w], mode="reference", x_override=xs[w][:k], mapping_override=active[w, :k])
        Jn = o.contact_jacobians(q[w], qd[w], tau[w], mode="reference", x_override=xs[w][:k], mapping_override=active[w, :k], q_noise=4e-16)
        for name in ("posPos", "posVel", "velPos", "velVel", "forceVel"):
            assert rel(Jg[name][w], J[name]) < 1e-8 + 20.0 * rel(Jn[name], J[name]), (scene, w, name)


@pytest.mark.parametrize("scene", ["two_box_stack", "walker_on_ground"])
def test_idempotent_forward_pass_leaves_the_cached_lcp_solution(scene):
    """forwardPass(world, idempotent=True) restores the constraint solver's cached solution (RestorableSnapshot.cpp:23,32 via
    NeuralUtils.cpp:29-33,61-62): with an idempotent call in between, the following steps are bit-identical to those without it."""
    make, states = SCENES[scene]
    desc = make()
    N = 64
    q, qd, tau = states(N, 17)

    def run(with_idempotent):
        world = BatchedWorld(desc, N)
        world.setPositions(cuda(q)); world.setVelocities(cuda(qd))
        out = []
        for t in range(3):
            world.setForces(cuda(tau))
            if with_idempotent and t == 1:
                q0, qd0 = world.getPositions().clone(), world.getVelocities().clone()
                # a different state, so that the idempotent step's own solution differs from the cached one
                world.setVelocities(world.getVelocities() * 0.5)
                snap = forwardPass(world, idempotent=True)
                assert torch.equal(world.getPositions(), q0) and torch.equal(world.getExternalForces(), cuda(tau))
                assert not torch.equal(snap.getPostStepVelocity(), world.getVelocities())
                world.setVelocities(qd0)
            forwardPass(world)
            out.append((world.getPositions().clone(), world.getVelocities().clone(), world.last_active_set.clone(), world.last_solver_status.clone()))
        return out

    for a, b in zip(run(False), run(True)):
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_dart_layer_autograd_through_contacts():
    """dart_layer(world, pos, vel, torque) with contacts: torch autograd pulls the loss back through ddb_backward; checked
    against central differences of the layer itself on worlds whose active set is stable under the probe."""
    from diffdart_b200.dart_layer import dart_layer
    make, states = SCENES["pendulum_on_floor"]
    desc = make()
    N = 32
    q, qd, tau = states(N, 4)
    world = BatchedWorld(desc, N)
    tq, tqd, tt = (cuda(a).requires_grad_(True) for a in (q, qd, tau))
    qn, qdn = dart_layer(world, tq, tqd, tt)
    rng = np.random.default_rng(0)
    cq, cqd = cuda(rng.normal(size=q.shape)), cuda(rng.normal(size=q.shape))
    ((qn * cq).sum() + (qdn * cqd).sum()).backward()
    active0 = world.last_active_set.cpu().numpy().copy()
    g = tt.grad.cpu().numpy()

    def loss(tau_np):
        a, b = dart_layer(world, cuda(q), cuda(qd), cuda(tau_np))
        return ((a * cq).sum(dim=1) + (b * cqd).sum(dim=1)).detach().cpu().numpy(), world.last_active_set.cpu().numpy().copy()

    eps = 1e-5
    checked = 0
    for d in range(q.shape[1]):
        tp, tm = tau.copy(), tau.copy()
        tp[:, d] += eps; tm[:, d] -= eps
        (lp, ap), (lm, am) = loss(tp), loss(tm)
        stable = np.all(ap == active0, axis=1) & np.all(am == active0, axis=1)
        fd = (lp - lm) / (2 * eps)
        checked += int(stable.sum())
        assert np.allclose(g[stable, d], fd[stable], rtol=1e-5, atol=1e-7)
    assert checked > N


def test_restitution_bounces_and_is_differentiated():
    """Restitution > 0: contacts closing faster than the bounce threshold leave with -e times their normal velocity, and the backward
    carries the bounce diagonals (covered against the oracle by the bouncing_boxes scene of the parametrised tests above; here the
    physics itself and central differences of the layer)."""
    from contact_scenes import bouncing_boxes
    from diffdart_b200.dart_layer import dart_layer
    desc = bouncing_boxes()
    q = np.array([[-2.0, 0.0, 0.499, 2.0, 0.0, 0.249]]); qd = np.array([[0.0, 0.0, -1.0, 0.0, 0.0, -0.05]]); tau = np.zeros((1, 6))
    world = BatchedWorld(desc, 1)
    tq, tqd, tt = (cuda(a).requires_grad_(True) for a in (q, qd, tau))
    qn, qdn = dart_layer(world, tq, tqd, tt)
    v = qdn.detach().cpu().numpy()[0]
    assert abs(v[2] - 0.8 * (1.0 + 9.81e-3)) < 1e-6      # bounced: -e times the pre-contact normal velocity (gravity acted for one step)
    assert abs(v[5]) < 1e-9                               # too slow to bounce: comes to rest
    qdn[0, 2].backward()
    g = tqd.grad.cpu().numpy()[0]
    assert abs(g[2] + 0.8) < 1e-6 and abs(g[5]) < 1e-12   # d v_z' / d v_z = -e through the bounce diagonal
