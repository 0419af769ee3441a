"""Shared helpers for the tests."""
import numpy as np

from diffdart_b200 import skeletons


def _exp_map_rot(w):
    """Rodrigues rotation of the vector w (angle = |w|)."""
    th = float(np.linalg.norm(w))
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / (th * th) * (K @ K)


def ref_equals(a, b, tol=1e-8):
    """The reference's own comparison rule (unittests/TestHelpers.hpp:75-106):
    absolute for |x| <= 1, relative above."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    return bool(np.all(np.abs(a - b) <= tol * scale))


def max_ref_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def config_inputs(name, N, seed=None):
    """Seeded synthetic inputs of SURVEY.md §8(d)."""
    if name == "cartpole":
        rng = np.random.default_rng(0 if seed is None else seed)
        q = np.stack([rng.uniform(-1, 1, N), rng.uniform(-np.pi, np.pi, N)], axis=1)
        qd = rng.uniform(-1, 1, (N, 2))
        tau = np.stack([rng.uniform(-10, 10, N), np.zeros(N)], axis=1)
        return q, qd, tau
    if name == "pendulum5":
        rng = np.random.default_rng(1234 if seed is None else seed)
        return rng.uniform(-1, 1, (N, 5)), rng.uniform(-1, 1, (N, 5)), rng.uniform(-1, 1, (N, 5))
    raise KeyError(name)


def random_inputs(n, N, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, (N, n)), rng.uniform(-1, 1, (N, n)), rng.uniform(-2, 2, (N, n))


def skewed_arm(n, seed, prismatic_every=3):
    """A chain with arbitrary joint axes, joint frames offset and rotated on both sides, off-centre masses and full inertia tensors,
    revolute and prismatic joints mixed: nothing in it is aligned with the canonical frames the kernels work in."""
    rng = np.random.default_rng(seed)
    s = skeletons.Skeleton("skewed")
    parent = None
    for i in range(n):
        prismatic = prismatic_every > 0 and (i % prismatic_every == 1)
        b = (s.createPrismaticJointAndBodyNodePair if prismatic else s.createRevoluteJointAndBodyNodePair)(parent)
        axis = rng.normal(size=3)
        b.setAxis(list(axis / np.linalg.norm(axis)))
        Rp, Rc = _exp_map_rot(rng.normal(size=3) * 0.7), _exp_map_rot(rng.normal(size=3) * 0.7)
        b.setTransformFromParentBodyNode(p=list(rng.uniform(-0.5, 0.5, 3)), R=Rp)
        b.setTransformFromChildBodyNode(p=list(rng.uniform(-0.3, 0.3, 3)), R=Rc)
        b.setMass(float(rng.uniform(0.5, 2.0))).setLocalCOM(list(rng.uniform(-0.2, 0.2, 3)))
        A = rng.normal(size=(3, 3))
        J = A @ A.T + 0.5 * np.eye(3)
        b.setMomentOfInertia(J[0, 0], J[1, 1], J[2, 2], J[0, 1], J[0, 2], J[1, 2])
        parent = b
    return skeletons.WorldDescription().addSkeleton(s).setGravity([0.3, -9.81, 1.1])
