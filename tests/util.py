"""Shared helpers for the tests."""
import numpy as np


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
