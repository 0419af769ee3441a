"""Seeded boxed-LCP instances shaped like contact problems (A = J M^-1 J^T, 3 rows per contact:
normal [0, inf) + two friction rows bounded by mu * x[normal]), incl. duplicated contacts (rank
deficiency, as four coplanar contacts under a box give) and frictionless contacts."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def golden_instances():
    cases = json.load(open(os.path.join(GOLDEN, "lcp_instances.json")))
    for c in cases:
        c["A"] = np.array(c["A"], dtype=np.float64)
    return cases


def random_contact_lcp(rng, nc=None, nd=None):
    nc = int(rng.integers(1, 9)) if nc is None else nc
    nd = int(rng.integers(3, 13)) if nd is None else nd
    n = 3 * nc
    J = rng.normal(size=(n, nd))
    if rng.uniform() < 0.5 and nc > 1:
        J[3:6] = J[0:3] + (0 if rng.uniform() < 0.5 else 1e-9 * rng.normal(size=(3, nd)))
    M = rng.normal(size=(nd, nd))
    Minv = np.linalg.inv(M @ M.T + np.eye(nd))
    A = J @ Minv @ J.T
    A = (A + A.T) / 2
    b = rng.normal(size=n)
    lo, hi, fi = np.zeros(n), np.zeros(n), -np.ones(n, dtype=np.int32)
    for k in range(nc):
        hi[3 * k] = np.inf
        mu = rng.choice([0.0, 0.5, 1.0])
        for t in (1, 2):
            lo[3 * k + t], hi[3 * k + t], fi[3 * k + t] = -mu, mu, 3 * k
    return A, b, lo, hi, fi


def pad_batch(problems, n_max):
    """Stack ragged problems into [N, n_max(, n_max)] arrays + n_rows."""
    N = len(problems)
    A = np.zeros((N, n_max, n_max))
    b, lo, hi = np.zeros((N, n_max)), np.zeros((N, n_max)), np.zeros((N, n_max))
    fi = -np.ones((N, n_max), dtype=np.int32)
    rows = np.zeros(N, dtype=np.int32)
    for k, (a, bb, l, h, f) in enumerate(problems):
        n = len(bb)
        A[k, :n, :n], b[k, :n], lo[k, :n], hi[k, :n], fi[k, :n], rows[k] = a, bb, l, h, f, n
    return A, b, lo, hi, fi, rows
