"""Lemke solver: the numpy restatement (oracle/lemke.py) is pinned on the reference's own instances
(unittests/unit/test_Lemke.cpp:39-185: err == 0 and validate()), the CUDA solver (ddb_lemke_solve) is compared with the
restatement: same return code, same number of pivots, solution within 1e-9, on the reference's instances and on seeded random
problems (ragged sizes, trivial problems, problems without a solution)."""
import json
import os

import numpy as np
import pytest

from oracle.lemke import lemke, validate

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lemke_instances.json")


def golden_cases():
    return [(c["name"], np.array(c["A"]).reshape(c["n"], c["n"]), np.array(c["b"])) for c in json.load(open(GOLDEN))]


def random_cases(seed, count, n_lo, n_hi):
    """Contact-like LCPs: M = J J^T + eps I (symmetric positive definite, Lemke terminates with a solution), q with mixed signs."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        n = int(rng.integers(n_lo, n_hi + 1))
        J = rng.normal(size=(n, max(2, n // 2 + 1)))
        M = J @ J.T + 1e-3 * np.eye(n)
        q = rng.normal(size=n) - 0.2
        out.append((M, q))
    return out


@pytest.mark.parametrize("name,A,b", golden_cases(), ids=[c[0] for c in golden_cases()])
def test_oracle_on_reference_instances(name, A, b):
    err, z, pivots = lemke(A, b)
    assert err == 0
    assert validate(A, z, b)
    w = A @ z + b
    assert np.all(z >= -1e-4) and np.all(w >= -1e-4) and np.all(np.abs(z * w) <= 1e-4)


def test_oracle_special_cases():
    # q >= 0: trivial solution without a pivot (Lemke.cpp:70-75)
    err, z, piv = lemke(np.eye(3), np.array([0.0, 1.0, 2.0]))
    assert err == 0 and not piv and np.all(z == 0)
    # no solution (M = 0 cannot lift a negative q): ray termination, zero vector (Lemke.cpp:201-205,318-321)
    err, z, piv = lemke(np.zeros((2, 2)), np.array([-1.0, 1.0]))
    assert err == 2 and np.all(z == 0)
    # random positive definite problems solve and validate
    for M, q in random_cases(3, 50, 1, 24):
        err, z, piv = lemke(M, q)
        assert err == 0 and validate(M, z, q)


@pytest.mark.gpu
def test_device_matches_oracle():
    import torch

    from diffdart_b200.lcp import lemke_solve

    cases = [(A, b) for _, A, b in golden_cases()] + random_cases(11, 300, 1, 40)
    cases += [(np.eye(3), np.array([0.0, 1.0, 2.0])), (np.zeros((2, 2)), np.array([-1.0, 1.0])), (np.zeros((4, 4)), -np.ones(4))]
    P, n_max = len(cases), max(len(q) for _, q in cases)
    M = np.zeros((P, n_max, n_max)); q = np.zeros((P, n_max)); rows = np.zeros(P, dtype=np.int32)
    for i, (A, b) in enumerate(cases):
        n = len(b)
        M[i, :n, :n] = A; q[i, :n] = b; rows[i] = n
    z, err, piv = lemke_solve(torch.tensor(M, device="cuda"), torch.tensor(q, device="cuda"), torch.tensor(rows, device="cuda"))
    z, err, piv = z.cpu().numpy(), err.cpu().numpy(), piv.cpu().numpy()
    solved = 0
    for i, (A, b) in enumerate(cases):
        oerr, oz, opiv = lemke(A, b)
        n = len(b)
        assert err[i] == oerr, (i, n, err[i], oerr)
        assert piv[i] == len(opiv), (i, n, piv[i], len(opiv))
        assert np.all(z[i, n:] == 0)
        assert np.max(np.abs(z[i, :n] - oz)) <= 1e-9 * max(1.0, np.max(np.abs(oz))), (i, n)
        if oerr == 0:
            assert validate(A, z[i, :n], b)
            solved += 1
    assert solved >= 300


@pytest.mark.gpu
def test_device_large_batch_properties():
    """16384 problems of 30 rows (the size of config 3's LCPs without friction): every one solves and satisfies the LCP conditions."""
    import torch

    from diffdart_b200.lcp import lemke_solve

    P, n = 16384, 30
    g = torch.Generator(device="cuda").manual_seed(5)
    J = torch.randn(P, n, 16, dtype=torch.float64, device="cuda", generator=g)
    M = J @ J.transpose(1, 2) + 1e-3 * torch.eye(n, dtype=torch.float64, device="cuda")
    q = torch.randn(P, n, dtype=torch.float64, device="cuda", generator=g) - 0.2
    z, err, piv = lemke_solve(M, q)
    assert int((err != 0).sum()) == 0
    w = torch.einsum("pij,pj->pi", M, z) + q
    assert float(z.min()) >= -1e-4 and float(w.min()) >= -1e-4 and float((z * w).abs().max()) <= 1e-4
    assert int(piv.max()) <= 1000 and int(piv.min()) >= 0
