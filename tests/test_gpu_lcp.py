"""GPU: batched Dantzig / PGS / validity kernels through the C ABI against the reference's own Dantzig solver
(oracle/_ref) and the CPU restatements. Bar: bit-exact x (integer-like parity for the pivoting solver)."""
import numpy as np
import pytest
import torch

from diffdart_b200.lcp import is_lcp_solution_valid, solve_boxed_lcp
from lcp_cases import golden_instances, pad_batch, random_contact_lcp
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def cuda(a, dtype=torch.float64):
    return torch.tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


@pytest.fixture(autouse=True)
def _quiet_ode(capfd):
    yield
    capfd.readouterr()


def _run(problems, solver="dantzig", early=False, x0=None):
    n_max = max(len(p[1]) for p in problems)
    A, b, lo, hi, fi, rows = pad_batch(problems, n_max)
    x, ok = solve_boxed_lcp(cuda(A), cuda(b), cuda(lo), cuda(hi), cuda(fi, torch.int32), solver=solver, early_termination=early,
                            x0=None if x0 is None else cuda(x0), n_rows=cuda(rows, torch.int32))
    return x.cpu().numpy(), ok.cpu().numpy(), (A, b, lo, hi, fi, rows)


def test_dantzig_reference_test_instances():
    cases = golden_instances()
    problems = [(c["A"], np.array(c["b"]), np.array(c["lo"]), np.array(c["hi"]), np.array(c["findex"], dtype=np.int32)) for c in cases]
    for early in (False, True):
        x, ok, _ = _run(problems, early=early)
        for k, c in enumerate(cases):
            want = c["ref_dantzig_early%d" % int(early)]
            assert bool(ok[k]) == want["ok"], c["name"]
            if want["ok"]:
                assert np.array_equal(x[k, :c["n"]], np.array(want["x"]), equal_nan=True), c["name"]


@pytest.mark.parametrize("early", [False, True])
def test_dantzig_random_contact_problems_bit_exact(early):
    rng = np.random.default_rng(42)
    problems = [random_contact_lcp(rng) for _ in range(1500)]
    x, ok, _ = _run(problems, early=early)
    for k, (A, b, lo, hi, fi) in enumerate(problems):
        rok, rx = orc.ref_dantzig(A, b, lo, hi, fi, early)
        assert bool(ok[k]) == rok, k
        if rok:
            assert np.array_equal(x[k, :len(b)], rx, equal_nan=True), k


def test_dantzig_max_size_60_rows():
    """Box-stack size: 20 contacts x 3 rows (k_max = 60)."""
    rng = np.random.default_rng(7)
    problems = [random_contact_lcp(rng, nc=20, nd=30) for _ in range(64)]
    x, ok, _ = _run(problems)
    for k, (A, b, lo, hi, fi) in enumerate(problems):
        rok, rx = orc.ref_dantzig(A, b, lo, hi, fi)
        assert bool(ok[k]) == rok and np.array_equal(x[k], rx, equal_nan=True)


def test_pgs_and_validity():
    rng = np.random.default_rng(3)
    problems = [random_contact_lcp(rng) for _ in range(300)]
    n_max = max(len(p[1]) for p in problems)
    x0 = np.zeros((len(problems), n_max))
    for k, p in enumerate(problems):
        if k % 2:
            x0[k, :len(p[1])] = rng.uniform(0, 0.1, size=len(p[1]))
    x, ok, (A, b, lo, hi, fi, rows) = _run(problems, solver="pgs", x0=x0)
    for k, (Ak, bk, lk, hk, fk) in enumerate(problems):
        n = len(bk)
        ook, ox = orc.pgs_solve(Ak, bk, lk, hk, fk, x0[k, :n])
        assert bool(ok[k]) == ook and np.array_equal(x[k, :n], ox, equal_nan=True)
    for ign in (False, True):
        v = is_lcp_solution_valid(cuda(A), cuda(x), cuda(b), cuda(lo), cuda(hi), cuda(fi, torch.int32), ign, cuda(rows, torch.int32)).cpu().numpy()
        for k, (Ak, bk, lk, hk, fk) in enumerate(problems):
            assert bool(v[k]) == orc.lcp_valid(Ak, x[k, :len(bk)], bk, lk, hk, fk, ign)


def test_ragged_and_empty():
    rng = np.random.default_rng(9)
    problems = [random_contact_lcp(rng, nc=c) for c in (1, 8, 1, 3)]
    n_max = 24
    A, b, lo, hi, fi, rows = pad_batch(problems, n_max)
    rows[2] = 0  # an empty problem: nothing to solve, reported ok, x untouched
    x, ok = solve_boxed_lcp(cuda(A), cuda(b), cuda(lo), cuda(hi), cuda(fi, torch.int32), n_rows=cuda(rows, torch.int32))
    x, ok = x.cpu().numpy(), ok.cpu().numpy()
    assert ok[2] and np.all(x[2] == 0)
    for k in (0, 1, 3):
        Ak, bk, lk, hk, fk = problems[k]
        _, rx = orc.ref_dantzig(Ak, bk, lk, hk, fk)
        assert np.array_equal(x[k, :len(bk)], rx, equal_nan=True) and np.all(x[k, len(bk):] == 0)
