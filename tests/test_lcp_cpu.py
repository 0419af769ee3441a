"""CPU: the device LCP source (diffdart_b200/csrc/lcp.cuh) compiled for the host and checked for
bit-identical results against the REFERENCE's own Dantzig solver (oracle/_ref, compiled unmodified from
/root/reference/dart/external/odelcpsolver) — on the literal instances of the reference's
unittests/unit/test_LCPUtils.cpp and on seeded contact-shaped problems. The GPU run of the same source is
checked in test_gpu_lcp.py; this test pins the arithmetic ORDER without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from lcp_cases import golden_instances, random_contact_lcp
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", os.path.join(HERE, "host_emulation"), "-s"], check=True)
    lib = C.CDLL(os.path.join(HERE, "host_emulation", "_build", "libemul.so"))

    def solve(A, b, lo, hi, fi, solver=0, early=0, x0=None):
        n = len(b)
        A = np.ascontiguousarray(A, dtype=np.float64)
        b, lo, hi = (np.array(a, dtype=np.float64) for a in (b, lo, hi))
        fi = np.array(fi, dtype=np.int32)
        x = np.zeros(n)
        x0p = None if x0 is None else np.array(x0, dtype=np.float64).ctypes.data_as(dp)
        ok = lib.emul_lcp_solve(n, A.ctypes.data_as(dp), b.ctypes.data_as(dp), lo.ctypes.data_as(dp), hi.ctypes.data_as(dp),
                                fi.ctypes.data_as(ip), x0p, solver, early, x.ctypes.data_as(dp))
        return bool(ok), x

    def valid(A, x, b, lo, hi, fi, ign=False):
        n = len(b)
        A, x, b, lo, hi = (np.ascontiguousarray(a, dtype=np.float64) for a in (A, x, b, lo, hi))
        fi = np.array(fi, dtype=np.int32)
        return bool(lib.emul_lcp_valid(n, A.ctypes.data_as(dp), x.ctypes.data_as(dp), b.ctypes.data_as(dp), hi.ctypes.data_as(dp),
                                       lo.ctypes.data_as(dp), fi.ctypes.data_as(ip), int(ign)))

    solve.valid = valid
    return solve


@pytest.fixture(autouse=True)
def _quiet_ode(capfd):
    yield
    capfd.readouterr()  # the reference's solver prints "LCP internal error, s <= 0" on degenerate problems


def test_dantzig_bit_identical_on_reference_test_instances(emul):
    for c in golden_instances():
        for early in (0, 1):
            ok, x = emul(c["A"], c["b"], c["lo"], c["hi"], c["findex"], 0, early)
            want = c["ref_dantzig_early%d" % early]
            assert ok == want["ok"], c["name"]
            if ok:
                assert np.array_equal(x, np.array(want["x"]), equal_nan=True), c["name"]


def test_dantzig_bit_identical_on_random_contact_problems(emul):
    rng = np.random.default_rng(0)
    for _ in range(600):
        A, b, lo, hi, fi = random_contact_lcp(rng)
        for early in (False, True):
            ok, x = emul(A, b, lo, hi, fi, 0, int(early))
            rok, rx = orc.ref_dantzig(A, b, lo, hi, fi, early)
            assert ok == rok
            if ok:
                assert np.array_equal(x, rx, equal_nan=True)


def test_dantzig_with_unbounded_rows(emul):
    """Rows with lo=-inf, hi=+inf are moved to the front and factored first (lcp.cpp:456-487). That block uses a
    plain LDL^T here (not the reference's blocked dFactorLDLT), so agreement is to rounding, not bit-exact."""
    rng = np.random.default_rng(5)
    for _ in range(50):
        A, b, lo, hi, fi = random_contact_lcp(rng, nc=3, nd=12)
        lo[0], hi[0] = -np.inf, np.inf
        lo[3], hi[3] = -np.inf, np.inf
        ok, x = emul(A, b, lo, hi, fi)
        rok, rx = orc.ref_dantzig(A, b, lo, hi, fi)
        assert ok == rok
        if np.all(np.isfinite(rx)) and np.abs(rx).max() < 1e6:  # ill-posed draws blow up identically; skip them
            assert np.allclose(x, rx, rtol=1e-9, atol=1e-11)


def test_pgs_and_validity_match_cpu_restatement(emul):
    rng = np.random.default_rng(1)
    for k in range(200):
        A, b, lo, hi, fi = random_contact_lcp(rng)
        x0 = rng.uniform(0, 0.1, size=len(b)) if k % 2 else None
        ok, x = emul(A, b, lo, hi, fi, solver=1, x0=x0)
        ook, ox = orc.pgs_solve(A, b, lo, hi, fi, x0)
        assert ok == ook and np.array_equal(x, ox, equal_nan=True)
        for ign in (False, True):
            assert emul.valid(A, x, b, lo, hi, fi, ign) == orc.lcp_valid(A, x, b, lo, hi, fi, ign)
        _, xd = orc.ref_dantzig(A, b, lo, hi, fi)
        if np.all(np.isfinite(xd)):
            assert emul.valid(A, xd, b, lo, hi, fi) == orc.lcp_valid(A, xd, b, lo, hi, fi)


def test_ode_friction_bound_quirk():
    """ODE freezes friction bounds at |hi*x[findex]| when the first friction row is reached and never
    refreshes them, so a Dantzig answer can violate isLCPSolutionValid and fall through to PGS."""
    A = np.array([[2.0, 0.5, 0.2], [0.5, 2.0, 0.1], [0.2, 0.1, 2.0]])
    b = np.array([1.0, 1.0, 1.0])
    ok, x = orc.ref_dantzig(A, b, [0, -0.5, -0.5], [np.inf, 0.5, 0.5], [-1, 0, 0])
    assert ok and x[0] > 0
