"""TEST INFRASTRUCTURE (never on the product path): CPU restatement of the reference's Lemke solver.

  lemke(M, q) -> (err, z, pivots)     dart::lcpsolver::Lemke        dart/lcpsolver/Lemke.cpp:61-337
  validate(M, z, q) -> bool           dart::lcpsolver::validate     dart/lcpsolver/Lemke.cpp:340-357

Standard LCP  w = M z + q >= 0, z >= 0, z.w = 0  by complementary pivoting with the artificial variable t = 2n.
The reference solves B d = Be with Eigen's `householderQr().solve` at every pivot (Lemke.cpp:192); numpy's `qr` is the same
unpivoted Householder factorisation (LAPACK dgeqrf), so the restatement agrees with it to rounding. Third-party boundary: Eigen 3.3.7
is not vendored; parity is pinned on the reference's own instances (unittests/unit/test_Lemke.cpp:39-185, expected:
err == 0 and validate()), committed as tests/golden/lemke_instances.json by tests/golden/make_lemke_golden.py.
`pivots` lists (entering, leaving) per iteration: the device solver must take the same path.
"""
import numpy as np

ZER_TOL, PIV_TOL, MAXITER = 1e-5, 1e-8, 1000  # Lemke.cpp:66-68


def _solve(B, rhs):
    Q, R = np.linalg.qr(B)
    with np.errstate(all="ignore"):
        return np.linalg.solve(R, Q.T @ rhs) if np.all(np.diag(R) != 0) else np.full(len(rhs), np.nan)


def validate(M, z, q, threshold=1e-4):
    w = M @ z + q
    return bool(np.all(w >= -threshold) and np.all(z >= -threshold) and np.all(np.abs(w * z) <= threshold))


def lemke(M, q):
    M, q = np.asarray(M, dtype=np.float64), np.asarray(q, dtype=np.float64)
    n = q.size
    pivots = []
    if q.min() >= 0:  # trivial solution (:70-75)
        return 0, np.zeros(n), pivots
    t = 2 * n
    x = q.copy()
    B = -np.eye(n)
    # initial basis = all slack variables w (bas[i] = n + i); the most negative one leaves for the artificial variable (:150-171)
    lvindex = int(np.argmax(-x))
    tval = -x[lvindex]
    bas = [n + i for i in range(n)]
    leaving = bas[lvindex]
    bas[lvindex] = t
    U = (x < 0).astype(np.float64)
    Be = -(B @ U)
    x = x + tval * U
    x[lvindex] = tval
    B[:, lvindex] = Be
    err, it = 0, 0
    for it in range(MAXITER):
        if leaving == t:
            break
        if leaving < n:
            entering = n + leaving
            Be = np.zeros(n)
            Be[leaving] = -1.0
        else:
            entering = leaving - n
            Be = M[:, entering].copy()
        d = _solve(B, Be)
        j = [i for i in range(n) if d[i] > PIV_TOL]
        if not j:
            err = 2  # ray termination (:201-205)
            break
        theta = min((x[i] + ZER_TOL) / d[i] for i in j)
        j = [i for i in j if x[i] / d[i] <= theta]
        if not j:
            err = 4
            break
        art = [i for i in j if bas[i] == t]
        if art:
            lvindex = art[-1]  # the reference keeps the last match of its scan (:247-251)
        else:
            best, lvindex = d[j[0]], j[0]
            for i in j:
                if d[i] - best > PIV_TOL:
                    best, lvindex = d[i], i
        leaving = bas[lvindex]
        ratio = x[lvindex] / d[lvindex]
        x = x - ratio * d
        x[lvindex] = ratio
        B[:, lvindex] = Be
        bas[lvindex] = entering
        pivots.append((entering, leaving))
    else:
        it = MAXITER
    if it >= MAXITER and leaving != t:
        err = 1
    z = np.zeros(n)
    if err == 0:
        for i, b in enumerate(bas):
            if b < n:  # (the reference writes into a 2n vector and keeps its head)
                z[b] = x[i]
        if not validate(M, z, q):
            err = 3
    return err, z, pivots
