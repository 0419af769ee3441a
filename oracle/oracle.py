"""TEST INFRASTRUCTURE ONLY — ctypes binding of the CPU oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. It loads
  oracle/_build/liboracle.so     our restatement of the reference arithmetic
  oracle/_ref/libodelcp_ref.so   the reference's own, unmodified ODE Dantzig solver
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libodelcp_ref.so")

_dp = C.POINTER(C.c_double)


def build():
    """Compile the oracle (and, where /root/reference exists, the reference Dantzig)."""
    subprocess.run(["make", "-C", HERE, "-s"], check=True, stdout=subprocess.DEVNULL)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _lib = C.CDLL(ORACLE_SO)
        _lib.ddo_create.restype = C.c_void_p
        _lib.ddo_create.argtypes = [C.c_void_p, C.c_void_p]
        _lib.ddo_destroy.argtypes = [C.c_void_p]
    return _lib


class OracleWorld:
    """CPU restatement of World::step + BackpropSnapshot for one topology; batched over worlds."""

    def __init__(self, world_description):
        self._topo, self._params, self._keep = world_description.to_ctypes()
        self.h = C.c_void_p(lib().ddo_create(C.byref(self._topo), C.byref(self._params)))
        self.n = self._topo.n_dofs
        self.nb = self._topo.n_bodies
        self.dt = self._params.time_step

    def __del__(self):
        if getattr(self, "h", None):
            lib().ddo_destroy(self.h)
            self.h = None

    @staticmethod
    def _c(a):
        return np.ascontiguousarray(a, dtype=np.float64)

    def step(self, q, qd, tau, threads=1):
        q, qd, tau = self._c(q), self._c(qd), self._c(tau)
        N = q.shape[0]
        qn, qdn = np.empty_like(q), np.empty_like(q)
        lib().ddo_step(self.h, N, threads, _ptr(q), _ptr(qd), _ptr(tau), _ptr(qn), _ptr(qdn))
        return qn, qdn

    def jacobians(self, q, qd, tau, mode="reference", threads=1):
        """mode 'reference': the reference's recipe incl. its finite differences; 'exact': long-double FD."""
        q, qd, tau = self._c(q), self._c(qd), self._c(tau)
        N, n = q.shape
        out = [np.empty((N, n, n)) for _ in range(5)]
        lib().ddo_jacobians(self.h, N, threads, 0 if mode == "reference" else 1, _ptr(q), _ptr(qd), _ptr(tau),
                            *[_ptr(o) for o in out])
        # stored [world][col][row] -> return as [world][row][col]
        names = ("posPos", "posVel", "velPos", "velVel", "forceVel")
        return {k: np.ascontiguousarray(o.transpose(0, 2, 1)) for k, o in zip(names, out)}

    def step_backprop(self, q, qd, tau, gqn, gqdn, mode="reference", threads=1, with_step=True):
        q, qd, tau, gqn, gqdn = map(self._c, (q, qd, tau, gqn, gqdn))
        N = q.shape[0]
        qn, qdn = (np.empty_like(q), np.empty_like(q)) if with_step else (None, None)
        gq, gqd, gtau = np.empty_like(q), np.empty_like(q), np.empty_like(q)
        lib().ddo_step_backprop(self.h, N, threads, 0 if mode == "reference" else 1, _ptr(q), _ptr(qd), _ptr(tau),
                                _ptr(gqn), _ptr(gqdn), _ptr(qn), _ptr(qdn), _ptr(gq), _ptr(gqd), _ptr(gtau))
        return qn, qdn, gq, gqd, gtau

    def inv_mass_matrix(self, q):
        q = self._c(q)
        N, n = q.shape
        out = np.empty((N, n, n))
        lib().ddo_inv_mass_matrix(self.h, N, _ptr(q), _ptr(out))
        return np.ascontiguousarray(out.transpose(0, 2, 1))

    def coriolis_gravity(self, q, qd):
        q, qd = self._c(q), self._c(qd)
        out = np.empty_like(q)
        lib().ddo_coriolis_gravity(self.h, q.shape[0], _ptr(q), _ptr(qd), _ptr(out))
        return out

    def forward_dynamics(self, q, qd, tau):
        q, qd, tau = self._c(q), self._c(qd), self._c(tau)
        out = np.empty_like(q)
        lib().ddo_forward_dynamics(self.h, q.shape[0], _ptr(q), _ptr(qd), _ptr(tau), _ptr(out))
        return out

    def body_transforms(self, q):
        q = self._c(q)
        out = np.empty((q.shape[0], self.nb, 12))
        lib().ddo_body_transforms(self.h, q.shape[0], _ptr(q), _ptr(out))
        return out

    def step_contacts(self, q, qd, tau, lcp_cache=None, kmax=256, cmax=64, gradients=False):
        """World::step with contacts (gradients disabled) for ONE world. Returns a dict with the next state and the
        intermediate quantities (contacts, boxed LCP, solution, solver tier mask)."""
        mask = lib().ddo_load_ref_libs(REF_SO.encode(), BOXBOX_SO.encode())
        if mask != 3:
            raise RuntimeError("oracle/_ref libraries missing (mask %d)" % mask)
        n = self.n
        q, qd, tau = (np.ascontiguousarray(a, dtype=np.float64).reshape(n) for a in (q, qd, tau))
        cache = np.zeros(kmax)
        ncache = C.c_int(0)
        if lcp_cache is not None:
            cache[:len(lcp_cache)] = lcp_cache
            ncache.value = len(lcp_cache)
        qn, qdn, pre = np.zeros(n), np.zeros(n), np.zeros(n)
        contacts = np.zeros((cmax, 10))
        A, b, lo, hi, x = np.zeros((kmax, kmax)), np.zeros(kmax), np.zeros(kmax), np.zeros(kmax), np.zeros(kmax)
        fi = -np.ones(kmax, dtype=np.int32)
        counts = np.zeros(4, dtype=np.int32)
        mapping = np.full(kmax, -9, dtype=np.int32)
        ipt = C.POINTER(C.c_int)
        lib().ddo_step_contacts(self.h, _ptr(q), _ptr(qd), _ptr(tau), _ptr(cache), C.byref(ncache), kmax, _ptr(qn), _ptr(qdn), _ptr(pre), cmax,
                                _ptr(contacts), _ptr(A), _ptr(b), _ptr(lo), _ptr(hi), fi.ctypes.data_as(ipt), _ptr(x), counts.ctypes.data_as(ipt), 1 if gradients else 0, mapping.ctypes.data_as(ipt))
        nc, k = int(counts[0]), int(counts[1])
        return {"q_next": qn, "qd_next": qdn, "pre_constraint_vel": pre, "contacts": contacts[:nc], "A": A[:k, :k], "b": b[:k], "lo": lo[:k],
                "hi": hi[:k], "findex": fi[:k], "x": x[:k], "n_groups": int(counts[2]), "tier_mask": int(counts[3]),
                "lcp_cache": cache[:ncache.value].copy(), "active_set": mapping[:k].copy()}

    def contact_jacobians(self, q, qd, tau, lcp_cache=None, mode="reference", eps=1e-6, x_override=None, mapping_override=None, q_noise=0.0):
        """forwardPass with contacts for ONE world + the BackpropSnapshot Jacobians through the contact solution.
        mode 'reference': the reference's recipe; 'frozen_fd': central differences of the frozen-active-set velocity map
        (collision detector re-run at every perturbed state). Matrices are returned [row][col]. x_override: evaluate the
        Jacobians at these LCP multipliers (all rows, group order) instead of the oracle's own solution — the multipliers of a
        rank-deficient group are fixed only up to the null space of A. mapping_override: the ConstraintMapping per row that goes
        with those multipliers (world row numbering). q_noise: relative noise applied to Q before its decomposition, to
        measure how strongly the pseudo-inverse formulas amplify rounding in this world."""
        mask = lib().ddo_load_ref_libs(REF_SO.encode(), BOXBOX_SO.encode())
        if mask != 3:
            raise RuntimeError("oracle/_ref libraries missing (mask %d)" % mask)
        n = self.n
        q, qd, tau = (np.ascontiguousarray(a, dtype=np.float64).reshape(n) for a in (q, qd, tau))
        cache = np.zeros(1) if lcp_cache is None else np.ascontiguousarray(lcp_cache, dtype=np.float64)
        nc = 0 if lcp_cache is None else len(lcp_cache)
        out = [np.zeros((n, n)) for _ in range(5)]
        info = np.zeros(4, dtype=np.int32)
        qn, qdn = np.zeros(n), np.zeros(n)
        xo = None if x_override is None else np.ascontiguousarray(x_override, dtype=np.float64)
        fc_ls = C.c_double(0.0)
        lib().ddo_contact_jacobians(self.h, 0 if mode == "reference" else 1, C.c_double(eps), _ptr(q), _ptr(qd), _ptr(tau), _ptr(cache), nc,
                                    *[_ptr(o) for o in out], info.ctypes.data_as(C.POINTER(C.c_int)), _ptr(qn), _ptr(qdn), _ptr(xo),
                                    0 if xo is None else len(xo), C.byref(fc_ls),
                                    None if mapping_override is None else np.ascontiguousarray(mapping_override, dtype=np.int32).ctypes.data_as(C.POINTER(C.c_int)),
                                    C.c_double(q_noise))
        names = ("posPos", "posVel", "velPos", "velVel", "forceVel")
        r = {k: np.ascontiguousarray(o.T) for k, o in zip(names, out)}
        r.update(num_clamping=int(info[0]), num_upper_bound=int(info[1]), pseudo_inverse=bool(info[2]), structure_kept=bool(info[3]),
                 q_next=qn, qd_next=qdn, fc_vs_least_squares=fc_ls.value)
        return r

    def spatial_inertias(self):
        out = np.empty((self.nb, 6, 6))
        lib().ddo_spatial_inertias(self.h, _ptr(out))
        return out


def transform_inertia(T12, I66):
    out = np.empty((6, 6))
    lib().ddo_transform_inertia(_ptr(np.ascontiguousarray(T12, dtype=np.float64)), _ptr(np.ascontiguousarray(I66, dtype=np.float64)), _ptr(out))
    return out


def exp_angular(s):
    out = np.empty((3, 3))
    lib().ddo_exp_angular(_ptr(np.ascontiguousarray(s, dtype=np.float64)), _ptr(out))
    return out


def exp_map_rot(s):
    out = np.empty((3, 3))
    lib().ddo_exp_map_rot(_ptr(np.ascontiguousarray(s, dtype=np.float64)), _ptr(out))
    return out


def log_map_rot(R):
    out = np.empty(3)
    lib().ddo_log_map_rot(_ptr(np.ascontiguousarray(R, dtype=np.float64)), _ptr(out))
    return out


def pgs_solve(A, b, lo, hi, findex, x0=None):
    """CPU restatement of PgsBoxedLcpSolver::solve (one problem). Returns (ok, x)."""
    n = len(b)
    A = np.ascontiguousarray(A, dtype=np.float64)
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    b, lo, hi = (np.array(a, dtype=np.float64) for a in (b, lo, hi))
    fi = np.array(findex, dtype=np.int32)
    ok = lib().ddo_pgs_solve(n, _ptr(A), _ptr(x), _ptr(b), _ptr(lo), _ptr(hi), fi.ctypes.data_as(C.POINTER(C.c_int)))
    return bool(ok), x


def cod_solve(A, b):
    """The oracle's minimum-norm least-squares solve; returns (x, rank)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    m, n = A.shape
    x = np.zeros(n)
    rank = lib().ddo_cod_solve(m, n, _ptr(A), _ptr(b), _ptr(x))
    return x, rank


def lcp_valid(A, x, b, lo, hi, findex, ignore_friction=False):
    """CPU restatement of LCPUtils::isLCPSolutionValid."""
    n = len(b)
    A, x, b, lo, hi = (np.ascontiguousarray(a, dtype=np.float64) for a in (A, x, b, lo, hi))
    fi = np.array(findex, dtype=np.int32)
    return bool(lib().ddo_lcp_valid(n, _ptr(A), _ptr(x), _ptr(b), _ptr(hi), _ptr(lo), fi.ctypes.data_as(C.POINTER(C.c_int)), int(ignore_friction)))


# ---------------------------------------------------------------------------
# the reference's own Dantzig solver
# ---------------------------------------------------------------------------
_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            build()
        if not os.path.exists(REF_SO):
            raise RuntimeError("oracle/_ref/libodelcp_ref.so is missing and /root/reference is not present to build it")
        _ref = C.CDLL(REF_SO)
    return _ref


def ref_dantzig(A, b, lo, hi, findex, early_termination=False):
    """Run the reference's dSolveLCP exactly as DantzigBoxedLcpSolver::solve calls it.
    A (n x n), returns (ok, x). Inputs are copied (the solver destroys A, b, lo, hi)."""
    r = ref_lib()
    n = len(b)
    pad = r.ddref_dpad(n)
    Ap = np.zeros((n, pad))
    Ap[:, :n] = A
    x = np.zeros(n)
    b = np.array(b, dtype=np.float64)
    lo = np.array(lo, dtype=np.float64)
    hi = np.array(hi, dtype=np.float64)
    fi = np.array(findex, dtype=np.int32)
    ok = r.ddref_dantzig_solve(n, _ptr(Ap), _ptr(x), _ptr(b), _ptr(lo), _ptr(hi), fi.ctypes.data_as(C.POINTER(C.c_int)),
                               1 if early_termination else 0)
    return bool(ok), x


BOXBOX_SO = os.path.join(HERE, "_ref", "libboxbox_ref.so")
# ---------------------------------------------------------------------------
# the reference's own box-box narrow phase (DARTCollide.cpp:49-1427 compiled through a small shim)
# ---------------------------------------------------------------------------
BOXBOX_SO = os.path.join(HERE, "_ref", "libboxbox_ref.so")
_bb = None


def ref_collide_box_box(size0, T0, size1, T1):
    """collideBoxBox of the reference. T: 12 doubles (R row-major, p). Returns an [n, 26] array:
    point, normal, edgeAClosest, edgeAFixed, edgeADir, edgeBClosest, edgeBFixed, edgeBDir, depth, type."""
    global _bb
    if _bb is None:
        if not os.path.exists(BOXBOX_SO):
            build()
        _bb = C.CDLL(BOXBOX_SO)
    out = np.zeros(26 * 16)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (size0, T0, size1, T1)]
    n = _bb.ddref_collide_box_box(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), _ptr(out), 16)
    return out[:26 * n].reshape(n, 26).copy()
