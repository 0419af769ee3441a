"""Batched boxed-LCP solvers on the GPU (the reference's ``BoxedLcpSolver::solve`` for N problems).

  solve_boxed_lcp(A, b, lo, hi, findex, solver="dantzig")   DantzigBoxedLcpSolver / PgsBoxedLcpSolver
  is_lcp_solution_valid(A, x, b, lo, hi, findex)            LCPUtils::isLCPSolutionValid

All tensors are CUDA tensors: A [N, n, n] float64, b/lo/hi/x [N, n] float64, findex [N, n] int32,
optional n_rows [N] int32 for ragged batches (rows beyond n_rows[i] are ignored and left untouched).
"""
import ctypes as C

import torch

from . import _abi

DANTZIG, PGS = 0, 1


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def solve_boxed_lcp(A, b, lo, hi, findex, solver="dantzig", early_termination=False, x0=None, n_rows=None):
    lib = _abi.load_library()
    N, n = b.shape
    dev = b.device
    A, b, lo, hi = (t.contiguous() for t in (A, b, lo, hi))
    findex = findex.to(torch.int32).contiguous()
    x = torch.zeros_like(b) if x0 is None else x0.clone().contiguous()
    ok = torch.zeros(N, dtype=torch.int32, device=dev)
    rc = lib.ddb_lcp_solve(dev.index or 0, _stream(dev), N, n, None if n_rows is None else n_rows.data_ptr(), A.data_ptr(), b.data_ptr(),
                           lo.data_ptr(), hi.data_ptr(), findex.data_ptr(), DANTZIG if solver == "dantzig" else PGS,
                           1 if early_termination else 0, x.data_ptr(), ok.data_ptr())
    if rc:
        raise RuntimeError("ddb_lcp_solve failed: %s" % _abi.ERRORS.get(rc, rc))
    return x, ok.bool()


def is_lcp_solution_valid(A, x, b, lo, hi, findex, ignore_friction_indices=False, n_rows=None):
    lib = _abi.load_library()
    N, n = b.shape
    dev = b.device
    A, x, b, lo, hi = (t.contiguous() for t in (A, x, b, lo, hi))
    findex = findex.to(torch.int32).contiguous()
    out = torch.zeros(N, dtype=torch.int32, device=dev)
    rc = lib.ddb_lcp_valid(dev.index or 0, _stream(dev), N, n, None if n_rows is None else n_rows.data_ptr(), A.data_ptr(), x.data_ptr(),
                           b.data_ptr(), lo.data_ptr(), hi.data_ptr(), findex.data_ptr(), 1 if ignore_friction_indices else 0, out.data_ptr())
    if rc:
        raise RuntimeError("ddb_lcp_valid failed: %s" % _abi.ERRORS.get(rc, rc))
    return out.bool()


def lemke_solve(M, q, n_rows=None):
    """dart::lcpsolver::Lemke for a batch of standard LCPs w = M z + q >= 0, z >= 0, z.w = 0 (Lemke.cpp:61-337).
    M [P, n, n], q [P, n] float64 CUDA tensors (n_rows [P] int32: size of each problem, default n). Returns (z [P, n], err [P] int32,
    pivots [P] int32); err as the reference: 0 solved, 1 iteration limit, 2 ray termination, 3 validation failed, 4 no pivot."""
    import torch

    lib = _abi.load_library()
    M, q = M.contiguous(), q.contiguous()
    P, n = q.shape
    z = torch.empty_like(q)
    err = torch.empty(P, dtype=torch.int32, device=q.device)
    piv = torch.empty(P, dtype=torch.int32, device=q.device)
    rc = lib.ddb_lemke_solve(q.device.index or 0, torch.cuda.current_stream(q.device).cuda_stream, P, n,
                             None if n_rows is None else n_rows.contiguous().data_ptr(), M.data_ptr(), q.data_ptr(), z.data_ptr(),
                             err.data_ptr(), piv.data_ptr())
    if rc != 0:
        raise RuntimeError("ddb_lemke_solve failed: %s" % _abi.ERRORS.get(rc, rc))
    return z, err, piv
