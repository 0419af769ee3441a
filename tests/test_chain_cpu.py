"""Host emulation of the chain kernels (diffdart_b200/csrc/chain.cuh: canonical joint frames, S = e_z) against the oracle.
The device headers are compiled for the host by tests/host_emulation (test infrastructure); the same sources run on the GPU,
where tests/test_gpu_parity.py repeats these comparisons through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from diffdart_b200 import skeletons
from oracle import oracle as orc
from util import max_ref_err

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def chain_step():
    subprocess.run(["make", "-C", os.path.join(HERE, "host_emulation"), "-s"], check=True)
    lib = C.CDLL(os.path.join(HERE, "host_emulation", "_build", "libemul.so"))
    dp = C.POINTER(C.c_double)

    def run(desc, q, qd, tau, gqn, gqdn, kind=1):
        topo, params, keep = desc.to_ctypes()
        n = len(q)
        outs = [np.zeros(n) for _ in range(5)]
        minv, jac = np.zeros((n, n)), np.zeros((5, n, n))
        rc = lib.emul_chain_step(C.byref(topo), C.byref(params), *[np.ascontiguousarray(a).ctypes.data_as(dp) for a in (q, qd, tau, gqn, gqdn)],
                                 *[o.ctypes.data_as(dp) for o in outs], minv.ctypes.data_as(dp), jac.ctypes.data_as(dp), kind)
        assert rc == 0, rc
        return outs, minv.T, jac.transpose(0, 2, 1)  # column-major per matrix -> [row][col]

    return run


from util import skewed_arm  # noqa: E402


CASES = {
    "cartpole": lambda: skeletons.create_cartpole(),
    "pendulum5": lambda: skeletons.create_multiarm_robot(5),
    "skewed3": lambda: skewed_arm(3, 1),
    "skewed5": lambda: skewed_arm(5, 2),
    "skewed5_revolute": lambda: skewed_arm(5, 3, prismatic_every=0),
}
ALL_REVOLUTE = {"pendulum5", "skewed5_revolute"}


@pytest.mark.parametrize("case", sorted(CASES))
def test_chain_kernels_match_oracle(chain_step, case):
    desc = CASES[case]()
    o = orc.OracleWorld(desc)
    n = desc.to_ctypes()[0].n_dofs
    rng = np.random.default_rng(5)
    for trial in range(6):
        q, qd, tau = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-2, 2, n)
        gqn, gqdn = rng.normal(size=n), rng.normal(size=n)
        # kind 1: revolute / prismatic mixed (selects); kind 2: the instantiation for revolute joints only
        kind = 2 if (case in ALL_REVOLUTE and trial % 2) else 1
        (qn, qdn, gq, gqd, gtau), minv, jac = chain_step(desc, q, qd, tau, gqn, gqdn, kind)
        oqn, oqdn, ogq, ogqd, ogtau = o.step_backprop(q[None], qd[None], tau[None], gqn[None], gqdn[None], mode="exact")
        assert max_ref_err(qn, oqn[0]) < 1e-12 and max_ref_err(qdn, oqdn[0]) < 1e-11, (case, trial)
        for name, a, b in (("dL/dq", gq, ogq[0]), ("dL/dqd", gqd, ogqd[0]), ("dL/dtau", gtau, ogtau[0])):
            assert max_ref_err(a, b) < 1e-8, (case, trial, name, max_ref_err(a, b))
        assert max_ref_err(minv, o.inv_mass_matrix(q[None])[0]) < 1e-10
        J = o.jacobians(q[None], qd[None], tau[None], "exact")
        for k, name in enumerate(("posPos", "posVel", "velPos", "velVel", "forceVel")):
            assert max_ref_err(jac[k], J[name][0]) < 1e-8, (case, trial, name, max_ref_err(jac[k], J[name][0]))
