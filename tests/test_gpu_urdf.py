"""GPU: a model that came in through the URDF loader (free root, revolute, continuous, fixed and prismatic joints) steps and
back-propagates like the oracle says (same bars as test_gpu_parity.py)."""
import numpy as np
import pytest

from diffdart_b200 import _abi
from diffdart_b200.skeletons import WorldDescription
from diffdart_b200.urdf_loader import load_urdf
from oracle import oracle as orc
from test_gpu_parity import check_against_oracle, run_gpu
from util import max_ref_err, ref_equals
from test_urdf_loader_cpu import ARM

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("free_root", [False, True])
def test_urdf_model_against_oracle(free_root):
    urdf = ARM % ({"root": "base", "root_inertial": '<inertial><mass value="4.0"/><inertia ixx="0.3" iyy="0.3" izz="0.3"/></inertial>'}
                  if free_root else {"root": "world", "root_inertial": ""})
    skel, _ = load_urdf(urdf)
    for b in skel.bodies:  # nothing touches in this test: contact-free dynamics
        b.shape_type = _abi.SHAPE_NONE
    desc = WorldDescription().addSkeleton(skel).setGravity([0.0, -9.81, 0.0])
    n = desc.getNumDofs()
    rng = np.random.default_rng(5)
    q, qd, tau = rng.uniform(-0.5, 0.5, (24, n)), rng.uniform(-1, 1, (24, n)), rng.uniform(-2, 2, (24, n))
    if not free_root:
        check_against_oracle(desc, q, qd, tau, seed=3)
        return
    # FreeJoint root: bars of test_gpu_parity.py::test_free_joint_backward (the exp/log-map blocks hold to 1e-8, the rest tighter)
    gq, gqd = rng.normal(size=q.shape), rng.normal(size=q.shape)
    qn, qdn, dq, dqd, dtau, J, Minv = run_gpu(desc, q, qd, tau, gq, gqd)
    o = orc.OracleWorld(desc)
    oqn, oqdn = o.step(q, qd, tau)
    assert ref_equals(qn, oqn, 1e-11) and ref_equals(qdn, oqdn, 1e-11)
    assert ref_equals(Minv, o.inv_mass_matrix(q), 1e-11)
    Je = o.jacobians(q, qd, tau, "exact", threads=4)
    for k in Je:
        assert ref_equals(J[k], Je[k], 1e-8), (k, max_ref_err(J[k], Je[k]))
    _, _, ogq, ogqd, ogtau = o.step_backprop(q, qd, tau, gq, gqd, mode="exact", threads=4, with_step=False)
    assert ref_equals(dq, ogq, 1e-8) and ref_equals(dqd, ogqd, 1e-8) and ref_equals(dtau, ogtau, 1e-11)
