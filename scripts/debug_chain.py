import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from diffdart_b200 import skeletons
from oracle import oracle as orc
from util import max_ref_err, random_inputs
from test_gpu_parity import run_gpu
np.set_printoptions(linewidth=200, precision=5)
for links in (3,):
    desc = skeletons.create_multiarm_robot(links)
    q, qd, tau = random_inputs(links, 4, links)
    rng = np.random.default_rng(links)
    gq, gqd = rng.normal(size=q.shape), rng.normal(size=q.shape)
    o = orc.OracleWorld(desc)
    _, _, ogq, ogqd, ogtau = o.step_backprop(q, qd, tau, gq, gqd, mode="exact", threads=4, with_step=False)
    for fg in ("0", "1"):
        os.environ["DDB_FORCE_GENERIC"] = fg
        qn, qdn, dq, dqd, dtau, J, Minv = run_gpu(desc, q, qd, tau, gq, gqd)
        print("generic" if fg == "1" else "chain")
        print(" dq err\n", dq - ogq)
        print(" dqd err\n", dqd - ogqd)
        print(" (dq - gq)/dt\n", (dq - gq) / 1e-3, "\n expected\n", (ogq - gq) / 1e-3)
