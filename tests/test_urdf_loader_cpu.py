"""URDF loader: host-side logic, checked on the CPU against the same model written through the builder API
(dart/utils/urdf/DartLoader.cpp is the wire-format reference) and stepped by the oracle."""
import numpy as np
import pytest

from diffdart_b200 import _abi
from diffdart_b200.skeletons import Skeleton, WorldDescription
from diffdart_b200.urdf_loader import load_urdf, _rpy
from oracle.oracle import OracleWorld

ARM = """<?xml version="1.0"?>
<robot name="arm">
  <link name="%(root)s">%(root_inertial)s</link>
  <link name="upper">
    <inertial><origin xyz="0.1 0 0.2" rpy="0.3 -0.2 0.5"/><mass value="2.5"/>
      <inertia ixx="0.4" ixy="0.01" ixz="-0.02" iyy="0.5" iyz="0.03" izz="0.6"/></inertial>
    <collision><origin xyz="0 0 0.25" rpy="0 0 0.1"/><geometry><box size="0.2 0.3 0.5"/></geometry></collision>
    <collision><geometry><cylinder radius="0.1" length="0.4"/></geometry></collision>
  </link>
  <link name="fore"><inertial><mass value="1.5"/><inertia ixx="0.2" iyy="0.2" izz="0.1"/></inertial>
    <collision><geometry><sphere radius="0.15"/></geometry></collision></link>
  <link name="tool"><inertial><origin xyz="0 0.05 0"/><mass value="0.3"/><inertia ixx="0.01" iyy="0.01" izz="0.01"/></inertial></link>
  <link name="slider"><inertial><mass value="0.8"/><inertia ixx="0.05" iyy="0.06" izz="0.07"/></inertial></link>
  <joint name="shoulder" type="revolute"><parent link="%(root)s"/><child link="upper"/>
    <origin xyz="0 0 0.5" rpy="0.1 0.2 0.3"/><axis xyz="0 1 0"/><limit lower="-1.5" upper="2.0" velocity="3" effort="40"/></joint>
  <joint name="elbow" type="continuous"><parent link="upper"/><child link="fore"/>
    <origin xyz="0 0 0.5"/><axis xyz="1 0 0"/><limit lower="0" upper="0" velocity="5" effort="20"/></joint>
  <joint name="wrist" type="fixed"><parent link="fore"/><child link="tool"/><origin xyz="0 0 0.3" rpy="0 0.4 0"/></joint>
  <joint name="rail" type="prismatic"><parent link="upper"/><child link="slider"/>
    <origin xyz="0.2 0 0"/><axis xyz="0 0 2"/><limit lower="0.1" upper="0.4" velocity="1" effort="10"/></joint>
</robot>"""


def hand_built(free_root):
    s = Skeleton("arm")
    root = None
    if free_root:
        root = s.createFreeJointAndBodyNodePair(None, "base").setMass(4.0).setMomentOfInertia(0.3, 0.3, 0.3)
    up = s.createRevoluteJointAndBodyNodePair(root, "upper").setAxis([0, 1, 0])
    up.setTransformFromParentBodyNode(_rpy(0.1, 0.2, 0.3), [0, 0, 0.5])
    R = _rpy(0.3, -0.2, 0.5)
    J = R @ np.array([[0.4, 0.01, -0.02], [0.01, 0.5, 0.03], [-0.02, 0.03, 0.6]]) @ R.T
    up.setMass(2.5).setLocalCOM([0.1, 0, 0.2]).setMomentOfInertia(J[0, 0], J[1, 1], J[2, 2], J[0, 1], J[0, 2], J[1, 2])
    up.setBoxShape([0.2, 0.3, 0.5], _rpy(0, 0, 0.1), [0, 0, 0.25])
    fo = s.createRevoluteJointAndBodyNodePair(up, "fore").setAxis([1, 0, 0])
    fo.setTransformFromParentBodyNode(p=[0, 0, 0.5])
    fo.setMass(1.5).setMomentOfInertia(0.2, 0.2, 0.1).setSphereShape(0.15)
    to = s.createWeldJointAndBodyNodePair(fo, "tool")
    to.setTransformFromParentBodyNode(_rpy(0, 0.4, 0), [0, 0, 0.3])
    to.setMass(0.3).setLocalCOM([0, 0.05, 0]).setMomentOfInertia(0.01, 0.01, 0.01)
    sl = s.createPrismaticJointAndBodyNodePair(up, "slider").setAxis([0, 0, 1])
    sl.setTransformFromParentBodyNode(p=[0.2, 0, 0])
    sl.setMass(0.8).setMomentOfInertia(0.05, 0.06, 0.07)
    return s


@pytest.mark.parametrize("free_root", [False, True])
def test_loaded_model_equals_hand_built(free_root):
    urdf = ARM % ({"root": "base", "root_inertial": '<inertial><mass value="4.0"/><inertia ixx="0.3" iyy="0.3" izz="0.3"/></inertial>'}
                  if free_root else {"root": "world", "root_inertial": ""})
    skel, report = load_urdf(urdf)
    ref = hand_built(free_root)
    assert [b.name for b in skel.bodies] == [b.name for b in ref.bodies]
    for a, b in zip(skel.bodies, ref.bodies):
        assert a.joint_type == b.joint_type and (a.parent is None) == (b.parent is None)
        if a.parent is not None:
            assert a.parent.index == b.parent.index
        for x, y in ((a.axis, b.axis), (a.T_parent_to_joint[0], b.T_parent_to_joint[0]), (a.T_parent_to_joint[1], b.T_parent_to_joint[1]),
                     (a.com, b.com), (a.moment, b.moment), (a.shape_size, b.shape_size), (a.shape_offset[0], b.shape_offset[0]),
                     (a.shape_offset[1], b.shape_offset[1])):
            assert np.allclose(x, y, atol=1e-15)
        assert a.mass == b.mass and a.shape_type == b.shape_type
    # limits (DartLoader.cpp:383-391): revolute keeps them, continuous drops the position limits
    by = {b.name: b for b in skel.bodies}
    assert by["upper"].position_lower[0] == -1.5 and by["upper"].position_upper[0] == 2.0
    assert by["upper"].velocity_upper[0] == 3 and by["upper"].force_lower[0] == -40
    assert np.isinf(by["fore"].position_lower[0]) and by["fore"].velocity_upper[0] == 5
    assert by["slider"].position_lower[0] == 0.1 and report["initial_positions"] == {"rail": 0.25}
    assert any("cylinder" in s for s in report["ignored"])
    # the flattened topology steps identically in the oracle
    wa, wb = WorldDescription().addSkeleton(skel), WorldDescription().addSkeleton(ref)
    n = wa.getNumDofs()
    assert n == wb.getNumDofs() == (9 if free_root else 3)
    rng = np.random.default_rng(1)
    q, qd, tau = rng.uniform(-0.5, 0.5, (4, n)), rng.uniform(-1, 1, (4, n)), rng.uniform(-2, 2, (4, n))
    # bodies carry collision shapes but nothing touches: compare the contact-free dynamics
    for w in (wa, wb):
        for b in w.skeletons[0].bodies:
            b.shape_type = _abi.SHAPE_NONE
    qa, qda = OracleWorld(wa).step(q, qd, tau)
    qb, qdb = OracleWorld(wb).step(q, qd, tau)
    assert np.array_equal(qa, qb) and np.array_equal(qda, qdb)
    topo, params, keep = wa.to_ctypes()
    assert topo.n_bodies == len(skel.bodies) and topo.n_dofs == n


def test_rejects_what_it_cannot_carry():
    with pytest.raises(ValueError, match="planar"):
        load_urdf('<robot name="r"><link name="a"/><link name="b"/><joint name="j" type="planar"><parent link="a"/><child link="b"/></joint></robot>')
    with pytest.raises(ValueError, match="mimic"):
        load_urdf('<robot name="r"><link name="a"/><link name="b"/><joint name="j" type="revolute"><parent link="a"/><child link="b"/>'
                  '<mimic joint="k"/></joint></robot>')
    with pytest.raises(ValueError, match="one root"):
        load_urdf('<robot name="r"><link name="a"/><link name="b"/></robot>')
    with pytest.raises(ValueError, match="not a URDF"):
        load_urdf("<sdf/>")


ATLAS = "/root/reference/data/sdf/atlas/atlas_v3_no_head.urdf"


@pytest.mark.skipif(not __import__("os").path.exists(ATLAS), reason="the reference tree is only mounted in the build container")
def test_atlas_topology_of_config_5():
    """atlas_v3_no_head.urdf = 27 revolute joints + FreeJoint root = 33 DOFs; its mesh colliders are reported, not guessed."""
    skel, report = load_urdf(ATLAS)
    w = WorldDescription().addSkeleton(skel).setGravity([0.0, -9.81, 0.0])
    assert w.getNumDofs() == 33 and sum(b.joint_type == _abi.JOINT_REVOLUTE for b in skel.bodies) == 27
    assert skel.bodies[0].joint_type == _abi.JOINT_FREE and skel.bodies[0].joint_name == "rootJoint"
    assert all("mesh" in s for s in report["ignored"]) and len(report["ignored"]) > 20
    topo, params, keep = w.to_ctypes()
    assert topo.n_bodies == 34
    rng = np.random.default_rng(13)
    q, qd = rng.uniform(-0.01, 0.01, (2, 33)), rng.uniform(-0.1, 0.1, (2, 33))
    qn, qdn = OracleWorld(w).step(q, qd, np.zeros((2, 33)))
    assert np.all(np.isfinite(qn)) and np.all(np.isfinite(qdn))
    # free fall of the whole robot: the root's linear acceleration in the world frame is gravity
    M = OracleWorld(w).inv_mass_matrix(q)
    assert np.all(np.linalg.eigvalsh(M[0]) > 0)
