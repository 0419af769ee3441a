"""diffdart_b200/skel_loader.py against the rules of the reference's SkelParser (dart/utils/SkelParser.cpp), on a .skel file written for
this repo and, when present, on the reference's own data/skel/half_cheetah.skel (BASELINE configs[3])."""
import math
import os

import numpy as np
import pytest

from contact_scenes import HALF_CHEETAH_SKEL, HOPPER_SKEL, half_cheetah_available
from diffdart_b200 import _abi
from diffdart_b200.skel_loader import load_skel
from oracle.oracle import OracleWorld


def test_hopper_fixture_follows_the_parser_rules():
    desc, report = load_skel(HOPPER_SKEL)
    assert desc.time_step == 0.002 and list(desc.gravity) == [0.0, -9.81, 0.0]            # <physics> (SkelParser.cpp:712-742)
    ground, hopper = desc.skeletons
    assert [b.name for b in ground.bodies] == ["ground"] and ground.bodies[0].joint_type == _abi.JOINT_WELD
    # creation order: file order of the joints, parents pulled forward (getNextJointAndNodePair :798-852)
    assert [b.name for b in hopper.bodies] == ["slide_x", "slide_y", "torso", "back_thigh", "back_shin", "head", "front_thigh", "front_shin"]
    assert [b.joint_name for b in hopper.bodies] == ["root_x", "root_y", "root_rot", "back_hip", "back_knee", "neck", "front_hip", "front_knee"]
    assert desc.getNumDofs() == 7 and desc.getNumBodies() == 9
    by = {b.name: b for b in hopper.bodies}
    # T_ParentBodyToJoint = parentWorld^-1 childWorld T_ChildBodyToJoint (:1573-1589); the skeleton frame (0, 0.1, 0) shifts every body
    assert np.allclose(by["slide_x"].T_parent_to_joint[1], [0, 0.6, 0])
    assert np.allclose(by["back_thigh"].T_parent_to_joint[1], [-0.25, 0.01, 0]) and np.allclose(by["back_thigh"].T_child_to_joint[1], [0, 0.01, 0])
    assert np.allclose(by["back_shin"].T_parent_to_joint[1], [0, -0.25, 0])
    assert np.allclose(by["head"].T_parent_to_joint[1], [0.35, 0.05, 0]) and by["head"].joint_type == _abi.JOINT_WELD
    assert np.allclose(by["slide_y"].axis, [0, 1, 0])                                       # Joint::setAxis normalises (0, 2, 0)
    assert by["back_shin"].position_lower[0] == -1.2 and by["back_shin"].position_upper[0] == 0.4 and np.isinf(by["front_shin"].position_upper[0])
    # inertia: mass 0 keeps the default moment I3; an explicit tensor is taken as given
    assert by["slide_x"].mass == 0.0 and np.array_equal(by["slide_x"].moment, np.eye(3))
    assert by["torso"].mass == 3.5 and np.allclose(by["torso"].com, [0.02, 0, 0]) and by["torso"].moment[0, 1] == 0.001
    # shapes: capsule (radius, height) about local z with the shape node's Euler-XYZ transform; equal-diameter ellipsoid = sphere
    assert by["torso"].shape_type == _abi.SHAPE_CAPSULE and list(by["torso"].shape_size[:2]) == [0.05, 0.5]
    assert np.allclose(by["torso"].shape_offset[0] @ [0, 0, 1], [1, 0, 0])                   # rotated about y by pi/2: axis along x
    assert np.allclose(by["back_shin"].shape_offset[0] @ [0, 0, 1], [0, -1, 0]) and np.allclose(by["back_shin"].shape_offset[1], [0, -0.1, 0])
    assert by["head"].shape_type == _abi.SHAPE_SPHERE and by["head"].shape_size[0] == 0.08
    assert ground.bodies[0].shape_type == _abi.SHAPE_BOX and list(ground.bodies[0].shape_size) == [200.0, 0.1, 4.0]
    assert any("further collision shape" in m for m in report["ignored"]) and any("collision_detector" in m for m in report["ignored"])
    assert report["initial_positions"] == {"hopper/back_knee": 0.05}


def test_unsupported_joint_kinds_raise():
    text = open(HOPPER_SKEL).read().replace('type="weld" name="neck"', 'type="ball" name="neck"')
    with pytest.raises(ValueError, match="ball"):
        load_skel(text)


def test_body_without_a_joint_becomes_a_free_root():
    text = """<skel version="1.0"><world name="w"><skeleton name="s">
      <body name="base"><transformation>0 1 0 0 0 0.5</transformation></body>
      <body name="arm"><transformation>0 1.5 0 0 0 0</transformation></body>
      <joint type="revolute" name="j"><parent>base</parent><child>arm</child><axis><xyz>0 0 1</xyz></axis></joint>
    </skeleton></world></skel>"""
    desc, _ = load_skel(text)
    b = desc.skeletons[0].bodies
    assert [x.name for x in b] == ["base", "arm"] and b[0].joint_type == _abi.JOINT_FREE and b[0].joint_name == "root"
    assert np.allclose(b[0].T_parent_to_joint[1], [0, 1, 0])          # FreeJoint root at the body's initial transform (:1040-1052)
    c, s = math.cos(0.5), math.sin(0.5)
    assert np.allclose(b[0].T_parent_to_joint[0], [[c, -s, 0], [s, c, 0], [0, 0, 1]])     # Euler XYZ, rotation about z
    assert np.allclose(b[1].T_parent_to_joint[1], np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]]) @ [0, 0.5, 0])
    assert desc.getNumDofs() == 7


def test_hopper_world_steps_on_its_capsule_ends_and_reports_pipe_contacts():
    desc, _ = load_skel(HOPPER_SKEL)
    o = OracleWorld(desc)
    q = np.zeros(7); q[1] = -0.114
    r = o.step_contacts(q, np.zeros(7), np.zeros(7), gradients=True)
    assert len(r["contacts"]) == 2 and len(r["x"]) == 6 and not r["tier_mask"] & 512
    assert np.allclose(r["contacts"][:, 3], 0.0, atol=1e-12)          # contact points on the ground plane, under the two shin ends
    assert np.allclose(sorted(r["contacts"][:, 2]), [-0.25, 0.25], atol=1e-9)
    # the torso capsule lying on the ground is a pipe contact (createCapsuleMeshContact): not carried, reported in bit 9
    q = np.zeros(7); q[1] = -0.552; q[3] = q[5] = 1.5
    r = o.step_contacts(q, np.zeros(7), np.zeros(7), gradients=True)
    assert r["tier_mask"] & 512


@pytest.mark.skipif(not half_cheetah_available(), reason="the reference's half_cheetah.skel has not been copied to baseline/_ref/skel")
def test_half_cheetah_skel_loads_as_the_reference_describes_it():
    desc, report = load_skel(HALF_CHEETAH_SKEL)
    assert desc.time_step == 0.002 and list(desc.gravity) == [0.0, -9.81, 0.0]            # half_cheetah.skel:5-6
    ground, cheetah = desc.skeletons
    assert desc.getNumDofs() == 9 and len(cheetah.bodies) == 10                            # 9 DOFs
    assert [b.joint_name for b in cheetah.bodies][:4] == ["j_pelvis_x", "j_pelvis_y", "j_pelvis_rot", "j_head"]
    assert [b.joint_type for b in cheetah.bodies][:4] == [_abi.JOINT_PRISMATIC, _abi.JOINT_PRISMATIC, _abi.JOINT_REVOLUTE, _abi.JOINT_WELD]
    assert all(b.shape_type == _abi.SHAPE_CAPSULE for b in cheetah.bodies[2:])
    assert abs(sum(b.mass for b in cheetah.bodies) - 14.0) < 0.3                           # the MuJoCo half-cheetah's 14 kg
    by = {b.name: b for b in cheetah.bodies}
    assert by["b_thigh"].position_lower[0] == -0.52 and by["b_thigh"].position_upper[0] == 1.05
    assert report["ignored"] == ["physics: <collision_detector>bullet</collision_detector> (this path has its own narrow phase)"]
    o = OracleWorld(desc)
    q = np.zeros(9); q[1] = -0.125
    r = o.step_contacts(q, np.zeros(9), np.zeros(9), gradients=True)
    assert len(r["contacts"]) == 2 and not r["tier_mask"] & 512                            # standing on the end caps of both feet
