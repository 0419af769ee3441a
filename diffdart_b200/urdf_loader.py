"""URDF -> Skeleton: the wire format of the reference's ``dart::utils::DartLoader`` for the parts this path uses.

  load_urdf(path or XML string) -> (Skeleton, report)     DartLoader::parseSkeleton / modelInterfaceToSkeleton
                                                          (dart/utils/urdf/DartLoader.cpp:190-260)
Mapping, as the reference does it:
  * a root link named "world" is the inertial frame and gets no body; any other root link hangs on a FreeJoint called
    "rootJoint" (DartLoader.cpp:204-238);
  * joints: revolute / continuous -> RevoluteJoint (continuous: position limits +-inf), prismatic -> PrismaticJoint,
    fixed -> WeldJoint, floating -> FreeJoint (DartLoader.cpp:422-478); the joint frame sits at <origin> in the parent link's
    frame and coincides with the child link's frame (T_ParentBodyToJoint = origin, T_ChildBodyToJoint = I, :378-380);
    <limit lower upper velocity effort> -> position / velocity / force limits (:383-391); axis default (1, 0, 0) (URDF spec);
  * <inertial>: mass, COM = origin xyz, moment rotated by origin rpy: J' = R J R^T (DartLoader.cpp:511-527);
  * <collision> <box size> / <sphere radius> with its <origin> -> the body's collision shape. This path carries one box or sphere
    per body: further collision elements, cylinders, capsules and meshes are listed in report["ignored"], never silently replaced.
  Joint damping / friction are not read: World::addSkeleton zeroes them on this path (World.cpp:698-732, SURVEY finding 4);
  planar joints and mimic joints raise.
Host-side logic only (xml.etree); the topology goes to the GPU through WorldDescription.to_ctypes as for the built-in fixtures.
"""
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

from . import _abi
from .skeletons import Skeleton


def _floats(text, n, default):
    if text is None:
        return np.array(default, dtype=np.float64)
    v = [float(t) for t in text.split()]
    if len(v) != n:
        raise ValueError("expected %d numbers, got %r" % (n, text))
    return np.array(v, dtype=np.float64)


def _rpy(r, p, y):
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _origin(elem):
    o = None if elem is None else elem.find("origin")
    if o is None:
        return np.eye(3), np.zeros(3)
    return _rpy(*_floats(o.get("rpy"), 3, (0, 0, 0))), _floats(o.get("xyz"), 3, (0, 0, 0))


_JOINT_KIND = {"revolute": _abi.JOINT_REVOLUTE, "continuous": _abi.JOINT_REVOLUTE, "prismatic": _abi.JOINT_PRISMATIC,
               "fixed": _abi.JOINT_WELD, "floating": _abi.JOINT_FREE}


def load_urdf(source):
    """Returns (Skeleton, report); report = {"ignored": [...], "initial_positions": {joint name: value}}."""
    text = open(source).read() if os.path.exists(str(source)) else source
    robot = ET.fromstring(text)
    if robot.tag != "robot":
        raise ValueError("not a URDF: root element is <%s>" % robot.tag)
    links = {l.get("name"): l for l in robot.findall("link")}
    joints = robot.findall("joint")
    children, child_names = {}, set()
    for j in joints:
        kind = j.get("type")
        if kind not in _JOINT_KIND:
            raise ValueError("joint %r: type %r is not supported on this path" % (j.get("name"), kind))
        if j.find("mimic") is not None:
            raise ValueError("joint %r: mimic joints are not supported on this path" % j.get("name"))
        parent, child = j.find("parent").get("link"), j.find("child").get("link")
        if parent not in links or child not in links:
            raise ValueError("joint %r refers to an unknown link" % j.get("name"))
        children.setdefault(parent, []).append(j)
        child_names.add(child)
    roots = [n for n in links if n not in child_names]
    if len(roots) != 1:
        raise ValueError("a URDF has exactly one root link, found %r" % roots)
    skel = Skeleton(robot.get("name", "robot"))
    report = {"ignored": [], "initial_positions": {}}

    def fill_body(body, link):
        inertial = link.find("inertial")
        if inertial is not None:
            R, com = _origin(inertial)
            mass = inertial.find("mass")
            body.setMass(float(mass.get("value")) if mass is not None else 1.0)
            body.setLocalCOM(com)
            i = inertial.find("inertia")
            if i is not None:
                g = lambda k: float(i.get(k, 0.0))
                J = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
                J = R @ J @ R.T
                body.setMomentOfInertia(J[0, 0], J[1, 1], J[2, 2], J[0, 1], J[0, 2], J[1, 2])
        taken = False
        for col in link.findall("collision"):
            geom = col.find("geometry")
            shape = None if geom is None or len(geom) == 0 else geom[0]
            R, p = _origin(col)
            if shape is not None and shape.tag == "box" and not taken:
                body.setBoxShape(_floats(shape.get("size"), 3, None), R, p)
                taken = True
            elif shape is not None and shape.tag == "sphere" and not taken:
                body.setSphereShape(float(shape.get("radius")), R, p)
                taken = True
            else:
                report["ignored"].append("link %s: collision <%s>%s" % (link.get("name"), "none" if shape is None else shape.tag,
                                                                        " (this path carries one box or sphere per body)" if taken else ""))

    def add_children(link_name, parent_body):
        for j in children.get(link_name, []):
            kind = j.get("type")
            body = skel.createJointAndBodyNodePair(_JOINT_KIND[kind], parent_body, j.find("child").get("link"))
            body.joint_name = j.get("name")
            R, p = _origin(j)
            body.setTransformFromParentBodyNode(R, p)
            axis = j.find("axis")
            if _JOINT_KIND[kind] in (_abi.JOINT_REVOLUTE, _abi.JOINT_PRISMATIC):
                a = _floats(None if axis is None else axis.get("xyz"), 3, (1, 0, 0))
                body.setAxis(a / np.linalg.norm(a))
                lim = j.find("limit")
                if lim is not None:
                    lo, hi = float(lim.get("lower", 0.0)), float(lim.get("upper", 0.0))
                    vel, eff = float(lim.get("velocity", "inf")), float(lim.get("effort", "inf"))
                    if kind == "continuous":
                        lo, hi = -np.inf, np.inf
                    body.position_lower[:], body.position_upper[:] = lo, hi
                    body.velocity_lower[:], body.velocity_upper[:] = -vel, vel
                    body.force_lower[:], body.force_upper[:] = -eff, eff
                    # zero outside the limits: the reference starts the joint inside them (DartLoader.cpp:393-411)
                    if lo > 0 or hi < 0:
                        report["initial_positions"][j.get("name")] = (lo + hi) / 2.0 if np.isfinite(lo) and np.isfinite(hi) else (lo if np.isfinite(lo) else hi)
            fill_body(body, links[j.find("child").get("link")])
            add_children(j.find("child").get("link"), body)

    root = roots[0]
    root_body = None
    if root != "world":
        root_body = skel.createFreeJointAndBodyNodePair(None, root)
        root_body.joint_name = "rootJoint"
        fill_body(root_body, links[root])
    add_children(root, root_body)
    return skel, report
