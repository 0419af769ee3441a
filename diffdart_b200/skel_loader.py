""".skel -> WorldDescription: the wire format of the reference's ``dart::utils::SkelParser`` for the parts this path uses.

  load_skel(path or XML string) -> (WorldDescription, report)          SkelParser::readWorld (dart/utils/SkelParser.cpp:697-796)

Mapping, as the reference does it:
  * <physics>: <time_step>, <gravity> (SkelParser.cpp:712-742); the <collision_detector> choice is not carried (this path has its own
    narrow phase) and is listed in the report;
  * <skeleton>: optional <transformation> = the frame its bodies' transformations are given in (:990-996);
  * <body>: <transformation> "x y z rx ry rz" (translation + Euler XYZ, XmlHelpers.cpp:346-379, Geometry.cpp:1260-1290) = the body's world
    transform at the zero pose; <inertia> <mass>, <offset> (local COM), optional <moment_of_inertia> (default moment I3 also when the
    mass is 0: Inertia.hpp:62-63, SkelParser.cpp:1141-1175); <gravity> (:1115-1118); <collision_shape> with its own <transformation>
    (relative to the body, :514-535): <box>, <sphere>, <ellipsoid> with three equal diameters (a sphere), <capsule> (radius, height; axis =
    local z). This path carries one box, sphere or capsule per body: further collision shapes, cylinders, cones, planes and meshes are
    listed in report["ignored"], never silently replaced;
  * <joint type="weld|revolute|prismatic|free">: <parent> ("world" = none), <child>, optional <transformation> = T_ChildBodyToJoint;
    T_ParentBodyToJoint = parentWorld^-1 * childWorld * T_ChildBodyToJoint (:1573-1589); <axis><xyz> (normalised by Joint::setAxis) with
    <limit> <lower> / <upper> -> position limits (:1976-1992); <init_pos> / <init_vel> are returned in the report (:2036-2056).
    Joint <dynamics> (damping, spring_stiffness, friction) is not read: World::addSkeleton zeroes springs and damping on this path
    (World.cpp:698-732, SURVEY finding 4); other joint kinds (ball, euler, universal, planar, translational, screw) raise;
  * creation order: joints in file order, a joint whose parent body has not been created yet is preceded by its parent's joint
    (getNextJointAndNodePair :798-852); DOFs are numbered in that order, skeletons in file order. A parent body without a joint of its own
    becomes the root on a FreeJoint named "root" (:1040-1052).
Host-side logic only (xml.etree); the topology goes to the GPU through WorldDescription.to_ctypes as for the built-in fixtures.
"""
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

from . import _abi
from .skeletons import Skeleton, WorldDescription


def _numbers(text, n):
    v = [float(t) for t in (text or "").split()]
    if len(v) != n:
        raise ValueError("expected %d numbers, got %r" % (n, text))
    return np.array(v, dtype=np.float64)


def _euler_xyz(a):
    """math::eulerXYZToMatrix (Geometry.cpp:1260-1290): R = Rx(a0) Ry(a1) Rz(a2)."""
    cx, sx, cy, sy, cz, sz = math.cos(a[0]), math.sin(a[0]), math.cos(a[1]), math.sin(a[1]), math.cos(a[2]), math.sin(a[2])
    return np.array([[cy * cz, -cy * sz, sy],
                     [cx * sz + cz * sx * sy, cx * cz - sx * sy * sz, -cy * sx],
                     [sx * sz - cx * cz * sy, cz * sx + cx * sy * sz, cx * cy]])


def _iso(elem):
    """<transformation> of elem as (R, p); identity when absent."""
    t = None if elem is None else elem.find("transformation")
    if t is None:
        return np.eye(3), np.zeros(3)
    v = _numbers(t.text, 6)
    return _euler_xyz(v[3:]), v[:3].copy()


def _mul(a, b):
    return a[0] @ b[0], a[0] @ b[1] + a[1]


def _inv(a):
    return a[0].T, -(a[0].T @ a[1])


_JOINT_KIND = {"weld": _abi.JOINT_WELD, "revolute": _abi.JOINT_REVOLUTE, "prismatic": _abi.JOINT_PRISMATIC, "free": _abi.JOINT_FREE}


def _text(elem, tag, default=None):
    e = None if elem is None else elem.find(tag)
    return default if e is None or e.text is None else e.text.strip()


def _read_shape(body, coll, report, skel_name):
    geom = coll.find("geometry")
    shape = None if geom is None or len(geom) == 0 else geom[0]
    R, p = _iso(coll)
    where = "%s/%s" % (skel_name, body.name)
    if _text(coll, "collidable") is not None and float(_text(coll, "collidable")) == 0.0:
        report["ignored"].append("%s: collision shape marked not collidable" % where)
        return False
    if shape is None:
        report["ignored"].append("%s: collision shape without geometry" % where)
    elif shape.tag == "box":
        body.setBoxShape(_numbers(_text(shape, "size"), 3), R, p)
        return True
    elif shape.tag == "sphere":
        body.setSphereShape(float(_text(shape, "radius")), R, p)
        return True
    elif shape.tag == "ellipsoid":
        d = _numbers(_text(shape, "size"), 3)
        if d[0] == d[1] == d[2]:  # EllipsoidShape::isSphere: the detector treats it as a sphere of radius size / 2
            body.setSphereShape(0.5 * d[0], R, p)
            return True
        report["ignored"].append("%s: collision <ellipsoid> with unequal diameters" % where)
    elif shape.tag == "capsule":
        body.setCapsuleShape(float(_text(shape, "radius")), float(_text(shape, "height")), R, p)
        return True
    else:
        report["ignored"].append("%s: collision <%s>" % (where, shape.tag))
    return False


def _read_skeleton(sk, report):
    name = sk.get("name", "skeleton")
    frame = _iso(sk)
    skel = Skeleton(name)
    bodies = {}
    for b in sk.findall("body"):
        bn = b.get("name")
        if bn in bodies:
            raise ValueError("skeleton %r: body name %r is not unique" % (name, bn))
        bodies[bn] = {"xml": b, "world": _mul(frame, _iso(b)) if b.find("transformation") is not None else frame}
    joints, order = {}, []
    for j in sk.findall("joint"):
        kind = j.get("type")
        if kind not in _JOINT_KIND:
            raise ValueError("skeleton %r, joint %r: type %r is not supported on this path" % (name, j.get("name"), kind))
        parent, child = _text(j, "parent"), _text(j, "child")
        if child not in bodies:
            raise ValueError("joint %r: unknown child body %r" % (j.get("name"), child))
        if parent == "world" and "world" not in bodies:
            parent = None
        if parent is not None and parent not in bodies:
            raise ValueError("joint %r: unknown parent body %r" % (j.get("name"), parent))
        if child in joints:
            report["ignored"].append("%s: joint %s discarded (body %s already has a parent joint)" % (name, j.get("name"), child))
            continue
        joints[child] = {"xml": j, "kind": kind, "parent": parent, "child": child}
        order.append(child)
    created = {}

    def create(child_name, jd):
        j, body_xml = jd["xml"], bodies[child_name]["xml"]
        parent_body = created.get(jd["parent"]) if jd["parent"] is not None else None
        body = skel.createJointAndBodyNodePair(_JOINT_KIND[jd["kind"]], parent_body, child_name)
        body.joint_name = j.get("name") if j is not None else "root"
        child_world = bodies[child_name]["world"]
        parent_world = bodies[jd["parent"]]["world"] if jd["parent"] is not None else (np.eye(3), np.zeros(3))
        c2j = _iso(j) if j is not None else (np.eye(3), np.zeros(3))
        p2j = _mul(_mul(_inv(parent_world), child_world), c2j)
        body.setTransformFromParentBodyNode(*p2j)
        body.setTransformFromChildBodyNode(*c2j)
        if jd["kind"] in ("revolute", "prismatic"):
            axis = j.find("axis")
            if axis is None or axis.find("xyz") is None:
                raise ValueError("joint %r is missing axis information" % j.get("name"))
            a = _numbers(_text(axis, "xyz"), 3)
            body.setAxis(a / np.linalg.norm(a))
            lim = axis.find("limit")
            if lim is not None:
                if lim.find("lower") is not None:
                    body.position_lower[:] = float(_text(lim, "lower"))
                if lim.find("upper") is not None:
                    body.position_upper[:] = float(_text(lim, "upper"))
            for tag, key in (("init_pos", "initial_positions"), ("init_vel", "initial_velocities")):
                if _text(j, tag) is not None:
                    report[key]["%s/%s" % (name, j.get("name"))] = float(_text(j, tag))
        inertia = body_xml.find("inertia")
        if inertia is not None:
            body.setMass(float(_text(inertia, "mass", "1")))
            moi = inertia.find("moment_of_inertia")
            if moi is not None:
                g = lambda k: float(_text(moi, k, "0"))
                body.setMomentOfInertia(g("ixx"), g("iyy"), g("izz"), g("ixy"), g("ixz"), g("iyz"))
            if inertia.find("offset") is not None:
                body.setLocalCOM(_numbers(_text(inertia, "offset"), 3))
        if _text(body_xml, "gravity") is not None:
            body.gravity_mode = _text(body_xml, "gravity").lower() in ("1", "true")
        taken = False
        for coll in body_xml.findall("collision_shape"):
            if taken:
                report["ignored"].append("%s/%s: further collision shape (this path carries one shape per body)" % (name, child_name))
            else:
                taken = _read_shape(body, coll, report, name)
        created[child_name] = body

    pending = list(order)
    while pending:
        cur = pending[0]
        # a joint whose parent body has not been created yet is preceded by its parent's joint (getNextJointAndNodePair)
        while True:
            parent = joints[cur]["parent"]
            if parent is None or parent in created:
                break
            if parent in joints:
                cur = parent
                continue
            # the parent body has no joint of its own: it becomes the root on a FreeJoint named "root"
            create(parent, {"xml": None, "kind": "free", "parent": None, "child": parent})
            break
        create(cur, joints[cur])
        pending.remove(cur)
    if _text(sk, "mobile") is not None and _text(sk, "mobile").lower() in ("0", "false") and skel.getNumDofs() > 0:
        report["ignored"].append("%s: <mobile>false</mobile> on a skeleton with joints that have DOFs" % name)
    return skel


def load_skel(source):
    """Returns (WorldDescription, report); report = {"ignored": [...], "initial_positions": {...}, "initial_velocities": {...}}."""
    text = open(source).read() if os.path.exists(str(source)) else source
    root = ET.fromstring(text)
    if root.tag != "skel":
        raise ValueError("not a .skel file: root element is <%s>" % root.tag)
    world = root.find("world")
    if world is None:
        raise ValueError(".skel file without a <world>")
    report = {"ignored": [], "initial_positions": {}, "initial_velocities": {}}
    desc = WorldDescription()
    phys = world.find("physics")
    if phys is not None:
        if _text(phys, "time_step") is not None:
            desc.setTimeStep(float(_text(phys, "time_step")))
        if _text(phys, "gravity") is not None:
            desc.setGravity(_numbers(_text(phys, "gravity"), 3))
        if _text(phys, "collision_detector") is not None:
            report["ignored"].append("physics: <collision_detector>%s</collision_detector> (this path has its own narrow phase)" % _text(phys, "collision_detector"))
    for sk in world.findall("skeleton"):
        desc.addSkeleton(_read_skeleton(sk, report))
    return desc, report
