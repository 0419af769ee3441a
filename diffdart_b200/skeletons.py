"""Host-side description of a world's skeleton topology.

A deliberately small mirror of the reference's authoring API
(``Skeleton::createJointAndBodyNodePair``, ``BodyNode::setMass`` …) — just enough to
state the fixtures the reference's tests and examples use on this path — that
flattens into the ``ddb_topology`` of include/ddb.h.  No physics happens here.

Fixtures restated from the reference:
  create_cartpole            unittests/comprehensive/test_Gradients.cpp:1254-1299
  create_multiarm_robot      unittests/TestHelpers.hpp:217-266 (N-link revolute arm)
  create_box_stack           examples/box_stacking/main.cpp:43-112,411-414
  create_block_on_ground     unittests/comprehensive/test_Gradients.cpp:218-290
"""
import ctypes as C
import math

import numpy as np

from . import _abi


def _iso(R=None, p=None):
    return (np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3),
            np.zeros(3) if p is None else np.asarray(p, dtype=np.float64).reshape(3))


class BodyNode:
    """One BodyNode plus its parent Joint."""

    def __init__(self, skeleton, index, parent, joint_type, name):
        self.skeleton = skeleton
        self.index = index  # index within the skeleton
        self.parent = parent  # BodyNode or None
        self.joint_type = joint_type
        self.name = name
        self.axis = np.array([0.0, 0.0, 1.0])  # RevoluteJoint default axis is z
        self.T_parent_to_joint = _iso()
        self.T_child_to_joint = _iso()
        # Inertia defaults: mass 1, COM 0, moment I3 (dart/dynamics/Inertia.hpp:62-63)
        self.mass = 1.0
        self.com = np.zeros(3)
        self.moment = np.eye(3)
        self.gravity_mode = True
        # BodyNodeAspect.hpp:47-48
        self.friction = 1.0
        self.restitution = 0.0
        self.shape_type = _abi.SHAPE_NONE
        self.shape_size = np.zeros(3)
        self.shape_offset = _iso()
        nd = self.num_dofs
        self.position_lower = np.full(nd, -np.inf)
        self.position_upper = np.full(nd, np.inf)
        self.velocity_lower = np.full(nd, -np.inf)
        self.velocity_upper = np.full(nd, np.inf)
        self.force_lower = np.full(nd, -np.inf)
        self.force_upper = np.full(nd, np.inf)

    @property
    def num_dofs(self):
        return {_abi.JOINT_WELD: 0, _abi.JOINT_REVOLUTE: 1, _abi.JOINT_PRISMATIC: 1, _abi.JOINT_FREE: 6}[self.joint_type]

    # --- the subset of the Joint / BodyNode setters the fixtures use ---
    def setAxis(self, axis):
        self.axis = np.asarray(axis, dtype=np.float64)
        return self

    def setTransformFromParentBodyNode(self, R=None, p=None):
        self.T_parent_to_joint = _iso(R, p)
        return self

    def setTransformFromChildBodyNode(self, R=None, p=None):
        self.T_child_to_joint = _iso(R, p)
        return self

    def setMass(self, m):
        self.mass = float(m)
        return self

    def setLocalCOM(self, com):
        self.com = np.asarray(com, dtype=np.float64)
        return self

    def setMomentOfInertia(self, ixx, iyy, izz, ixy=0.0, ixz=0.0, iyz=0.0):
        self.moment = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]], dtype=np.float64)
        return self

    def setFrictionCoeff(self, mu):
        self.friction = float(mu)
        return self

    def setRestitutionCoeff(self, e):
        self.restitution = float(e)
        return self

    def setBoxShape(self, size, R=None, p=None):
        self.shape_type = _abi.SHAPE_BOX
        self.shape_size = np.asarray(size, dtype=np.float64)
        self.shape_offset = _iso(R, p)
        return self

    def setSphereShape(self, radius, R=None, p=None):
        self.shape_type = _abi.SHAPE_SPHERE
        self.shape_size = np.array([radius, 0.0, 0.0])
        self.shape_offset = _iso(R, p)
        return self


    def setCapsuleShape(self, radius, height, R=None, p=None):
        """CapsuleShape(radius, height): axis = local z, `height` is the length of the cylinder part (dart/dynamics/CapsuleShape.hpp)."""
        self.shape_type = _abi.SHAPE_CAPSULE
        self.shape_size = np.array([radius, height, 0.0])
        self.shape_offset = _iso(R, p)
        return self


class Skeleton:
    def __init__(self, name="skeleton"):
        self.name = name
        self.bodies = []

    def createJointAndBodyNodePair(self, joint_type, parent=None, name=None):
        b = BodyNode(self, len(self.bodies), parent, joint_type, name or "body_%d" % len(self.bodies))
        self.bodies.append(b)
        return b

    def createRevoluteJointAndBodyNodePair(self, parent=None, name=None):
        return self.createJointAndBodyNodePair(_abi.JOINT_REVOLUTE, parent, name)

    def createPrismaticJointAndBodyNodePair(self, parent=None, name=None):
        return self.createJointAndBodyNodePair(_abi.JOINT_PRISMATIC, parent, name)

    def createWeldJointAndBodyNodePair(self, parent=None, name=None):
        return self.createJointAndBodyNodePair(_abi.JOINT_WELD, parent, name)

    def createFreeJointAndBodyNodePair(self, parent=None, name=None):
        return self.createJointAndBodyNodePair(_abi.JOINT_FREE, parent, name)

    def getNumDofs(self):
        return sum(b.num_dofs for b in self.bodies)


class WorldDescription:
    """Skeletons + World-level parameters (World.cpp:67-94); flattens to ddb_topology."""

    def __init__(self):
        self.skeletons = []
        self.gravity = np.array([0.0, 0.0, -9.81])
        self.time_step = 1e-3
        self.contact_clipping_depth = 0.03
        self.max_contacts = 0  # 0 = the library default (64 contacts, 192 LCP rows per world); sizes the contact scratch and the per-slot record

    def setGravity(self, g):
        self.gravity = np.asarray(g, dtype=np.float64)
        return self

    def setMaxContacts(self, n):
        """Capacity of the contact stage per world (ddb_params.reserved[0]); worlds that exceed it report bit 8 in their status."""
        self.max_contacts = int(n)
        return self

    def setTimeStep(self, dt):
        self.time_step = float(dt)
        return self

    def addSkeleton(self, skel):
        self.skeletons.append(skel)
        return self

    def getNumDofs(self):
        return sum(s.getNumDofs() for s in self.skeletons)

    def getNumBodies(self):
        return sum(len(s.bodies) for s in self.skeletons)

    def to_ctypes(self, layout=_abi.LAYOUT_WORLD_MAJOR, n_snapshot_slots=1):
        """Returns (Topology, Params, keepalive) — keepalive owns the arrays the structs point at."""
        nb, nd = self.getNumBodies(), self.getNumDofs()
        if nb > _abi.DDB_MAX_BODIES or nd > _abi.DDB_MAX_DOFS:
            raise ValueError("topology exceeds DDB_MAX_BODIES/DDB_MAX_DOFS")
        bodies = (_abi.Body * max(nb, 1))()
        lims = [np.empty(nd) for _ in range(6)]
        flat, dof = 0, 0
        for si, skel in enumerate(self.skeletons):
            base = flat
            for b in skel.bodies:
                cb = bodies[flat]
                cb.parent = -1 if b.parent is None else base + b.parent.index
                if b.parent is not None and b.parent.index >= b.index:
                    raise ValueError("bodies must be listed parent-before-child")
                cb.joint_type = b.joint_type
                cb.dof_offset = dof
                cb.skeleton = si
                cb.gravity_mode = 1 if b.gravity_mode else 0
                cb.shape_type = b.shape_type
                cb.axis[:] = list(b.axis)
                for dst, (R, p) in ((cb.parent_to_joint, b.T_parent_to_joint), (cb.child_to_joint, b.T_child_to_joint),
                                    (cb.shape_offset, b.shape_offset)):
                    dst.r[:] = list(R.reshape(9))
                    dst.p[:] = list(p)
                cb.mass = b.mass
                cb.com[:] = list(b.com)
                m = b.moment
                cb.moment[:] = [m[0, 0], m[1, 1], m[2, 2], m[0, 1], m[0, 2], m[1, 2]]
                cb.friction = b.friction
                cb.restitution = b.restitution
                cb.shape_size[:] = list(b.shape_size)
                k = b.num_dofs
                for arr, src in zip(lims, (b.position_lower, b.position_upper, b.velocity_lower, b.velocity_upper,
                                           b.force_lower, b.force_upper)):
                    arr[dof:dof + k] = src
                dof += k
                flat += 1
        topo = _abi.Topology()
        topo.n_bodies, topo.n_dofs = nb, nd
        topo.bodies = C.cast(bodies, C.POINTER(_abi.Body))
        dp = C.POINTER(C.c_double)
        (topo.position_lower, topo.position_upper, topo.velocity_lower, topo.velocity_upper, topo.force_lower,
         topo.force_upper) = [a.ctypes.data_as(dp) for a in lims]
        params = _abi.default_params()
        params.gravity[:] = list(self.gravity)
        params.time_step = self.time_step
        params.contact_clipping_depth = self.contact_clipping_depth
        params.layout = layout
        params.n_snapshot_slots = n_snapshot_slots
        params.reserved[0] = self.max_contacts
        return topo, params, (bodies, lims)


# ---------------------------------------------------------------------------
# Fixtures
# ---------------------------------------------------------------------------

def create_cartpole():
    """Config 1: prismatic x sled + revolute z pole, child offset (0,-0.5,0), g = (0,-9.81,0)."""
    world = WorldDescription().setGravity([0.0, -9.81, 0.0])
    cartpole = Skeleton("cartpole")
    sled = cartpole.createPrismaticJointAndBodyNodePair(None, "sled").setAxis([1, 0, 0])
    arm = cartpole.createRevoluteJointAndBodyNodePair(sled, "arm").setAxis([0, 0, 1])
    arm.setTransformFromChildBodyNode(p=[0.0, -0.5, 0.0])
    world.addSkeleton(cartpole)
    return world


def create_multiarm_robot(num_links, gravity=(0.0, -9.81, 0.0)):
    """Config 2: N revolute-z links, parent->joint offset (0,1,0) for i>0, mass 1, moment I3."""
    world = WorldDescription().setGravity(list(gravity))
    arm = Skeleton("arm")
    parent = None
    for i in range(num_links):
        b = arm.createRevoluteJointAndBodyNodePair(parent, "arm_%d" % i).setAxis([0, 0, 1])
        if parent is not None:
            b.setTransformFromParentBodyNode(p=[0.0, 1.0, 0.0])
        b.setMass(1.0)
        parent = b
    world.addSkeleton(arm)
    return world


def create_random_tree(num_bodies, seed, free_root=False, gravity=(0.3, -9.81, 0.7)):
    """A branching tree of revolute/prismatic joints with skewed axes, offsets, COMs and
    inertias — exercises every term of the recursions (not a reference fixture)."""
    rng = np.random.default_rng(seed)
    world = WorldDescription().setGravity(list(gravity))
    skel = Skeleton("tree")

    def rand_rot():
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        t = rng.uniform(-1.0, 1.0)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        return np.eye(3) + math.sin(t) * K + (1 - math.cos(t)) * K @ K

    for i in range(num_bodies):
        parent = None if i == 0 else skel.bodies[int(rng.integers(0, i))]
        if i == 0 and free_root:
            b = skel.createFreeJointAndBodyNodePair(None)
        elif rng.uniform() < 0.25:
            b = skel.createPrismaticJointAndBodyNodePair(parent)
        else:
            b = skel.createRevoluteJointAndBodyNodePair(parent)
        ax = rng.normal(size=3)
        b.setAxis(ax / np.linalg.norm(ax))
        b.setTransformFromParentBodyNode(rand_rot(), rng.uniform(-0.5, 0.5, size=3))
        b.setTransformFromChildBodyNode(rand_rot(), rng.uniform(-0.3, 0.3, size=3))
        b.setMass(rng.uniform(0.5, 2.0))
        b.setLocalCOM(rng.uniform(-0.2, 0.2, size=3))
        A = rng.normal(size=(3, 3))
        mom = A @ A.T * 0.1 + np.eye(3) * rng.uniform(0.2, 1.0)
        b.setMomentOfInertia(mom[0, 0], mom[1, 1], mom[2, 2], mom[0, 1], mom[0, 2], mom[1, 2])
    world.addSkeleton(skel)
    return world


def create_box_stack(num_boxes=5):
    """Config 3 topology: welded floor 10x10x0.01 at z=-0.005 + FreeJoint boxes 1x1x0.5
    (examples/box_stacking/main.cpp:43-112); gravity is the World default (0,0,-9.81)."""
    world = WorldDescription()
    floor = Skeleton("floor")
    fb = floor.createWeldJointAndBodyNodePair(None, "floor")
    fb.setTransformFromParentBodyNode(p=[0.0, 0.0, -0.005])
    fb.setBoxShape([10.0, 10.0, 0.01])
    world.addSkeleton(floor)
    for i in range(num_boxes):
        s = Skeleton("box_%d" % i)
        b = s.createFreeJointAndBodyNodePair(None, "box_%d" % i)
        b.setBoxShape([1.0, 1.0, 0.5])
        b.setMass(1.0)
        world.addSkeleton(s)
    return world
