"""Contact scenes built from Revolute / Prismatic / Weld joints (FreeJoint bodies come later): seeded states in which
contacts exist with depth in (0, 0.03)."""
import numpy as np

from diffdart_b200 import skeletons


def _floor(world, thickness=1.0, mu=1.0, name="floor"):
    floor = skeletons.Skeleton(name)
    fb = floor.createWeldJointAndBodyNodePair(None).setTransformFromParentBodyNode(p=[0, 0, -0.5 * thickness])
    fb.setBoxShape([10, 10, thickness]).setFrictionCoeff(mu)
    world.addSkeleton(floor)


def _xyz_body(world, name, mass=2.0, mu=1.0, box=None, sphere=None, restitution=0.0):
    """A body that translates along x, y, z (three prismatic joints with light carrier links)."""
    s = skeletons.Skeleton(name)
    bx = s.createPrismaticJointAndBodyNodePair(None).setAxis([1, 0, 0]).setMass(1e-3).setMomentOfInertia(1e-3, 1e-3, 1e-3)
    by = s.createPrismaticJointAndBodyNodePair(bx).setAxis([0, 1, 0]).setMass(1e-3).setMomentOfInertia(1e-3, 1e-3, 1e-3)
    bz = s.createPrismaticJointAndBodyNodePair(by).setAxis([0, 0, 1]).setMass(mass).setFrictionCoeff(mu).setRestitutionCoeff(restitution)
    if box is not None:
        bz.setBoxShape(box)
    if sphere is not None:
        bz.setSphereShape(sphere)
    world.addSkeleton(s)


def sliding_box(mu=1.0):
    """Box sliding on the floor: 4 contacts, 12 rows, sliding friction."""
    w = skeletons.WorldDescription()
    _floor(w, mu=mu)
    _xyz_body(w, "box", box=[1, 1, 1], mu=mu)
    return w


def sliding_box_states(N, seed):
    rng = np.random.default_rng(seed)
    q = np.stack([rng.uniform(-1, 1, N), rng.uniform(-1, 1, N), 0.5 - rng.uniform(1e-4, 5e-3, N)], axis=1)
    qd = np.stack([rng.uniform(-1, 1, N), rng.uniform(-1, 1, N), rng.uniform(-0.5, 0.2, N)], axis=1)
    tau = rng.uniform(-3, 3, (N, 3))
    return q, qd, tau


def two_box_stack():
    """Two translating boxes stacked on the floor: one constrained group, 4 + 4 contacts."""
    w = skeletons.WorldDescription()
    _floor(w)
    _xyz_body(w, "lower", box=[1, 1, 0.5])
    _xyz_body(w, "upper", mass=1.0, box=[0.6, 0.8, 0.5])
    return w


def two_box_stack_states(N, seed):
    rng = np.random.default_rng(seed)
    pen = lambda: rng.uniform(1e-4, 3e-3, N)
    z0 = 0.25 - pen()
    q = np.stack([rng.uniform(-.2, .2, N), rng.uniform(-.2, .2, N), z0,
                  rng.uniform(-.1, .1, N), rng.uniform(-.05, .05, N), z0 + 0.5 - pen()], axis=1)
    qd = rng.uniform(-0.3, 0.3, (N, 6))
    tau = rng.uniform(-1, 1, (N, 6))
    return q, qd, tau


def separate_groups():
    """Two boxes far apart on the floor (two constrained groups) + a sphere resting on the floor (frictionless)."""
    w = skeletons.WorldDescription()
    _floor(w)
    _xyz_body(w, "a", box=[1, 1, 1])
    _xyz_body(w, "b", box=[0.5, 0.5, 0.5], mass=0.7)
    _xyz_body(w, "ball", sphere=0.3, mass=0.5, mu=0.0)
    return w


def separate_groups_states(N, seed):
    rng = np.random.default_rng(seed)
    pen = lambda: rng.uniform(1e-4, 3e-3, N)
    q = np.stack([rng.uniform(-3, -2, N), rng.uniform(-1, 1, N), 0.5 - pen(),
                  rng.uniform(2, 3, N), rng.uniform(-1, 1, N), 0.25 - pen(),
                  rng.uniform(-.5, .5, N), rng.uniform(3, 4, N), 0.3 - pen()], axis=1)
    qd = rng.uniform(-0.5, 0.5, (N, 9))
    tau = rng.uniform(-1, 1, (N, 9))
    return q, qd, tau


def pendulum_on_floor():
    """Two-link revolute pendulum (y axis) whose tip box touches the floor."""
    w = skeletons.WorldDescription()
    _floor(w)
    s = skeletons.Skeleton("pendulum")
    a = s.createRevoluteJointAndBodyNodePair(None).setAxis([0, 1, 0]).setTransformFromParentBodyNode(p=[0, 0, 1.5])
    a.setLocalCOM([0, 0, -0.4])
    b = s.createRevoluteJointAndBodyNodePair(a).setAxis([0, 1, 0]).setTransformFromParentBodyNode(p=[0, 0, -0.8])
    b.setLocalCOM([0, 0, -0.3]).setBoxShape([0.2, 0.2, 0.2], p=[0, 0, -0.6])
    w.addSkeleton(s)
    return w


def pendulum_on_floor_states(N, seed):
    rng = np.random.default_rng(seed)
    # both joints near zero: tip box bottom at z = 1.5 - 0.8 - 0.6 - 0.1 = 0.0 -> lower the angles slightly for contact
    q = rng.uniform(-0.02, 0.02, (N, 2))
    qd = rng.uniform(-0.5, 0.5, (N, 2))
    tau = rng.uniform(-1, 1, (N, 2))
    return q, qd, tau


def box_stack(n=5):
    """BASELINE config 3: examples/box_stacking — welded floor + n FreeJoint boxes 1x1x0.5, mu = 1."""
    return skeletons.create_box_stack(n)


def box_stack_states(N, seed, n=5):
    """resting stack with 1 mm overlap per interface, x,y ~ U(-0.05,0.05), yaw ~ U(-0.1,0.1),
    qd ~ U(-0.1,0.1), tau = 0. FreeJoint coordinates are [logMap(R); p]."""
    rng = np.random.default_rng(seed)
    q = np.zeros((N, 6 * n))
    for i in range(n):
        q[:, 6 * i + 2] = rng.uniform(-0.1, 0.1, N)
        q[:, 6 * i + 3] = rng.uniform(-0.05, 0.05, N)
        q[:, 6 * i + 4] = rng.uniform(-0.05, 0.05, N)
        q[:, 6 * i + 5] = 0.25 + 0.5 * i - 1e-3 * (i + 1)
    qd = rng.uniform(-0.1, 0.1, (N, 6 * n))
    tau = np.zeros((N, 6 * n))
    return q, qd, tau


def free_box_on_floor():
    """One FreeJoint box dropped on the floor at an arbitrary small tilt (vertex / edge / face contacts)."""
    return skeletons.create_box_stack(1)


def free_box_on_floor_states(N, seed):
    rng = np.random.default_rng(seed)
    q = np.zeros((N, 6))
    q[:, 0:3] = rng.uniform(-0.05, 0.05, (N, 3))
    q[:, 3:5] = rng.uniform(-1, 1, (N, 2))
    q[:, 5] = 0.25 + rng.uniform(-2e-3, 2e-2, N)
    qd = rng.uniform(-0.5, 0.5, (N, 6))
    tau = rng.uniform(-1, 1, (N, 6))
    return q, qd, tau


def box_on_pedestal():
    """A FreeJoint box resting, yawed and slightly tilted, on a narrow welded pedestal: the manifold mixes VERTEX_FACE /
    FACE_VERTEX and EDGE_EDGE contacts, and the tilt keeps the reference-face choice of dBoxBox away from a tie."""
    w = skeletons.WorldDescription()
    ped = skeletons.Skeleton("pedestal")
    pb = ped.createWeldJointAndBodyNodePair(None).setTransformFromParentBodyNode(p=[0, 0, -0.5])
    pb.setBoxShape([0.9, 0.9, 1.0])
    w.addSkeleton(ped)
    s = skeletons.Skeleton("box")
    b = s.createFreeJointAndBodyNodePair(None, "box")
    b.setBoxShape([1.0, 1.0, 0.5]).setMass(1.0)
    w.addSkeleton(s)
    return w


def box_on_pedestal_states(N, seed):
    rng = np.random.default_rng(seed)
    q = np.zeros((N, 6))
    q[:, 0:2] = rng.uniform(-0.004, 0.004, (N, 2))
    q[:, 2] = rng.uniform(0.25, 0.6, N)
    q[:, 3:5] = rng.uniform(-0.05, 0.05, (N, 2))
    q[:, 5] = 0.25 - rng.uniform(2e-3, 4e-3, N)
    qd = rng.uniform(-0.2, 0.2, (N, 6))
    tau = rng.uniform(-1, 1, (N, 6))
    return q, qd, tau


def sphere_pendulum_on_floor():
    """Two-link revolute pendulum whose tip sphere touches the floor (SPHERE_BOX / BOX_SPHERE contact gradients), next to
    a second one-link pendulum whose tip sphere touches the first sphere (SPHERE_SPHERE)."""
    w = skeletons.WorldDescription()
    _floor(w)
    s = skeletons.Skeleton("pendulum")
    a = s.createRevoluteJointAndBodyNodePair(None).setAxis([0, 1, 0]).setTransformFromParentBodyNode(p=[0, 0, 1.494])
    a.setLocalCOM([0, 0, -0.4])
    b = s.createRevoluteJointAndBodyNodePair(a).setAxis([0, 1, 0]).setTransformFromParentBodyNode(p=[0, 0, -0.8])
    b.setLocalCOM([0, 0, -0.3]).setSphereShape(0.1, p=[0, 0, -0.6])
    w.addSkeleton(s)
    t = skeletons.Skeleton("second")
    c = t.createRevoluteJointAndBodyNodePair(None).setAxis([0, 1, 0]).setTransformFromParentBodyNode(p=[0.195, 0, 1.5])
    c.setLocalCOM([0, 0, -0.7]).setSphereShape(0.1, p=[0, 0, -1.38])
    w.addSkeleton(t)
    return w


def sphere_pendulum_on_floor_states(N, seed):
    rng = np.random.default_rng(seed)
    q = rng.uniform(-0.01, 0.01, (N, 3))
    qd = rng.uniform(-0.5, 0.5, (N, 3))
    tau = rng.uniform(-1, 1, (N, 3))
    return q, qd, tau


def fixed_base_arm_on_floor():
    """A fixed-base arm as DartLoader builds it from a URDF (world -> fixed -> base): the welded base box sits 2 mm inside the floor.
    BodyNode::isReactive is per body (BodyNode.cpp:1999-2018): the base depends on no generalized coordinate, so its contacts with the
    floor never become constraints, while the tip box of the two revolute links does touch the floor and does."""
    w = skeletons.WorldDescription()
    _floor(w)
    s = skeletons.Skeleton("arm")
    base = s.createWeldJointAndBodyNodePair(None).setTransformFromParentBodyNode(p=[0, 0, 0.098])
    base.setBoxShape([0.3, 0.3, 0.2])
    a = s.createRevoluteJointAndBodyNodePair(base).setAxis([0, 1, 0]).setTransformFromParentBodyNode(p=[0, 0, 1.402])
    a.setLocalCOM([0, 0, -0.4])
    b = s.createRevoluteJointAndBodyNodePair(a).setAxis([0, 1, 0]).setTransformFromParentBodyNode(p=[0, 0, -0.8])
    b.setLocalCOM([0, 0, -0.3]).setBoxShape([0.2, 0.2, 0.2], p=[0, 0, -0.6])
    w.addSkeleton(s)
    return w


def fixed_base_arm_on_floor_states(N, seed):
    rng = np.random.default_rng(seed)
    q = rng.uniform(-0.02, 0.02, (N, 2))   # as pendulum_on_floor: the tip box dips into the floor (bodies of one skeleton do not collide)
    qd = rng.uniform(-0.5, 0.5, (N, 2))
    tau = rng.uniform(-1, 1, (N, 2))
    return q, qd, tau


def bouncing_boxes():
    """Two translating boxes with restitution 0.8 / 0.9 hitting the floor (e = 0.8 * 1.0 resp. 0.9 * 1.0 with the floor's coefficient 1):
    contacts whose closing speed times e exceeds DART_BOUNCING_VELOCITY_THRESHOLD bounce (ContactConstraint.cpp:395-418), the
    others do not; the clamping rows that bounced carry 1 + e in getBounceDiagonals."""
    w = skeletons.WorldDescription()
    _floor(w)
    for sk in w.skeletons:
        for b in sk.bodies:
            b.setRestitutionCoeff(1.0)
    _xyz_body(w, "a", box=[1, 1, 1], restitution=0.8)
    _xyz_body(w, "b", box=[0.5, 0.5, 0.5], mass=0.7, restitution=0.9)
    return w


def bouncing_boxes_states(N, seed):
    rng = np.random.default_rng(seed)
    pen = lambda: rng.uniform(1e-4, 3e-3, N)
    q = np.stack([rng.uniform(-3, -2, N), rng.uniform(-1, 1, N), 0.5 - pen(),
                  rng.uniform(2, 3, N), rng.uniform(-1, 1, N), 0.25 - pen()], axis=1)
    qd = rng.uniform(-0.5, 0.5, (N, 6))
    qd[:, 2] = -rng.uniform(0.0, 1.5, N)     # some slower than 0.1 / e (no bounce), most faster
    qd[:, 5] = -rng.uniform(0.3, 2.0, N)
    tau = rng.uniform(-1, 1, (N, 6))
    return q, qd, tau


SCENES = {
    "bouncing_boxes": (bouncing_boxes, bouncing_boxes_states),
    "fixed_base_arm_on_floor": (fixed_base_arm_on_floor, fixed_base_arm_on_floor_states),
    "box_stack": (box_stack, box_stack_states),
    "free_box_on_floor": (free_box_on_floor, free_box_on_floor_states),
    "box_on_pedestal": (box_on_pedestal, box_on_pedestal_states),
    "sliding_box": (sliding_box, sliding_box_states),
    "two_box_stack": (two_box_stack, two_box_stack_states),
    "separate_groups": (separate_groups, separate_groups_states),
    "pendulum_on_floor": (pendulum_on_floor, pendulum_on_floor_states),
    "sphere_pendulum_on_floor": (sphere_pendulum_on_floor, sphere_pendulum_on_floor_states),
}


# ---- BASELINE configs[4]: Atlas with box feet on the ground box ----
# The model files are the reference's own data (data/sdf/atlas/atlas_v3_box_colliders.urdf, ground.urdf). They are not part of
# this repo: the scene exists when they have been copied to baseline/_ref/atlas/ (git-ignored; it travels to the GPU box):
#   mkdir -p baseline/_ref/atlas && cp /root/reference/data/sdf/atlas/{atlas_v3_box_colliders,ground}.urdf baseline/_ref/atlas/
import math
import os

ATLAS_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "atlas")


def atlas_available():
    return all(os.path.exists(os.path.join(ATLAS_DIR, f)) for f in ("atlas_v3_box_colliders.urdf", "ground.urdf"))


def atlas_on_ground():
    from diffdart_b200.urdf_loader import load_urdf

    atlas, _ = load_urdf(os.path.join(ATLAS_DIR, "atlas_v3_box_colliders.urdf"))
    ground, _ = load_urdf(os.path.join(ATLAS_DIR, "ground.urdf"))
    return skeletons.WorldDescription().addSkeleton(atlas).addSkeleton(ground).setGravity([0.0, -9.81, 0.0])


def atlas_on_ground_states(N, seed):
    """Pose of unittests/unit/test_AtlasGradients.cpp:222-223 (root rotated by -pi/2 about x, lowered by 0.01) + U(-0.01, 0.01),
    qd ~ U(-0.1, 0.1), tau = 0."""
    rng = np.random.default_rng(seed)
    q = np.zeros((N, 33))
    q[:, 0] = -0.5 * math.pi
    q[:, 4] = -0.01
    q += rng.uniform(-0.01, 0.01, (N, 33))
    return q, rng.uniform(-0.1, 0.1, (N, 33)), np.zeros((N, 33))


# ---- BASELINE configs[3]: a half-cheetah-style URDF walker on the ground (tests/fixtures/walker.urdf, written for this repo) ----
WALKER_URDF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", "walker.urdf")


def walker_on_ground(max_contacts=16):
    from diffdart_b200.urdf_loader import load_urdf

    walker, _ = load_urdf(WALKER_URDF)
    w = skeletons.WorldDescription().addSkeleton(walker).setMaxContacts(max_contacts)
    _floor(w)
    return w


def walker_on_ground_states(N, seed):
    """Standing pose with both feet 0-4 mm inside the floor (foot boxes reach 0.06 m below the ankles: back leg 0.59 + 0.06, front leg
    0.48 + 0.06, so the torso is pitched to level them), small joint noise, qd ~ U(-0.3, 0.3), joint torques ~ U(-5, 5) on the legs."""
    rng = np.random.default_rng(seed)
    q = np.zeros((N, 9))
    pitch = np.arcsin((0.65 - 0.54) / 1.0)          # front hip lower than the back hip
    q[:, 2] = pitch + rng.uniform(-0.002, 0.002, N)
    q[:, 1] = 0.595 - rng.uniform(5e-4, 4e-3, N)     # ankles at 0.06 - depth: the foot boxes dip 0.5-4 mm into the floor
    q[:, 0] = rng.uniform(-0.5, 0.5, N)
    q[:, 3:] = rng.uniform(-0.01, 0.01, (N, 6))
    q[:, 3:6] -= q[:, 2:3] * np.array([1.0, 0.0, 0.0])   # legs hang vertically: hip angle cancels the pitch
    q[:, 6:9] -= q[:, 2:3] * np.array([1.0, 0.0, 0.0])
    qd = rng.uniform(-0.3, 0.3, (N, 9))
    tau = np.zeros((N, 9))
    tau[:, 3:] = rng.uniform(-5, 5, (N, 6))
    return q, qd, tau


# ---- capsules: a planar two-legged hopper in the reference's .skel format (tests/fixtures/capsule_hopper.skel, written for this repo) ----
HOPPER_SKEL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", "capsule_hopper.skel")


def capsule_hopper_on_ground(max_contacts=8):
    from diffdart_b200.skel_loader import load_skel

    return load_skel(HOPPER_SKEL)[0].setMaxContacts(max_contacts)


def capsule_hopper_on_ground_states(N, seed):
    """Standing on the end caps of both shins, 0-5 mm inside the ground (shin ends reach 0.11 below the root height at the zero pose),
    small joint angles so that the cap stays the closest feature; qd ~ U(-0.3, 0.3), joint torques ~ U(-3, 3)."""
    rng = np.random.default_rng(seed)
    q = np.zeros((N, 7))
    q[:, 0] = rng.uniform(-0.5, 0.5, N)
    q[:, 1] = -0.1105 - rng.uniform(0.0, 5e-3, N)
    q[:, 2:] = rng.uniform(-0.05, 0.05, (N, 5))
    q[:, 3] -= q[:, 2]
    q[:, 5] -= q[:, 2]
    qd = rng.uniform(-0.3, 0.3, (N, 7))
    tau = np.zeros((N, 7))
    tau[:, 3:] = rng.uniform(-3, 3, (N, 4))
    return q, qd, tau


SCENES["capsule_hopper_on_ground"] = (capsule_hopper_on_ground, capsule_hopper_on_ground_states)

HALF_CHEETAH_SKEL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "skel", "half_cheetah.skel")


def half_cheetah_available():
    """The reference's data/skel/half_cheetah.skel (BASELINE configs[3]) is not part of this repo: the scene exists when it has been copied
    to baseline/_ref/skel/ (git-ignored; it travels to the GPU box):
      mkdir -p baseline/_ref/skel && cp /root/reference/data/skel/half_cheetah.skel baseline/_ref/skel/"""
    return os.path.exists(HALF_CHEETAH_SKEL)


def half_cheetah_on_ground(max_contacts=8):
    from diffdart_b200.skel_loader import load_skel

    return load_skel(HALF_CHEETAH_SKEL)[0].setMaxContacts(max_contacts)


def half_cheetah_on_ground_states(N, seed):
    """The model's zero pose lowered until the end caps of both foot capsules are 0-6 mm inside the ground (they reach 0.10-0.12 below
    the root at the zero pose), small joint noise, qd ~ U(-0.3, 0.3), joint torques ~ U(-20, 20) on the six leg joints
    (the reference's half_cheetah_bench.py drives the same joints)."""
    rng = np.random.default_rng(seed)
    q = np.zeros((N, 9))
    q[:, 0] = rng.uniform(-0.5, 0.5, N)
    q[:, 1] = -0.120 - rng.uniform(0.0, 6e-3, N)
    q[:, 2:] = rng.uniform(-0.02, 0.02, (N, 7))
    qd = rng.uniform(-0.3, 0.3, (N, 9))
    tau = np.zeros((N, 9))
    tau[:, 3:] = rng.uniform(-20, 20, (N, 6))
    return q, qd, tau


# scenes only the GPU tests run (the host emulation of a 35-body world is slow)
GPU_SCENES = dict(SCENES)
GPU_SCENES["walker_on_ground"] = (walker_on_ground, walker_on_ground_states)
if atlas_available():
    GPU_SCENES["atlas_on_ground"] = (atlas_on_ground, atlas_on_ground_states)
if half_cheetah_available():
    GPU_SCENES["half_cheetah_on_ground"] = (half_cheetah_on_ground, half_cheetah_on_ground_states)
