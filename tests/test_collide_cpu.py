"""CPU: the device narrow-phase source (diffdart_b200/csrc/collide.cuh) compiled for the host is bit-identical to the
REFERENCE's own dBoxBox (dart/collision/dart/DARTCollide.cpp:49-1427 compiled through oracle/ref_boxbox_shim_*.h)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from collide_cases import box_pair
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emul_box_box():
    subprocess.run(["make", "-C", os.path.join(HERE, "host_emulation"), "-s"], check=True)
    lib = C.CDLL(os.path.join(HERE, "host_emulation", "_build", "libemul.so"))

    def f(s0, T0, s1, T1):
        out = np.zeros(26 * 8)
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (s0, T0, s1, T1)]
        n = lib.emul_box_box(a[0].ctypes.data_as(dp), a[1].ctypes.data_as(dp), a[2].ctypes.data_as(dp), a[3].ctypes.data_as(dp), out.ctypes.data_as(dp))
        return out[:26 * n].reshape(n, 26)

    return f


def test_reference_box_box_known_answers():
    """1x1x0.5 box resting 1 mm deep on the 10x10x0.01 floor of examples/box_stacking: 4 FACE_VERTEX contacts at the
    box's bottom corners, normal -z (from box toward floor... object 2 -> object 1), depth 1e-3."""
    T0 = np.concatenate([np.eye(3).reshape(9), [0, 0, -0.005]])
    T1 = np.concatenate([np.eye(3).reshape(9), [0.1, -0.2, 0.249]])
    c = orc.ref_collide_box_box([10, 10, 0.01], T0, [1, 1, 0.5], T1)
    assert c.shape == (4, 26)
    assert np.allclose(c[:, 24], 1e-3, atol=1e-15) and np.all(c[:, 25] == 2)
    assert np.allclose(c[:, 3:6], [0, 0, -1])
    assert np.allclose(sorted(map(tuple, np.round(c[:, 0:2], 12))), sorted([(-0.4, -0.7), (-0.4, 0.3), (0.6, -0.7), (0.6, 0.3)]))
    # two equal boxes stacked with a yaw: the overlap polygon has 8 vertices, all EDGE_EDGE
    th = 0.1
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    c = orc.ref_collide_box_box([1, 1, 0.5], np.concatenate([np.eye(3).reshape(9), [0, 0, 0.25]]), [1, 1, 0.5],
                                np.concatenate([R.reshape(9), [0.02, 0.01, 0.749]]))
    assert c.shape[0] == 8 and np.all(c[:, 25] == 3)


def test_device_box_box_bit_identical_to_reference(emul_box_box):
    rng = np.random.default_rng(0)
    hist = {}
    for t in range(6000):
        s0, T0, s1, T1 = box_pair(rng, t % 4)
        want = orc.ref_collide_box_box(s0, T0, s1, T1)
        got = emul_box_box(s0, T0, s1, T1)
        hist[len(want)] = hist.get(len(want), 0) + 1
        assert got.shape == want.shape and np.array_equal(got, want), t
    assert all(k in hist for k in range(0, 8))  # every contact count 0..7 was exercised


def test_capsule_box_against_a_brute_force_of_the_geometry():
    """Capsule against box (collide.cuh: capsuleBox, the sphere branch of the reference's collideCapsuleBox / collideBoxCapsule,
    DARTCollide.cpp:4217-4361). The reference picks the end sphere from libccd's MPR position (libccd is not in the reference tree:
    parity unpinned), so the closed-form choice is validated geometrically: a dense sampling of the capsule's axis must have its nearest
    point to the box at the end that was chosen; the contact then is the sphere-box contact of that end (point on the box, unit normal,
    depth = radius - distance); where the nearest point is strictly inside the axis the pair is reported as not carried (-1), or as a
    miss when the box is farther away than the radius."""
    subprocess.run(["make", "-C", os.path.join(HERE, "host_emulation"), "-s"], check=True)
    lib = C.CDLL(os.path.join(HERE, "host_emulation", "_build", "libemul.so"))
    rng = np.random.default_rng(4)
    ip = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)

    def rot(w):
        th = np.linalg.norm(w)
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K

    def box_dist(p, half):  # distance of points p [n, 3] (box frame) to the box
        return np.linalg.norm(np.maximum(np.abs(p) - half, 0.0), axis=1)

    seen = {"end": 0, "pipe": 0, "miss": 0}
    for t in range(3000):
        size = rng.uniform(0.3, 2.0, 3)
        Rb, pb = rot(rng.normal(size=3)), rng.uniform(-0.5, 0.5, 3)
        r, h = rng.uniform(0.03, 0.2), rng.uniform(0.1, 1.0)
        Rc = rot(rng.normal(size=3))
        # place the capsule near the box surface: centre at a random direction, at about the box's extent
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        pc = pb + Rb @ (d * (0.5 * np.linalg.norm(size) * rng.uniform(0.5, 1.2)))
        Tb, Tc = np.concatenate([Rb.reshape(9), pb]), np.concatenate([Rc.reshape(9), pc])
        for capsule_first in (True, False):
            out = np.zeros(26 * 8)
            args = (3, ip([r, h, 0]), ip(Tc), 1, ip(size), ip(Tb)) if capsule_first else (1, ip(size), ip(Tb), 3, ip([r, h, 0]), ip(Tc))
            n = lib.emul_collide_pair(*args, out.ctypes.data_as(dp))
            ts = np.linspace(-0.5 * h, 0.5 * h, 2001)
            axis_pts = pc + np.outer(ts, Rc[:, 2])
            dist = box_dist((axis_pts - pb) @ Rb, 0.5 * size)
            k = int(np.argmin(dist))
            if n == 1:
                seen["end"] += 1
                c = out[:26]
                centre = c[9:12]                                  # sphereCenter travels in edgeAFixed
                k = 0 if np.allclose(centre, axis_pts[0], atol=1e-12) else 2000
                assert np.allclose(centre, axis_pts[k], atol=1e-12) and dist[k] <= dist.min() + 1e-12, t   # the chosen end attains the minimum
                if dist[k] > 1e-9:                                # centre outside the box: point = nearest box point, depth = r - distance
                    assert abs(np.linalg.norm(c[3:6]) - 1) < 1e-12 and abs(c[24] - (r - dist[k])) < 1e-9, t
                    assert box_dist(((c[0:3] - pb) @ Rb)[None], 0.5 * size)[0] < 1e-9
                    sign = 1.0 if capsule_first else -1.0         # normal from object 2 to object 1
                    assert np.allclose(c[3:6], sign * (centre - c[0:3]) / dist[k], atol=1e-9)
                assert int(c[25]) in ((4, 1) if capsule_first else (5, 2))
            elif n == 0:
                seen["miss"] += 1
                assert dist.min() > r - 1e-6, t
            else:
                seen["pipe"] += 1
                assert n == -1 and dist.min() <= r + 1e-6 and (0 < k < 2000 or dist[k] == 0.0 or abs(dist[0] - dist[2000]) < 1e-6), t
    assert min(seen.values()) > 50, seen
