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
