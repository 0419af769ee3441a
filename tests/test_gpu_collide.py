"""GPU: ddb_collide_pairs against the reference's own box-box (bit-exact) and numpy restatements of the sphere cases."""
import ctypes as C

import numpy as np
import pytest
import torch

from collide_cases import box_pair, np_box_sphere, np_sphere_sphere, rot
from diffdart_b200 import _abi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def cuda(a, dtype=torch.float64):
    return torch.tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


def collide(shape1, size1, T1, shape2, size2, T2):
    lib = _abi.load_library()
    N = len(shape1)
    out = torch.zeros(N, 8, 26, dtype=torch.float64, device="cuda")
    cnt = torch.zeros(N, dtype=torch.int32, device="cuda")
    a = [cuda(shape1, torch.int32), cuda(size1), cuda(T1), cuda(shape2, torch.int32), cuda(size2), cuda(T2)]
    rc = lib.ddb_collide_pairs(0, None, N, *[t.data_ptr() for t in a], out.data_ptr(), cnt.data_ptr())
    assert rc == 0
    return out.cpu().numpy(), cnt.cpu().numpy()


def test_box_box_bit_exact_vs_reference():
    rng = np.random.default_rng(11)
    pairs = [box_pair(rng, t % 4) for t in range(8000)]
    s0, T0, s1, T1 = (np.array(x) for x in zip(*pairs))
    out, cnt = collide(np.ones(len(pairs)), s0, T0, np.ones(len(pairs)), s1, T1)
    for k, (a, b, c, d) in enumerate(pairs):
        want = orc.ref_collide_box_box(a, b, c, d)
        assert cnt[k] == len(want), k
        assert np.array_equal(out[k, :cnt[k]], want), k


def test_sphere_cases():
    rng = np.random.default_rng(12)
    N = 3000
    shape1, shape2 = np.zeros(N, dtype=np.int32), np.zeros(N, dtype=np.int32)
    size1, size2, T1, T2 = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 12)), np.zeros((N, 12))
    want = []
    for k in range(N):
        kind = k % 3
        Ra, Rb = rot(rng), rot(rng)
        pa = rng.uniform(-.2, .2, 3)
        if kind == 0:  # sphere-sphere
            r0, r1 = rng.uniform(0.2, 1, 2)
            pb = pa + rng.normal(size=3) * 0.6 * (0 if k % 31 == 0 else 1)
            shape1[k], shape2[k], size1[k, 0], size2[k, 0] = 2, 2, r0, r1
            want.append(("ss", np_sphere_sphere(r0, pa, r1, pb)))
        else:
            size = rng.uniform(0.4, 1.2, 3)
            r = rng.uniform(0.1, 0.6)
            c = pa + Ra @ (rng.uniform(-1, 1, 3) * (0.5 * size + 0.8 * r))
            if kind == 1:  # box (1) - sphere (2)
                shape1[k], shape2[k], size1[k], size2[k, 0], pb, Rb = 1, 2, size, r, c, Rb
                want.append(("bs", np_box_sphere(size, Ra, pa, r, c, False)))
            else:  # sphere (1) - box (2)
                shape1[k], shape2[k], size1[k, 0], size2[k] = 2, 1, r, size
                want.append(("sb", np_box_sphere(size, Ra, pa, r, c, True)))
                T1[k], T2[k] = np.concatenate([Rb.reshape(9), c]), np.concatenate([Ra.reshape(9), pa])
                continue
        T1[k], T2[k] = np.concatenate([Ra.reshape(9), pa]), np.concatenate([Rb.reshape(9), pb])
    out, cnt = collide(shape1, size1, T1, shape2, size2, T2)
    hits = 0
    for k, (kind, w) in enumerate(want):
        if w is None:
            assert cnt[k] == 0, (k, kind)
            continue
        hits += 1
        assert cnt[k] == 1, (k, kind)
        assert np.allclose(out[k, 0, 0:3], w[0], rtol=1e-10, atol=1e-12) and np.allclose(out[k, 0, 3:6], w[1], rtol=1e-7, atol=1e-9)
        assert np.isclose(out[k, 0, 24], w[2], rtol=1e-10, atol=1e-12)
        assert out[k, 0, 25] == (6 if kind == "ss" else w[3])
    assert hits > N // 4
