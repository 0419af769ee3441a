"""Seeded box / sphere poses for the narrow-phase tests + numpy restatements of the sphere cases."""
import math

import numpy as np


def rot(rng, small=False):
    a = rng.normal(size=3)
    a /= np.linalg.norm(a)
    t = rng.uniform(-0.2, 0.2) if small else rng.uniform(-3, 3)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(t) * K + (1 - math.cos(t)) * K @ K


def box_pair(rng, mode):
    """mode 0: arbitrary poses; 1: axis-aligned stack; 2: yawed stack (the box-stack config); 3: slightly tilted stack."""
    s0, s1 = rng.uniform(0.3, 1.5, 3), rng.uniform(0.3, 1.5, 3)
    if mode == 0:
        R0, R1 = rot(rng), rot(rng)
        p0 = rng.uniform(-.3, .3, 3)
        p1 = p0 + rng.uniform(-1, 1, 3)
    elif mode == 1:
        R0 = R1 = np.eye(3)
        p0 = np.zeros(3)
        p1 = np.array([rng.uniform(-.3, .3), rng.uniform(-.3, .3), 0.5 * (s0[2] + s1[2]) - rng.uniform(0, 0.02)])
    elif mode == 2:
        th = rng.uniform(-0.3, 0.3)
        R0 = np.eye(3)
        R1 = np.array([[math.cos(th), -math.sin(th), 0], [math.sin(th), math.cos(th), 0], [0, 0, 1]])
        p0 = np.zeros(3)
        p1 = np.array([rng.uniform(-.2, .2), rng.uniform(-.2, .2), 0.5 * (s0[2] + s1[2]) - rng.uniform(0, 0.02)])
    else:
        R0, R1 = rot(rng, True), rot(rng, True)
        p0 = np.zeros(3)
        p1 = np.array([rng.uniform(-.2, .2), rng.uniform(-.2, .2), 0.5 * (s0[2] + s1[2]) - rng.uniform(-0.01, 0.03)])
    return s0, np.concatenate([R0.reshape(9), p0]), s1, np.concatenate([R1.reshape(9), p1])


def np_sphere_sphere(r0, c0, r1, c1):
    """collideSphereSphere (DARTCollide.cpp:1735-1800): returns (point, normal, depth) or None."""
    rsum = r0 + r1
    n = c0 - c1
    nsq = n @ n
    if nsq > rsum * rsum:
        return None
    point = (r1 / rsum) * c0 + (r0 / rsum) * c1
    if nsq < 1e-6:
        return point, np.zeros(3), rsum
    d = math.sqrt(nsq)
    return point, n * (1.0 / d), rsum - d


def np_box_sphere(size, Rb, pb, r, c0, sphere_first):
    """collideBoxSphere / collideSphereBox (DARTCollide.cpp:1429-1733): (point, normal, depth, type) or None."""
    half = 0.5 * size
    p = Rb.T @ (c0 - pb)
    inside = True
    for k in range(3):
        if p[k] < -half[k]:
            p[k], inside = -half[k], False
        if p[k] > half[k]:
            p[k], inside = half[k], False

    def side():
        m = half - np.abs(p)
        idx = 0
        if m[1] < m[0]:
            idx = 1
        if m[2] < m[idx]:
            idx = 2
        return idx, m[idx]

    def axis_normal(idx):
        e = np.zeros(3)
        e[idx] = (1.0 if p[idx] > 0 else -1.0) if sphere_first else (-1.0 if p[idx] > 0 else 1.0)
        return Rb @ e

    if inside:
        idx, m = side()
        return c0, axis_normal(idx), m + r, (1 if sphere_first else 2)
    cp = Rb @ p + pb
    n = (c0 - cp) if sphere_first else (cp - c0)
    mag = np.linalg.norm(n)
    if r - mag < 0:
        return None
    if mag > 1e-6:
        n = n * (1.0 / mag)
    else:
        n = axis_normal(side()[0])
    return cp, n, r - mag, (4 if sphere_first else 5)
