"""Generates tests/golden/*.json from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box, the committed JSON files do).

  geometry.json   inputs/outputs of the reference's scalar-expanded math::transformInertia and
                  math::expAngular (dart/math/Geometry.cpp:1668-1702,1769-1851). Those two functions
                  are written as plain scalar expressions over _T(i,j) / _I(i,j) / _s[i], so their
                  bodies are transliterated token-for-token into Python at generation time and
                  evaluated on seeded random inputs — the outputs are the reference's arithmetic.
  lcp_instances.json  the literal (A, x, lo, hi, b, fIndex) boxed-LCP instances held by the
                  reference's own tests (unittests/unit/test_LCPUtils.cpp:51-759), together with
                  the answer of the reference's unmodified ODE Dantzig solver
                  (oracle/_ref/libodelcp_ref.so) on each.

usage: python tests/golden/make_golden.py
"""
import json
import math
import os
import re
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def _function_body(src, signature):
    i = src.index(signature)
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        if src[k] == "{":
            depth += 1
        elif src[k] == "}":
            depth -= 1
            if depth == 0:
                return src[j + 1:k]
        k += 1


def geometry():
    src = open(os.path.join(REF, "dart/math/Geometry.cpp")).read()
    rng = np.random.default_rng(20261018)
    out = {"transformInertia": [], "expAngular": []}

    # --- transformInertia ---
    body = _function_body(src, "Inertia transformInertia(const Eigen::Isometry3d& _T, const Inertia& _I)")
    body = re.sub(r"//.*", "", body)
    stmts = [s.strip() for s in body.split(";") if s.strip()]
    prog = []
    for s in stmts:
        s = " ".join(s.split())
        if s.startswith("Inertia ret") or s.startswith("return") or "triangularView" in s:
            continue
        s = re.sub(r"^double ", "", s)
        s = re.sub(r"_T\((\d), (\d)\)", r"T[\1][\2]", s)
        s = re.sub(r"_I\((\d), (\d)\)", r"I[\1][\2]", s)
        s = re.sub(r"ret\((\d), (\d)\)", r"ret[\1][\2]", s)
        prog.append(s)
    code = "\n".join(prog)
    for _ in range(8):
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        th = rng.uniform(-3, 3)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * K @ K
        p = rng.uniform(-1, 1, size=3)
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = p
        B = rng.normal(size=(6, 6))
        I = B @ B.T + np.eye(6)
        env = {"T": T.tolist(), "I": I.tolist(), "ret": np.eye(6).tolist()}
        exec(code, {}, env)
        ret = np.array(env["ret"])
        ret = np.triu(ret) + np.triu(ret, 1).T  # the function fills the upper triangle, then mirrors
        out["transformInertia"].append({"T": np.concatenate([R.reshape(9), p]).tolist(), "I": I.tolist(), "out": ret.tolist()})

    # --- expAngular ---
    body = _function_body(src, "Eigen::Isometry3d expAngular(const Eigen::Vector3d& _s)")
    body = re.sub(r"//.*", "", body)
    body = body.replace("Eigen::Isometry3d ret = Eigen::Isometry3d::Identity();", "")
    body = re.sub(r"double (\w+)\[\] = \{(.*?)\};", r"\1 = [\2];", body, flags=re.S)
    body = re.sub(r"double (\w+) =", r"\1 =", body)
    body = body.replace("_s[", "s[").replace("return ret;", "")
    body = re.sub(r"ret\((\d), (\d)\)", r"ret[\1][\2]", body)
    # the if/else on theta
    m = re.search(r"if \(theta > DART_EPSILON\)\s*\{(.*?)\}\s*else\s*\{(.*?)\}", body, flags=re.S)
    big, small = m.group(1), m.group(2)
    pre, post = body[:m.start()], body[m.end():]

    def clean(block):
        return "\n".join(" ".join(x.split()) for x in block.split(";") if x.strip())

    for k in range(10):
        s = rng.normal(size=3) * (1e-7 if k < 2 else rng.uniform(0.1, 3.0))
        env = {"s": s.tolist(), "ret": np.eye(3).tolist(), "sqrt": math.sqrt, "cos": math.cos, "sin": math.sin}
        exec(clean(pre), {}, env)
        exec(clean(big if env["theta"] > 1e-6 else small), {}, env)
        exec(clean(post), {}, env)
        out["expAngular"].append({"s": s.tolist(), "out": env["ret"]})
    return out


def lcp_instances():
    from oracle import oracle as orc
    src = open(os.path.join(REF, "unittests/unit/test_LCPUtils.cpp")).read()
    tests = re.split(r"\nTEST\(LCP_UTILS, (\w+)\)", src)
    cases = []
    inf = float("inf")

    def values(txt):
        txt = re.sub(r"//.*", "", txt)
        txt = txt.replace("std::numeric_limits<double>::infinity()", "inf")
        return [float(eval(t, {"inf": inf})) for t in txt.replace("\n", " ").split(",") if t.strip()]

    for name, body in zip(tests[1::2], tests[2::2]):
        got = {}
        for var in ("A", "x", "lo", "hi", "b", "fIndex"):
            m = re.search(r"^\s*%s\s*<<(.*?);" % var, body, flags=re.S | re.M)
            if m:
                got[var] = values(m.group(1))
        if not all(k in got for k in ("A", "lo", "hi", "b", "fIndex")):
            continue
        n = len(got["b"])
        if len(got["A"]) != n * n:
            continue
        A = np.array(got["A"]).reshape(n, n)
        fi = [int(v) for v in got["fIndex"]]
        case = {"name": name, "n": n, "A": A.tolist(), "x_cached": got.get("x", [0.0] * n), "lo": got["lo"], "hi": got["hi"],
                "b": got["b"], "findex": fi}
        for early in (False, True):
            ok, x = orc.ref_dantzig(A, got["b"], got["lo"], got["hi"], fi, early_termination=early)
            case["ref_dantzig_early%d" % int(early)] = {"ok": ok, "x": [float(v) for v in x]}
        cases.append(case)
    return cases


def _dump(name, obj):
    def enc(o):
        if isinstance(o, float) and math.isinf(o):
            return "Infinity" if o > 0 else "-Infinity"
        raise TypeError

    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1, allow_nan=True)
    print("wrote", name)


if __name__ == "__main__":
    _dump("geometry.json", geometry())
    _dump("lcp_instances.json", lcp_instances())
