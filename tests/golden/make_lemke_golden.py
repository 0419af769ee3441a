"""Writes tests/golden/lemke_instances.json: the literal (A, b) instances of the reference's Lemke tests
(unittests/unit/test_Lemke.cpp:39-185; each expects err == 0 and validate(A, f, b)), read from the reference tree where it lies.
Run in the build container (needs /root/reference): python tests/golden/make_lemke_golden.py"""
import json, os, re

SRC = "/root/reference/unittests/unit/test_Lemke.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lemke_instances.json")
text = open(SRC).read()
cases = []
for m in re.finditer(r"TEST\(Lemke, (\w+)\)(.*?)err = dart::lcpsolver::Lemke", text, flags=re.S):
    name, body = m.group(1), m.group(2)
    n = int(re.search(r"A\.resize\((\d+),", body).group(1))
    num = lambda s: [float(v) for v in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", s)]
    A = num(re.search(r"A\s*<<(.*?);", body, flags=re.S).group(1))
    b = num(re.search(r"b\s*<<(.*?);", body, flags=re.S).group(1))
    assert len(A) == n * n and len(b) == n, (name, len(A), len(b))
    cases.append({"name": name, "source": "unittests/unit/test_Lemke.cpp", "n": n, "A": A, "b": b, "expect_err": 0, "expect_validate": True})
json.dump(cases, open(OUT, "w"), indent=1)
print("wrote", OUT, [c["name"] for c in cases])
