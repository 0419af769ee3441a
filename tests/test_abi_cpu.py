"""The C-ABI library loads and exports every symbol include/ddb.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

from diffdart_b200 import _abi, skeletons

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ddb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddb_[a-z_0-9]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    lib = _abi.load_library()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libddb.so does not export %s" % n
    assert sorted(_abi.EXPORTED_SYMBOLS) == names
    assert lib.ddb_abi_version() == 1


def test_default_params_match_reference_world_defaults():
    lib = _abi.load_library()
    p = _abi.Params()
    lib.ddb_default_params(C.byref(p))
    assert list(p.gravity) == [0.0, 0.0, -9.81]  # World.cpp:71
    assert p.time_step == 1e-3  # World.cpp:72
    assert p.contact_clipping_depth == 0.03  # World.cpp:84


def test_struct_layout_matches_header():
    # sizeof(ddb_body) from the header's field list: 6 int32 + (3 + 12 + 12 + 1 + 3 + 6 + 1 + 1 + 3 + 12) doubles
    assert C.sizeof(_abi.Body) == 6 * 4 + 54 * 8
    assert C.sizeof(_abi.Transform) == 96
    assert C.sizeof(_abi.Params) == 5 * 8 + 8 * 4


def test_no_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _abi.load_library()
    topo, params, keep = skeletons.create_cartpole().to_ctypes()
    h = _abi._vp()
    rc = lib.ddb_create(C.byref(topo), C.byref(params), 4, 0, C.byref(h))
    assert rc == 3  # DDB_ERR_NO_DEVICE: there is no CPU path
    from diffdart_b200.world import BatchedWorld

    with pytest.raises(RuntimeError):
        BatchedWorld(skeletons.create_cartpole(), 4)


def test_topology_flattening():
    d = skeletons.create_box_stack(5)
    topo, params, keep = d.to_ctypes()
    assert topo.n_bodies == 6 and topo.n_dofs == 30
    assert topo.bodies[0].joint_type == _abi.JOINT_WELD and topo.bodies[0].parent == -1
    assert [topo.bodies[i].dof_offset for i in range(1, 6)] == [0, 6, 12, 18, 24]
    d = skeletons.create_multiarm_robot(5)
    topo, params, keep = d.to_ctypes()
    assert [topo.bodies[i].parent for i in range(5)] == [-1, 0, 1, 2, 3]
    assert list(params.gravity) == [0.0, -9.81, 0.0]
