"""ctypes mirror of include/ddb.h and the loader of the CUDA library.

The library is the product: if ``libddb.so`` is missing this module raises — there
is no CPU fallback (the CPU oracle under ``oracle/`` is test infrastructure only).
"""
import ctypes as C
import os

DDB_MAX_BODIES = 48
DDB_MAX_DOFS = 48

JOINT_WELD, JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_FREE = 0, 1, 2, 3
SHAPE_NONE, SHAPE_BOX, SHAPE_SPHERE, SHAPE_CAPSULE = 0, 1, 2, 3
LAYOUT_WORLD_MAJOR, LAYOUT_DOF_MAJOR = 0, 1

ERRORS = {1: "DDB_ERR_INVALID", 2: "DDB_ERR_CUDA", 3: "DDB_ERR_NO_DEVICE", 4: "DDB_ERR_STATE"}


class Transform(C.Structure):
    _fields_ = [("r", C.c_double * 9), ("p", C.c_double * 3)]


class Body(C.Structure):
    _fields_ = [
        ("parent", C.c_int32),
        ("joint_type", C.c_int32),
        ("dof_offset", C.c_int32),
        ("skeleton", C.c_int32),
        ("gravity_mode", C.c_int32),
        ("shape_type", C.c_int32),
        ("axis", C.c_double * 3),
        ("parent_to_joint", Transform),
        ("child_to_joint", Transform),
        ("mass", C.c_double),
        ("com", C.c_double * 3),
        ("moment", C.c_double * 6),
        ("friction", C.c_double),
        ("restitution", C.c_double),
        ("shape_size", C.c_double * 3),
        ("shape_offset", Transform),
    ]


class Topology(C.Structure):
    _fields_ = [
        ("n_bodies", C.c_int32),
        ("n_dofs", C.c_int32),
        ("bodies", C.POINTER(Body)),
        ("position_lower", C.POINTER(C.c_double)),
        ("position_upper", C.POINTER(C.c_double)),
        ("velocity_lower", C.POINTER(C.c_double)),
        ("velocity_upper", C.POINTER(C.c_double)),
        ("force_lower", C.POINTER(C.c_double)),
        ("force_upper", C.POINTER(C.c_double)),
    ]


class Params(C.Structure):
    _fields_ = [
        ("gravity", C.c_double * 3),
        ("time_step", C.c_double),
        ("contact_clipping_depth", C.c_double),
        ("layout", C.c_int32),
        ("n_snapshot_slots", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


def default_params():
    """Reference defaults: World.cpp:67-94."""
    p = Params()
    p.gravity[:] = [0.0, 0.0, -9.81]
    p.time_step = 1e-3
    p.contact_clipping_depth = 0.03
    p.layout = LAYOUT_WORLD_MAJOR
    p.n_snapshot_slots = 1
    return p


_LIB = None
LIB_PATH = os.environ.get("DDB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libddb.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p


def load_library():
    """Load libddb.so (built by ``__graft_entry__.build()`` / ``make -C diffdart_b200/csrc``)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "diffdart_b200: %s not found — build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    lib.ddb_abi_version.restype = C.c_int
    lib.ddb_default_params.argtypes = [C.POINTER(Params)]
    lib.ddb_default_params.restype = None
    lib.ddb_create.argtypes = [C.POINTER(Topology), C.POINTER(Params), C.c_int32, C.c_int32, C.POINTER(_vp)]
    lib.ddb_destroy.argtypes = [_vp]
    lib.ddb_last_error.argtypes = [_vp]
    lib.ddb_last_error.restype = C.c_char_p
    lib.ddb_num_worlds.argtypes = [_vp]
    lib.ddb_num_dofs.argtypes = [_vp]
    lib.ddb_forward.argtypes = [_vp, _vp] + [_vp] * 9
    lib.ddb_backward.argtypes = [_vp, _vp] + [_vp] * 5
    lib.ddb_jacobians.argtypes = [_vp, _vp] + [_vp] * 5
    lib.ddb_inv_mass_matrix.argtypes = [_vp, _vp, _vp]
    lib.ddb_snapshot_select.argtypes = [_vp, C.c_int32]
    lib.ddb_snapshot_current.argtypes = [_vp]
    lib.ddb_snapshot_reserve.argtypes = [_vp, C.c_int32, _ip]
    lib.ddb_forward_host.argtypes = [_vp] + [_vp] * 5
    lib.ddb_backward_host.argtypes = [_vp] + [_vp] * 5
    lib.ddb_jacobians_host.argtypes = [_vp] + [_vp] * 5
    lib.ddb_launch_count.argtypes = [_vp]
    lib.ddb_rollout_forward.argtypes = [_vp, _vp, C.c_int32] + [_vp] * 5
    lib.ddb_rollout_forward.restype = C.c_int
    lib.ddb_rollout_backward.argtypes = [_vp, _vp, C.c_int32] + [_vp] * 5
    lib.ddb_rollout_backward.restype = C.c_int
    lib.ddb_set_lanes.argtypes = [_vp, C.c_int32]
    lib.ddb_set_lanes.restype = C.c_int
    lib.ddb_lcp_solve.argtypes = [C.c_int32, _vp, C.c_int32, C.c_int32] + [_vp] * 6 + [C.c_int32, C.c_int32, _vp, _vp]
    lib.ddb_lcp_solve.restype = C.c_int
    lib.ddb_lcp_valid.argtypes = [C.c_int32, _vp, C.c_int32, C.c_int32] + [_vp] * 7 + [C.c_int32, _vp]
    lib.ddb_lcp_valid.restype = C.c_int
    lib.ddb_lemke_solve.argtypes = [C.c_int32, _vp, C.c_int32, C.c_int32] + [_vp] * 6
    lib.ddb_lemke_solve.restype = C.c_int
    lib.ddb_contact_capacity.argtypes = [_vp, _ip, _ip]
    lib.ddb_contact_capacity.restype = C.c_int
    lib.ddb_lcp_cache_get.argtypes = [_vp, _vp, _vp, _vp]
    lib.ddb_lcp_cache_get.restype = C.c_int
    lib.ddb_lcp_cache_set.argtypes = [_vp, _vp, _vp, _vp]
    lib.ddb_lcp_cache_set.restype = C.c_int
    lib.ddb_contact_debug_host.argtypes = [_vp, _vp, C.c_int32] + [_vp] * 9
    lib.ddb_contact_debug_host.restype = C.c_int
    lib.ddb_collide_pairs.argtypes = [C.c_int32, _vp, C.c_int32] + [_vp] * 8
    lib.ddb_collide_pairs.restype = C.c_int
    lib.ddb_measure_fp64_peak.argtypes = [C.c_int32, _dp]
    lib.ddb_measure_fp64_peak.restype = C.c_int
    lib.ddb_launch_count.restype = C.c_int64
    for name in (
        "ddb_create", "ddb_destroy", "ddb_forward", "ddb_backward", "ddb_jacobians", "ddb_inv_mass_matrix",
        "ddb_snapshot_select", "ddb_snapshot_current", "ddb_snapshot_reserve", "ddb_forward_host", "ddb_backward_host",
        "ddb_jacobians_host", "ddb_num_worlds", "ddb_num_dofs",
    ):
        getattr(lib, name).restype = C.c_int
    if lib.ddb_abi_version() != 1:
        raise RuntimeError("diffdart_b200: libddb.so ABI version mismatch")
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = [
    "ddb_default_params", "ddb_abi_version", "ddb_create", "ddb_destroy", "ddb_last_error", "ddb_num_worlds",
    "ddb_num_dofs", "ddb_forward", "ddb_backward", "ddb_jacobians", "ddb_inv_mass_matrix", "ddb_snapshot_select",
    "ddb_snapshot_current", "ddb_snapshot_reserve", "ddb_forward_host", "ddb_backward_host", "ddb_jacobians_host",
    "ddb_launch_count", "ddb_measure_fp64_peak", "ddb_lcp_solve", "ddb_lcp_valid", "ddb_collide_pairs", "ddb_contact_capacity", "ddb_contact_debug_host",
    "ddb_contact_pgs_sweeps_host", "ddb_contact_pgs_rows_host", "ddb_contact_profile_host", "ddb_rollout_forward", "ddb_rollout_backward", "ddb_set_lanes", "ddb_lemke_solve",
    "ddb_lcp_cache_get", "ddb_lcp_cache_set",
]
