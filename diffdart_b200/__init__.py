"""diffdart_b200 — B200-native batched differentiable rigid-body timestep.

A drop-in for the path behind the reference's ``dart::neural::forwardPass`` /
``BackpropSnapshot`` / Python ``dart_layer(world, pos, vel, forces)``, executed for
N independent worlds sharing one skeleton topology by hand-written sm_100a CUDA
kernels behind the C ABI in ``include/ddb.h``.
"""
from . import _abi, skeletons  # noqa: F401

__all__ = ["_abi", "skeletons", "world", "dart_layer"]
