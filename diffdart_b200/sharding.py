"""Sharding of independent worlds across GPUs.

Worlds never exchange data inside a step or a rollout, so the only distributed pieces are
(a) the static partition of the N worlds over the ranks and (b) an optional all-reduce of
gradients with respect to parameters shared by all worlds (e.g. ``lossWrtMass``,
dart/neural/NeuralUtils.hpp:24-30, or policy weights) once per backward sweep.
"""
import torch
import torch.distributed as dist


def shard_range(n_worlds, rank, world_size):
    """Contiguous block [lo, hi) of worlds owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_worlds, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank, world_size):
    """Rows of a [N, ...] batch owned by ``rank``."""
    lo, hi = shard_range(t.shape[0], rank, world_size)
    return t[lo:hi]


def allreduce_shared_gradient(per_world_grad):
    """Sum a [N_local, p] per-world gradient over the local worlds, then over the ranks
    (``ncclAllReduce(sum, float64)`` over NVLink with the nccl backend; gloo on CPU).
    The payload is p doubles, so it is latency-bound: call it once per backward sweep."""
    g = per_world_grad.sum(dim=0)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
    return g
