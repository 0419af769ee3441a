"""``dart_layer(world, pos, vel, torque, pointer=None)`` for a batch of worlds.

Source-compatible with python/diffdart/dart_layer.py:13-96 of the reference: same call,
same return ``(next_pos, next_vel)``, same backward ``(None, dL/dpos, dL/dvel, dL/dtorque,
None)``. ``world`` is a :class:`diffdart_b200.world.BatchedWorld`; tensors are float64 CUDA
tensors of shape ``[N, n]`` (``[n]`` with a single world reproduces the reference's call).
"""
import torch

from .world import BackpropSnapshot, LossGradient, forwardPass  # noqa: F401


class BackpropSnapshotPointer:
    def __init__(self):
        self.backprop_snapshot = None


class DartLayer(torch.autograd.Function):
    """One differentiable timestep of every world as a PyTorch layer."""

    @staticmethod
    def forward(ctx, world, pos, vel, torque, snapshot_pointer):
        world.setPositions(pos.detach())
        world.setVelocities(vel.detach())
        world.setForces(torque.detach())
        backprop_snapshot = forwardPass(world)
        ctx.backprop_snapshot = backprop_snapshot
        ctx.world = world
        if snapshot_pointer is not None:
            snapshot_pointer.backprop_snapshot = backprop_snapshot
        return world.getPositions(), world.getVelocities()

    @staticmethod
    def backward(ctx, grad_pos, grad_vel):
        backprop_snapshot = ctx.backprop_snapshot
        world = ctx.world
        nextTimestepGrad = LossGradient()
        nextTimestepGrad.lossWrtPosition = grad_pos.detach()
        nextTimestepGrad.lossWrtVelocity = grad_vel.detach()
        thisTimestepGrad = LossGradient()
        backprop_snapshot.backprop(world, thisTimestepGrad, nextTimestepGrad)
        return (None, thisTimestepGrad.lossWrtPosition, thisTimestepGrad.lossWrtVelocity, thisTimestepGrad.lossWrtTorque, None)


def dart_layer(world, pos, vel, torque, pointer=None):
    """Forward pass on ``world``, storing what the backward pass needs."""
    return DartLayer.apply(world, pos, vel, torque, pointer)
