"""T-step rollouts of a BatchedWorld: the caller of the hot path in the reference's trajectory code.

  dart_rollout(world, q0, qd0, taus) -> (poses, vels)     differentiable; poses[t], vels[t] = state after step t
      forward  = SingleShot::getSnapshots        (dart/trajectory/SingleShot.cpp:719-760): T forwardPass calls, one retained
                 BackpropSnapshot per step
      backward = SingleShot::backpropGradientWrt (SingleShot.cpp:650-700): the loss gradient w.r.t. every step's state (what
                 DartTorchLossFn hands back, python/diffdart/dart_torch_loss_fn.py:36-61) is pulled back step by step
  Both run inside libddb.so (ddb_rollout_forward / ddb_rollout_backward): one call per sweep, chunks of worlds side by side on the
  handle's lanes. Tensors: q0, qd0 [N, n]; taus, poses, vels [T, N, n]; float64 CUDA.
"""
import torch

from .world import _check


class _Rollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, world, q0, qd0, taus):
        T = taus.shape[0]
        q0c, qd0c, tc = world._as_batch(q0), world._as_batch(qd0), taus.to(dtype=torch.float64, device=world.device).contiguous()
        if tuple(tc.shape) != (T, world.N, world.n):
            raise ValueError("taus must be [T, %d, %d], got %s" % (world.N, world.n, tuple(tc.shape)))
        poses, vels = torch.empty_like(tc), torch.empty_like(tc)
        rc = world._lib.ddb_rollout_forward(world._h, world._stream(), T, q0c.data_ptr(), qd0c.data_ptr(), tc.data_ptr(), poses.data_ptr(),
                                            vels.data_ptr())
        _check(world._lib, world._h, rc, "ddb_rollout_forward")
        ctx.world, ctx.T = world, T
        return poses, vels

    @staticmethod
    def backward(ctx, g_poses, g_vels):
        world, T = ctx.world, ctx.T
        z = lambda g: torch.zeros(T, world.N, world.n, dtype=torch.float64, device=world.device) if g is None else g.contiguous()
        gp, gv = z(g_poses), z(g_vels)
        gq0 = torch.empty(world.N, world.n, dtype=torch.float64, device=world.device)
        gqd0, gtau = torch.empty_like(gq0), torch.empty_like(gp)
        rc = world._lib.ddb_rollout_backward(world._h, world._stream(), T, gp.data_ptr(), gv.data_ptr(), gq0.data_ptr(), gqd0.data_ptr(),
                                             gtau.data_ptr())
        _check(world._lib, world._h, rc, "ddb_rollout_backward")
        return None, gq0, gqd0, gtau


def dart_rollout(world, q0, qd0, taus):
    """Roll every world of `world` T = taus.shape[0] steps forward from (q0, qd0); returns (poses [T,N,n], vels [T,N,n]).
    The world's own state and the snapshot slots 0..T-1 of its handle are overwritten (slots held by BackpropSnapshot
    objects from forwardPass must not be in use)."""
    poses, vels = _Rollout.apply(world, q0, qd0, taus)
    world._q, world._qd = poses[-1].detach(), vels[-1].detach()
    return poses, vels
