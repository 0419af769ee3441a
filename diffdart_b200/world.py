"""Host-side mirror of the reference's interface for the batched path.

  BatchedWorld        N clones of one ``simulation::World`` sharing a skeleton topology
                      (World.cpp:109-157 ``clone()``); state accessors as World.cpp:1027-1229
  forwardPass         ``dart.neural.forwardPass(world, idempotent=False)``
                      (python/_diffdart/neural/NeuralUtils.cpp:84-88; dart/neural/NeuralUtils.cpp:26-66)
  BackpropSnapshot    python/_diffdart/neural/BackpropSnapshot.cpp:42-146
  LossGradient        python/_diffdart/neural/NeuralUtils.cpp:48-56

Tensors are float64 CUDA tensors of shape [N, n] (or [n] when N == 1, matching the
single-world reference). All compute happens in libddb.so; torch only owns device
memory and the stream. Without the CUDA library or a GPU this module raises.
"""
import ctypes as C

import torch

from . import _abi


def _check(lib, h, rc, what):
    if rc != 0:
        msg = lib.ddb_last_error(h).decode() if h else ""
        raise RuntimeError("%s failed: %s %s" % (what, _abi.ERRORS.get(rc, rc), msg))


class LossGradient:
    """dart.neural.LossGradient: lossWrtPosition / lossWrtVelocity / lossWrtTorque."""

    def __init__(self):
        self.lossWrtPosition = None
        self.lossWrtVelocity = None
        self.lossWrtTorque = None
        self.lossWrtMass = None


class BatchedWorld:
    def __init__(self, description, n_worlds=1, device=0, n_snapshot_slots=1):
        if not torch.cuda.is_available():
            raise RuntimeError("diffdart_b200 needs a CUDA device: there is no CPU fallback")
        self._lib = _abi.load_library()
        self.description = description
        self._topo, self._params, self._keep = description.to_ctypes(_abi.LAYOUT_WORLD_MAJOR, n_snapshot_slots)
        self._h = _abi._vp()
        rc = self._lib.ddb_create(C.byref(self._topo), C.byref(self._params), n_worlds, device, C.byref(self._h))
        if rc != 0:
            msg = self._lib.ddb_last_error(self._h).decode() if self._h else ""
            if self._h:
                self._lib.ddb_destroy(self._h)
            self._h = None
            raise RuntimeError("ddb_create failed: %s %s" % (_abi.ERRORS.get(rc, rc), msg))
        self.device = torch.device("cuda", device)
        self.N = n_worlds
        self.n = self._topo.n_dofs
        self._n_slots = n_snapshot_slots
        self._free_slots = list(range(n_snapshot_slots))
        z = lambda: torch.zeros(self.N, self.n, dtype=torch.float64, device=self.device)
        self._q, self._qd, self._tau = z(), z(), z()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ddb_destroy(self._h)
            self._h = None

    # ---- World accessors (World.cpp:1027-1229) ----
    def getNumDofs(self):
        return self.n

    def getNumWorlds(self):
        return self.N

    def getTimeStep(self):
        return self._params.time_step

    def _as_batch(self, t):
        t = torch.as_tensor(t, dtype=torch.float64, device=self.device)
        if t.dim() == 1:
            t = t.unsqueeze(0)
        if tuple(t.shape) != (self.N, self.n):
            raise ValueError("expected a [%d, %d] tensor, got %s" % (self.N, self.n, tuple(t.shape)))
        return t.contiguous()

    def _out(self, t):
        return t[0] if self._squeeze else t

    def setPositions(self, q):
        self._squeeze = torch.as_tensor(q).dim() == 1
        self._q = self._as_batch(q)

    def setVelocities(self, qd):
        self._qd = self._as_batch(qd)

    def setExternalForces(self, tau):
        self._tau = self._as_batch(tau)

    setForces = setExternalForces  # dart_layer.py:35 calls setForces; the binding exposes setExternalForces

    _squeeze = False
    last_active_set = None
    last_solver_status = None

    def getPositions(self):
        return self._out(self._q)

    def getVelocities(self):
        return self._out(self._qd)

    def getExternalForces(self):
        return self._out(self._tau)

    # ---- snapshot slots ----
    def _acquire_slot(self):
        if not self._free_slots:
            n_out = C.c_int32(0)
            rc = self._lib.ddb_snapshot_reserve(self._h, self._n_slots * 2, C.byref(n_out))
            _check(self._lib, self._h, rc, "ddb_snapshot_reserve")
            self._free_slots = list(range(self._n_slots, n_out.value))
            self._n_slots = n_out.value
        return self._free_slots.pop(0)

    def _release_slot(self, slot):
        self._free_slots.append(slot)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- raw batched calls ----
    def step_raw(self, q, qd, tau, slot, idempotent=False):
        qn, qdn = torch.empty_like(q), torch.empty_like(q)
        rc = self._lib.ddb_snapshot_select(self._h, slot)
        _check(self._lib, self._h, rc, "ddb_snapshot_select")
        cm, km = C.c_int32(0), C.c_int32(0)
        self._lib.ddb_contact_capacity(self._h, C.byref(cm), C.byref(km))
        active = status = None
        if km.value > 0:
            # contact worlds: the solver's cached solution (World::getCachedLCPSolution) is carried by the handle;
            # the active set (ConstraintMapping per LCP row) and the solver tier are returned per world
            active = torch.empty(self.N, km.value, dtype=torch.int32, device=self.device)
            status = torch.empty(self.N, dtype=torch.int32, device=self.device)
        cache = nrows = None
        if idempotent and km.value > 0:
            # forwardPass(world, idempotent=True) restores the solver's cached solution (RestorableSnapshot.cpp:23,32): the step
            # works on a scratch copy of the cache, which is dropped, so the handle's own cache stays as it was
            cache = torch.empty(self.N, km.value, dtype=torch.float64, device=self.device)
            nrows = torch.empty(self.N, dtype=torch.int32, device=self.device)
            rc = self._lib.ddb_lcp_cache_get(self._h, self._stream(), cache.data_ptr(), nrows.data_ptr())
            _check(self._lib, self._h, rc, "ddb_lcp_cache_get")
        rc = self._lib.ddb_forward(self._h, self._stream(), q.data_ptr(), qd.data_ptr(), tau.data_ptr(), qn.data_ptr(),
                                   qdn.data_ptr(), None if cache is None else cache.data_ptr(),
                                   None if nrows is None else nrows.data_ptr(), None if active is None else active.data_ptr(),
                                   None if status is None else status.data_ptr())
        _check(self._lib, self._h, rc, "ddb_forward")
        self.last_active_set, self.last_solver_status = active, status
        return qn, qdn

    def setLanes(self, n_lanes):
        """Number of internal streams the rollout and host-buffer entry points spread chunks of worlds over."""
        _check(self._lib, self._h, self._lib.ddb_set_lanes(self._h, int(n_lanes)), "ddb_set_lanes")

    def launch_count(self):
        return int(self._lib.ddb_launch_count(self._h))


class BackpropSnapshot:
    """Handle on the step retained in one snapshot slot of a BatchedWorld."""

    def __init__(self, world, slot, pre, post):
        self._world = world
        self._slot = slot
        self._pre = pre
        self._post = post

    def __del__(self):
        w = getattr(self, "_world", None)
        if w is not None and self._slot is not None:
            w._release_slot(self._slot)
            self._slot = None

    def _select(self):
        w = self._world
        _check(w._lib, w._h, w._lib.ddb_snapshot_select(w._h, self._slot), "ddb_snapshot_select")

    def backprop(self, world, thisTimestepLoss, nextTimestepLoss, perfLog=None, exploreAlternateStrategies=False):
        """BackpropSnapshot::backprop (BackpropSnapshot.cpp:115-186): fills thisTimestepLoss."""
        w = self._world
        gqn = w._as_batch(nextTimestepLoss.lossWrtPosition)
        gqdn = w._as_batch(nextTimestepLoss.lossWrtVelocity)
        gq, gqd, gtau = torch.empty_like(gqn), torch.empty_like(gqn), torch.empty_like(gqn)
        self._select()
        rc = w._lib.ddb_backward(w._h, w._stream(), gqn.data_ptr(), gqdn.data_ptr(), gq.data_ptr(), gqd.data_ptr(), gtau.data_ptr())
        _check(w._lib, w._h, rc, "ddb_backward")
        sq = torch.as_tensor(nextTimestepLoss.lossWrtPosition).dim() == 1
        thisTimestepLoss.lossWrtPosition = gq[0] if sq else gq
        thisTimestepLoss.lossWrtVelocity = gqd[0] if sq else gqd
        thisTimestepLoss.lossWrtTorque = gtau[0] if sq else gtau

    def _jacobians(self, which):
        w = self._world
        outs = {k: torch.empty(w.N, w.n, w.n, dtype=torch.float64, device=w.device) for k in which}
        ptr = lambda k: outs[k].data_ptr() if k in outs else None
        self._select()
        rc = w._lib.ddb_jacobians(w._h, w._stream(), ptr("posPos"), ptr("posVel"), ptr("velPos"), ptr("velVel"), ptr("forceVel"))
        _check(w._lib, w._h, rc, "ddb_jacobians")
        # the ABI returns [world][col][row]; expose [world][row][col]
        res = {k: v.transpose(1, 2) for k, v in outs.items()}
        if w._squeeze:
            res = {k: v[0] for k, v in res.items()}
        return res

    def getJacobians(self):
        return self._jacobians(("posPos", "posVel", "velPos", "velVel", "forceVel"))

    def getPosPosJacobian(self, world=None, perfLog=None):
        return self._jacobians(("posPos",))["posPos"]

    def getPosVelJacobian(self, world=None, perfLog=None):
        return self._jacobians(("posVel",))["posVel"]

    def getVelPosJacobian(self, world=None, perfLog=None):
        return self._jacobians(("velPos",))["velPos"]

    def getVelVelJacobian(self, world=None, perfLog=None):
        return self._jacobians(("velVel",))["velVel"]

    def getForceVelJacobian(self, world=None, perfLog=None):
        return self._jacobians(("forceVel",))["forceVel"]

    def getInvMassMatrix(self, world=None):
        w = self._world
        out = torch.empty(w.N, w.n, w.n, dtype=torch.float64, device=w.device)
        self._select()
        rc = w._lib.ddb_inv_mass_matrix(w._h, w._stream(), out.data_ptr())
        _check(w._lib, w._h, rc, "ddb_inv_mass_matrix")
        out = out.transpose(1, 2)
        return out[0] if w._squeeze else out

    def getMassMatrix(self, world=None):
        return torch.linalg.inv(self.getInvMassMatrix())

    def getPreStepPosition(self):
        return self._world._out(self._pre[0])

    def getPreStepVelocity(self):
        return self._world._out(self._pre[1])

    def getPreStepTorques(self):
        return self._world._out(self._pre[2])

    def getPostStepPosition(self):
        return self._world._out(self._post[0])

    def getPostStepVelocity(self):
        return self._world._out(self._post[1])

    def getPostStepTorques(self):
        return self._world._out(self._pre[2])


def forwardPass(world, idempotent=False):
    """dart.neural.forwardPass: take one step of every world, return the snapshot.
    idempotent=True leaves the world untouched: positions, velocities, forces and the constraint solver's cached LCP solution
    (NeuralUtils.cpp:29-33,61-62 through RestorableSnapshot.cpp:18-33)."""
    q, qd, tau = world._q, world._qd, world._tau
    slot = world._acquire_slot()
    qn, qdn = world.step_raw(q, qd, tau, slot, idempotent=idempotent)
    snap = BackpropSnapshot(world, slot, (q, qd, tau), (qn, qdn))
    if not idempotent:
        world._q, world._qd = qn, qdn
        # World::step(_resetCommand=true) clears the commands/forces (World.cpp:252-257)
        world._tau = torch.zeros_like(tau)
    return snap
