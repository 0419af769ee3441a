"""GPU: the T-step rollout entry points (ddb_rollout_forward / ddb_rollout_backward, the shape of SingleShot::getSnapshots and
SingleShot::backpropGradientWrt, dart/trajectory/SingleShot.cpp:650-760) and the chunk-pipelined host-buffer path.

Worlds are independent and every chunk runs the same kernels on the same numbers, so splitting a batch over lanes must not
change a single bit; against the CPU oracle the bars are those of test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest
import torch

from diffdart_b200 import _abi, skeletons
from diffdart_b200.rollout import dart_rollout
from diffdart_b200.world import BatchedWorld, LossGradient, forwardPass
from oracle import oracle as orc
from util import random_inputs, ref_equals, max_ref_err

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def stepwise(desc, q0, qd0, taus, gp, gv):
    """The same rollout through forwardPass / BackpropSnapshot.backprop, one call per step."""
    T, N = taus.shape[0], q0.shape[0]
    world = BatchedWorld(desc, N, n_snapshot_slots=T)
    world.setPositions(cuda(q0)); world.setVelocities(cuda(qd0))
    snaps, poses, vels = [], [], []
    for t in range(T):
        world.setForces(cuda(taus[t]))
        snaps.append(forwardPass(world))
        poses.append(world.getPositions().clone()); vels.append(world.getVelocities().clone())
    run_q = torch.zeros_like(poses[0]); run_qd = torch.zeros_like(poses[0])
    gtau = [None] * T
    for t in reversed(range(T)):
        nxt, this = LossGradient(), LossGradient()
        nxt.lossWrtPosition, nxt.lossWrtVelocity = cuda(gp[t]) + run_q, cuda(gv[t]) + run_qd
        snaps[t].backprop(world, this, nxt)
        run_q, run_qd, gtau[t] = this.lossWrtPosition, this.lossWrtVelocity, this.lossWrtTorque
    return torch.stack(poses), torch.stack(vels), run_q, run_qd, torch.stack(gtau)


def rollout(desc, q0, qd0, taus, gp, gv, lanes):
    N = q0.shape[0]
    world = BatchedWorld(desc, N)
    world.setLanes(lanes)
    a, b, c = cuda(q0).requires_grad_(), cuda(qd0).requires_grad_(), cuda(taus).requires_grad_()
    poses, vels = dart_rollout(world, a, b, c)
    ((poses * cuda(gp)).sum() + (vels * cuda(gv)).sum()).backward()
    return poses.detach(), vels.detach(), a.grad, b.grad, c.grad


@pytest.mark.parametrize("N,lanes", [(64, 1), (1000, 3), (4096, 2), (5000, 4)])
def test_rollout_equals_stepwise_bitwise(N, lanes):
    desc = skeletons.create_multiarm_robot(5)
    T, n = 6, 5
    q0, qd0, _ = random_inputs(n, N, 5)
    rng = np.random.default_rng(6)
    taus, gp, gv = rng.uniform(-1, 1, (T, N, n)), rng.normal(size=(T, N, n)), rng.normal(size=(T, N, n))
    ref = stepwise(desc, q0, qd0, taus, gp, gv)
    got = rollout(desc, q0, qd0, taus, gp, gv, lanes)
    for name, r, g in zip(("poses", "vels", "dL/dq0", "dL/dqd0", "dL/dtau"), ref, got):
        # the running gradient is added inside the backward kernel (a + b, the same sum the stepwise path forms with torch)
        assert torch.equal(r, g), (name, float((r - g).abs().max()))


def test_rollout_dof_major_layout():
    """The same rollout through the raw ABI with DOF-major vectors x[dof*N + world] (every step of a sequence is one such block)."""
    lib = _abi.load_library()
    desc = skeletons.create_multiarm_robot(4)
    N, n, T = 300, 4, 5
    q0, qd0, _ = random_inputs(n, N, 41)
    rng = np.random.default_rng(42)
    taus, gp, gv = rng.uniform(-1, 1, (T, N, n)), rng.normal(size=(T, N, n)), rng.normal(size=(T, N, n))
    ref = rollout(desc, q0, qd0, taus, gp, gv, 2)
    topo, params, keep = desc.to_ctypes(_abi.LAYOUT_DOF_MAJOR)
    h = _abi._vp()
    assert lib.ddb_create(C.byref(topo), C.byref(params), N, 0, C.byref(h)) == 0
    tr = lambda a: cuda(np.swapaxes(a, -1, -2))  # [.., N, n] -> [.., n, N]
    dq0, dqd0, dt, dgp, dgv = tr(q0), tr(qd0), tr(taus), tr(gp), tr(gv)
    poses, vels, gtau = torch.empty_like(dt), torch.empty_like(dt), torch.empty_like(dt)
    gq0, gqd0 = torch.empty_like(dq0), torch.empty_like(dq0)
    assert lib.ddb_rollout_forward(h, None, T, dq0.data_ptr(), dqd0.data_ptr(), dt.data_ptr(), poses.data_ptr(), vels.data_ptr()) == 0
    assert lib.ddb_rollout_backward(h, None, T, dgp.data_ptr(), dgv.data_ptr(), gq0.data_ptr(), gqd0.data_ptr(), gtau.data_ptr()) == 0
    torch.cuda.synchronize()
    for name, r, g in zip(("poses", "vels", "dL/dq0", "dL/dqd0", "dL/dtau"), ref, (poses, vels, gq0, gqd0, gtau)):
        assert torch.equal(r, g.transpose(-1, -2)), name
    lib.ddb_destroy(h)


def test_rollout_against_oracle():
    desc = skeletons.create_multiarm_robot(3)
    N, n, T = 48, 3, 10
    q0, qd0, _ = random_inputs(n, N, 31)
    rng = np.random.default_rng(32)
    taus, gp, gv = rng.uniform(-1, 1, (T, N, n)), rng.normal(size=(T, N, n)), rng.normal(size=(T, N, n))
    poses, vels, gq0, gqd0, gtau = rollout(desc, q0, qd0, taus, gp, gv, 2)
    o = orc.OracleWorld(desc)
    qs, qds = [q0], [qd0]
    for t in range(T):
        a, b = o.step(qs[-1], qds[-1], taus[t])
        qs.append(a); qds.append(b)
    assert ref_equals(poses.cpu().numpy(), np.stack(qs[1:]), 1e-11)
    assert ref_equals(vels.cpu().numpy(), np.stack(qds[1:]), 1e-11)
    gq, gqd = np.zeros_like(q0), np.zeros_like(q0)
    for t in reversed(range(T)):
        _, _, gq, gqd, gt = o.step_backprop(qs[t], qds[t], taus[t], gp[t] + gq, gv[t] + gqd, mode="exact", with_step=False)
        assert ref_equals(gtau[t].cpu().numpy(), gt, 1e-9), (t, max_ref_err(gtau[t].cpu().numpy(), gt))
    assert ref_equals(gq0.cpu().numpy(), gq, 1e-9), max_ref_err(gq0.cpu().numpy(), gq)
    assert ref_equals(gqd0.cpu().numpy(), gqd, 1e-9), max_ref_err(gqd0.cpu().numpy(), gqd)


def test_rollout_with_contacts_equals_stepwise():
    """Handles with a contact stage run the rollout step by step on the caller's stream: the same launches as the stepwise path."""
    desc = skeletons.create_box_stack(2)
    N, T, n = 32, 4, 12
    rng = np.random.default_rng(3)
    q0 = np.zeros((N, n)); qd0 = rng.uniform(-0.1, 0.1, (N, n))
    for i in range(2):
        q0[:, 6 * i + 3] = rng.uniform(-0.05, 0.05, N); q0[:, 6 * i + 4] = rng.uniform(-0.05, 0.05, N)
        q0[:, 6 * i + 5] = 0.25 + 0.5 * i - 1e-3 * (i + 1)
    taus = np.zeros((T, N, n)); gp, gv = rng.normal(size=(T, N, n)), rng.normal(size=(T, N, n))
    ref = stepwise(desc, q0, qd0, taus, gp, gv)
    got = rollout(desc, q0, qd0, taus, gp, gv, 2)
    for name, r, g in zip(("poses", "vels", "dL/dq0", "dL/dqd0", "dL/dtau"), ref, got):
        assert torch.allclose(r, g, rtol=0, atol=1e-12), (name, float((r - g).abs().max()))


@pytest.mark.parametrize("zero_copy,N", [(True, 10000), (False, 10000), (True, 77)])
def test_host_path_equals_device_path(zero_copy, N, monkeypatch):
    """Host-buffer entry points on a world-major batch. Pinned host buffers: the kernels move their rows between host memory and
    shared memory themselves (HOSTIO kernels, one launch per call); with DDB_HOST_ZEROCOPY=0 (and for pageable buffers): chunks of
    worlds pipelined over the lanes (copy-in, kernel, copy-out overlap). Every call must equal the device path bit for bit on the
    data that is in the host buffers at that moment; N = 77 leaves the last CTA partly filled with an odd number of doubles."""
    monkeypatch.setenv("DDB_HOST_ZEROCOPY", "1" if zero_copy else "0")
    lib = _abi.load_library()
    desc = skeletons.create_multiarm_robot(5)
    n = 5
    topo, params, keep = desc.to_ctypes()
    h = _abi._vp()
    assert lib.ddb_create(C.byref(topo), C.byref(params), N, 0, C.byref(h)) == 0
    hin = [torch.empty(N, n, dtype=torch.float64).pin_memory() for _ in range(5)]
    hout = [torch.empty(N, n, dtype=torch.float64).pin_memory() for _ in range(5)]
    dout = [torch.empty(N, n, dtype=torch.float64, device="cuda") for _ in range(5)]
    launches = []
    for call in range(3):
        q, qd, tau = random_inputs(n, N, 9 + call)
        g = np.random.default_rng(20 + call).normal(size=(2, N, n))
        for t, a in zip(hin, (q, qd, tau, g[0], g[1])):
            t.copy_(torch.tensor(a))
        for t in hout:
            t.zero_()
        l0 = lib.ddb_launch_count(h)
        assert lib.ddb_forward_host(h, *[t.data_ptr() for t in hin[:3]], hout[0].data_ptr(), hout[1].data_ptr()) == 0, lib.ddb_last_error(h)
        assert lib.ddb_backward_host(h, hin[3].data_ptr(), hin[4].data_ptr(), *[t.data_ptr() for t in hout[2:]]) == 0, lib.ddb_last_error(h)
        launches.append(lib.ddb_launch_count(h) - l0)
        din = [t.cuda() for t in hin]
        assert lib.ddb_forward(h, None, *[t.data_ptr() for t in din[:3]], dout[0].data_ptr(), dout[1].data_ptr(), None, None, None, None) == 0
        assert lib.ddb_backward(h, None, din[3].data_ptr(), din[4].data_ptr(), *[t.data_ptr() for t in dout[2:]]) == 0
        torch.cuda.synchronize()
        for a, b in zip(hout, dout):
            assert torch.equal(a, b.cpu()), call
    if zero_copy:
        assert launches[0] == 2  # one HOSTIO launch each
    else:
        assert launches[0] >= 4  # forward and backward each ran as several chunk launches
    # pageable host memory works the same way (the copies are then staged by the driver)
    pin = [np.ascontiguousarray(t.numpy().copy()) for t in hin]
    pout = [np.empty((N, n)) for _ in range(5)]
    for _ in range(3):
        assert lib.ddb_forward_host(h, *[a.ctypes.data for a in pin[:3]], pout[0].ctypes.data, pout[1].ctypes.data) == 0
        assert lib.ddb_backward_host(h, pin[3].ctypes.data, pin[4].ctypes.data, *[a.ctypes.data for a in pout[2:]]) == 0
        for a, b in zip(pout, dout):
            assert np.array_equal(a, b.cpu().numpy())
    lib.ddb_destroy(h)


def test_host_path_zero_copy_clips_at_the_limits():
    """clipLossGradientsToBounds inside the HOSTIO backward (it clips its shared-memory rows, not the host buffers): equal to the
    device path bit for bit on a chain with finite limits and inputs sitting on them."""
    lib = _abi.load_library()
    desc = skeletons.create_multiarm_robot(3)
    bodies = desc.skeletons[0].bodies
    bodies[0].position_lower[:] = -0.5
    bodies[1].velocity_upper[:] = 0.25
    bodies[2].force_lower[:] = -1.0
    N, n = 300, 3
    q, qd, tau = random_inputs(n, N, 9)
    q[:150, 0] = -0.5
    qd[100:250, 1] = 0.25
    tau[::2, 2] = -1.0
    g = np.random.default_rng(4).normal(size=(2, N, n))
    topo, params, keep = desc.to_ctypes()
    h = _abi._vp()
    assert lib.ddb_create(C.byref(topo), C.byref(params), N, 0, C.byref(h)) == 0
    hin = [torch.tensor(a, dtype=torch.float64).pin_memory() for a in (q, qd, tau, g[0], g[1])]
    hout = [torch.zeros(N, n, dtype=torch.float64).pin_memory() for _ in range(5)]
    dout = [torch.empty(N, n, dtype=torch.float64, device="cuda") for _ in range(5)]
    assert lib.ddb_forward_host(h, *[t.data_ptr() for t in hin[:3]], hout[0].data_ptr(), hout[1].data_ptr()) == 0, lib.ddb_last_error(h)
    assert lib.ddb_backward_host(h, hin[3].data_ptr(), hin[4].data_ptr(), *[t.data_ptr() for t in hout[2:]]) == 0, lib.ddb_last_error(h)
    assert lib.ddb_launch_count(h) == 2
    din = [t.cuda() for t in hin]
    assert lib.ddb_forward(h, None, *[t.data_ptr() for t in din[:3]], dout[0].data_ptr(), dout[1].data_ptr(), None, None, None, None) == 0
    assert lib.ddb_backward(h, None, din[3].data_ptr(), din[4].data_ptr(), *[t.data_ptr() for t in dout[2:]]) == 0
    torch.cuda.synchronize()
    for a, b in zip(hout, dout):
        assert torch.equal(a, b.cpu())
    assert (hout[2][:150, 0] <= 0).all() and (hout[3][100:250, 1] >= 0).all() and (hout[4][::2, 2] <= 0).all()
    assert (hout[2][:150, 0] == 0).any()  # some gradients were actually clipped
    lib.ddb_destroy(h)


def test_step_captured_in_a_cuda_graph_on_a_side_stream():
    """ddb_forward + ddb_backward captured into a CUDA graph on a stream other than the one the handle was first used on (the topology
    upload was ordered there): replays equal the direct calls bit for bit, and nothing in the entry points breaks the capture."""
    lib = _abi.load_library()
    desc = skeletons.create_multiarm_robot(5)
    N, n = 2048, 5
    topo, params, keep = desc.to_ctypes()
    h = _abi._vp()
    assert lib.ddb_create(C.byref(topo), C.byref(params), N, 0, C.byref(h)) == 0
    q, qd, tau = random_inputs(n, N, 3)
    g = np.random.default_rng(8).normal(size=(2, N, n))
    din = [torch.tensor(a, dtype=torch.float64, device="cuda") for a in (q, qd, tau, g[0], g[1])]
    ref = [torch.empty(N, n, dtype=torch.float64, device="cuda") for _ in range(5)]
    out = [torch.zeros(N, n, dtype=torch.float64, device="cuda") for _ in range(5)]

    def step(sp, o):
        assert lib.ddb_forward(h, sp, *[t.data_ptr() for t in din[:3]], o[0].data_ptr(), o[1].data_ptr(), None, None, None, None) == 0, lib.ddb_last_error(h)
        assert lib.ddb_backward(h, sp, din[3].data_ptr(), din[4].data_ptr(), *[t.data_ptr() for t in o[2:]]) == 0, lib.ddb_last_error(h)

    step(C.c_void_p(torch.cuda.current_stream().cuda_stream), ref)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        step(C.c_void_p(side.cuda_stream), out)
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    for a, b in zip(out, ref):
        assert torch.equal(a, b)
    lib.ddb_destroy(h)


def test_repeated_rollout_calls_replay_a_cuda_graph_with_the_same_bits(monkeypatch):
    """A contact-free rollout called again with the same buffers is recorded into a CUDA graph on its second call and replayed from then
    on (ddb_api.cu: rolloutGraph). Every call must give the bits of the direct launches (DDB_ROLLOUT_GRAPHS=0), also after the inputs in
    the buffers have changed, and a call with other buffers must not reuse the graph."""
    lib = _abi.load_library()
    desc = skeletons.create_multiarm_robot(5)
    N, n, T = 3000, 5, 5
    rng = np.random.default_rng(12)
    mk = lambda *shape: torch.tensor(rng.uniform(-1, 1, shape), dtype=torch.float64, device="cuda")

    def run(graphs, calls):
        monkeypatch.setenv("DDB_ROLLOUT_GRAPHS", "1" if graphs else "0")
        topo, params, keep = desc.to_ctypes()
        h = _abi._vp()
        assert lib.ddb_create(C.byref(topo), C.byref(params), N, 0, C.byref(h)) == 0
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        q0, qd0, taus, gp, gv = (torch.empty(*s, dtype=torch.float64, device="cuda") for s in ((N, n), (N, n), (T, N, n), (T, N, n), (T, N, n)))
        outs = [torch.zeros(T, N, n, dtype=torch.float64, device="cuda") for _ in range(3)] + [torch.zeros(N, n, dtype=torch.float64, device="cuda") for _ in range(2)]
        res = []
        for src in calls:
            for dst, a in zip((q0, qd0, taus, gp, gv), src):
                dst.copy_(a)
            assert lib.ddb_rollout_forward(h, sp, T, q0.data_ptr(), qd0.data_ptr(), taus.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr()) == 0, lib.ddb_last_error(h)
            assert lib.ddb_rollout_backward(h, sp, T, gp.data_ptr(), gv.data_ptr(), outs[3].data_ptr(), outs[4].data_ptr(), outs[2].data_ptr()) == 0, lib.ddb_last_error(h)
            torch.cuda.synchronize()
            res.append([o.clone() for o in outs])
        # other output buffers: must not replay the graph that writes into `outs`
        other = [torch.zeros_like(o) for o in outs]
        assert lib.ddb_rollout_forward(h, sp, T, q0.data_ptr(), qd0.data_ptr(), taus.data_ptr(), other[0].data_ptr(), other[1].data_ptr()) == 0
        assert lib.ddb_rollout_backward(h, sp, T, gp.data_ptr(), gv.data_ptr(), other[3].data_ptr(), other[4].data_ptr(), other[2].data_ptr()) == 0
        torch.cuda.synchronize()
        res.append(other)
        launches = lib.ddb_launch_count(h)
        lib.ddb_destroy(h)
        return res, launches

    calls = [(mk(N, n), mk(N, n), mk(T, N, n), mk(T, N, n), mk(T, N, n)) for _ in range(4)]
    direct, l0 = run(False, calls)
    graphed, l1 = run(True, calls)
    assert l0 == l1  # the launch count is the number of kernels that ran, replayed or not
    for a, b in zip(direct, graphed):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert not torch.equal(direct[0][0], direct[1][0])  # (the calls really differ)
