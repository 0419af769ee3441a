#!/usr/bin/env python
"""bench.py — world-steps/s (forward step + analytic backward VJP) of the batched differentiable timestep.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the CPU path on the host cores

Workload (BASELINE.json configs[1]): 65536 independent 5-link pendulums per GPU (no contact), one
"step" = one forward step + its analytic backward over the whole batch. The K timed steps run as the library's
T-step rollout with backprop (ddb_rollout_forward + ddb_rollout_backward, T = 8: the call pattern of the reference's
trajectory code, SingleShot.cpp:650-760), K/8 rollouts back to back; the same step issued as separate
ddb_forward + ddb_backward calls (dart_layer forward + backward semantics) is timed next to it ("single_step").
Weak scaling: every rank owns 65536 worlds; there is no inter-GPU traffic inside a step.
One JSON line is printed by rank 0 (see the contract in the task description).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "world-steps/sec incl. analytic backprop"
UNIT = "world-steps/s"
N_WORLDS = 65536
N_DOFS = 5
WORKLOAD = ("pendulum5: 65536 five-link revolute pendulums per GPU, no contact (BASELINE configs[1]); step = forward step + analytic backward VJP, "
            "run as 8-step rollouts with backprop through ddb_rollout_forward / ddb_rollout_backward")
# Algorithmic HBM bytes per world, SURVEY.md §8(d): forward reads q,qd,tau (3n) and writes q',qd' (2n) = 40n B;
# backward reads dL/dq',dL/dqd' (2n) + the pre-step state (3n) and writes dL/dq,dL/dqd,dL/dtau (3n) = 64n B.
# (The design additionally moves a 208 B/body snapshot each way on purpose — DESIGN.md §4 — which shows up in
# roofline.traffic, not in the algorithmic bytes.)
FWD_BYTES = 40 * N_DOFS
BWD_BYTES = 64 * N_DOFS
# DRAM traffic per whole-batch launch from the `ncu --set full` captures of the round-2 chain kernels (profiles/r02a_ncu_full_summary.csv),
# bytes: dram__bytes_read.sum + dram__bytes_write.sum. (The forward leaves most of its 68 MB snapshot dirty in L2 when it ends.)
FWD_TRAFFIC = 7.93e6 + 15.90e6
BWD_TRAFFIC = 73.45e6 + 1.73e6
# FP64 flops per world of the revolute-chain<5> kernels as executed (ncu: smsp__sass_thread_inst_executed_op_{dfma x2, dmul, dadd} per thread,
# profiles/r02a_ncu_full_summary.csv; the static SASS counts of profiles/r02_flops.md are 2 % lower): DFMA = 2, DADD / DMUL = 1
FWD_FLOPS = 3687
BWD_FLOPS = 3060


def inputs(N, seed=1234):
    rng = np.random.default_rng(seed)
    return [rng.uniform(-1, 1, (N, N_DOFS)) for _ in range(3)]


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = get(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_world_steps(n_worlds, threads, reps=1):
    """Times the CPU path (the oracle's restatement of World::step + BackpropSnapshot::backprop with the
    reference's own finite-difference recipes) on n_worlds worlds; returns world-steps/s."""
    from diffdart_b200 import skeletons
    from oracle.oracle import OracleWorld

    o = OracleWorld(skeletons.create_multiarm_robot(N_DOFS))
    q, qd, tau = inputs(n_worlds)
    g = np.random.default_rng(7).normal(size=(2, n_worlds, N_DOFS))
    o.step_backprop(q[:256], qd[:256], tau[:256], g[0][:256], g[1][:256], threads=threads)  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        o.step_backprop(q, qd, tau, g[0], g[1], mode="reference", threads=threads)
    dt = time.perf_counter() - t0
    return n_worlds * reps / dt, dt


# ---- second workload (reported next to the headline, not instead of it): BASELINE configs[2], contact-rich box stacking ----
CONTACT_WORLDS = 16384
CONTACT_WORKLOAD = ("box_stack: 16384 worlds per GPU of 5 FreeJoint boxes 1x1x0.5 stacked on a welded floor (BASELINE configs[2], SURVEY 8(d) C3: "
                    "1 mm overlap per interface, x,y ~ U(-0.05,0.05), yaw ~ U(-0.1,0.1), qd ~ U(-0.1,0.1), mu = 1, e = 0); step = forwardPass with the "
                    "contact solve (collision, LCP assembly, short-circuit / Dantzig / PGS chain, classification) + backward through the contact solution")


def box_stack_states(N, seed, n=5):
    rng = np.random.default_rng(seed)
    q = np.zeros((N, 6 * n))
    for i in range(n):
        q[:, 6 * i + 2] = rng.uniform(-0.1, 0.1, N)
        q[:, 6 * i + 3] = rng.uniform(-0.05, 0.05, N)
        q[:, 6 * i + 4] = rng.uniform(-0.05, 0.05, N)
        q[:, 6 * i + 5] = 0.25 + 0.5 * i - 1e-3 * (i + 1)
    qd = rng.uniform(-0.1, 0.1, (N, 6 * n))
    return q, qd, np.zeros((N, 6 * n))


def contact_bench(local_rank, rank, steps=4, warmup=3):
    """Forward + backward of the box-stacking batch through the Python host layer; returns (ms per step, tier histogram, rows)."""
    import torch

    from diffdart_b200 import skeletons
    from diffdart_b200.world import BatchedWorld, LossGradient, forwardPass

    N = CONTACT_WORLDS
    q, qd, tau = box_stack_states(N, 7 + rank)
    dev = torch.device("cuda", local_rank)
    world = BatchedWorld(skeletons.create_box_stack(5), N, device=local_rank)
    cu = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    world.setPositions(cu(q)); world.setVelocities(cu(qd))
    tq = cu(tau)
    g = torch.randn(N, q.shape[1], dtype=torch.float64, device=dev)
    nxt, this = LossGradient(), LossGradient()
    nxt.lossWrtPosition, nxt.lossWrtVelocity = g, g.clone()

    def step():
        world.setForces(tq)
        snap = forwardPass(world)  # the rollout advances: every step sees new contacts and the solver's cached solution
        snap.backprop(world, this, nxt)

    for _ in range(warmup):  # warm-up steps populate the LCP cache
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    status = world.last_solver_status.cpu().numpy()
    tiers = {str(int(k)): int(v) for k, v in zip(*np.unique(status, return_counts=True))}
    return e0.elapsed_time(e1) / steps, tiers, N


ATLAS_WORLDS = 4096
ATLAS_WORKLOAD = ("atlas: 4096 worlds per GPU of the Atlas humanoid (27 revolute joints + FreeJoint root = 33 DOFs, 35 bodies, box feet) standing on the ground "
                  "box (BASELINE configs[4], SURVEY 8(d) C5: pose of test_AtlasGradients.cpp:222-223 + U(-0.01,0.01), qd ~ U(-0.1,0.1), tau = 0); step = "
                  "forwardPass with the contact solve + backward through the contact solution; at N > 1 the gradient w.r.t. a parameter shared by all "
                  "worlds (here: a torque bias, 33 doubles) is summed over the local worlds and all-reduced over NCCL once per backward")


def atlas_bench(local_rank, rank, world_size, steps=5, warmup=2):
    """Config 5 when the reference's URDFs are present in baseline/_ref/atlas (tests/contact_scenes.py says how to put them there)."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from contact_scenes import atlas_available, atlas_on_ground, atlas_on_ground_states
    from diffdart_b200.sharding import allreduce_shared_gradient
    from diffdart_b200.world import BatchedWorld, LossGradient, forwardPass

    if not atlas_available():
        return None
    N = ATLAS_WORLDS
    q, qd, tau = atlas_on_ground_states(N, 13 + rank)
    dev = torch.device("cuda", local_rank)
    world = BatchedWorld(atlas_on_ground(), N, device=local_rank)
    cu = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    world.setPositions(cu(q)); world.setVelocities(cu(qd))
    tq = cu(tau)
    g = torch.randn(N, q.shape[1], dtype=torch.float64, device=dev)
    nxt, this = LossGradient(), LossGradient()
    nxt.lossWrtPosition, nxt.lossWrtVelocity = g, g.clone()
    shared = None

    def step():
        nonlocal shared
        world.setForces(tq)
        snap = forwardPass(world)
        snap.backprop(world, this, nxt)
        shared = allreduce_shared_gradient(this.lossWrtTorque)  # d loss / d (torque bias shared by all worlds)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    e0.record()
    for _ in range(20):
        allreduce_shared_gradient(this.lossWrtTorque)
    e1.record()
    torch.cuda.synchronize(dev)
    status = world.last_solver_status.cpu().numpy()
    return {"ms_per_step": ms, "allreduce_us": e0.elapsed_time(e1) / 20 * 1e3, "finite": bool(torch.isfinite(shared).all()),
            "tiers": {str(int(k)): int(v) for k, v in zip(*np.unique(status, return_counts=True))}}


WALKER_WORLDS, WALKER_STEPS = 32768, 20
WALKER_WORKLOAD = ("walker: 32768 worlds per GPU of a half-cheetah-style URDF walker (9 DOFs, box feet; tests/fixtures/walker.urdf) on the ground (BASELINE "
                   "configs[3] style), one 20-step trajectory rollout with backprop through the contact solutions via dart_rollout (loss = |q_T|^2 + |qd_T|^2); "
                   "scripts/walker_rollout_bench.py runs the 100-step version")


def walker_bench(local_rank, rank):
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from contact_scenes import walker_on_ground, walker_on_ground_states
    from diffdart_b200.rollout import dart_rollout
    from diffdart_b200.world import BatchedWorld

    N, T = WALKER_WORLDS, WALKER_STEPS
    dev = torch.device("cuda", local_rank)
    q, qd, tau = walker_on_ground_states(N, 11 + rank)
    world = BatchedWorld(walker_on_ground(), N, device=local_rank)
    cu = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    q0, qd0 = cu(q).requires_grad_(), cu(qd).requires_grad_()
    taus = cu(np.broadcast_to(tau, (T, N, 9)).copy()).requires_grad_()

    def run():
        for t in (q0, qd0, taus):
            t.grad = None
        poses, vels = dart_rollout(world, q0, qd0, taus)
        ((poses[-1] ** 2).sum() + (vels[-1] ** 2).sum()).backward()

    run()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize(dev)
    return {"ms_per_rollout": e0.elapsed_time(e1), "finite": bool(torch.isfinite(taus.grad).all() and torch.isfinite(q0.grad).all())}


CARTPOLE_WORLDS = 1 << 20
CARTPOLE_WORKLOAD = ("cartpole: 2^20 worlds per GPU of the reference's cartpole fixture (prismatic cart + revolute pole, no contact; BASELINE configs[0] as a "
                     "batch, SURVEY 8(d) C1: q0 ~ U(-1,1), q1 ~ U(-pi,pi), qd ~ U(-1,1), tau = (U(-10,10), 0)); step = forward step + analytic backward VJP, "
                     "8-step rollouts with backprop through ddb_rollout_forward / ddb_rollout_backward")


def cartpole_bench(lib, local_rank, rank, lanes, rollouts=6):
    """BASELINE configs[0] as a batch: the north star's 'contact-free 2-link cartpole batch'."""
    import torch

    from diffdart_b200 import _abi, skeletons

    N, n, T = CARTPOLE_WORLDS, 2, 8
    dev = torch.device("cuda", local_rank)
    topo, params, keep = skeletons.create_cartpole().to_ctypes(n_snapshot_slots=T)
    h = _abi._vp()
    if lib.ddb_create(C.byref(topo), C.byref(params), N, local_rank, C.byref(h)):
        raise RuntimeError(lib.ddb_last_error(h).decode())
    rng = np.random.default_rng(rank)
    mk = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    q0 = mk(np.stack([rng.uniform(-1, 1, N), rng.uniform(-np.pi, np.pi, N)], axis=1))
    qd0 = mk(rng.uniform(-1, 1, (N, n)))
    taus = mk(np.stack([rng.uniform(-10, 10, (T, N)), np.zeros((T, N))], axis=2))
    gp, gv = torch.randn(T, N, n, dtype=torch.float64, device=dev), torch.randn(T, N, n, dtype=torch.float64, device=dev)
    poses, vels, gtau = torch.empty_like(taus), torch.empty_like(taus), torch.empty_like(taus)
    gq0, gqd0 = torch.empty_like(q0), torch.empty_like(q0)
    sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    lib.ddb_set_lanes(h, lanes)

    def roll():
        r1 = lib.ddb_rollout_forward(h, sp, T, q0.data_ptr(), qd0.data_ptr(), taus.data_ptr(), poses.data_ptr(), vels.data_ptr())
        r2 = lib.ddb_rollout_backward(h, sp, T, gp.data_ptr(), gv.data_ptr(), gq0.data_ptr(), gqd0.data_ptr(), gtau.data_ptr())
        if r1 or r2:
            raise RuntimeError(lib.ddb_last_error(h).decode())

    for _ in range(2):
        roll()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rollouts):
        roll()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / (rollouts * T)
    lib.ddb_destroy(h)
    return ms


CHEETAH_WORLDS, CHEETAH_STEPS = 32768, 20
CHEETAH_WORKLOAD = ("half_cheetah: 32768 worlds per GPU of the reference's data/skel/half_cheetah.skel (9 DOFs, capsule links on a box ground, dt = 0.002; "
                    "BASELINE configs[3]) loaded through diffdart_b200.skel_loader, one 20-step trajectory rollout with backprop through the contact "
                    "solutions via dart_rollout (loss = |q_T|^2 + |qd_T|^2), joint torques ~ U(-20, 20) on the six leg joints")


def cheetah_bench(local_rank, rank):
    """BASELINE configs[3] on the reference's own model file, present when it has been copied to baseline/_ref/skel (tests/contact_scenes.py)."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from contact_scenes import half_cheetah_available, half_cheetah_on_ground, half_cheetah_on_ground_states
    from diffdart_b200.rollout import dart_rollout
    from diffdart_b200.world import BatchedWorld

    if not half_cheetah_available():
        return None
    N, T = CHEETAH_WORLDS, CHEETAH_STEPS
    dev = torch.device("cuda", local_rank)
    q, qd, tau = half_cheetah_on_ground_states(N, 17 + rank)
    world = BatchedWorld(half_cheetah_on_ground(), N, device=local_rank)
    cu = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    q0, qd0 = cu(q).requires_grad_(), cu(qd).requires_grad_()
    taus = cu(np.broadcast_to(tau, (T, N, 9)).copy()).requires_grad_()

    def run():
        for t in (q0, qd0, taus):
            t.grad = None
        poses, vels = dart_rollout(world, q0, qd0, taus)
        ((poses[-1] ** 2).sum() + (vels[-1] ** 2).sum()).backward()

    run()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    # solver tier masks of the last forward step the handle ran (bit 9: a capsule pair in a pipe configuration, not carried)
    world._lib.ddb_rollout_forward(world._h, None, 1, q0.data_ptr(), qd0.data_ptr(), taus.data_ptr(), torch.empty_like(q0).data_ptr(), torch.empty_like(q0).data_ptr())
    counts = np.zeros((N, 4), dtype=np.int32)
    world._lib.ddb_contact_debug_host(world._h, counts.ctypes.data_as(C.c_void_p), -1, None, None, None, None, None, None, None, None, None)
    return {"ms_per_rollout": ms, "finite": bool(torch.isfinite(taus.grad).all() and torch.isfinite(q0.grad).all()),
            "worlds_with_uncarried_pipe_contact_at_last_step": int(((counts[:, 3] & 512) != 0).sum())}


def contact_cpu_baseline(n_worlds=1024):
    """The oracle on the same workload: forwardPass with contacts + the reference's Jacobian recipe, one core, a few worlds."""
    from diffdart_b200 import skeletons
    from oracle.oracle import OracleWorld

    o = OracleWorld(skeletons.create_box_stack(5))
    q, qd, tau = box_stack_states(n_worlds, 7)
    t0 = time.perf_counter()
    for w in range(n_worlds):
        o.contact_jacobians(q[w], qd[w], tau[w], mode="reference")
    dt = time.perf_counter() - t0
    return n_worlds / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    # worlds per step: a bounded sample of the 65536-world batch, sized so that K steps stay within 3e6 .. 8e6 world-steps (tens of seconds on 16 threads)
    # (not below 4096 worlds: smaller steps are dominated by starting the thread pool, which would understate the CPU arm)
    sample = int(min(16384, max(4096, (200 * 16384 // max(args.steps, 1)) // 256 * 256)))
    for _ in range(max(args.warmup, 0)):
        cpu_world_steps(2048, threads)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        cpu_world_steps(sample, threads)
        total += sample
    elapsed = time.perf_counter() - t0
    value = total / elapsed
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": "%d of the 65536 worlds per step" % sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d worlds/step x %d steps; CPU restatement of the reference arithmetic (reference not buildable here: no Eigen)" % (sample, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    from diffdart_b200 import _abi, skeletons

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _abi.load_library()
    desc = skeletons.create_multiarm_robot(N_DOFS)
    N, n = N_WORLDS, N_DOFS
    SETS = 8  # rotating buffer sets: 8 x (26 MB of inputs/outputs + 68 MB snapshot slot) = 750 MB >> 126 MB L2: every step is cold
    topo, params, keep = desc.to_ctypes(n_snapshot_slots=SETS)
    h = _abi._vp()
    rc = lib.ddb_create(C.byref(topo), C.byref(params), N, local_rank, C.byref(h))
    if rc:
        raise RuntimeError("ddb_create failed: %s" % lib.ddb_last_error(h).decode())
    q, qd, tau = inputs(N, 1234 + rank)
    g = np.random.default_rng(99 + rank).normal(size=(2, N, n))
    mk = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    sets = []
    for s in range(SETS):
        ins = [mk(np.roll(a, s, axis=0)) for a in (q, qd, tau, g[0], g[1])]
        outs = [torch.empty(N, n, dtype=torch.float64, device=dev) for _ in range(5)]
        sets.append((ins, outs))
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)

    def step(s):
        (a, b, c, d, e), (o0, o1, o2, o3, o4) = sets[s % SETS]
        lib.ddb_snapshot_select(h, s % SETS)
        r1 = lib.ddb_forward(h, sp, a.data_ptr(), b.data_ptr(), c.data_ptr(), o0.data_ptr(), o1.data_ptr(), None, None, None, None)
        r2 = lib.ddb_backward(h, sp, d.data_ptr(), e.data_ptr(), o2.data_ptr(), o3.data_ptr(), o4.data_ptr())
        if r1 or r2:
            raise RuntimeError(lib.ddb_last_error(h).decode())

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- the rollout: T = SETS steps forward (one snapshot slot per step), then the reverse sweep ----
    T = SETS
    rng = np.random.default_rng(5 + rank)
    taus = mk(rng.uniform(-1, 1, (T, N, n)))
    gpos, gvel = mk(rng.normal(size=(T, N, n))), mk(rng.normal(size=(T, N, n)))
    poses, vels, gtaus = torch.empty_like(taus), torch.empty_like(taus), torch.empty_like(taus)
    q0d, qd0d = mk(q), mk(qd)
    gq0, gqd0 = torch.empty_like(q0d), torch.empty_like(q0d)
    lib.ddb_set_lanes(h, args.lanes)

    def rollout(t_steps):
        r1 = lib.ddb_rollout_forward(h, sp, t_steps, q0d.data_ptr(), qd0d.data_ptr(), taus.data_ptr(), poses.data_ptr(), vels.data_ptr())
        r2 = lib.ddb_rollout_backward(h, sp, t_steps, gpos.data_ptr(), gvel.data_ptr(), gq0.data_ptr(), gqd0.data_ptr(), gtaus.data_ptr())
        if r1 or r2:
            raise RuntimeError(lib.ddb_last_error(h).decode())

    def run_steps(k):
        while k > 0:
            rollout(min(T, k))
            k -= T

    sampler = ClockSampler(local_rank)
    run_steps(max(args.warmup, 3))
    run_steps(min(args.steps, 64))  # (untimed: the rollout lengths the timed region will use have all been issued once)
    for s in range(SETS):
        step(s)
    barrier()
    # ---- timed region: exactly K steps, device-resident inputs ----
    launches0 = lib.ddb_launch_count(h)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev[0].record(stream)
    run_steps(args.steps)
    ev[1].record(stream)
    barrier()
    elapsed_ms = ev[0].elapsed_time(ev[1])
    launches = int(lib.ddb_launch_count(h) - launches0)
    # ---- the same K steps as separate ddb_forward + ddb_backward calls on one stream (dart_layer's two calls) ----
    barrier()
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev2[0].record(stream)
    for s in range(args.steps):
        step(s)
    ev2[1].record(stream)
    barrier()
    single_ms = ev2[0].elapsed_time(ev2[1])
    # keep the GPU under the same load a little longer if the timed region was too short to sample clocks
    t_end = time.time() + (0.3 if elapsed_ms < 300 else 0.0)
    s = 0
    while time.time() < t_end:
        step(s)
        s += 1
        if s % 64 == 0:
            torch.cuda.synchronize(dev)
    torch.cuda.synchronize(dev)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # ---- per-kernel durations (same launches, events around each kernel) ----
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(min(args.steps, 64))]
    for i, e3 in enumerate(kev):
        (a, b, c, d, e), (o0, o1, o2, o3, o4) = sets[i % SETS]
        lib.ddb_snapshot_select(h, i % SETS)
        e3[0].record(stream)
        lib.ddb_forward(h, sp, a.data_ptr(), b.data_ptr(), c.data_ptr(), o0.data_ptr(), o1.data_ptr(), None, None, None, None)
        e3[1].record(stream)
        lib.ddb_backward(h, sp, d.data_ptr(), e.data_ptr(), o2.data_ptr(), o3.data_ptr(), o4.data_ptr())
        e3[2].record(stream)
    torch.cuda.synchronize(dev)
    fwd_ms = float(np.mean([e3[0].elapsed_time(e3[1]) for e3 in kev]))
    bwd_ms = float(np.mean([e3[1].elapsed_time(e3[2]) for e3 in kev]))
    # ---- second column of SURVEY 8(d): forward + the five materialised Jacobians (BM_Cartpole_Jacobians semantics) ----
    jac = [torch.empty(N, n, n, dtype=torch.float64, device=dev) for _ in range(5)]
    jac_steps = max(3, min(args.steps, 40))

    def jac_step(i):
        (a, b, c, d, e), (o0, o1, o2, o3, o4) = sets[i % SETS]
        lib.ddb_snapshot_select(h, i % SETS)
        r1 = lib.ddb_forward(h, sp, a.data_ptr(), b.data_ptr(), c.data_ptr(), o0.data_ptr(), o1.data_ptr(), None, None, None, None)
        r2 = lib.ddb_jacobians(h, sp, *[t.data_ptr() for t in jac])
        if r1 or r2:
            raise RuntimeError(lib.ddb_last_error(h).decode())

    for i in range(3):
        jac_step(i)
    barrier()
    jev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    jev[0].record(stream)
    for i in range(jac_steps):
        jac_step(i)
    jev[1].record(stream)
    barrier()
    jac_ms = jev[0].elapsed_time(jev[1]) / jac_steps
    # ---- e2e: host buffers through the C ABI's host entry points, copies inside the timed region ----
    hin = [torch.tensor(a, dtype=torch.float64).pin_memory() for a in (q, qd, tau, g[0], g[1])]
    hout = [torch.empty(N, n, dtype=torch.float64).pin_memory() for _ in range(5)]
    e2e_steps = max(3, min(args.steps, 50))

    def e2e_step():
        r1 = lib.ddb_forward_host(h, hin[0].data_ptr(), hin[1].data_ptr(), hin[2].data_ptr(), hout[0].data_ptr(), hout[1].data_ptr())
        r2 = lib.ddb_backward_host(h, hin[3].data_ptr(), hin[4].data_ptr(), hout[2].data_ptr(), hout[3].data_ptr(), hout[4].data_ptr())
        if r1 or r2:
            raise RuntimeError(lib.ddb_last_error(h).decode())

    lib.ddb_snapshot_select(h, 0)
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    # ---- BASELINE configs[0] as a batch (its own handle): 2^20 cartpoles ----
    try:
        cartpole_ms = cartpole_bench(lib, local_rank, rank, args.lanes)
    except Exception as exc:  # optional workload: report, never fail the headline
        cartpole_ms = float("nan")
        print("cartpole batch failed: %s" % str(exc)[:200], file=sys.stderr)
    # ---- second workload: contact-rich box stacking (its own handle) ----
    contact_ms, contact_tiers = float("nan"), {}
    if not args.no_contact:
        contact_ms, contact_tiers, _ = contact_bench(local_rank, rank)
    atlas = None
    if not args.no_contact:
        try:
            atlas = atlas_bench(local_rank, rank, world_size)
        except Exception as exc:  # optional workload: report, never fail the headline
            atlas = {"error": str(exc)[:200]}
    atlas_ms = atlas["ms_per_step"] if atlas and "ms_per_step" in atlas else float("nan")
    walker = None
    if not args.no_contact:
        try:
            walker = walker_bench(local_rank, rank)
        except Exception as exc:  # optional workload: report, never fail the headline
            walker = {"error": str(exc)[:200]}
    walker_ms = walker["ms_per_rollout"] if walker and "ms_per_rollout" in walker else float("nan")
    cheetah = None
    if not args.no_contact:
        try:
            cheetah = cheetah_bench(local_rank, rank)
        except Exception as exc:  # optional workload: report, never fail the headline
            cheetah = {"error": str(exc)[:200]}
    cheetah_ms = cheetah["ms_per_rollout"] if cheetah and "ms_per_rollout" in cheetah else float("nan")
    # ---- reduce over ranks (max time) ----
    t = torch.tensor([elapsed_ms, e2e_s * 1e3, fwd_ms, bwd_ms, contact_ms, single_ms, atlas_ms, jac_ms, walker_ms, cartpole_ms, cheetah_ms], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, fwd_ms, bwd_ms, contact_ms, single_ms, atlas_ms, jac_ms, walker_ms, cartpole_ms, cheetah_ms = [float(x) for x in t.cpu()]
    peak = ctypes_double()
    lib.ddb_measure_fp64_peak(local_rank, C.byref(peak))
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        else:
            hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        value = world_size * N * args.steps / (elapsed_ms * 1e-3)
        dom = ("k_backward", bwd_ms, BWD_BYTES, BWD_FLOPS, BWD_TRAFFIC) if bwd_ms >= fwd_ms else ("k_forward", fwd_ms, FWD_BYTES, FWD_FLOPS, FWD_TRAFFIC)
        achieved = dom[2] * N / (dom[1] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "worlds_per_gpu": N, "n_dofs": n, "dt": 1e-3, "rollout_steps": T, "lanes": args.lanes,
                       "cache": "every rollout writes and re-reads 8 snapshot slots + 8 steps of inputs/outputs (590 MB >> 126 MB L2): steps read cold data"},
            "clocks": sampler.result(),
            "e2e": {"value": world_size * N * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 5 * N * n * 8,
                    "d2h_bytes_per_step": 5 * N * n * 8, "steps": e2e_steps,
                    "path": "ddb_forward_host + ddb_backward_host with pinned host buffers (the kernels read and write the pinned buffers themselves through shared memory; DDB_HOST_ZEROCOPY=0 selects the cudaMemcpyAsync chunk pipeline)"},
            "gpu_launches": launches,
            "kernels": {"k_forward_ms": fwd_ms, "k_backward_ms": bwd_ms, "note": "whole-batch launches timed alone on one stream, rotating cold sets"},
            "single_step": {"value": world_size * N * args.steps / (single_ms * 1e-3), "unit": UNIT, "ms_per_step": single_ms / args.steps,
                            "path": "ddb_forward + ddb_backward per step on one stream, 8 rotating cold sets"},
            "forward_plus_jacobians": {"value": world_size * N / (jac_ms * 1e-3), "unit": "world-steps/s", "ms_per_step": jac_ms,
                                       "path": "ddb_forward + ddb_jacobians (pos-pos, pos-vel, vel-pos, vel-vel, force-vel as [N, n, n]) per step: the shape of "
                                               "BM_Cartpole_Jacobians (unittests/benchmarks/bench_Jacobians.cpp:131-139) on config 2"},
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": dom[4], "peak_source": peak_src,
                         "note": "FP64-pipe bound, not HBM bound: see roofline_fp64 and DESIGN.md §4"},
            "roofline_fp64": {"peak": peak.value, "unit": "TFLOP/s", "peak_source": "ddb_measure_fp64_peak (DFMA micro-benchmark, this run)",
                              "kernel": dom[0], "achieved": dom[3] * N / (dom[1] * 1e-3) / 1e12, "frac": dom[3] * N / (dom[1] * 1e-3) / 1e12 / peak.value,
                              "whole_step_achieved": (FWD_FLOPS + BWD_FLOPS) * N * args.steps / (elapsed_ms * 1e-3) / 1e12,
                              "whole_step_frac": (FWD_FLOPS + BWD_FLOPS) * N * args.steps / (elapsed_ms * 1e-3) / 1e12 / peak.value,
                              "note": "flops per world as executed (ncu, profiles/r02a_ncu_full_summary.csv); whole_step = the timed rollout (chunks of both kernels overlap on the lanes)"},
        }
        if cartpole_ms == cartpole_ms:
            cp_bytes = 104 * 2  # 104 n bytes per world-step (VJP)
            line["cartpole_batch"] = {"workload": CARTPOLE_WORKLOAD, "value": world_size * CARTPOLE_WORLDS / (cartpole_ms * 1e-3), "unit": UNIT,
                                      "ms_per_step": cartpole_ms, "worlds_per_gpu": CARTPOLE_WORLDS, "scaling": "weak",
                                      "algorithmic_gb_per_s": cp_bytes * CARTPOLE_WORLDS / (cartpole_ms * 1e-3) / 1e9,
                                      "snapshot_gb_per_s": 2 * 2 * 13 * 16 * CARTPOLE_WORLDS / (cartpole_ms * 1e-3) / 1e9,
                                      "note": "snapshot_gb_per_s = the 208 B per body the forward writes and the backward reads back (the traffic that binds this small model)"}
        if not args.no_contact:
            line["contact_rich"] = {"workload": CONTACT_WORKLOAD, "value": world_size * CONTACT_WORLDS / (contact_ms * 1e-3), "unit": UNIT,
                                    "ms_per_step": contact_ms, "worlds_per_gpu": CONTACT_WORLDS, "scaling": "weak",
                                    "solver_tier_mask_hist_rank0": contact_tiers,
                                    "tier_bits": "1 short-circuit through the cached solution, 2 Dantzig, 4 PGS, 8 friction dropped"}
            if world_size == 1 and not args.no_cpu_baseline:
                try:
                    cv, cs = contact_cpu_baseline()
                    line["contact_rich"]["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": 1, "kind": "port",
                                                            "sample": "1024 worlds, forwardPass + the reference's Jacobian recipe on one core (%.1f s)" % cs}
                except Exception as exc:  # the oracle's reference libraries travel prebuilt; report instead of failing the headline
                    line["contact_rich"]["cpu_baseline"] = {"unavailable": str(exc)[:200]}
        if atlas is not None:
            if "ms_per_step" in atlas:
                line["atlas_contact"] = {"workload": ATLAS_WORKLOAD, "value": world_size * ATLAS_WORLDS / (atlas_ms * 1e-3), "unit": UNIT, "ms_per_step": atlas_ms,
                                         "worlds_per_gpu": ATLAS_WORLDS, "scaling": "weak", "shared_gradient_allreduce_us_rank0": atlas["allreduce_us"],
                                         "solver_tier_mask_hist_rank0": atlas["tiers"], "finite": atlas["finite"]}
            else:
                line["atlas_contact"] = atlas
        if walker is not None:
            if "ms_per_rollout" in walker:
                line["walker_rollout"] = {"workload": WALKER_WORKLOAD, "value": world_size * WALKER_WORLDS * WALKER_STEPS / (walker_ms * 1e-3), "unit": UNIT,
                                          "ms_per_rollout": walker_ms, "worlds_per_gpu": WALKER_WORLDS, "rollout_steps": WALKER_STEPS, "scaling": "weak",
                                          "finite": walker["finite"]}
            else:
                line["walker_rollout"] = walker
        if cheetah is not None:
            if "ms_per_rollout" in cheetah:
                line["half_cheetah_rollout"] = {"workload": CHEETAH_WORKLOAD, "value": world_size * CHEETAH_WORLDS * CHEETAH_STEPS / (cheetah_ms * 1e-3), "unit": UNIT,
                                                "ms_per_rollout": cheetah_ms, "worlds_per_gpu": CHEETAH_WORLDS, "rollout_steps": CHEETAH_STEPS, "scaling": "weak",
                                                "finite": cheetah["finite"],
                                                "worlds_with_uncarried_pipe_contact_at_first_step_rank0": cheetah["worlds_with_uncarried_pipe_contact_at_last_step"]}
            else:
                line["half_cheetah_rollout"] = cheetah
        if world_size == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v1, s1 = cpu_world_steps(N_WORLDS, threads)
            reps = int(min(200, max(1, round(12.0 / max(s1, 1e-3)))))  # about 12 s of CPU work
            v, secs = cpu_world_steps(N_WORLDS, threads, reps=reps)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d passes over the same 65536-world batch (%.1f s); CPU restatement of the reference arithmetic incl. its finite-difference Jacobian recipes (reference not buildable here: no Eigen)" % (reps, secs)}
        print(json.dumps(line))
    lib.ddb_destroy(h)
    if world_size > 1:
        dist.destroy_process_group()
    return 0


def ctypes_double():
    return C.c_double(0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-contact", action="store_true", help="skip the second (contact-rich) workload")
    ap.add_argument("--lanes", type=int, default=2, help="internal streams the rollout spreads chunks of worlds over")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
