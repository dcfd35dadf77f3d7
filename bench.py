#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched step() hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W              # our arm (default)
    python bench.py --impl reference --gpus 1 --steps K --warmup W
    torchrun ... bench.py --gpus N ...                          # one rank per GPU

A "step" is one sync step of one batch: every env of the pool advances once.  Workload at
N=1: BASELINE.json configs[1], CartPole-v1 with num_envs=65536 on one B200 (weak scaling for
N>1: 65536 envs per GPU, env ids sharded, one NCCL all-gather of the output columns per
step).  `value` is measured with actions and outputs resident in HBM (device-resident C-ABI
path, K single-step launches replayed from a CUDA graph); `e2e` is the same metric through
the reference-facing API (envpool_b200.make(...).step(numpy)) with host buffers, H2D/D2H
inside the timed region.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# task id -> (engine task, registered max_episode_steps, iopt, #discrete actions or None)
TASKS = {
    "CartPole-v1": ("CartPole", 500, -1, 2),
    "Pendulum-v1": ("Pendulum", 200, 1, None),
    "Acrobot-v1": ("Acrobot", 500, -1, 3),
    "MountainCar-v0": ("MountainCar", 200, -1, 3),
    "FrozenLake-v1": ("FrozenLake", 100, 4, 4),
    "Catch-v0": ("Catch", -1, -1, 3),
    "Taxi-v3": ("Taxi", 200, -1, 6),
    "HalfCheetah-v4": ("HalfCheetah", 1000, -1, None),
}
METRIC = "env steps/sec (whole box)"
L2_BYTES = 126 * 1024 * 1024


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the bench runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.samples = []
        self.proc = None
        self.t = None
        self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                smax = float(parts[1])
                if t0 - 0.05 <= ts <= t1 + 0.05:
                    sm.append(float(parts[0]))
                    power.append(float(parts[2]))
                    for nm, v in zip(names, parts[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples_under_load": len(sm),
                "power_w_max": max(power) if power else None}


def bind_to_gpu_numa(local: int):
    """Pin this process (and the pinned host buffers it allocates from now on) to the CPUs of
    the NUMA node the GPU hangs off: a host thread on the far socket pays the inter-socket
    hop on every pinned-memory copy and every doorbell.  Returns what was done."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return {"pci": bdf, "node": node, "bound": False}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"pci": bdf, "node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"pci": bdf, "node": node, "bound": True, "cpus": len(cpus)}
    except Exception as exc:  # never fatal: the bench runs unbound
        return {"bound": False, "error": f"{type(exc).__name__}: {exc}"}


def make_action_stream(torch, task, n, device, min_bytes):
    """[T, N, ...] synthetic actions on the device, larger than L2 so no row is re-read
    while it could still be cached."""
    eng, _, _, nact = TASKS[task]
    row = n * (48 if eng == "HalfCheetah" else 4)
    T = max(64, -(-min_bytes // row))
    T = -(-T // 64) * 64
    g = torch.Generator(device=device)
    g.manual_seed(1)
    if eng == "HalfCheetah":
        a = torch.rand((T, n, 6), generator=g, device=device, dtype=torch.float64) * 2 - 1
    elif nact is None:
        a = (torch.rand((T, n, 1), generator=g, device=device, dtype=torch.float32) * 4 - 2)
    else:
        a = torch.randint(0, nact, (T, n), generator=g, device=device, dtype=torch.int32)
    return a


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the env-step engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from envpool_b200 import _capi

    eng, ms, iopt, _ = TASKS[args.task]
    n = args.num_envs
    pool = _capi.CPool(eng, n, seed=0, max_episode_steps=ms, iopt=iopt, device=local,
                       precision=args.precision, env_id_offset=rank * n)
    stream = torch.cuda.ExternalStream(pool.stream, device=dev)
    actions = make_action_stream(torch, args.task, n, dev, 2 * L2_BYTES)
    T = actions.shape[0]
    K, W = args.steps, args.warmup
    use_graph = not args.no_graph

    def run_steps(count, t0=0):
        """`count` sync steps of this rank's shard; actions cycle through the [T, N] stream
        starting at row t0.  Env-id sharding needs no data-path collective."""
        if count <= T - t0:
            pool.step_many_device(actions, t0, count, use_graph=use_graph)
            return
        q, r = divmod(count, T)
        for _ in range(q):
            pool.step_many_device(actions, 0, T, use_graph=use_graph)
        if r:
            pool.step_many_device(actions, 0, r, use_graph=use_graph)

    # A short timed region (the driver passes --steps 20) replays one K-step graph.  It reads
    # action rows [t_timed, t_timed + K) -- rows the lead-in does not touch and the L2 flush
    # has evicted -- so the action stream is as cold as in a long run.
    t_timed = (T // 2) if K <= T // 2 else 0
    flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)

    pool.reset_device()
    pool.sync()
    sampler = ClockSampler(local)
    if rank == 0 and not args.profile:
        sampler.start()
    t_load0 = time.time()
    # warm-up (also captures the CUDA graphs); stretched so that clocks are sampled under
    # load for ~1 s before the timed region starts
    run_steps(W)
    pool.sync()
    lead_reps = 2
    if args.profile and use_graph:   # graphs captured outside the timed region
        run_steps(K)
        run_steps(K, t_timed)
        pool.sync()
    if not args.profile:
        t_w = time.time()
        while True:  # also instantiates every CUDA graph the timed region will replay
            run_steps(K)
            run_steps(K, t_timed)
            pool.sync()
            if time.time() - t_w >= 1.0:
                break
        # how many K-step replays keep the GPU busy for ~0.5 ms (the lead-in below)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run_steps(K)
        e1.record(stream)
        torch.cuda.synchronize()
        lead_reps = int(min(200, max(2, 0.5 / max(e0.elapsed_time(e1), 1e-3))))
    launches0 = pool.launch_count
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Timed region: L2 flush, then an UNTIMED lead-in of the same launches, ev0, the K timed
    # steps, ev1 -- enqueued back to back with no host synchronisation in between, so the GPU
    # is still busy with the lead-in while the host enqueues the timed launches and the host's
    # launch latency stays outside ev0..ev1 whatever --steps is.
    with torch.cuda.stream(stream):
        flush.fill_(1)
    lead_launches = 0
    if not args.profile:
        for _ in range(lead_reps):
            run_steps(K)
        lead_launches = pool.launch_count - launches0
    ev0.record(stream)
    run_steps(K, t_timed)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_load1 = time.time()
    ms_total = ev0.elapsed_time(ev1)
    launches = pool.launch_count - launches0 - lead_launches
    if world > 1:
        tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    clocks = sampler.stop(t_load0, t_load1) if rank == 0 and not args.profile else None
    ms_per_step = ms_total / K
    value = n * world * K / (ms_total * 1e-3)

    result = None
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        bpe = pool.bytes_per_env_step
        achieved = bpe * n / (ms_per_step * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(f"{args.task}:{n}:{args.precision}")
        except Exception:
            pass
        result = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == "f64" else "f32", "data": "synthetic",
            "config": {
                "workload": f"{args.task} sync num_envs={n} per GPU x {world} GPU"
                            + (" (BASELINE.json configs[1])"
                               if (args.task, n) == ("CartPole-v1", 65536) else ""),
                "api": "device-resident C-ABI single-step kernel, one launch per step"
                       + (", CUDA-graph replay" if use_graph else "")
                       + ("; env ids sharded over ranks, no data-path collective in `value` "
                          "(see with_allgather for the exchange step)" if world > 1 else ""),
                "l2": f"L2 flushed (256 MiB fill) before the timed region; action stream "
                      f"{actions.numel() * actions.element_size() >> 20} MiB > 126 MiB L2, "
                      f"the timed steps read rows the untimed lead-in does not touch; the "
                      f"recurrent env state and output slab ({bpe * n >> 10} KiB) stay on "
                      f"chip by construction at this num_envs",
                "timing": f"CUDA events on the launching stream around exactly {K} steps, "
                          f"preceded by {lead_reps} untimed replays enqueued without a host "
                          f"sync (host launch latency outside the window)",
                "precision": args.precision, "seed": 0,
            },
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "bytes_per_env_step": bpe, "peak_source": peak_src,
                         "kernel": "step_kernel<%s>" % eng},
            "clocks": clocks,
        }
    # ---- the exchange step of north_star: all-gather of the packed outputs per step --------
    if world > 1 and not args.profile:
        ag = run_allgather(args, torch, dist, pool, actions, dev, stream, world)
        px = run_peer_exchange(args, torch, dist, pool, actions, dev, world, rank)
        if rank == 0:
            # headline exchange number = the engine's own NVLink peer exchange; the NCCL
            # all-gather of the same slab is kept beside it as the library baseline
            if "value" in px:
                result["with_allgather"] = px
                result["with_allgather_nccl"] = ag
            else:
                result["with_allgather"] = ag
                result["with_allgather_peer"] = px
    # ---- fused rollout API: T steps per launch, state in registers ----------------------
    if world == 1 and (not args.profile or args.profile_rollout):
        ro = run_rollout(args, torch, pool, actions, dev)
        if rank == 0:
            result["rollout"] = ro
    # ---- e2e through the public API with host buffers --------------------------------
    if not args.profile:
        e2e = run_e2e(args, torch, local, rank, world)
        if rank == 0:
            result["e2e"] = e2e
    if rank == 0 and world == 1 and not args.profile and not args.no_cpu:
        try:
            _, nt, sweep = sweep_reference_threads(args)
        except Exception as exc:  # the sweep is a courtesy to the baseline, never fatal
            nt, sweep = 0, {"error": f"{type(exc).__name__}: {exc}"}
        result["cpu_baseline"] = cpu_baseline(args, budget_s=args.cpu_seconds, num_threads=nt)
        result["cpu_baseline"]["threads_swept"] = sweep
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_allgather(args, torch, dist, pool, actions, dev, stream, world):
    """Every step followed by ONE NCCL all-gather of the packed output slab, so that every
    GPU holds the outputs of all `world * num_envs` envs (north_star's exchange step)."""
    from envpool_b200._capi import _torch_view
    from envpool_b200.sharded import all_gather_packed

    n, T = pool.n, actions.shape[0]
    slab = _torch_view(pool.outputs_device_ptr(), (pool.slab_bytes,), torch.uint8, dev.index)
    full = torch.empty((world, pool.slab_bytes), dtype=torch.uint8, device=dev)
    steps = int(min(max(args.steps, 64), 2048)) // 64 * 64
    chunk = 64
    side = torch.cuda.Stream(device=dev)

    def body(k0):
        for k in range(chunk):
            pool.step_device(actions[(k0 + k) % T],
                             stream=torch.cuda.current_stream(dev).cuda_stream)
            all_gather_packed(slab, full)

    # warm-up (also initialises NCCL's channels for this size), then capture `chunk` steps
    # -- step kernel + all-gather each -- into one CUDA graph so that the loop is not bound by
    # the host's launch rate
    mode = "cuda-graph replay of 64-step chunks"
    with torch.cuda.stream(side):
        body(0)
        torch.cuda.synchronize()
        graph = None
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                body(0)
            graph = g
        except Exception as exc:  # capture unsupported: fall back to eager launches
            mode = f"eager launches (graph capture failed: {type(exc).__name__})"
            torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(side)
        for c in range(steps // chunk):
            if graph is not None:
                graph.replay()
            else:
                body(c * chunk)
        ev1.record(side)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    tt = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    gathered = pool.slab_bytes * (world - 1)
    return {"value": n * world * steps / (ms * 1e-3), "unit": "env-steps/s", "steps": steps,
            "ms_per_step": ms / steps, "allgather_bytes_in_per_gpu_per_step": gathered,
            "nvlink_gbs_in_per_gpu": gathered * steps / (ms * 1e-3) / 1e9,
            "api": "step_device + one ncclAllGather of the packed output slab per step; " + mode}


def run_peer_exchange(args, torch, dist, pool, actions, dev, world, rank):
    """Every step written into this rank's slice of the gather buffer and pushed to all
    peers by the engine itself (epb_step_exchange_device + epb_exchange_wait: CUDA-IPC
    mapped peer memory, NVLink stores, sequence flags) -- no library collective."""
    T = actions.shape[0]

    def all_ok(ok):
        f = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        return bool(f.item())

    err, handle = "", None
    try:
        handle = pool.exchange_init(world, rank)
    except Exception as exc:
        err = f"exchange_init: {exc}"
    if not all_ok(handle is not None):
        return {"unavailable": err or "a peer failed exchange_init"}
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    try:
        pool.exchange_attach_ipc(handles)
    except Exception as exc:
        err = f"exchange_attach_ipc: {exc}"
    if not all_ok(not err):
        return {"unavailable": err or "a peer failed exchange_attach_ipc"}

    steps = int(min(max(args.steps, 64), 2048)) // 64 * 64
    chunk = 64  # even: the gather halves alternate with the step count
    side = torch.cuda.Stream(device=dev)

    def body(k0):
        for k in range(chunk):
            pool.step_exchange(actions[(k0 + k) % T], stream=side.cuda_stream)
            pool.exchange_wait(stream=side.cuda_stream)

    mode = "cuda-graph replay of 64-step chunks"
    with torch.cuda.stream(side):
        body(0)
        torch.cuda.synchronize()
        graph = None
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                body(0)
            graph = g
        except Exception as exc:
            mode = f"eager launches (graph capture failed: {type(exc).__name__})"
            torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(side)
        for c in range(steps // chunk):
            if graph is not None:
                graph.replay()
            else:
                body(c * chunk)
        ev1.record(side)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    tt = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    pushed, timed_out = pool.exchange_status()
    bad = torch.tensor([1 if timed_out else 0], device=dev, dtype=torch.int32)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX)
    if bad.item():
        return {"unavailable": "a peer wait hit its time bound"}
    dist.barrier()
    gathered = pool.slab_bytes * (world - 1)
    return {"value": pool.n * world * steps / (ms * 1e-3), "unit": "env-steps/s", "steps": steps,
            "ms_per_step": ms / steps, "allgather_bytes_in_per_gpu_per_step": gathered,
            "nvlink_gbs_in_per_gpu": gathered * steps / (ms * 1e-3) / 1e9,
            "steps_pushed_total": pushed,
            "api": "epb_step_exchange_device + epb_exchange_wait: step kernel writes its gather "
                   "slice, push kernel stores it into every peer over NVLink (CUDA IPC), "
                   "sequence-flag wait; " + mode}


def run_rollout(args, torch, pool, actions, dev):
    """epb_rollout_device: T consecutive sync steps of all envs in ONE launch (the fused
    random-action rollout of north_star); every step's outputs are written to [T, N, ...]
    arrays in HBM, actions are read from the [T, N] stream, env state lives in registers."""
    n = pool.n
    out_row = sum(k.row_bytes for k in pool.keys)
    T = int(max(4, min(64, (2 << 30) // max(1, out_row * n), actions.shape[0])))
    tdt = {np.dtype(np.int32): torch.int32, np.dtype(np.float32): torch.float32,
           np.dtype(np.float64): torch.float64, np.dtype(np.bool_): torch.bool}
    cols = [torch.empty((T, n) + k.shape, dtype=tdt[k.dtype], device=dev) for k in pool.keys]
    stream = torch.cuda.ExternalStream(pool.stream, device=dev)
    launches = max(3, min(200, args.steps // T))
    nrows = actions.shape[0] // T
    for i in range(3):
        pool.rollout_device(actions[(i % nrows) * T:], T, cols)
    pool.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(launches):
        pool.rollout_device(actions[(i % nrows) * T:], T, cols)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    steps = launches * T
    state_bytes = pool.bytes_per_env_step - pool.action_key.row_bytes - out_row
    bpe = pool.action_key.row_bytes + out_row + state_bytes / T
    peak, _ = measured_peak_hbm()
    achieved = bpe * n * steps / (ms * 1e-3) / 1e9
    return {"value": n * steps / (ms * 1e-3), "unit": "env-steps/s", "T_per_launch": T,
            "launches": launches, "ms_per_step": ms / steps,
            "bytes_per_env_step": bpe,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak},
            "api": "epb_rollout_device: T steps per launch, outputs [T,N,...] written to HBM"}


def run_e2e(args, torch, local, rank, world):
    """Same metric through envpool_b200.make(...).step(numpy): per step the action batch is
    copied host->device (via pinned staging) and every state column comes back device->host
    into a pinned slab that the returned numpy arrays view."""
    import envpool_b200

    eng, ms, iopt, nact = TASKS[args.task]
    n = args.num_envs
    env = envpool_b200.make(args.task, env_type="gymnasium", num_envs=n, seed=0,
                            device=local, precision=args.precision, env_id_offset=rank * n)
    rng = np.random.default_rng(1)
    Ta = 64
    if eng == "HalfCheetah":
        acts = rng.uniform(-1, 1, size=(Ta, n, 6))
    elif nact is None:
        acts = rng.uniform(-2, 2, size=(Ta, n, 1)).astype(np.float32)
    else:
        acts = rng.integers(0, nact, size=(Ta, n)).astype(np.int32)
    env.reset()
    E = int(min(max(args.steps, 50), 400))
    for t in range(20):
        env.step(acts[t % Ta])
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    checksum = 0.0
    for t in range(E):
        obs, rew, term, trunc, info = env.step(acts[t % Ta])
        checksum += float(rew[0])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([dt], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    dp = env.device_pool
    return {"value": n * world * E / dt, "unit": "env-steps/s", "steps": E,
            "h2d_bytes_per_step": int(acts[0].nbytes), "d2h_bytes_per_step": int(dp.slab_bytes),
            "ms_per_step": dt / E * 1e3,
            "api": "envpool_b200.make(task,'gymnasium').step(numpy actions) -> numpy"}


def cpu_baseline(args, budget_s=15.0, steps=None, warmup=3, num_threads=0):
    """The reference's own CPU thread pool (oracle/_ref = its AsyncEnvPool + env headers
    compiled from /root/reference) timed on this box's host cores; falls back to the scalar
    oracle port when _ref is absent."""
    eng, ms, iopt, nact = TASKS[args.task]
    n = args.num_envs
    rng = np.random.default_rng(1)
    from oracle import ref_lib

    cores = os.cpu_count() or 1
    if ref_lib.available() and eng != "HalfCheetah":
        Ta = 16
        if nact is None:
            acts = rng.uniform(-2, 2, size=(Ta, n, 1)).astype(np.float32)
        else:
            acts = rng.integers(0, nact, size=(Ta, n)).astype(np.int32)
        pool = ref_lib.RefPool(eng, n, seed=0, max_episode_steps=ms, iopt=iopt,
                               num_threads=num_threads)
        if steps is None:
            probe = pool.bench(acts, 1, 3) / 3
            steps = int(min(max(budget_s / max(probe, 1e-6), 5), 5000))
        dt = pool.bench(acts, warmup, steps)
        pool.close()
        used = min(n, cores) if num_threads <= 0 else num_threads
        return {"value": n * steps / dt, "unit": "env-steps/s", "cores": used,
                "kind": "reference", "ms_per_step": dt / steps * 1e3,
                "sample": f"{steps} sync steps of the same {n}-env batch workload, "
                          f"AsyncEnvPool num_threads={used} of {cores} hw threads"
                          + (" (its default: min(batch, hw threads))" if num_threads <= 0
                             else "")}
    from oracle.oracle_lib import OraclePool

    pool = OraclePool(eng, n, seed=0, max_episode_steps=ms, iopt=iopt)
    if eng == "HalfCheetah":
        acts = rng.uniform(-1, 1, size=(4, n, 6))
    elif nact is None:
        acts = rng.uniform(-2, 2, size=(4, n, 1)).astype(np.float32)
    else:
        acts = rng.integers(0, nact, size=(4, n)).astype(np.int32)
    pool.reset()
    t0 = time.perf_counter()
    pool.step(acts[0])
    probe = time.perf_counter() - t0
    if steps is None:
        steps = int(min(max(budget_s / max(probe, 1e-6), 2), 2000))
    t0 = time.perf_counter()
    for t in range(steps):
        pool.step(acts[t % 4])
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "env-steps/s", "cores": 1, "kind": "port",
            "ms_per_step": dt / steps * 1e3,
            "sample": f"{steps} sync steps of the same {n}-env batch, scalar C port"}


def sweep_reference_threads(args):
    """The reference's default thread count (every hardware thread) is not its best on a
    many-core host: each dequeue is serialised by one semaphore
    (action_buffer_queue.h:71-80) and the workers spin.  Probe a few thread counts on a
    moderate batch and give the reference the best one -- its own benchmark script takes
    num_threads as a knob too (benchmark/test_envpool.py).  Returns (best probe result,
    best num_threads (0 = the default), {threads: env-steps/s})."""
    probe_args = argparse.Namespace(**vars(args))
    probe_args.num_envs = min(args.num_envs, 16384)
    probe = cpu_baseline(probe_args, steps=3, warmup=1)
    cores = os.cpu_count() or 1
    sweep = {str(probe["cores"]): probe["value"]}
    best = 0
    if probe["kind"] == "reference":
        cand = {max(1, cores // d) for d in (2, 4, 8, 16)} | {1, 4}
        for nt in sorted(cand - {probe["cores"]}):
            r = cpu_baseline(probe_args, steps=3, warmup=1, num_threads=nt)
            sweep[str(nt)] = r["value"]
            if r["value"] > probe["value"]:
                probe, best = r, nt
    return probe, best, {k: round(v) for k, v in sweep.items()}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (oracle/_ref = its
    AsyncEnvPool + env headers compiled from /root/reference) on this box's host cores, same
    metric / config keys as our arm.  A "step" is one Send/Recv of a batch; when K full-size
    batches would not finish in a few minutes, each step is a bounded sample of the workload
    (a smaller batch of the same env, same thread count) -- said in `cpu_baseline.sample`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, args.warmup
    n_full = args.num_envs * world   # our arm at this --gpus steps this many envs per step
    budget_s = 150.0
    # probe the per-env-step cost at a moderate batch, then size the per-step sample
    probe_args = argparse.Namespace(**vars(args))
    probe_args.num_envs = min(n_full, 16384)
    try:
        probe, best_threads, sweep = sweep_reference_threads(probe_args)
    except Exception as exc:  # fall back to the reference's default thread count
        probe, best_threads = cpu_baseline(probe_args, steps=3, warmup=1), 0
        sweep = {"error": f"{type(exc).__name__}: {exc}"}
    rate = max(probe["value"], 1.0)
    n_s = int(min(n_full, max(64, rate * budget_s / max(K + W, 1))))
    if n_s < n_full:
        n_s = 1 << (n_s.bit_length() - 1)   # power of two, >= 64
    args.num_envs = n_s
    cb = cpu_baseline(args, steps=K, warmup=W, num_threads=best_threads)
    cb["threads_swept"] = sweep
    if n_s < n_full:
        cb["sample"] = (f"each of the {K} steps is a {n_s}-env batch (bounded sample of the "
                        f"{n_full}-env workload so that the run ends in minutes); "
                        + cb["sample"])
    out = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "env-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.task} sync num_envs={n_full} "
                               + ("(BASELINE.json configs[1]) " if
                                  (args.task, n_full) == ("CartPole-v1", 65536) else "")
                               + f"on the reference CPU thread pool, {cb['cores']} host threads",
                   "note": "the reference has no GPU path; one process uses every host core, "
                           "so the value does not scale with --gpus"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=2000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--task", default="CartPole-v1", choices=sorted(TASKS))
    ap.add_argument("--num-envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--profile", action="store_true",
                    help="kernel loop only (for ncu): no clocks sampler, e2e or cpu legs")
    ap.add_argument("--profile-rollout", action="store_true",
                    help="with --profile: also run the fused rollout leg (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 20000 and args.warmup == 2000:
            args.steps, args.warmup = 100, 3
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
