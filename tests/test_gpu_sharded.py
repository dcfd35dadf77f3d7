"""GPU, 2 ranks = 2 processes: ShardedPool -- each rank steps its env-id block and every rank
ends up with the oracle's full batch (toy_text: bit-exact).

* Two GPUs present: one rank per GPU over NCCL; first the library all-gather of the packed
  outputs, then the engine's own peer exchange (CUDA IPC handles, NVLink stores, sequence
  flags), both against the oracle.
* One GPU (the driver's single-GPU box): both processes share device 0.  CUDA IPC maps a
  peer process's allocation on the same device just as well, so the whole cross-process
  exchange protocol (IPC attach, ring slots, credit / data / ack flags, receiver-side
  re-expansion of the common columns, the captured chain with the waits on a parallel branch)
  runs; only the transport is local memory instead of NVLink.  NCCL refuses two ranks on one
  device, so the rendezvous is gloo and the NCCL cross-check is left to the 2-GPU case."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret, ndev):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    dev = rank % ndev
    torch.cuda.set_device(dev)
    if ndev >= world:
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world)
    from envpool_b200._capi import _torch_view
    from envpool_b200.sharded import ShardedPool, packed_views
    from oracle.oracle_lib import OraclePool

    n = 4096
    rng = np.random.default_rng(0)
    ok = True
    if ndev >= world:
        pool = ShardedPool("FrozenLake-v1", n, seed=5, device=dev)
        orc = OraclePool("FrozenLake", n, seed=5, max_episode_steps=100, iopt=4)
        pool.reset_device()
        want = orc.reset()
        for t in range(25):
            for packed in (True, False):
                full = pool.all_gather(packed=packed)
                torch.cuda.synchronize()
                for k, w in want.items():
                    g = full[k].reshape((n,) + tuple(full[k].shape[2:] if packed else
                                                     full[k].shape[1:])).cpu().numpy()
                    ok &= bool(np.array_equal(g, w))
            a = rng.integers(0, 4, size=n).astype(np.int32)
            pool.step_device(torch.from_numpy(a[pool.offset:pool.offset + pool.count]).cuda())
            want = orc.step(a)
    # the engine's own peer exchange (CUDA IPC + peer stores) against the oracle
    for task, eng, kw, n_act, tol in (
            ("FrozenLake-v1", "FrozenLake", dict(max_episode_steps=100, iopt=4), 4, 0.0),
            ("CartPole-v1", "CartPole", dict(max_episode_steps=500), 2, 1e-6)):
        pool2 = ShardedPool(task, n, seed=5, device=dev)
        orc2 = OraclePool(eng, n, seed=5, **kw)
        pool2.enable_peer_exchange()
        want = orc2.reset()
        full = pool2.reset_exchange()

        def check(full, want):
            good = True
            for k, w in want.items():
                g = full[k].reshape((n,) + tuple(full[k].shape[2:])).cpu().numpy()
                if tol == 0.0 or g.dtype.kind in "ib":
                    good &= bool(np.array_equal(g, w))
                else:
                    good &= bool(np.all(np.abs(g - w) <= tol * (1 + np.abs(w))))
            return good

        for t in range(25):
            pool2.pool.sync()
            ok &= check(full, want)
            a = rng.integers(0, n_act, size=n).astype(np.int32)
            full = pool2.step_exchange(
                torch.from_numpy(a[pool2.offset:pool2.offset + pool2.count]).cuda())
            want = orc2.step(a)
        pool2.pool.sync()
        ok &= check(full, want)
        # the captured chain: K exchanged steps, waits on a parallel graph branch
        T, K = 16, 8
        acts = rng.integers(0, n_act, size=(T, n)).astype(np.int32)
        d_acts = torch.from_numpy(
            np.ascontiguousarray(acts[:, pool2.offset:pool2.offset + pool2.count])).cuda()
        for rep in range(3):
            ptr = pool2.pool.step_exchange_many(d_acts, (rep * K) % T, K, use_graph=True)
            for k in range(K):
                want = orc2.step(acts[(rep * K + k) % T])
            pool2.pool.sync()
            raw = _torch_view(ptr, (pool2.world, pool2.pool.exchange_slice_bytes), torch.uint8,
                              pool2.pool.device)
            ok &= check(packed_views(raw, pool2.pool.keys, pool2.count), want)
        steps, timed_out = pool2.pool.exchange_status()
        ok &= (steps == 26 + 3 * K) and not timed_out
        dist.barrier()
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_sharded_pool_matches_oracle():
    import torch
    import torch.multiprocessing as mp

    ndev = torch.cuda.device_count()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret, ndev)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret[0] is True and ret[1] is True
