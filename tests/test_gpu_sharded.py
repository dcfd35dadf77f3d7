"""GPU, 2 ranks (skipped on a single-GPU box): ShardedPool over NCCL -- each rank steps its
env-id block, one all-gather of the packed outputs per step, and the gathered batch equals
the oracle's full batch (toy_text: bit-exact); then the same through the engine's own peer
exchange (CUDA IPC handles, NVLink stores, sequence flags)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world, device_id=torch.device("cuda", rank))
    from envpool_b200.sharded import ShardedPool
    from oracle.oracle_lib import OraclePool

    n = 4096
    pool = ShardedPool("FrozenLake-v1", n, seed=5, device=rank)
    orc = OraclePool("FrozenLake", n, seed=5, max_episode_steps=100, iopt=4)
    rng = np.random.default_rng(0)
    ok = True
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
    # the engine's own peer exchange (CUDA IPC + NVLink stores) must agree with both
    pool2 = ShardedPool("FrozenLake-v1", n, seed=5, device=rank)
    orc2 = OraclePool("FrozenLake", n, seed=5, max_episode_steps=100, iopt=4)
    pool2.enable_peer_exchange()
    want = orc2.reset()
    full = pool2.reset_exchange()
    for t in range(25):
        pool2.pool.sync()
        for k, w in want.items():
            g = full[k].reshape((n,) + tuple(full[k].shape[2:])).cpu().numpy()
            ok &= bool(np.array_equal(g, w))
        a = rng.integers(0, 4, size=n).astype(np.int32)
        full = pool2.step_exchange(
            torch.from_numpy(a[pool2.offset:pool2.offset + pool2.count]).cuda())
        want = orc2.step(a)
    pool2.pool.sync()
    steps, timed_out = pool2.pool.exchange_status()
    ok &= (steps == 26) and not timed_out
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_pool_matches_oracle():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret[0] is True and ret[1] is True
