"""Body of tests/test_gpu_exchange.py::test_captured_exchange_chain_runs_ahead_of_the_waits
(run as a subprocess, see there): two pools play rank 0 / 1 on one device; captured exchange
chains of several lengths against the oracle's full batch, then the same steps as direct
launches, bit-identical."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from helpers import assert_batch_equal  # noqa: E402


def views(pool, ptr, world, n_local):
    import torch

    from envpool_b200._capi import _torch_view
    from envpool_b200.sharded import packed_views

    full = _torch_view(ptr, (world, pool.exchange_slice_bytes), torch.uint8, pool.device)
    return {k: v.reshape((world * n_local,) + tuple(v.shape[2:])).cpu().numpy()
            for k, v in packed_views(full, pool.keys, n_local).items()}


def main(task):
    import torch

    from envpool_b200._capi import CPool
    from oracle.oracle_lib import OraclePool

    kw, n_act, tol = (dict(max_episode_steps=9), 2, 1e-6) if task == "CartPole" else (dict(), 3, 0.0)
    n, world, T = 3000, 2, 24
    rng = np.random.default_rng(4)
    acts = rng.integers(0, n_act, size=(T, world * n)).astype(np.int32)
    d_acts = [torch.from_numpy(np.ascontiguousarray(acts[:, r * n:(r + 1) * n])).cuda()
              for r in range(world)]

    def make():
        pools = [CPool(task, n, seed=3, env_id_offset=r * n, **kw) for r in range(world)]
        for r, p in enumerate(pools):
            p.exchange_init(world, r)
        bases = [p.exchange_base() for p in pools]
        for p in pools:
            p.exchange_attach(bases)
        for p in pools:
            p.step_exchange(None)
        for p in pools:
            p.exchange_wait()
        for p in pools:
            p.sync()
        return pools

    pools, plain = make(), make()
    orc = OraclePool(task, world * n, seed=3, **kw)
    orc.reset()
    assert pools[0].exchange_depth >= 3
    t = 0
    for K in (8, 8, 4, 12, 8):
        ptrs = [p.step_exchange_many(d_acts[r], t % T, K, use_graph=True)
                for r, p in enumerate(pools)]
        for p in pools:
            p.sync()
        ptrs2 = [p.step_exchange_many(d_acts[r], t % T, K, use_graph=False)
                 for r, p in enumerate(plain)]
        for p in plain:
            p.sync()
        for k in range(K):
            want = orc.step(acts[(t + k) % T])
        t += K
        for r, p in enumerate(pools):
            got = views(p, ptrs[r], world, n)
            assert_batch_equal(got, want, task, tol, f"{task} rank {r} after {t} steps")
            got2 = views(plain[r], ptrs2[r], world, n)
            assert_batch_equal(got2, got, task, 0.0, f"{task} rank {r}: direct vs captured")
    for p in pools + plain:
        steps, timed_out = p.exchange_status()
        assert steps == 1 + t and not timed_out, (steps, timed_out)
        p.close()
    print("CHAIN OK", task, t)


if __name__ == "__main__":
    main(sys.argv[1])
