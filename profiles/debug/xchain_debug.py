"""Debug: two pools on one device, captured exchange chains of various K; prints wall time
and exchange status per chain."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("ENVPOOL_B200_EXCHANGE_TIMEOUT_S", "2")
import torch  # noqa: E402

from envpool_b200._capi import CPool  # noqa: E402

task, kw, n_act = sys.argv[1], {}, 2
if task == "CartPole":
    kw = dict(max_episode_steps=9)
Ks = [int(x) for x in sys.argv[2].split(",")]
use_graph = sys.argv[3] != "direct" if len(sys.argv) > 3 else True
n, world, T = 3000, 2, 24
rng = np.random.default_rng(4)
acts = rng.integers(0, n_act, size=(T, world * n)).astype(np.int32)
d_acts = [torch.from_numpy(np.ascontiguousarray(acts[:, r * n:(r + 1) * n])).cuda()
          for r in range(world)]
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
t = 0
for K in Ks:
    t0 = time.time()
    for r, p in enumerate(pools):
        p.step_exchange_many(d_acts[r], t % T, K, use_graph=use_graph)
        print(f"  K={K} rank {r} enqueued after {time.time() - t0:.3f}s", flush=True)
    for p in pools:
        p.sync()
    t += K
    print(f"K={K} done in {time.time() - t0:.3f}s status", [p.exchange_status() for p in pools],
          flush=True)
