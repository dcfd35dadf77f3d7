"""Where the time of the host-buffer step goes, and how stable it is (run on the GPU box).
Per-step wall times (min / median / p90 / max) of envpool_b200.make().step(numpy) and of the
bare C-ABI send+recv, with and without binding the process to the GPU's NUMA node."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def stats(v):
    v = np.sort(np.asarray(v)) * 1e6
    return {"min": round(float(v[0]), 1), "med": round(float(v[len(v) // 2]), 1),
            "p90": round(float(v[int(len(v) * 0.9)]), 1), "max": round(float(v[-1]), 1),
            "mean": round(float(v.mean()), 1)}


def main():
    bind = len(sys.argv) > 1 and sys.argv[1] == "bind"
    info = {"bind": bind}
    if bind:
        info["numa"] = bench.bind_to_gpu_numa(0)
    info["affinity_cpus"] = len(os.sched_getaffinity(0))
    import torch  # noqa: F401

    import envpool_b200

    N = 65536
    env = envpool_b200.make("CartPole-v1", env_type="gymnasium", num_envs=N, seed=0)
    env.reset()
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 2, size=(16, N)).astype(np.int32)
    for rep in range(3):
        for t in range(30):
            env.step(acts[t % 16])
        ts = []
        for t in range(200):
            t0 = time.perf_counter()
            env.step(acts[t % 16])
            ts.append(time.perf_counter() - t0)
        info[f"step_rep{rep}"] = stats(ts)
    conv = env._from(acts[0], None)
    ts_send, ts_recv, ts_to = [], [], []
    for t in range(200):
        t0 = time.perf_counter()
        env._send(conv)
        t1 = time.perf_counter()
        st = env._recv()
        t2 = time.perf_counter()
        env._to(st, False, True)
        t3 = time.perf_counter()
        ts_send.append(t1 - t0)
        ts_recv.append(t2 - t1)
        ts_to.append(t3 - t2)
    info["_send"] = stats(ts_send)
    info["_recv"] = stats(ts_recv)
    info["_to"] = stats(ts_to)
    print(json.dumps(info))


if __name__ == "__main__":
    main()
