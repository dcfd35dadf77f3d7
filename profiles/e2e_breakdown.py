"""Where the time of one host-buffer step goes (run on the GPU box).  Prints a table; used to
decide what to optimise in the make().step(numpy) path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import envpool_b200  # noqa: E402


def timeit(fn, n=300):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    env = envpool_b200.make("CartPole-v1", env_type="gymnasium", num_envs=N, seed=0)
    env.reset()
    a = np.random.default_rng(0).integers(0, 2, size=N).astype(np.int32)
    conv = env._from(a, None)
    print(f"N={N}")
    print("step() total              %8.1f us" % timeit(lambda: env.step(a)))
    print("_from (python convert)    %8.1f us" % timeit(lambda: env._from(a, None)))

    def sr():
        env._send(conv)
        return env._recv()

    print("_send + _recv (pybind)    %8.1f us" % timeit(sr))
    st = sr()
    print("_to (python unflatten)    %8.1f us" % timeit(lambda: env._to(st, False, True)))
    from envpool_b200 import _capi

    dp = env.device_pool
    import ctypes

    ids = np.arange(N, dtype=np.int32)

    def c_send_recv():
        _capi._check(dp.lib.epb_send(dp.h, a.ctypes.data, ids.ctypes.data, N))
        slab, n, r0 = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        _capi._check(dp.lib.epb_recv_slab_ex(dp.h, ctypes.byref(slab), ctypes.byref(r0),
                                             ctypes.byref(n)))
        _capi._check(dp.lib.epb_release_slab(dp.h, slab))

    print("epb_send+recv (C ABI)     %8.1f us" % timeit(c_send_recv))

    def c_send_only():
        _capi._check(dp.lib.epb_send(dp.h, a.ctypes.data, ids.ctypes.data, N))

    t_send = timeit(c_send_only, 100)
    dp.sync()
    # drain
    for _ in range(120):
        slab, n, r0 = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        dp.lib.epb_recv_slab_ex(dp.h, ctypes.byref(slab), ctypes.byref(r0), ctypes.byref(n))
        dp.lib.epb_release_slab(dp.h, slab)
    print("epb_send host time only   %8.1f us (enqueue; GPU work overlaps)" % t_send)
    import torch

    x = torch.empty(dp.slab_bytes, dtype=torch.uint8, device="cuda")
    h = torch.empty(dp.slab_bytes, dtype=torch.uint8).pin_memory()

    def d2h():
        h.copy_(x, non_blocking=True)
        torch.cuda.synchronize()

    print("bare D2H of one slab      %8.1f us (%d bytes)" % (timeit(d2h), dp.slab_bytes))


if __name__ == "__main__":
    main()
