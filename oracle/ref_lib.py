"""TEST INFRASTRUCTURE ONLY: ctypes wrapper over oracle/_ref/libenvpool_ref.so.

That library is the reference's own AsyncEnvPool + env headers compiled from
/root/reference (oracle/Makefile `ref`, oracle/ref_harness/ref_driver.cc).  Used
to (1) generate tests/golden/*.npz, (2) pin the C restatement (ep_oracle.c) and
(3) time the reference CPU thread pool for bench.py's reference arm.  The product
package envpool_b200 never imports this module.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libenvpool_ref.so")

# (action dtype, action trailing dims) per reference env; state keys after the 8
# common ones (envpool/core/env_spec.h:37-43) with dtype and trailing shape.
COMMON_KEYS = [
    ("info:env_id", np.int32, ()),
    ("info:players.env_id", np.int32, ()),
    ("elapsed_step", np.int32, ()),
    ("done", np.bool_, ()),
    ("reward", np.float32, ()),
    ("discount", np.float32, ()),
    ("step_type", np.int32, ()),
    ("trunc", np.bool_, ()),
]
ENV_TABLE = {
    "CartPole": dict(act=(np.int32, ()), keys=[("obs", np.float32, (4,))]),
    "Pendulum": dict(act=(np.float32, (1,)), keys=[("obs", np.float32, (3,))]),
    "Acrobot": dict(act=(np.int32, ()), keys=[("obs", np.float32, (6,)),
                                             ("info:state", np.float32, (2,))]),
    "MountainCar": dict(act=(np.int32, ()), keys=[("obs", np.float32, (2,))]),
    "MountainCarContinuous": dict(act=(np.float32, (1,)),
                                  keys=[("obs", np.float32, (2,))]),
    "FrozenLake": dict(act=(np.int32, ()), keys=[("obs", np.int32, ())]),
    "Catch": dict(act=(np.int32, ()), keys=[("obs", np.float32, (10, 5))]),
    "Taxi": dict(act=(np.int32, ()), keys=[("obs", np.int32, ())]),
    "NChain": dict(act=(np.int32, ()), keys=[("obs", np.int32, ())]),
    "CliffWalking": dict(act=(np.int32, ()), keys=[("obs", np.int32, ()),
                                                  ("info:prob", np.float32, ())]),
    "Blackjack": dict(act=(np.int32, ()), keys=[("obs", np.int32, (3,))]),
}


def available() -> bool:
    return os.path.exists(REF_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(REF_SO)
        L.ref_create.restype = ctypes.c_void_p
        L.ref_create.argtypes = [ctypes.c_char_p] + [ctypes.c_int] * 5
        L.ref_destroy.argtypes = [ctypes.c_void_p]
        L.ref_reset.argtypes = [ctypes.c_void_p]
        L.ref_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ref_num_keys.argtypes = [ctypes.c_void_p]
        L.ref_key_bytes.restype = ctypes.c_uint64
        L.ref_key_bytes.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.ref_copy.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.ref_bench.restype = ctypes.c_double
        L.ref_bench.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


class RefPool:
    """The reference AsyncEnvPool<Env> in sync mode (batch_size == num_envs)."""

    def __init__(self, task, num_envs, seed=42, max_episode_steps=-1, iopt=-1,
                 num_threads=0):
        self.task = task
        self.n = num_envs
        self.table = ENV_TABLE[task]
        self.h = lib().ref_create(task.encode(), num_envs, num_threads, seed,
                                  max_episode_steps, iopt)
        if not self.h:
            raise RuntimeError(f"ref_create({task}) failed")
        self.keys = COMMON_KEYS + self.table["keys"]

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _collect(self):
        out = {}
        assert lib().ref_num_keys(self.h) == len(self.keys)
        for k, (name, dt, shape) in enumerate(self.keys):
            arr = np.empty((self.n,) + tuple(shape), dtype=dt)
            assert lib().ref_key_bytes(self.h, k) == arr.nbytes, (name, arr.nbytes)
            lib().ref_copy(self.h, k, arr.ctypes.data)
            out[name] = arr
        return out

    def reset(self):
        lib().ref_reset(self.h)
        return self._collect()

    def step(self, action):
        dt, shape = self.table["act"]
        a = np.ascontiguousarray(action, dtype=dt).reshape((self.n,) + shape)
        lib().ref_step(self.h, a.ctypes.data)
        return self._collect()

    def bench(self, actions, warmup, steps):
        """actions: [T, N, ...] stream; returns seconds for `steps` timed steps."""
        dt, shape = self.table["act"]
        a = np.ascontiguousarray(actions, dtype=dt)
        return lib().ref_bench(self.h, a.ctypes.data, a.shape[0], warmup, steps)
